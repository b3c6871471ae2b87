// Epilogue instruction-mix probe (clock64-based, one CTA per SM, 8 or 16 warps): how many clocks per
// warp-instruction the packed fp32 FMA (FFMA2) takes alone and when co-issued with the other pipes the
// polynomial epilogues use (3-input max on ALU, ex2 on MUFU, bf16x2 pack, broadcast LDS.128).  Stand-alone.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pk(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float lo(unsigned long long v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float hi(unsigned long long v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }

#define CH 8
// OP 0: FFMA2 only (8 per inner)        1: FFMA (scalar, 8)           2: 8 FFMA2 + 4 fmax3
//    3: 8 FFMA2 + 2 ex2                 4: 8 FFMA2 + 4 cvt.bf16x2     5: 8 FFMA2 + 4 (2 IADD + PRMT)
//    6: 8 FFMA2 + 4 LDS.128 broadcast   7: 4 cvt.bf16x2 only          8: 8 ex2 only
//    9: 8 FFMA2 + 8 ex2                10: 8 FFMA2 + 4 FADD2
template <int OP>
__global__ void __launch_bounds__(512, 1) k_mix(unsigned long long* out, float* sink, float seed, int iters) {
    __shared__ float4 tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = make_float4(seed, seed * 0.5f, 1.f, 2.f);
    __syncthreads();
    unsigned long long a[CH];
    float m[4], e[CH];
    uint32_t u[4];
    const unsigned long long b = pk(seed * 0.999f, seed * 0.998f), c = pk(0.25f, 0.125f);
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = pk(seed + i, seed - i); e[i] = -0.1f * i - threadIdx.x * 1e-4f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -1e30f; u[i] = i; }
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (OP == 1) {
#pragma unroll
            for (int i = 0; i < CH; ++i) e[i] = fmaf(e[i], seed, 0.25f);
        } else if (OP != 7 && OP != 8) {
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = ffma2(a[i], b, c);
        }
        if (OP == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = fmaxf(fmaxf(m[i], lo(a[2 * i])), hi(a[2 * i + 1]));
        }
        if (OP == 3) {
#pragma unroll
            for (int i = 0; i < 2; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
        }
        if (OP == 8 || OP == 9) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
        }
        if (OP == 4 || OP == 7) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t p;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi(a[i])), "f"(lo(a[i])));
                u[i] ^= p;
            }
        }
        if (OP == 5) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t x = static_cast<uint32_t>(a[i]) + 0x8000u, y = static_cast<uint32_t>(a[i] >> 32) + 0x8000u;
                u[i] ^= __byte_perm(x, y, 0x7632);
            }
        }
        if (OP == 6) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = tab[(it + i) & 63];
                m[i] += v.x;
                u[i] ^= __float_as_uint(v.w);
            }
        }
        if (OP == 10) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned long long d;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a[i]), "l"(a[i + 4]));
                a[i] = d;
            }
        }
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += lo(a[i]) + hi(a[i]) + e[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += m[i] + __uint_as_float(u[i]);
    if (s == 12345.678f) sink[0] = s;
}

template <int OP>
void run(const char* name, int warps) {
    unsigned long long* d; float* s;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
    const int iters = 2048;
    k_mix<OP><<<148, warps * 32>>>(d, s, 1.0f, 64);
    cudaDeviceSynchronize();
    k_mix<OP><<<148, warps * 32>>>(d, s, 1.0f, iters);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    // clocks per inner iteration per SMSP-resident warp set: warps/4 warps share one scheduler
    printf("%-36s warps=%2d  %7.2f clk per inner iteration per warp-slot (x%d warps/SMSP)  [%s]\n", name, warps,
           c / iters / (warps / 4), warps / 4, cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    for (int w : {8, 16}) {
        run<0>("8 FFMA2", w);
        run<1>("8 FFMA", w);
        run<10>("8 FFMA2 + 4 FADD2", w);
        run<2>("8 FFMA2 + 4 fmax3", w);
        run<3>("8 FFMA2 + 2 ex2", w);
        run<9>("8 FFMA2 + 8 ex2", w);
        run<8>("8 ex2", w);
        run<4>("8 FFMA2 + 4 cvt.bf16x2", w);
        run<7>("4 cvt.bf16x2", w);
        run<5>("8 FFMA2 + 4 (2 IADD + PRMT)", w);
        run<6>("8 FFMA2 + 4 LDS.128", w);
    }
    return 0;
}
