// Per-SM instruction-throughput probe for the epilogue design: lane-ops per clock per SM for the
// candidate instructions (which pipe they land on decides the sweep epilogues).  Stand-alone.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) k_op(float* out, float seed, int iters) {
    float a[CHAINS];
    float b = seed * 1.0001f + threadIdx.x * 1e-6f, c = seed * 0.5f;
    int ia[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + i + threadIdx.x * 1e-3f; ia[i] = __float_as_int(a[i]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) a[i] = a[i] + b;                                   // FADD
            if (OP == 1) a[i] = fmaf(a[i], b, c);                           // FFMA
            if (OP == 2) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));   // FMNMX 2-input
            if (OP == 3) a[i] = fmaxf(fmaxf(a[i], b), c + i);               // FMNMX3 candidate
            if (OP == 4) asm volatile("max.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(__float_as_int(b) + it));  // IMNMX
            if (OP == 5) { if (b + it > a[i]) a[i] = b + it; }              // FSETP + FSEL (+FADD)
            if (OP == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));     // MUFU.EX2
            if (OP == 7) {                                                  // cvt.rn.bf16x2.f32 (pack)
                uint32_t p;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(a[i]), "f"(b));
                a[i] = __uint_as_float(p | 0x3f000000u);
            }
            if (OP == 8) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));     // MUFU.RCP
            if (OP == 9) { ia[i] = max(max(ia[i], __float_as_int(b) + it), __float_as_int(c) + i); }   // VIMNMX3?
            if (OP == 10) a[i] = a[i] * b;                                  // FMUL
            if (OP == 11) { a[i] = fmaf(a[i], b, c); ia[i] = max(ia[i], __float_as_int(a[i])); }  // FFMA + IMNMX mix
            if (OP == 12) { a[i] = fmaf(a[i], b, c); asm volatile("max.f32 %0, %0, %1;" : "+f"(b) : "f"(a[i])); }  // FFMA + FMNMX mix
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + __int_as_float(ia[i]);
    if (s == 12345.678f) out[0] = s + b;
}

template <int OP>
double run(const char* name, int ops_per_inner, int nsm, double ghz_hint) {
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ctas = nsm * 4;   // 4 x 256 threads per SM = 32 warps/SM
    k_op<OP><<<ctas, 256>>>(d, 1.0f, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_op<OP><<<ctas, 256>>>(d, 1.0f, ITERS);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double laneops = (double)ctas * 256 * ITERS * CHAINS * ops_per_inner;
    double per_sm_per_ns = laneops / nsm / (ms * 1e6);
    printf("%-28s %8.3f ms  %7.2f lane-ops/ns/SM  (~%6.1f /clk @%.2f GHz)\n", name, ms, per_sm_per_ns,
           per_sm_per_ns / ghz_hint, ghz_hint);
    cudaFree(d);
    return per_sm_per_ns;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz / 1e6;
    printf("%s, %d SMs, max clock %.3f GHz (per-clk figures assume max clock)\n", p.name, nsm, ghz);
    run<0>("FADD", 1, nsm, ghz);
    run<1>("FFMA", 1, nsm, ghz);
    run<10>("FMUL", 1, nsm, ghz);
    run<2>("max.f32 (2-input)", 1, nsm, ghz);
    run<3>("fmaxf(fmaxf()) (3-input?)", 2, nsm, ghz);
    run<4>("max.s32", 1, nsm, ghz);
    run<9>("max(max()) s32 (3-input?)", 2, nsm, ghz);
    run<5>("FADD+FSETP+FSEL", 1, nsm, ghz);
    run<6>("ex2.approx", 1, nsm, ghz);
    run<8>("rcp.approx", 1, nsm, ghz);
    run<7>("cvt.rn.bf16x2.f32", 1, nsm, ghz);
    run<11>("FFMA + max.s32 pair", 1, nsm, ghz);
    run<12>("FFMA + max.f32 pair", 1, nsm, ghz);
    return 0;
}
