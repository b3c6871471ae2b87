// Why do back-to-back tcgen05.mma run at ~110 clk instead of 64 inside the sweep kernels?  One CTA per SM:
// warp 8 lane 0 issues hoisted-descriptor 128x128x16 SS MMAs into TMEM columns [0,256); 8 other warps run an
// epilogue-like loop with a selectable mix: tcgen05.ld of a 32-column chunk (columns [256,512)) and/or `nf`
// packed FFMA2 per loaded pair.  Reports clocks per MMA.  Stand-alone.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "dcl_ptx.cuh"
using namespace dcl;

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <bool kLd, int kNf>
__global__ void __launch_bounds__(288, 1) k_mix(unsigned long long* out, float* sink, int n_mma, int ep_iters, int ep_warps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * kTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 8) tmem_alloc<512>(smem_u32(&slot));
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    unsigned long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
            const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + kTileBytes);
            uint64_t dA[8], dB[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { dA[k] = ftile_desc_kmajor(sA, k); dB[k] = ftile_desc_kmajor(sB, k); }
            for (int i = 0; i < n_mma / 8; ++i) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ss(tmem + (i & 1) * 128, dA[k], dB[k], idesc, k > 0);
            }
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            out[blockIdx.x * 2] = clock64() - t0;
        }
    } else if (warp < ep_warps) {
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t cbase = 256 + (warp >> 2) * 128;
        unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
        const unsigned long long c1 = 0x3f8000003f800000ull;
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0x3f000000u + j + threadIdx.x;
        for (int it = 0; it < ep_iters; ++it) {
            if (kLd) {
                tmem_ld32(tmem + lane_off + cbase + (it & 3) * 32, v);
                tmem_ld_wait();
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                unsigned long long s = (static_cast<unsigned long long>(v[2 * j + 1]) << 32) | v[2 * j];
                unsigned long long e = s;
#pragma unroll
                for (int f = 0; f < kNf; ++f) e = ffma2(e, s, c1);
                acc[j & 3] ^= e;
            }
        }
        if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = clock64() - t0;
        if ((acc[0] ^ acc[1] ^ acc[2] ^ acc[3]) == 0x12345ull) sink[0] = 1.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

template <bool kLd, int kNf>
void run(const char* name, int ep_warps) {
    unsigned long long* d; float* s;
    cudaMalloc(&d, 148 * 16); cudaMalloc(&s, 4);
    cudaMemset(d, 0, 148 * 16);
    const int smem = 2 * kTileBytes;
    cudaFuncSetAttribute(k_mix<kLd, kNf>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_mix<kLd, kNf><<<148, 288, smem>>>(d, s, 64, 16, ep_warps);
    cudaDeviceSynchronize();
    const int n_mma = 8192;
    // epilogue iterations sized to outlast the MMA stream in every configuration
    k_mix<kLd, kNf><<<148, 288, smem>>>(d, s, n_mma, 40000, ep_warps);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cm = 0, ce = 0;
    for (int i = 0; i < 148; ++i) { cm += h[2 * i]; ce += h[2 * i + 1]; }
    cm /= 148; ce /= 148;
    printf("%-52s ep_warps=%d  %7.1f clk/MMA   epilogue: %7.1f clk per 32-col chunk  [%s]\n", name, ep_warps, cm / n_mma,
           ep_warps ? ce / 40000 : 0.0, cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    run<false, 0>("MMA alone", 0);
    run<false, 3>("MMA + FFMA2 only (3 per pair)", 8);
    run<false, 6>("MMA + FFMA2 only (6 per pair)", 8);
    run<true, 0>("MMA + tcgen05.ld only (back to back)", 8);
    run<true, 3>("MMA + ld + 3 FFMA2 per pair (fwd-like)", 8);
    run<true, 6>("MMA + ld + 6 FFMA2 per pair (bwd-like)", 8);
    run<true, 3>("MMA + ld + 3 FFMA2 per pair, 4 warps", 4);
    run<true, 12>("MMA + ld + 12 FFMA2 per pair (low ld duty)", 8);
    return 0;
}
