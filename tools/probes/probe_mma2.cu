// tcgen05.mma issue-rate probe with descriptors hoisted out of the loop (as the real kernels do):
// clocks per 128x128x16 / 128x256x16 SS MMA and per 128x128x16 TS MMA, alone and interleaved
// (8 x SS then 8 x TS: the backward's per-tile pattern).  One CTA per SM.  Stand-alone.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "dcl_ptx.cuh"
using namespace dcl;

// MODE 0: SS N=128   1: SS N=256   2: TS N=128   3: 8xSS(N=128) + 8xTS   4: 4xSS(N=256, two tiles) + 2x8xTS
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_mma(unsigned long long* out, int n_groups) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * kTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + kTileBytes);   // B: up to 64 KiB (256 rows)
        const uint32_t id128 = umma_idesc_bf16(128, 128, 0, 0), id256 = umma_idesc_bf16(128, 256, 0, 0),
                       idts = umma_idesc_bf16(128, 128, 0, 1);
        uint64_t dA[8], dB[8], dBm[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            dA[k] = ftile_desc_kmajor(sA, k);
            // N = 256: each 64-channel panel holds 256 rows x 128 B = 32 KiB
            dB[k] = (MODE == 1 || MODE == 4) ? umma_smem_desc(sB + (k >> 2) * 32768 + (k & 3) * 32, 16, 1024)
                                             : ftile_desc_kmajor(sB, k);
            dBm[k] = ftile_desc_mnmajor(sB, k);
        }
        unsigned long long t0 = clock64();
        for (int g = 0; g < n_groups; ++g) {
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ss(tmem, dA[k], dB[k], id128, k > 0);
            } else if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ss(tmem, dA[k], dB[k], id256, k > 0);
            } else if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tmem, tmem + 256 + k * 8, dBm[k], idts, 1);
            } else if (MODE == 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ss(tmem + 128, dA[k], dB[k], id128, k > 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tmem, tmem + 256 + k * 8, dBm[k], idts, 1);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ss(tmem + 256, dA[k], dB[k], id256, k > 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tmem, tmem + 128 + k * 8, dBm[k], idts, 1);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_ts(tmem, tmem + 192 + k * 8, dBm[k], idts, 1);
            }
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE>
void run(const char* name, double tiles_per_group) {
    unsigned long long* d;
    cudaMalloc(&d, 148 * 8);
    const int smem = 4 * kTileBytes;
    cudaFuncSetAttribute(k_mma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_mma<MODE><<<148, 128, smem>>>(d, 16);
    cudaDeviceSynchronize();
    const int groups = 1024;
    k_mma<MODE><<<148, 128, smem>>>(d, groups);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-52s %8.1f clk/group  = %7.1f clk per 128x128x128 tile-MMA  [%s]\n", name, c / groups,
           c / groups / tiles_per_group, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<0>("SS 128x128x16 x8 (one S tile)", 1);
    run<1>("SS 128x256x16 x8 (two S tiles)", 2);
    run<2>("TS 128x128x16 x8 (one dF tile)", 1);
    run<3>("SS N=128 x8 + TS x8 (backward tile)", 2);
    run<4>("SS N=256 x8 + 2 x TS x8 (two backward tiles)", 4);
    return 0;
}
