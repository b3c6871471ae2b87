// Does mbarrier polling by the epilogue warps slow the tensor pipe?  MMA thread: 8 x (128x128x16 TS or SS MMA) per
// tile, then tcgen05.commit to a per-buffer barrier.  8 epilogue warps wait on that barrier (MODE 0: all lanes
// try_wait in a loop; 1: lane 0 only + __syncwarp; 2: all lanes, try_wait with a 20 us suspend hint), read the tile
// with 4 tcgen05.ld and do 3 FFMA2 per pair.  Reports clocks per tile of 8 MMAs (512 = tensor floor).  Stand-alone.
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
__device__ __forceinline__ uint32_t try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok;
}
template <int MODE>
__device__ __forceinline__ void wait_mode(uint32_t bar, uint32_t parity) {
    if (MODE == 0) {
        mbar_wait(bar, parity);
    } else if (MODE == 1) {
        if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
        __syncwarp();
    } else {
        while (!try_wait_hint(bar, parity, 20000)) {}
    }
}

template <int MODE, bool kTS, bool kEpi>
__global__ void __launch_bounds__(288, 1) k_poll(unsigned long long* out, float* sink, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * kTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 8) tmem_alloc<512>(smem_u32(&slot));
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); }
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
            for (int i = 0; i < n_tiles; ++i) {
                // no back-pressure: the epilogue is faster than the MMA stream and simply follows it
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (kTS) umma_ts(tmem + 128 + (i & 1) * 128, tmem + k * 8, dB[k], idesc, k > 0);
                    else umma_ss(tmem + 128 + (i & 1) * 128, dA[k], dB[k], idesc, k > 0);
                }
                tc_commit(smem_u32(&bars[i & 1]));
            }
            tc_commit(smem_u32(&bars[2]));
            mbar_wait(smem_u32(&bars[2]), 0);
            out[blockIdx.x] = clock64() - t0;
        }
    } else {
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
        const unsigned long long c1 = 0x3f8000003f800000ull;
        for (int i = (warp >> 2); i < n_tiles; i += 2) {          // group g takes tiles of parity g
            wait_mode<MODE>(smem_u32(&bars[i & 1]), (i >> 1) & 1);
            tc_fence_after();
            if (kEpi) {
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + 128 + (i & 1) * 128 + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        unsigned long long s = (static_cast<unsigned long long>(v[2 * j + 1]) << 32) | v[2 * j];
                        unsigned long long e = ffma2(s, s, c1);
                        e = ffma2(e, s, c1);
                        acc[j & 3] = ffma2(e, s, acc[j & 3]);
                    }
                }
            }
        }
        if ((acc[0] ^ acc[1] ^ acc[2] ^ acc[3]) == 0x12345ull) sink[0] = 1.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

template <int MODE, bool kTS, bool kEpi>
void run(const char* name) {
    unsigned long long* d; float* s;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
    const int smem = 2 * kTileBytes;
    cudaFuncSetAttribute(k_poll<MODE, kTS, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_poll<MODE, kTS, kEpi><<<148, 288, smem>>>(d, s, 8);
    cudaDeviceSynchronize();
    const int n = 1024;
    k_poll<MODE, kTS, kEpi><<<148, 288, smem>>>(d, s, n);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-64s %7.1f clk per tile (8 MMAs)  [%s]\n", name, c / n, cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    run<0, false, false>("SS, epilogue warps only wait (all lanes poll)");
    run<1, false, false>("SS, epilogue warps only wait (lane 0 polls)");
    run<2, false, false>("SS, epilogue warps only wait (20 us suspend hint)");
    run<0, false, true>("SS, wait (all lanes) + ld + FFMA2");
    run<1, false, true>("SS, wait (lane 0) + ld + FFMA2");
    run<2, false, true>("SS, wait (hint) + ld + FFMA2");
    run<0, true, true>("TS, wait (all lanes) + ld + FFMA2");
    run<1, true, true>("TS, wait (lane 0) + ld + FFMA2");
    run<2, true, true>("TS, wait (hint) + ld + FFMA2");
    return 0;
}
