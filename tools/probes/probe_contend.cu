// Does tcgen05.ld traffic slow tcgen05.mma (shared TMEM port)?  One CTA per SM: warp 8 lane 0 issues
// back-to-back 128x128x16 bf16 MMAs (SS or TS form) into TMEM columns [0,128); 0/4/8 other warps
// stream tcgen05.ld over columns [256,512).  Reports clocks per MMA and ld bytes/clk.  Stand-alone.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "dcl_ptx.cuh"
using namespace dcl;

__global__ void __launch_bounds__(288, 1)
k_contend(unsigned long long* out, float* sink, int n_mma, int ld_warps, int ld_iters, int ts_form, int st_too) {
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
        if (lane == 0 && n_mma > 0) {
            const uint32_t idesc_ss = umma_idesc_bf16(128, 128, 0, 0), idesc_ts = umma_idesc_bf16(128, 128, 0, 1);
            const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + kTileBytes);
            for (int i = 0; i < n_mma; ++i) {
                const int k = i & 7;
                if (ts_form) umma_ts(tmem, tmem + 128 + k * 8, ftile_desc_mnmajor(sB, k), idesc_ts, 1);
                else umma_ss(tmem, ftile_desc_kmajor(sA, k), ftile_desc_kmajor(sB, k), idesc_ss, 1);
            }
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            out[blockIdx.x * 2] = clock64() - t0;
        }
    } else if (warp < ld_warps) {
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t cbase = 256 + (warp >> 2) * 128;
        uint32_t acc = 0;
        for (int it = 0; it < ld_iters; ++it) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_off + cbase + (it & 3) * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
            if (st_too) {
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = acc + j;
                tmem_st16(tmem + lane_off + cbase + (it & 3) * 16, w);
            }
        }
        if (st_too) tmem_st_wait();
        if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = clock64() - t0;
        if (acc == 0x12345u) sink[0] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

// MMA stream (warp 8) + bulk-load stream (warp 7 lane 0: 32 KiB cp.async.bulk, `depth` in flight)
__global__ void __launch_bounds__(288, 1)
k_tma(unsigned long long* out, const uint8_t* src, size_t src_bytes, int n_mma, int n_loads, int depth, int ts_form) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar, lbar[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * kTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 8) tmem_alloc<512>(smem_u32(&slot));
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&lbar[i]), 1);
        mbar_fence_init();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    unsigned long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0 && n_mma > 0) {
            const uint32_t idesc_ss = umma_idesc_bf16(128, 128, 0, 0), idesc_ts = umma_idesc_bf16(128, 128, 0, 1);
            const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + kTileBytes);
            for (int i = 0; i < n_mma; ++i) {
                const int k = i & 7;
                if (ts_form) umma_ts(tmem, tmem + 128 + k * 8, ftile_desc_mnmajor(sB, k), idesc_ts, 1);
                else umma_ss(tmem, ftile_desc_kmajor(sA, k), ftile_desc_kmajor(sB, k), idesc_ss, 1);
            }
            tc_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            out[blockIdx.x * 2] = clock64() - t0;
        }
    } else if (warp == 7) {
        if (lane == 0 && n_loads > 0) {
            const size_t ntiles = src_bytes / kTileBytes;
            size_t tile = (static_cast<size_t>(blockIdx.x) * 7919u) % ntiles;
            for (int i = 0; i < n_loads + depth; ++i) {
                const int sl = i % depth;
                if (i >= depth) mbar_wait(smem_u32(&lbar[sl]), ((i / depth) - 1) & 1);
                if (i < n_loads) {
                    mbar_arrive_expect_tx(smem_u32(&lbar[sl]), kTileBytes);
                    tma_bulk_g2s(smem_u32(smem + (2 + sl) * kTileBytes), src + tile * kTileBytes, kTileBytes, smem_u32(&lbar[sl]));
                    tile = (tile + 149) % ntiles;
                }
            }
            out[blockIdx.x * 2 + 1] = clock64() - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

void run_tma(const char* name, int n_mma, int n_loads, int depth, int ts) {
    unsigned long long* d; uint8_t* src;
    const size_t src_bytes = 16u << 20;      // 16 MiB: L2 resident like the anchor set at N = 65536
    cudaMalloc(&d, 148 * 16); cudaMalloc(&src, src_bytes);
    cudaMemset(d, 0, 148 * 16); cudaMemset(src, 0x3c, src_bytes);
    const int smem = (2 + 4) * kTileBytes;
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_tma<<<148, 288, smem>>>(d, src, src_bytes, n_mma ? 64 : 0, n_loads ? 16 : 0, depth, ts);
    cudaDeviceSynchronize();
    k_tma<<<148, 288, smem>>>(d, src, src_bytes, n_mma, n_loads, depth, ts);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cm = 0, cl = 0;
    for (int i = 0; i < 148; ++i) { cm += h[2 * i]; cl += h[2 * i + 1]; }
    cm /= 148; cl /= 148;
    printf("%-44s", name);
    if (n_mma) printf("  %7.1f clk/MMA", cm / n_mma);
    if (n_loads) printf("  load %6.1f B/clk/SM (%.0f clk per 32 KiB tile)", (double)n_loads * kTileBytes / cl, cl / n_loads);
    printf("  [%s]\n", cudaGetErrorString(e));
    cudaFree(d); cudaFree(src);
}

void run(const char* name, int n_mma, int ld_warps, int ld_iters, int ts, int st) {
    unsigned long long* d; float* s;
    cudaMalloc(&d, 148 * 16); cudaMalloc(&s, 4);
    cudaMemset(d, 0, 148 * 16);
    const int smem = 2 * kTileBytes;
    cudaFuncSetAttribute(k_contend, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_contend<<<148, 288, smem>>>(d, s, n_mma ? 64 : 0, ld_warps, ld_warps ? 64 : 0, ts, st);
    cudaDeviceSynchronize();
    k_contend<<<148, 288, smem>>>(d, s, n_mma, ld_warps, ld_iters, ts, st);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cm = 0, cl = 0;
    for (int i = 0; i < 148; ++i) { cm += h[2 * i]; cl += h[2 * i + 1]; }
    cm /= 148; cl /= 148;
    printf("%-44s", name);
    if (n_mma) printf("  %7.1f clk/MMA", cm / n_mma);
    if (ld_warps) printf("  ld %6.1f B/clk/SM (%.0f clk)", (double)ld_warps * 32 * 128 * ld_iters / cl, cl);
    printf("  [%s]\n", cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    run("MMA SS alone", 4096, 0, 0, 0, 0);
    run("MMA TS alone", 4096, 0, 0, 1, 0);
    run("ld alone, 4 warps", 0, 4, 4096, 0, 0);
    run("ld alone, 8 warps", 0, 8, 4096, 0, 0);
    run("MMA SS + ld 4 warps", 4096, 4, 4096, 0, 0);
    run("MMA SS + ld 8 warps", 4096, 8, 8192, 0, 0);
    run("MMA TS + ld 8 warps", 4096, 8, 8192, 1, 0);
    run("MMA SS + ld+st 8 warps", 4096, 8, 8192, 0, 1);
    run("MMA SS + ld 8 warps (few ld: 1/4 duty)", 4096, 8, 1024, 0, 0);
    run_tma("bulk loads alone, depth 1", 0, 1024, 1, 0);
    run_tma("bulk loads alone, depth 2", 0, 1024, 2, 0);
    run_tma("bulk loads alone, depth 4", 0, 1024, 4, 0);
    run_tma("MMA SS + bulk loads depth 4", 8192, 1024, 4, 0);
    run_tma("MMA TS + bulk loads depth 4", 8192, 1024, 4, 1);
    run_tma("MMA SS + bulk loads depth 2", 8192, 1024, 2, 0);
    return 0;
}
