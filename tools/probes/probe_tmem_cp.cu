// tcgen05.cp probe: can an F-tile (128 rows x 128 bf16, two 128-byte-swizzled K-major panels, as the gather kernel
// writes it and TMA lands it) be copied from shared memory into the TMEM layout the TS-form MMA expects for its A
// operand (lane = row, 32-bit column c = bf16 pair (2c, 2c+1)) by eight tcgen05.cp.128x256b, one per K = 16 slice,
// with the same descriptors the MMA uses for those slices?  If so the sweeps can stage their row blocks through TMA +
// cp instead of global loads + tcgen05.st in the epilogue warps.  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I. probe_tmem_cp.cu -o probe_tmem_cp && ./probe_tmem_cp
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "dcl_ptx.cuh"
using namespace dcl;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_probe(const uint32_t* __restrict__ tile_words, uint32_t* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kTileBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = tile_words[i];
    if (warp == 0) tmem_alloc<128>(smem_u32(&slot));
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t sT = smem_u32(smem);
#pragma unroll
        for (int k = 0; k < 8; ++k) tmem_cp_128x256b(tmem + k * 8, ftile_desc_kmajor(sT, k));
        tc_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    uint32_t v[32];
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld32(tmem + lane_off + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

int main() {
    // logical tile: element (row r, channel d) = bf16 bits (r << 7 | d) -- every element distinct
    std::vector<uint32_t> words(kTileBytes / 4);
    auto put = [&](int r, int d, uint16_t bits) {
        const int panel = d / 64, dd = d % 64, chunk = dd / 8, e = dd % 8;
        const size_t byte = static_cast<size_t>(panel) * kHalfBytes + static_cast<size_t>(r) * 128 + ((chunk ^ (r & 7)) << 4) + e * 2;
        reinterpret_cast<uint16_t*>(words.data())[byte / 2] = bits;
    };
    for (int r = 0; r < 128; ++r)
        for (int d = 0; d < 128; ++d) put(r, d, static_cast<uint16_t>((r << 7) | d));
    uint32_t *d_in, *d_out;
    cudaMalloc(&d_in, kTileBytes);
    cudaMalloc(&d_out, 128 * 64 * 4);
    cudaMemcpy(d_in, words.data(), kTileBytes, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0xff, 128 * 64 * 4);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes + 1024);
    k_probe<<<1, 128, kTileBytes + 1024>>>(d_in, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> got(128 * 64);
    cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 64; ++c) {
            const uint32_t want = static_cast<uint32_t>((r << 7) | (2 * c)) | (static_cast<uint32_t>((r << 7) | (2 * c + 1)) << 16);
            if (got[r * 64 + c] != want) {
                if (bad < 8) printf("lane %3d col %2d: got %08x (row %u ch %u | row %u ch %u) want %08x\n", r, c, got[r * 64 + c],
                                    (got[r * 64 + c] & 0xffff) >> 7, got[r * 64 + c] & 127, (got[r * 64 + c] >> 16) >> 7,
                                    (got[r * 64 + c] >> 16) & 127, want);
                ++bad;
            }
        }
    printf("tcgen05.cp.128x256b x8 with the MMA's K-slice descriptors: %ld of %d words differ from the TS-form A layout -> %s\n",
           bad, 128 * 64, bad ? "MISMATCH" : "OK");
    return bad ? 2 : 0;
}
