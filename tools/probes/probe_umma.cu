// Stand-alone hardware probe (no torch): validates the F-tile image + UMMA descriptor
// conventions in dcl_ptx.cuh on a real B200 before the production kernels rely on them.
//   phase 1: S = F_I * F_J^T            (SS form, both operands K-major SW128, via cp.async.bulk)
//   phase 2: O = bf16(S*scale) * F_J    (TS form: A from TMEM, B = same smem tile, MN-major)
// Usage: probe_umma [variant]   variant 0 = library convention; 1..3 = alternates for diagnosis.
// Exit code 0 iff both phases match a host fp64 reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "dcl_ptx.cuh"

using namespace dcl;

struct ProbeCfg {
    uint32_t mn_lbo, mn_sbo, mn_kstep;  // MN-major B descriptor parameters (bytes)
    uint32_t idesc_ts;                  // instruction descriptor for phase 2
    float scale;
};

__global__ void __launch_bounds__(192, 1)
probe_kernel(const uint8_t* __restrict__ tiles, float* __restrict__ S_out, float* __restrict__ O_out,
             ProbeCfg cfg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // [0,32K) F_I, [32K,64K) F_J, then barriers
    uint8_t* sI = smem;
    uint8_t* sJ = smem + kTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t bar_load = smem_u32(&bars[0]);
    const uint32_t bar_s = smem_u32(&bars[1]);
    const uint32_t bar_p = smem_u32(&bars[2]);
    const uint32_t bar_o = smem_u32(&bars[3]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        tmem_alloc<512>(smem_u32(tmem_slot));
    }
    if (threadIdx.x == 32) {
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 128);
        mbar_init(bar_o, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tS = tmem, tP = tmem + 128, tO = tmem + 256;

    if (warp == 0 && lane == 0) {
        mbar_arrive_expect_tx(bar_load, 2 * kTileBytes);
        tma_bulk_g2s(smem_u32(sI), tiles, kTileBytes, bar_load);
        tma_bulk_g2s(smem_u32(sJ), tiles + kTileBytes, kTileBytes, bar_load);
    } else if (warp == 1 && lane == 0) {
        mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            umma_ss(tS, ftile_desc_kmajor(smem_u32(sI), k), ftile_desc_kmajor(smem_u32(sJ), k), idesc,
                    k > 0);
        tc_commit(bar_s);
        // phase 2
        mbar_wait(bar_p, 0);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t bd = umma_smem_desc(smem_u32(sJ) + k * cfg.mn_kstep, cfg.mn_lbo, cfg.mn_sbo);
            umma_ts(tO, tP + k * 8, bd, cfg.idesc_ts, k > 0);
        }
        tc_commit(bar_o);
    } else if (warp >= 2) {
        const int q = warp & 3;              // TMEM lane quadrant this warp may touch
        const int row = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        mbar_wait(bar_s, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tS + lane_off + c0, v);
            tmem_ld_wait();
            uint32_t p[16];
#pragma unroll
            for (int j = 0; j < 32; ++j) S_out[row * 128 + c0 + j] = __uint_as_float(v[j]);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2 * j]) * cfg.scale,
                                                         __uint_as_float(v[2 * j + 1]) * cfg.scale);
                p[j] = *reinterpret_cast<uint32_t*>(&b);
            }
            tmem_st16(tP + lane_off + c0 / 2, p);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_p);
        mbar_wait(bar_o, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tO + lane_off + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) O_out[row * 128 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                  \
        }                                                                              \
    } while (0)

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    ProbeCfg cfg;
    cfg.scale = 1.0f / 64.0f;
    cfg.mn_lbo = kHalfBytes; cfg.mn_sbo = 1024; cfg.mn_kstep = 2048;
    cfg.idesc_ts = umma_idesc_bf16(128, 128, 0, 1);
    if (variant == 1) { cfg.mn_lbo = 1024; cfg.mn_sbo = kHalfBytes; }
    if (variant == 2) { cfg.idesc_ts = umma_idesc_bf16(128, 128, 0, 0); }  // wrong on purpose: K-major B
    if (variant == 3) { cfg.mn_lbo = kHalfBytes; cfg.mn_sbo = 1024; cfg.mn_kstep = 32; }

    const int R = 256;
    std::vector<float> F(R * kDim);
    uint32_t s = 12345u;
    for (auto& x : F) { s = s * 1664525u + 1013904223u; x = ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; }
    std::vector<uint8_t> img(2 * kTileBytes);
    std::vector<float> Fb(R * kDim);
    for (int r = 0; r < R; ++r)
        for (int d = 0; d < kDim; ++d) {
            __nv_bfloat16 b = __float2bfloat16(F[r * kDim + d]);
            Fb[r * kDim + d] = __bfloat162float(b);
            uint32_t off = (r / 128) * kTileBytes + ftile_offset(r % 128, d);
            *reinterpret_cast<__nv_bfloat16*>(&img[off]) = b;
        }
    uint8_t* d_img; float *d_S, *d_O;
    CK(cudaMalloc(&d_img, img.size()));
    CK(cudaMalloc(&d_S, 128 * 128 * 4));
    CK(cudaMalloc(&d_O, 128 * 128 * 4));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_S, 0xff, 128 * 128 * 4));
    CK(cudaMemset(d_O, 0xff, 128 * 128 * 4));
    const int smem = 2 * kTileBytes + 256;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 192, smem>>>(d_img, d_S, d_O, cfg);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> S(128 * 128), O(128 * 128);
    CK(cudaMemcpy(S.data(), d_S, S.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(O.data(), d_O, O.size() * 4, cudaMemcpyDeviceToHost));

    double errS = 0, maxS = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            double ref = 0;
            for (int d = 0; d < kDim; ++d) ref += (double)Fb[i * kDim + d] * Fb[(128 + j) * kDim + d];
            errS = fmax(errS, fabs(ref - S[i * 128 + j]));
            maxS = fmax(maxS, fabs(ref));
        }
    double errO = 0, maxO = 0;
    std::vector<float> P(128 * 128);
    for (int i = 0; i < 128 * 128; ++i) P[i] = __bfloat162float(__float2bfloat16(S[i] * cfg.scale));
    for (int i = 0; i < 128; ++i)
        for (int d = 0; d < kDim; ++d) {
            double ref = 0;
            for (int k = 0; k < 128; ++k) ref += (double)P[i * 128 + k] * Fb[(128 + k) * kDim + d];
            errO = fmax(errO, fabs(ref - O[i * 128 + d]));
            maxO = fmax(maxO, fabs(ref));
        }
    printf("variant %d: S max|err| %.3e (max|ref| %.3e)   O max|err| %.3e (max|ref| %.3e)\n", variant,
           errS, maxS, errO, maxO);
    printf("  S[0][0..3] = %g %g %g %g   O[0][0..3] = %g %g %g %g\n", S[0], S[1], S[2], S[3], O[0], O[1],
           O[2], O[3]);
    bool okS = errS <= 1e-4 * fmax(maxS, 1.0), okO = errO <= 1e-4 * fmax(maxO, 1.0);
    printf("  phase1(SS,K-major) %s   phase2(TS,MN-major B) %s\n", okS ? "PASS" : "FAIL",
           okO ? "PASS" : "FAIL");
    return (okS && okO) ? 0 : 1;
}
