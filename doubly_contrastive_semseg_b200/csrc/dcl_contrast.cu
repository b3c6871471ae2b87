// N x N row-normalised contrast (reference utils/loss.py:339-389 pixel term, :175-204 image term)
// as tcgen05/TMEM tile sweeps; the N x N matrix only ever exists as 128x128 fp32 tiles in TMEM.
//
// Math, in raw dot-product units s_ij = f_i . f_j (a_ij = s_ij / T):
//   sweep A : smax_i = max_j s_ij, S1 = sum_j (s_ij - c_i), S2 = sum_j (s_ij - c_i)^2   (c_i = |f_i|^2)
//             -> kappa_i = 1 / max(sqrt(sum_j (s_ij - smax_i)^2), T*1e-12)   so   l_ij = (s_ij - smax_i) kappa_i
//   sweep B : Den_i = sum_{den(i,j)} exp(l_ij),  Bt_i = sum_{den} exp(l_ij) * l_ij*log2(e)
//             den = different label (pixel) | j != i (image)
//   sweep C : over tiles that can hold positives only: P_i, sum_pos lp_ij, sum_pos 1/(E+Den), sum_pos l/(E+Den)
//   finalize: per-row loss, Q_i, R_i and the per-row constants the backward consumes
//   backward: G_ik = dS_ik + dS_ki recomputed per tile -> bf16 in TMEM -> dF_I += G_IJ F_J (TS-form MMA)
//
// Kernel shape: 192 threads = warp 0 (TMA producer lane + TMEM allocator), warp 1 (MMA issuer lane),
// warps 2..5 (epilogue; warp w owns TMEM lanes 32*(w%4)..+31, one thread per anchor row).  Two CTAs
// are co-resident per SM (256 TMEM columns and ~100 KiB of shared memory each) so one CTA's
// MMA/TMA latency hides under the other's CUDA-core epilogue.  Work = the flattened list of
// (row block, column block) tiles cut into gridDim.x equal contiguous ranges; a range that crosses
// a row-block boundary flushes a deterministic partial ("segment") instead of using atomics.
#include <cfloat>
#include <cuda_bf16.h>
#include "dcl_common.cuh"
#include "dcl_ptx.cuh"

namespace dcl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 192;
constexpr int kTmemCols = 256;

// ---------------------------------------------------------------------------------------------
// flattened-range partition of nI x nJ tiles over G CTAs
struct Part {
    long long total;
    int G, nJ;
    __host__ __device__ long long begin(int c) const { return static_cast<long long>(c) * total / G; }
    // CTA whose range contains flat tile index x
    __host__ __device__ int cta_of(long long x) const {
        return static_cast<int>(((x + 1) * G + total - 1) / total - 1);
    }
    __host__ __device__ int first_cta(int I) const { return cta_of(static_cast<long long>(I) * nJ); }
    __host__ __device__ int nseg(int I) const {
        return cta_of(static_cast<long long>(I + 1) * nJ - 1) - first_cta(I) + 1;
    }
};

struct Params {
    const uint8_t* tiles;    // [nJ] F-tiles
    const int32_t* y;        // [nJ*128]
    const float* sqnorm;     // [nJ*128]
    const int2* blk_range;   // [nJ] (min,max) label of the valid rows of a block; (INT_MAX,-1) if none
    const int32_t* blk_nvalid;
    int nJ, rb0, nI, n_valid, mode;
    float T, Tb;
    Part part;               // partition for sweeps A, B and the backward
    int maxseg;
    int splitc;              // sweep C: CTAs per row block
    float4* pA;              // [nI][maxseg][128] (max(s-c), S1, S2, -)
    float2* pB;              // [nI][maxseg][128] (Den, Bt)
    float4* pC;              // [nI][splitc][128] (P, sum lp | sum l, sum inv, sum inv*l)
    float* pD;               // [nI][maxseg][128][128] dF partials
    float4* colA;            // [nJ*128] (a, b, p, q)
    float4* colB;            // [nJ*128] (wn, Den, y bits, 0)
    float* rowloss;          // [nJ*128]
    float* blockloss;        // [nI]
    float* loss_sum;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// tcgen05.wait::ld with the destination registers threaded through as in/out operands, so the
// compiler cannot schedule a use of the asynchronously written registers above the wait.
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                   "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                   "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                   "+r"(v[31])
                 :
                 : "memory");
}

// Walk one 128-column fp32 accumulator row in four 32-column chunks, the next chunk's TMEM load in
// flight while the current one is processed.  fn(c0, v) sees columns [c0, c0+32).
template <class Fn>
__device__ __forceinline__ void for_each_chunk(uint32_t taddr, Fn&& fn) {
    uint32_t va[32], vb[32];
    tmem_ld32(taddr, va);
    tmem_ld_wait_on(va);
    tmem_ld32(taddr + 32, vb);
    fn(0, va);
    tmem_ld_wait_on(vb);
    tmem_ld32(taddr + 64, va);
    fn(32, vb);
    tmem_ld_wait_on(va);
    tmem_ld32(taddr + 96, vb);
    fn(64, va);
    tmem_ld_wait_on(vb);
    fn(96, vb);
}

__device__ __forceinline__ bool ranges_overlap(int2 a, int2 b) { return a.x <= b.y && b.x <= a.y; }

// Row constants derived from sweep A partials: t_ij = fma(s_ij, a, b) = l_ij * log2(e).
struct RowA {
    float a, b, kappa, smax;
};
__device__ __forceinline__ RowA combine_A(const Params& p, int I, int r) {
    const int ns = p.part.nseg(I);
    float mx = -FLT_MAX;
    double S1 = 0.0, S2 = 0.0;
    for (int s = 0; s < ns; ++s) {
        float4 v = p.pA[(static_cast<size_t>(I) * p.maxseg + s) * 128 + r];
        mx = fmaxf(mx, v.x);
        S1 += v.y;
        S2 += v.z;
    }
    const float c = p.sqnorm[(p.rb0 + I) * 128 + r];
    const double d = mx;                                         // smax - c
    double S2s = S2 - 2.0 * d * S1 + static_cast<double>(p.n_valid) * d * d;
    if (!(S2s > 0.0)) S2s = 0.0;
    const float rT = fmaxf(static_cast<float>(sqrt(S2s)), p.T * 1e-12f);   // F.normalize eps, loss.py:366
    RowA o;
    o.kappa = 1.0f / rT;
    o.smax = c + mx;
    o.a = o.kappa * kLog2e;
    o.b = -o.smax * o.a;
    return o;
}
__device__ __forceinline__ float2 combine_B(const Params& p, int I, int r) {
    const int ns = p.part.nseg(I);
    float den = 0.f, bt = 0.f;
    for (int s = 0; s < ns; ++s) {
        float2 v = p.pB[(static_cast<size_t>(I) * p.maxseg + s) * 128 + r];
        den += v.x;
        bt += v.y;
    }
    return make_float2(den, bt);
}

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up (bytes from a 1024-aligned base)
struct SmemSweep {
    static constexpr int kI = 0;
    static constexpr int kJ = kTileBytes;                 // 2 slots
    static constexpr int kY = 3 * kTileBytes;             // 2 x 512 B labels of the column block
    static constexpr int kBar = kY + 1024;                // full[2] empty[2] tfull[2] ifull iempty
    static constexpr int kTmem = kBar + 64;
    static constexpr int kBytes = kTmem + 16 + 1024;      // + alignment slack
};
struct SmemBwd {
    static constexpr int kI = 0;
    static constexpr int kJ = kTileBytes;                 // 2 slots
    static constexpr int kCA = 3 * kTileBytes;            // 2 x 2 KiB colA
    static constexpr int kCB = kCA + 4096;                // 2 x 2 KiB colB
    static constexpr int kBar = kCB + 4096;               // full[2] empty[2] tfull pfull dfull dempty ifull iempty
    static constexpr int kTmem = kBar + 96;
    static constexpr int kBytes = kTmem + 16 + 1024;
};

enum { SWEEP_A = 0, SWEEP_B = 1, SWEEP_C = 2 };

// Tile sequence of one CTA.  A/B/backward: contiguous flat range.  C: the CTA's share of the
// tiles of ONE row block whose label range overlaps the row block's.  All three warp roles run an
// identical copy of this iterator, which is what keeps their barrier phases in step.
template <bool kRelevantOnly>
struct TileIter {
    long long t, t1;
    int nJ, I, J, r, s, splitc;
    int2 rI;
    const int2* blk;
    __device__ TileIter(const Params& p) {
        nJ = p.nJ;
        if (kRelevantOnly) {
            I = blockIdx.x / p.splitc;
            s = blockIdx.x % p.splitc;
            splitc = p.splitc;
            J = 0;
            r = 0;
            blk = p.blk_range;
            rI = blk[p.rb0 + I];
        } else {
            t = p.part.begin(blockIdx.x);
            t1 = p.part.begin(blockIdx.x + 1);
        }
    }
    // next tile -> (I, J); `last_of_seg` = no further tile of this CTA shares the row block
    __device__ bool next(int& oI, int& oJ, bool& last_of_seg) {
        if (kRelevantOnly) {
            while (J < nJ) {
                int j = J++;
                if (ranges_overlap(rI, blk[j])) {
                    if ((r++ % splitc) == s) {
                        oI = I;
                        oJ = j;
                        last_of_seg = false;      // single segment; the caller flushes at the end
                        return true;
                    }
                }
            }
            return false;
        } else {
            if (t >= t1) return false;
            oI = static_cast<int>(t / nJ);
            oJ = static_cast<int>(t % nJ);
            ++t;
            last_of_seg = (t >= t1) || (static_cast<int>(t / nJ) != oI);
            return true;
        }
    }
};

// =============================================================================================
// Sweeps A / B / C
// =============================================================================================
template <int kSweep, int kMode>
__global__ void __launch_bounds__(kThreads, 2) k_sweep(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sI = base + SmemSweep::kI;
    const uint32_t sJ = base + SmemSweep::kJ;
    const uint32_t sY = base + SmemSweep::kY;
    const int32_t* sYg = reinterpret_cast<const int32_t*>(gen + SmemSweep::kY);
    const uint32_t bar = base + SmemSweep::kBar;
    const uint32_t b_full = bar, b_empty = bar + 16, b_tfull = bar + 32, b_ifull = bar + 48,
                   b_iempty = bar + 56;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemSweep::kTmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
    if (threadIdx.x == 32) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 5);      // MMA commit + 4 epilogue warps
            mbar_init(b_tfull + 8 * s, 1);
        }
        mbar_init(b_ifull, 1);
        mbar_init(b_iempty, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            TileIter<kSweep == SWEEP_C> iter(p);
            int I, J, curI = -1, it = 0, seg = 0;
            bool last;
            while (iter.next(I, J, last)) {
                if (I != curI) {
                    mbar_wait(b_iempty, (seg & 1) ^ 1);
                    mbar_arrive_expect_tx(b_ifull, kTileBytes);
                    tma_bulk_g2s(sI, p.tiles + static_cast<size_t>(p.rb0 + I) * kTileBytes, kTileBytes,
                                 b_ifull);
                    curI = I;
                    ++seg;
                }
                const int slot = it & 1;
                mbar_wait(b_empty + 8 * slot, ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(b_full + 8 * slot, kTileBytes + 512);
                tma_bulk_g2s(sJ + slot * kTileBytes, p.tiles + static_cast<size_t>(J) * kTileBytes,
                             kTileBytes, b_full + 8 * slot);
                tma_bulk_g2s(sY + slot * 512, p.y + static_cast<size_t>(J) * 128, 512, b_full + 8 * slot);
                ++it;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            TileIter<kSweep == SWEEP_C> iter(p);
            const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
            int I, J, curI = -1, it = 0, seg = 0;
            bool last;
            while (iter.next(I, J, last)) {
                if (I != curI) {
                    mbar_wait(b_ifull, seg & 1);
                    curI = I;
                    ++seg;
                }
                const int slot = it & 1;
                mbar_wait(b_full + 8 * slot, (it >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ss(tmem + slot * 128, ftile_desc_kmajor(sI, k),
                            ftile_desc_kmajor(sJ + slot * kTileBytes, k), idesc, k > 0);
                tc_commit(b_empty + 8 * slot);
                tc_commit(b_tfull + 8 * slot);
                if (kSweep != SWEEP_C && last) tc_commit(b_iempty);
                ++it;
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue
        const int q = warp & 3;
        const int r = q * 32 + lane;                       // row inside the block == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        TileIter<kSweep == SWEEP_C> iter(p);
        int I, J, curI = -1, it = 0;
        bool last;
        // per-row state
        float acc0[4], acc1[4], acc2[4], acc3[4];
        float cshift = 0.f, ra = 0.f, rb = 0.f, rden = 1.f;
        int yi = -1, gi = -1;
        int2 rI = make_int2(0, -1);

        auto reset = [&]() {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc0[u] = (kSweep == SWEEP_A) ? -FLT_MAX : 0.f;
                acc1[u] = acc2[u] = acc3[u] = 0.f;
            }
        };
        auto flush = [&](int fI) {
            if (kSweep == SWEEP_A) {
                const int seg = blockIdx.x - p.part.first_cta(fI);
                float mx = fmaxf(fmaxf(acc0[0], acc0[1]), fmaxf(acc0[2], acc0[3]));
                p.pA[(static_cast<size_t>(fI) * p.maxseg + seg) * 128 + r] =
                    make_float4(mx, (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]),
                                (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]), 0.f);
            } else if (kSweep == SWEEP_B) {
                const int seg = blockIdx.x - p.part.first_cta(fI);
                p.pB[(static_cast<size_t>(fI) * p.maxseg + seg) * 128 + r] =
                    make_float2((acc0[0] + acc0[1]) + (acc0[2] + acc0[3]),
                                (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]));
            } else {
                p.pC[(static_cast<size_t>(fI) * p.splitc + (blockIdx.x % p.splitc)) * 128 + r] =
                    make_float4((acc0[0] + acc0[1]) + (acc0[2] + acc0[3]),
                                (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]),
                                (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]),
                                (acc3[0] + acc3[1]) + (acc3[2] + acc3[3]));
            }
        };
        auto begin_segment = [&](int nI_) {
            gi = (p.rb0 + nI_) * 128 + r;
            yi = p.y[gi];
            rI = p.blk_range[p.rb0 + nI_];
            if (kSweep == SWEEP_A) {
                cshift = p.sqnorm[gi];
            } else {
                RowA ra_ = combine_A(p, nI_, r);
                ra = ra_.a;
                rb = ra_.b;
                if (kSweep == SWEEP_C) rden = combine_B(p, nI_, r).x;
            }
            reset();
        };

        if (kSweep == SWEEP_C) {               // C always owns exactly one row block, even if it
            curI = blockIdx.x / p.splitc;      // turns out to have no tile: its slot must be written
            begin_segment(curI);
        }
        while (iter.next(I, J, last)) {
            if (I != curI) {
                if (curI >= 0) flush(curI);
                curI = I;
                begin_segment(I);
            }
            const int slot = it & 1;
            const int par = (it >> 1) & 1;
            mbar_wait(b_full + 8 * slot, par);             // labels of the column block have landed
            mbar_wait(b_tfull + 8 * slot, par);
            tc_fence_after();
            const int32_t* ys = sYg + slot * 128;
            const int col0 = J * 128;
            const int2 rJ = p.blk_range[J];
            const bool all_valid = p.blk_nvalid[J] == 128;
            const uint32_t taddr = tmem + slot * 128 + lane_off;

            if (kSweep == SWEEP_A) {
                if (all_valid) {
                    for_each_chunk(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(v[j]) - cshift;
                            acc0[j & 3] = fmaxf(acc0[j & 3], x);
                            acc1[j & 3] += x;
                            acc2[j & 3] = fmaf(x, x, acc2[j & 3]);
                        }
                    });
                } else {
                    for_each_chunk(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(v[j]) - cshift;
                            if (ys[c0 + j] >= 0) {
                                acc0[j & 3] = fmaxf(acc0[j & 3], x);
                                acc1[j & 3] += x;
                                acc2[j & 3] = fmaf(x, x, acc2[j & 3]);
                            }
                        }
                    });
                }
            } else if (kSweep == SWEEP_B) {
                const bool fast = all_valid && (kMode == DCL_MODE_PIXEL ? !ranges_overlap(rI, rJ)
                                                                        : (p.rb0 + I) != J);
                if (fast) {
                    for_each_chunk(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = fmaf(__uint_as_float(v[j]), ra, rb);
                            float e = ex2f(t);
                            acc0[j & 3] += e;
                            acc1[j & 3] = fmaf(e, t, acc1[j & 3]);
                        }
                    });
                } else {
                    for_each_chunk(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = fmaf(__uint_as_float(v[j]), ra, rb);
                            float e = ex2f(t);
                            const int yj = ys[c0 + j];
                            const bool den = (yj >= 0) && (kMode == DCL_MODE_PIXEL ? (yj != yi)
                                                                                   : (col0 + c0 + j != gi));
                            if (den) {
                                acc0[j & 3] += e;
                                acc1[j & 3] = fmaf(e, t, acc1[j & 3]);
                            }
                        }
                    });
                }
            } else {
                for_each_chunk(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int yj = ys[c0 + j];
                        const bool pos = (yj == yi) && (col0 + c0 + j != gi);
                        float t = fmaf(__uint_as_float(v[j]), ra, rb);
                        float l = t * kLn2;
                        if (kMode == DCL_MODE_PIXEL) {
                            float d = ex2f(t) + rden;
                            float inv = rcpf(d);
                            float lp = fmaf(-kLn2, lg2f(d), l);
                            if (pos) {
                                acc0[j & 3] += 1.f;
                                acc1[j & 3] += lp;
                                acc2[j & 3] += inv;
                                acc3[j & 3] = fmaf(inv, l, acc3[j & 3]);
                            }
                        } else {
                            if (pos) {
                                acc0[j & 3] += 1.f;
                                acc1[j & 3] += l;
                            }
                        }
                    }
                });
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_empty + 8 * slot);
            ++it;
        }
        if (curI >= 0) flush(curI);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

// =============================================================================================
// finalize: partials -> per-row loss and backward constants (one thread per local row)
// =============================================================================================
template <int kMode>
__global__ void __launch_bounds__(128) k_finalize(const Params p) {
    const int I = blockIdx.x, r = threadIdx.x;
    const int gi = (p.rb0 + I) * 128 + r;
    const int yi = p.y[gi];
    float rl = 0.f;
    float4 cA = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cB = make_float4(0.f, 1.f, __int_as_float(-1), 0.f);
    if (yi >= 0) {
        RowA ra = combine_A(p, I, r);
        float2 db = combine_B(p, I, r);
        float P = 0.f, SL = 0.f, SI = 0.f, SIL = 0.f;
        for (int s = 0; s < p.splitc; ++s) {
            float4 v = p.pC[(static_cast<size_t>(I) * p.splitc + s) * 128 + r];
            P += v.x; SL += v.y; SI += v.z; SIL += v.w;
        }
        const float ratio = p.T / p.Tb;
        const float c = ratio / static_cast<float>(p.n_valid);
        const float w = -c / P;                       // P == 0 -> NaN, as in the reference (loss.py:383)
        const float Bl = db.y * kLn2;                 // sum_den E*l
        float Q, R, wn;
        if (kMode == DCL_MODE_PIXEL) {
            rl = -ratio * SL / P;
            Q = w * SI;
            R = w * db.x * SIL - Q * Bl;
            wn = ra.kappa * w * db.x;
        } else {
            rl = -ratio * (SL - P * logf(db.x)) / P;
            Q = -c / db.x;
            R = w * SL - Q * Bl;
            wn = ra.kappa * w;
        }
        cA = make_float4(ra.a, ra.b, -ra.kappa * R * kLn2, -ra.kappa * Q);
        cB = make_float4(wn, db.x, __int_as_float(yi), 0.f);
    }
    p.colA[gi] = cA;
    p.colB[gi] = cB;
    p.rowloss[gi] = rl;
    // deterministic block sum of the row losses
    __shared__ float red[128];
    red[r] = rl;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (r < s) red[r] += red[r + s];
        __syncthreads();
    }
    if (r == 0) p.blockloss[I] = red[0];
}

__global__ void k_loss_sum(const float* blockloss, int nI, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < nI; ++i) s += blockloss[i];
        *out = s;
    }
}

__global__ void __launch_bounds__(128) k_block_ranges(const int32_t* y, int2* blk_range, int32_t* blk_nvalid) {
    const int J = blockIdx.x, r = threadIdx.x;
    const int v = y[J * 128 + r];
    int lo = v >= 0 ? v : INT_MAX, hi = v >= 0 ? v : -1, n = v >= 0 ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    __shared__ int slo[4], shi[4], sn[4];
    if ((r & 31) == 0) { slo[r >> 5] = lo; shi[r >> 5] = hi; sn[r >> 5] = n; }
    __syncthreads();
    if (r == 0) {
        blk_range[J] = make_int2(min(min(slo[0], slo[1]), min(slo[2], slo[3])),
                                 max(max(shi[0], shi[1]), max(shi[2], shi[3])));
        blk_nvalid[J] = sn[0] + sn[1] + sn[2] + sn[3];
    }
}

// =============================================================================================
// Backward: per tile  S = F_I F_J^T  ->  G (bf16, TMEM, aliasing S)  ->  dF_I += G F_J
// =============================================================================================
template <int kMode>
__global__ void __launch_bounds__(kThreads, 2) k_backward(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sI = base + SmemBwd::kI;
    const uint32_t sJ = base + SmemBwd::kJ;
    const uint32_t sCA = base + SmemBwd::kCA;
    const uint32_t sCB = base + SmemBwd::kCB;
    const float4* gCA = reinterpret_cast<const float4*>(gen + SmemBwd::kCA);
    const float4* gCB = reinterpret_cast<const float4*>(gen + SmemBwd::kCB);
    const uint32_t bar = base + SmemBwd::kBar;
    const uint32_t b_full = bar, b_empty = bar + 16, b_tfull = bar + 32, b_pfull = bar + 40,
                   b_dfull = bar + 48, b_dempty = bar + 56, b_ifull = bar + 64, b_iempty = bar + 72;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemBwd::kTmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
    if (threadIdx.x == 32) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 1);
        }
        mbar_init(b_tfull, 1);
        mbar_init(b_pfull, 4);
        mbar_init(b_dfull, 1);
        mbar_init(b_dempty, 4);
        mbar_init(b_ifull, 1);
        mbar_init(b_iempty, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tS = tmem, tP = tmem, tD = tmem + 128;     // G (bf16, 64 columns) aliases S

    if (warp == 0) {
        if (lane == 0) {
            TileIter<false> iter(p);
            int I, J, curI = -1, it = 0, seg = 0;
            bool last;
            while (iter.next(I, J, last)) {
                if (I != curI) {
                    mbar_wait(b_iempty, (seg & 1) ^ 1);
                    mbar_arrive_expect_tx(b_ifull, kTileBytes);
                    tma_bulk_g2s(sI, p.tiles + static_cast<size_t>(p.rb0 + I) * kTileBytes, kTileBytes,
                                 b_ifull);
                    curI = I;
                    ++seg;
                }
                const int slot = it & 1;
                mbar_wait(b_empty + 8 * slot, ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(b_full + 8 * slot, kTileBytes + 4096);
                tma_bulk_g2s(sJ + slot * kTileBytes, p.tiles + static_cast<size_t>(J) * kTileBytes,
                             kTileBytes, b_full + 8 * slot);
                tma_bulk_g2s(sCA + slot * 2048, p.colA + static_cast<size_t>(J) * 128, 2048,
                             b_full + 8 * slot);
                tma_bulk_g2s(sCB + slot * 2048, p.colB + static_cast<size_t>(J) * 128, 2048,
                             b_full + 8 * slot);
                ++it;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            TileIter<false> iter(p);
            const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
            const uint32_t idesc_d = umma_idesc_bf16(128, 128, 0, 1);   // B = F_J, MN-major
            int I, J, curI = -1, it = 0, seg = 0;
            bool last;
            while (iter.next(I, J, last)) {
                const bool first = (I != curI);
                if (first) {
                    mbar_wait(b_ifull, seg & 1);
                    if (seg > 0) mbar_wait(b_dempty, (seg - 1) & 1);   // previous dF drained
                    curI = I;
                    ++seg;
                }
                const int slot = it & 1;
                mbar_wait(b_full + 8 * slot, (it >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ss(tS, ftile_desc_kmajor(sI, k), ftile_desc_kmajor(sJ + slot * kTileBytes, k),
                            idesc_s, k > 0);
                tc_commit(b_tfull);
                mbar_wait(b_pfull, it & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ts(tD, tP + k * 8, ftile_desc_mnmajor(sJ + slot * kTileBytes, k), idesc_d,
                            (!first) || k > 0);
                tc_commit(b_empty + 8 * slot);
                if (last) {
                    tc_commit(b_dfull);
                    tc_commit(b_iempty);
                }
                ++it;
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        TileIter<false> iter(p);
        int I, J, curI = -1, it = 0, seg = 0;
        bool last;
        float4 rA = make_float4(0.f, 0.f, 0.f, 0.f), rB = make_float4(0.f, 1.f, 0.f, 0.f);
        int yi = -1, gi = -1;
        int2 rI = make_int2(0, -1);
        while (iter.next(I, J, last)) {
            if (I != curI) {
                curI = I;
                gi = (p.rb0 + I) * 128 + r;
                rA = p.colA[gi];
                rB = p.colB[gi];
                yi = __float_as_int(rB.z);
                rI = p.blk_range[p.rb0 + I];
            }
            const int slot = it & 1;
            mbar_wait(b_full + 8 * slot, (it >> 1) & 1);
            mbar_wait(b_tfull, it & 1);
            tc_fence_after();
            const float4* cA = gCA + slot * 128;
            const float4* cB = gCB + slot * 128;
            const int col0 = J * 128;
            // fast tile: every pair is a plain "denominator" pair in both directions (no positives,
            // no self pair); padding needs no mask here because padded F rows are zero
            const bool fast = !ranges_overlap(rI, p.blk_range[J]) &&
                              (kMode == DCL_MODE_PIXEL || (p.rb0 + I) != J);
            if (fast) {
                for_each_chunk(tS + lane_off, [&](int c0, const uint32_t (&v)[32]) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float g[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float s = __uint_as_float(v[j + u]);
                            const float4 ck = cA[c0 + j + u];
                            const float ti = fmaf(s, rA.x, rA.y);
                            const float tk = fmaf(s, ck.x, ck.y);
                            float acc = ti * rA.z;
                            acc = fmaf(tk, ck.z, acc);
                            acc = fmaf(ex2f(ti), rA.w, acc);
                            acc = fmaf(ex2f(tk), ck.w, acc);
                            g[u] = acc;
                        }
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(g[0], g[1]);
                        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    tmem_st16(tP + lane_off + (c0 >> 1), pk);
                });
            } else {
                for_each_chunk(tS + lane_off, [&](int c0, const uint32_t (&v)[32]) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float g[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float s = __uint_as_float(v[j + u]);
                            const float4 ck = cA[c0 + j + u];
                            const float4 dk = cB[c0 + j + u];
                            const int yk = __float_as_int(dk.z);
                            const float ti = fmaf(s, rA.x, rA.y);
                            const float tk = fmaf(s, ck.x, ck.y);
                            const float ei = ex2f(ti), ek = ex2f(tk);
                            const bool same = (yk == yi);
                            const bool notself = (col0 + c0 + j + u) != gi;
                            const bool pos = same && notself;
                            const bool den = (kMode == DCL_MODE_PIXEL) ? !same : notself;
                            float acc = ti * rA.z;
                            acc = fmaf(tk, ck.z, acc);
                            if (den) {
                                acc = fmaf(ei, rA.w, acc);
                                acc = fmaf(ek, ck.w, acc);
                            }
                            if (pos) {
                                if (kMode == DCL_MODE_PIXEL) {
                                    acc = fmaf(rB.x, rcpf(ei + rB.y), acc);
                                    acc = fmaf(dk.x, rcpf(ek + dk.y), acc);
                                } else {
                                    acc += rB.x + dk.x;
                                }
                            }
                            g[u] = acc;
                        }
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(g[0], g[1]);
                        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    tmem_st16(tP + lane_off + (c0 >> 1), pk);
                });
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_pfull);
            ++it;
            if (last) {
                // drain the finished dF_I partial
                mbar_wait(b_dfull, seg & 1);
                tc_fence_after();
                const int sidx = blockIdx.x - p.part.first_cta(I);
                float* out = p.pD + ((static_cast<size_t>(I) * p.maxseg + sidx) * 128 + r) * 128;
                for_each_chunk(tD + lane_off, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(out + c0 + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                        __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                });
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_dempty);
                ++seg;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

// dF[row] = sum over segments of the partials
__global__ void __launch_bounds__(256) k_reduce_dF(const Params p, float* __restrict__ dF) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);     // one warp per row
    const int lane = threadIdx.x & 31;
    if (row >= p.nI * 128) return;
    const int I = row >> 7, r = row & 127;
    const int ns = p.part.nseg(I);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ns; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(
            p.pD + ((static_cast<size_t>(I) * p.maxseg + s) * 128 + r) * 128 + lane * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dF + static_cast<size_t>(row) * 128 + lane * 4) = acc;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Layout {
    Part part;
    int maxseg, splitc;
    size_t off_range, off_nvalid, off_pA, off_pB, off_pC, off_pD, off_bl, bytes;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Layout make_layout(int nI, int nJ) {
    Layout L;
    const int ctas = 2 * sm_count();
    L.part.total = static_cast<long long>(nI) * nJ;
    L.part.nJ = nJ;
    L.part.G = static_cast<int>(L.part.total < ctas ? L.part.total : ctas);
    const long long q = L.part.total / L.part.G;              // >= 1 tiles per CTA
    L.maxseg = static_cast<int>((nJ + q - 1) / q + 1);
    int sc = (ctas + nI - 1) / nI;
    L.splitc = sc < 1 ? 1 : (sc > 8 ? 8 : sc);
    size_t o = 0;
    L.off_range = o;  o = align_up(o + sizeof(int2) * nJ, 256);
    L.off_nvalid = o; o = align_up(o + sizeof(int32_t) * nJ, 256);
    L.off_pA = o;     o = align_up(o + sizeof(float4) * static_cast<size_t>(nI) * L.maxseg * 128, 256);
    L.off_pB = o;     o = align_up(o + sizeof(float2) * static_cast<size_t>(nI) * L.maxseg * 128, 256);
    L.off_pC = o;     o = align_up(o + sizeof(float4) * static_cast<size_t>(nI) * L.splitc * 128, 256);
    L.off_bl = o;     o = align_up(o + sizeof(float) * nI, 256);
    L.off_pD = o;     o = align_up(o + sizeof(float) * static_cast<size_t>(nI) * L.maxseg * 128 * 128, 256);
    L.bytes = o;
    return L;
}

static Params make_params(const Layout& L, const void* tiles, const int32_t* y, const float* sqnorm,
                          int nJ, int rb0, int nI, int n_valid, int mode, float T, float Tb, void* ws) {
    Params p{};
    uint8_t* w = static_cast<uint8_t*>(ws);
    p.tiles = static_cast<const uint8_t*>(tiles);
    p.y = y;
    p.sqnorm = sqnorm;
    p.blk_range = reinterpret_cast<const int2*>(w + L.off_range);
    p.blk_nvalid = reinterpret_cast<const int32_t*>(w + L.off_nvalid);
    p.nJ = nJ; p.rb0 = rb0; p.nI = nI; p.n_valid = n_valid; p.mode = mode;
    p.T = T; p.Tb = Tb;
    p.part = L.part;
    p.maxseg = L.maxseg;
    p.splitc = L.splitc;
    p.pA = reinterpret_cast<float4*>(w + L.off_pA);
    p.pB = reinterpret_cast<float2*>(w + L.off_pB);
    p.pC = reinterpret_cast<float4*>(w + L.off_pC);
    p.pD = reinterpret_cast<float*>(w + L.off_pD);
    p.blockloss = reinterpret_cast<float*>(w + L.off_bl);
    return p;
}

template <class K>
static int set_smem(K kernel, int bytes) {
    DCL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return 0;
}

template <int kMode>
static int run_fwd(Params p, const Layout& L, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (int e = set_smem(k_sweep<SWEEP_A, kMode>, SmemSweep::kBytes)) return e;
        if (int e = set_smem(k_sweep<SWEEP_B, kMode>, SmemSweep::kBytes)) return e;
        if (int e = set_smem(k_sweep<SWEEP_C, kMode>, SmemSweep::kBytes)) return e;
        configured = true;
    }
    k_block_ranges<<<p.nJ, 128, 0, st>>>(p.y, const_cast<int2*>(p.blk_range),
                                          const_cast<int32_t*>(p.blk_nvalid));
    DCL_LAUNCH_CHECK("k_block_ranges");
    k_sweep<SWEEP_A, kMode><<<L.part.G, kThreads, SmemSweep::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_sweep<A>");
    k_sweep<SWEEP_B, kMode><<<L.part.G, kThreads, SmemSweep::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_sweep<B>");
    k_sweep<SWEEP_C, kMode><<<p.nI * L.splitc, kThreads, SmemSweep::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_sweep<C>");
    k_finalize<kMode><<<p.nI, 128, 0, st>>>(p);
    DCL_LAUNCH_CHECK("k_finalize");
    k_loss_sum<<<1, 32, 0, st>>>(p.blockloss, p.nI, p.loss_sum);
    DCL_LAUNCH_CHECK("k_loss_sum");
    return 0;
}

template <int kMode>
static int run_bwd(Params p, const Layout& L, float* dF, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (int e = set_smem(k_backward<kMode>, SmemBwd::kBytes)) return e;
        configured = true;
    }
    k_block_ranges<<<p.nJ, 128, 0, st>>>(p.y, const_cast<int2*>(p.blk_range),
                                          const_cast<int32_t*>(p.blk_nvalid));
    DCL_LAUNCH_CHECK("k_block_ranges");
    k_backward<kMode><<<L.part.G, kThreads, SmemBwd::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_backward");
    k_reduce_dF<<<(p.nI * 128 + 7) / 8, 256, 0, st>>>(p, dF);
    DCL_LAUNCH_CHECK("k_reduce_dF");
    return 0;
}

}  // namespace dcl

using namespace dcl;

extern "C" size_t dcl_contrast_workspace_bytes(int nI, int nJ) {
    if (nI <= 0 || nJ <= 0) return 0;
    return make_layout(nI, nJ).bytes;
}

static int check_args(const void* tiles, const int32_t* y, int nJ, int rb0, int nI, int mode, void* ws,
                      size_t ws_bytes, const Layout& L) {
    if (!tiles || !y || !ws) return fail(DCL_ERR_ARG, "null pointer argument");
    if (nJ <= 0 || nI <= 0 || rb0 < 0 || rb0 + nI > nJ)
        return fail(DCL_ERR_ARG, "bad block range rb0=%d nI=%d nJ=%d", rb0, nI, nJ);
    if (mode != DCL_MODE_PIXEL && mode != DCL_MODE_SUPCON) return fail(DCL_ERR_ARG, "bad mode %d", mode);
    if (reinterpret_cast<uintptr_t>(tiles) % 128 || reinterpret_cast<uintptr_t>(y) % 16 ||
        reinterpret_cast<uintptr_t>(ws) % 256)
        return fail(DCL_ERR_ARG, "tiles must be 128-byte, y 16-byte, workspace 256-byte aligned");
    if (ws_bytes < L.bytes)
        return fail(DCL_ERR_WORKSPACE, "workspace %zu < required %zu", ws_bytes, L.bytes);
    return 0;
}

extern "C" int dcl_contrast_fwd(const void* tiles, const int32_t* y, const float* sqnorm, int nJ, int rb0,
                                int nI, int n_valid, int mode, float temperature,
                                float base_temperature, void* workspace, size_t workspace_bytes,
                                float* colA, float* colB, float* rowloss, float* loss_sum,
                                void* stream) {
    if (int e = dcl_check_device()) return e;
    if (nJ <= 0 || nI <= 0) return fail(DCL_ERR_ARG, "empty problem");
    Layout L = make_layout(nI, nJ);
    if (int e = check_args(tiles, y, nJ, rb0, nI, mode, workspace, workspace_bytes, L)) return e;
    if (!sqnorm || !colA || !colB || !rowloss || !loss_sum || n_valid <= 0)
        return fail(DCL_ERR_ARG, "null output or n_valid <= 0");
    if (reinterpret_cast<uintptr_t>(colA) % 16 || reinterpret_cast<uintptr_t>(colB) % 16)
        return fail(DCL_ERR_ARG, "colA/colB must be 16-byte aligned");
    Params p = make_params(L, tiles, y, sqnorm, nJ, rb0, nI, n_valid, mode, temperature,
                           base_temperature, workspace);
    p.colA = reinterpret_cast<float4*>(colA);
    p.colB = reinterpret_cast<float4*>(colB);
    p.rowloss = rowloss;
    p.loss_sum = loss_sum;
    return mode == DCL_MODE_PIXEL ? run_fwd<DCL_MODE_PIXEL>(p, L, as_stream(stream))
                                  : run_fwd<DCL_MODE_SUPCON>(p, L, as_stream(stream));
}

extern "C" int dcl_contrast_bwd(const void* tiles, const int32_t* y, const float* colA, const float* colB,
                                int nJ, int rb0, int nI, int mode, void* workspace,
                                size_t workspace_bytes, float* dF, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (nJ <= 0 || nI <= 0) return fail(DCL_ERR_ARG, "empty problem");
    Layout L = make_layout(nI, nJ);
    if (int e = check_args(tiles, y, nJ, rb0, nI, mode, workspace, workspace_bytes, L)) return e;
    if (!colA || !colB || !dF) return fail(DCL_ERR_ARG, "null pointer argument");
    if (reinterpret_cast<uintptr_t>(colA) % 16 || reinterpret_cast<uintptr_t>(colB) % 16 ||
        reinterpret_cast<uintptr_t>(dF) % 16)
        return fail(DCL_ERR_ARG, "colA/colB/dF must be 16-byte aligned");
    Params p = make_params(L, tiles, y, nullptr, nJ, rb0, nI, 1, mode, 1.f, 1.f, workspace);
    p.colA = reinterpret_cast<float4*>(const_cast<float*>(colA));
    p.colB = reinterpret_cast<float4*>(const_cast<float*>(colB));
    return mode == DCL_MODE_PIXEL ? run_bwd<DCL_MODE_PIXEL>(p, L, dF, as_stream(stream))
                                  : run_bwd<DCL_MODE_SUPCON>(p, L, dF, as_stream(stream));
}
