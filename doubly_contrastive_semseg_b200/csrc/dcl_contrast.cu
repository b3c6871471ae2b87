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
// Kernel shape (one persistent CTA per SM, 320 threads): warps 0..7 = two epilogue groups of four
// warps (warp w owns TMEM lanes 32*(w%4)..+31, one thread per anchor row), warp 8 = TMA producer lane
// + TMEM allocator, warp 9 = MMA issuer lane.  Shared-memory tile slots (4-deep ring) and TMEM
// accumulator stages (2) are decoupled, so the L2->smem->MMA latency of a tile is hidden under the
// epilogues of the tiles before it.
//   sweeps  : a CTA owns a PAIR of row blocks (M = 256): every F_J tile fetched from L2 feeds two
//             128x128 S tiles, one per epilogue group; TMEM = 2 stages x (2 x 128) columns.
//   backward: a CTA owns one row block; the two groups alternate column tiles (S double-buffered,
//             G written over S, dF accumulator in a third 128-column region).
// Work = the flattened list of (row unit, column block) tiles cut into gridDim.x equal contiguous
// ranges; a range that crosses a row-unit boundary flushes a deterministic partial ("segment")
// instead of using atomics, so results are bit-reproducible.
#include <cfloat>
#include <cuda_bf16.h>
#include "dcl_common.cuh"
#include "dcl_ptx.cuh"

namespace dcl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 320;
constexpr int kTmemCols = 512;
// The SM's warp arbiter favours high warp ids: the two single-lane pipeline drivers get the highest
// ids so they are never starved by the eight issue-bound epilogue warps (0..7).
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kMaxBlocks = 1024;   // column blocks whose label ranges are cached in shared memory

// ---------------------------------------------------------------------------------------------
// flattened-range partition of nU row units x nJ column blocks over G CTAs
struct Part {
    long long total;
    int G, nJ;
    __host__ __device__ long long begin(int c) const { return static_cast<long long>(c) * total / G; }
    // CTA whose range contains flat tile index x
    __host__ __device__ int cta_of(long long x) const {
        return static_cast<int>(((x + 1) * G + total - 1) / total - 1);
    }
    __host__ __device__ int first_cta(int U) const { return cta_of(static_cast<long long>(U) * nJ); }
    __host__ __device__ int nseg(int U) const {
        return cta_of(static_cast<long long>(U + 1) * nJ - 1) - first_cta(U) + 1;
    }
};

struct Params {
    const uint8_t* tiles;    // [nJ] F-tiles
    const int32_t* y;        // [nJ*128]
    const float* sqnorm;     // [nJ*128]
    int nJ, rb0, nI, nP, n_valid, mode;
    float T, Tb;
    Part partS;              // sweeps A/B: units = pairs of row blocks
    Part partD;              // backward:   units = row blocks
    int maxsegS, maxsegD;
    int splitc;              // sweep C: CTAs per row-block pair
    float4* pA;              // [nI][maxsegS][128] (max(s-c), S1, S2, -)
    float2* pB;              // [nI][maxsegS][128] (Den, Bt)
    float4* pC;              // [nI][splitc][128] (P, sum lp | sum l, sum inv, sum inv*l)
    float* pD;               // [nI][maxsegD][128][128] dF partials
    float4* colA;            // [nJ*128] (a, b, p, q)
    float4* colB;            // [nJ*128] (wn, Den, y bits, 0)
    float* rowloss;          // [nJ*128]
    float* blockloss;        // [nI]
    float* loss_sum;
    unsigned int* ticket;    // finalize's last-block counter (zeroed by sweep A)
    long long* trace;        // diagnostics only: per-role clock64 stamps of CTA 0 (dcl_debug_trace)
    int debug;               // diagnostics only (dcl_debug_flags): 1 skip epilogue math, 2 skip S MMAs, 4 skip dF MMAs
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^t for t in [-1.4427, 0] (l = t*ln2 lies in [-1, 0]: the rows are unit-normalised, loss.py:366) on the
// FMA pipe: minimax-fitted polynomials, no range reduction needed.  The MUFU pipe (16 ex2/clk/SM) is
// the epilogues' bottleneck, so a fixed 3-of-8 share of the exponentials goes here instead.
// Max relative error 1.6e-5 (degree 4, forward sums) / 3.2e-4 (degree 3, backward).
__device__ __forceinline__ float ex2_poly4(float t) {
    float p = 0.005771667696535587f;
    p = fmaf(p, t, 0.05083702877163887f);
    p = fmaf(p, t, 0.23775413632392883f);
    p = fmaf(p, t, 0.6926645040512085f);
    return fmaf(p, t, 0.9999837875366211f);
}
__device__ __forceinline__ float ex2_poly3(float t) {
    float p = 0.033237408846616745f;
    p = fmaf(p, t, 0.22061625123023987f);
    p = fmaf(p, t, 0.6871119737625122f);
    return fmaf(p, t, 0.9996766448020935f);
}
// j is a compile-time (unrolled) column index: DCL_POLY_* of every 8 columns take the polynomial
#ifndef DCL_POLY_FWD
#define DCL_POLY_FWD 3
#endif
#ifndef DCL_POLY_BWD
#define DCL_POLY_BWD 3
#endif
#define DCL_EX2_FWD(t, j) ((((j) & 7) < DCL_POLY_FWD) ? ex2_poly4(t) : ex2f(t))
#define DCL_EX2_BWD(t, j) ((((j) & 7) < DCL_POLY_BWD) ? ex2_poly3(t) : ex2f(t))

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

// Walk `kChunks` consecutive 32-column chunks of an fp32 accumulator row, the next chunk's TMEM
// load in flight while the current one is processed.  fn(c0, v) sees columns [c0, c0+32).
template <int kChunks, class Fn>
__device__ __forceinline__ void for_each_chunk(uint32_t taddr, Fn&& fn) {
    static_assert(kChunks == 2 || kChunks == 4, "2 or 4 chunks");
    uint32_t va[32], vb[32];
    tmem_ld32(taddr, va);
    tmem_ld_wait_on(va);
    tmem_ld32(taddr + 32, vb);
    fn(0, va);
    tmem_ld_wait_on(vb);
    if (kChunks == 4) {
        tmem_ld32(taddr + 64, va);
        fn(32, vb);
        tmem_ld_wait_on(va);
        tmem_ld32(taddr + 96, vb);
        fn(64, va);
        tmem_ld_wait_on(vb);
        fn(96, vb);
    } else {
        fn(32, vb);
    }
}

// diagnostics: stamp (role, tile, event) for CTA 0's first 32 tiles
__device__ __forceinline__ void trace_stamp(const Params& p, int role, int it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 32) p.trace[(role * 32 + it) * 4 + ev] = clock64();   // caller enables it around ONE launch
}

__device__ __forceinline__ bool ranges_overlap(int2 a, int2 b) { return a.x <= b.y && b.x <= a.y; }

// Label range (min,max over valid rows; (INT_MAX,-1) if none) and valid-row count of every column
// block, computed by the whole CTA into shared memory before the roles split.
__device__ __forceinline__ void compute_block_info(const int32_t* __restrict__ y, int nJ, int2* sRange,
                                                   int* sNv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int J = warp; J < nJ; J += nwarps) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(y + static_cast<size_t>(J) * 128) + lane);
        const int e[4] = {v.x, v.y, v.z, v.w};
        int lo = INT_MAX, hi = -1, n = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (e[u] >= 0) { lo = min(lo, e[u]); hi = max(hi, e[u]); ++n; }
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (lane == 0) { sRange[J] = make_int2(lo, hi); sNv[J] = n; }
    }
}

// Row constants derived from sweep A partials: t_ij = fma(s_ij, a, b) = l_ij * log2(e).
struct RowA {
    float a, b, kappa, smax;
};
__device__ __forceinline__ RowA combine_A(const Params& p, int I, int r) {
    const int ns = p.partS.nseg(I >> 1);
    float mx = -FLT_MAX;
    double S1 = 0.0, S2 = 0.0;
    for (int s = 0; s < ns; ++s) {
        float4 v = p.pA[(static_cast<size_t>(I) * p.maxsegS + s) * 128 + r];
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
    const int ns = p.partS.nseg(I >> 1);
    float den = 0.f, bt = 0.f;
    for (int s = 0; s < ns; ++s) {
        float2 v = p.pB[(static_cast<size_t>(I) * p.maxsegS + s) * 128 + r];
        den += v.x;
        bt += v.y;
    }
    return make_float2(den, bt);
}

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up (bytes from a 1024-aligned base)
struct SmemSweep {
    static constexpr int kI = 0;                          // 2 row-block tiles
    static constexpr int kJ = 2 * kTileBytes;             // 4 slots
    static constexpr int kY = 6 * kTileBytes;             // 8 x 512 B labels of column blocks
    static constexpr int kRange = kY + 8 * 512;           // int2[kMaxBlocks]
    static constexpr int kNv = kRange + 8 * kMaxBlocks;   // int[kMaxBlocks]
    static constexpr int kBar = kNv + 4 * kMaxBlocks;     // full[4] empty[4] tfull[2] tempty[2] ifull iempty yfull[8]
    static constexpr int kTmem = kBar + 256;
    static constexpr int kBytes = kTmem + 16 + 1024;      // + alignment slack
};
struct SmemBwd {
    static constexpr int kSlots = 5;                      // F_J tile + its column constants
    static constexpr int kStages = 3;                     // TMEM S/G stages (G aliases its S)
    static constexpr int kI = 0;
    static constexpr int kJ = kTileBytes;
    static constexpr int kCA = (1 + kSlots) * kTileBytes; // kSlots x 2 KiB colA
    static constexpr int kCB = kCA + kSlots * 2048;       // kSlots x 2 KiB colB
    static constexpr int kRange = kCB + kSlots * 2048;
    static constexpr int kBar = kRange + 8 * kMaxBlocks;  // full[5] empty[5] tfull[3] pfull[3] dfull dempty ifull iempty
    static constexpr int kTmem = kBar + 256;
    static constexpr int kBytes = kTmem + 16 + 1024;
};

enum { SWEEP_A = 0, SWEEP_B = 1, SWEEP_C = 2 };

// Tile sequence of one CTA: (row unit U, column block J).  Flat mode: contiguous range of the
// flattened U-major list.  Relevant mode (sweep C): CTA = (pair, split s) walks the column blocks
// whose label range overlaps either row block of the pair and keeps every splitc-th one.  All warp
// roles run an identical copy of this iterator, which keeps their barrier phases in step.
template <bool kRelevantOnly>
struct TileIter {
    int left;                 // flat mode: tiles remaining
    int nJ, U, J, r, s, splitc;
    int2 r0, r1;
    const int2* rng;
    __device__ TileIter(const Params& p, const Part& part, const int2* sRange) {
        nJ = p.nJ;
        rng = sRange;
        left = 0;
        U = J = r = s = 0;
        splitc = 1;
        r0 = r1 = make_int2(INT_MAX, -1);
        if (kRelevantOnly) {
            U = blockIdx.x / p.splitc;
            s = blockIdx.x % p.splitc;
            splitc = p.splitc;
            r0 = sRange[p.rb0 + 2 * U];
            if (2 * U + 1 < p.nI) r1 = sRange[p.rb0 + 2 * U + 1];
        } else {
            const long long t0 = part.begin(blockIdx.x), t1 = part.begin(blockIdx.x + 1);
            left = static_cast<int>(t1 - t0);
            U = static_cast<int>(t0 / nJ);            // the only divisions: once per CTA and role
            J = static_cast<int>(t0 - static_cast<long long>(U) * nJ);
        }
    }
    __device__ bool next(int& oU, int& oJ, bool& last_of_seg) {
        if (kRelevantOnly) {
            while (J < nJ) {
                const int j = J++;
                const int2 rj = rng[j];
                if (ranges_overlap(r0, rj) || ranges_overlap(r1, rj)) {
                    if ((r++ % splitc) == s) {
                        oU = U;
                        oJ = j;
                        last_of_seg = false;
                        return true;
                    }
                }
            }
            return false;
        } else {
            if (left <= 0) return false;
            oU = U;
            oJ = J;
            --left;
            if (++J == nJ) { J = 0; ++U; }
            last_of_seg = (left == 0) || (J == 0);
            return true;
        }
    }
};

// =============================================================================================
// Sweeps A / B / C
// =============================================================================================
template <int kSweep, int kMode>
__global__ void __launch_bounds__(kThreads, 1) k_sweep(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sI = base + SmemSweep::kI;
    const uint32_t sJ = base + SmemSweep::kJ;
    const uint32_t sY = base + SmemSweep::kY;
    const int32_t* sYg = reinterpret_cast<const int32_t*>(gen + SmemSweep::kY);
    int2* sRange = reinterpret_cast<int2*>(gen + SmemSweep::kRange);
    int* sNv = reinterpret_cast<int*>(gen + SmemSweep::kNv);
    const uint32_t bar = base + SmemSweep::kBar;
    // tfull / tempty are per (stage, group): index st * 2 + g
    const uint32_t b_full = bar, b_empty = bar + 32, b_tfull = bar + 64, b_tempty = bar + 96,
                   b_ifull = bar + 128, b_iempty = bar + 136, b_yfull = bar + 144;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemSweep::kTmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == kProducerWarp) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
    if (threadIdx.x == kMmaWarp * 32) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 1);      // MMA commit
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(b_tfull + 8 * s, 1);
            mbar_init(b_tempty + 8 * s, 4);     // the 4 warps of one epilogue group
        }
        mbar_init(b_ifull, 1);
        mbar_init(b_iempty, 1);
        // Labels of a column block ride their own 8-deep ring + barriers: an epilogue warp may lag
        // the producer by more than one phase of a 4-deep tile slot (1-bit parity would alias),
        // but never by 8 tiles (slot it&7 is refilled only after MMA(it+4), i.e. epilogue(it+2)).
        for (int s = 0; s < 8; ++s) mbar_init(b_yfull + 8 * s, 1);
        mbar_fence_init();
    }
    if (kSweep == SWEEP_A && blockIdx.x == 0 && threadIdx.x == 0) *p.ticket = 0u;
    compute_block_info(p.y, p.nJ, sRange, sNv);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kProducerWarp) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            TileIter<kSweep == SWEEP_C> iter(p, p.partS, sRange);
            int U, J, curU = -1, it = 0, seg = 0;
            bool last;
            while (iter.next(U, J, last)) {
                if (U != curU) {
                    const bool two = 2 * U + 1 < p.nI;
                    mbar_wait(b_iempty, (seg & 1) ^ 1);
                    mbar_arrive_expect_tx(b_ifull, two ? 2 * kTileBytes : kTileBytes);
                    tma_bulk_g2s(sI, p.tiles + static_cast<size_t>(p.rb0 + 2 * U) * kTileBytes,
                                 two ? 2 * kTileBytes : kTileBytes, b_ifull);   // the pair is contiguous
                    curU = U;
                    ++seg;
                }
                const int slot = it & 3;
                trace_stamp(p, 0, it, 0);
                mbar_wait(b_empty + 8 * slot, ((it >> 2) & 1) ^ 1);
                trace_stamp(p, 0, it, 1);
                mbar_arrive_expect_tx(b_full + 8 * slot, kTileBytes);
                tma_bulk_g2s(sJ + slot * kTileBytes, p.tiles + static_cast<size_t>(J) * kTileBytes,
                             kTileBytes, b_full + 8 * slot);
                mbar_arrive_expect_tx(b_yfull + 8 * (it & 7), 512);
                tma_bulk_g2s(sY + (it & 7) * 512, p.y + static_cast<size_t>(J) * 128, 512,
                             b_yfull + 8 * (it & 7));
                ++it;
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            TileIter<kSweep == SWEEP_C> iter(p, p.partS, sRange);
            const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
            int U, J, curU = -1, it = 0, seg = 0;
            bool last, two = false;
            while (iter.next(U, J, last)) {
                if (U != curU) {
                    mbar_wait(b_ifull, seg & 1);
                    two = 2 * U + 1 < p.nI;
                    curU = U;
                    ++seg;
                }
                const int slot = it & 3, st = it & 1;
                trace_stamp(p, 1, it, 0);
                mbar_wait(b_full + 8 * slot, (it >> 2) & 1);
                trace_stamp(p, 1, it, 1);
                const uint32_t par_e = ((it >> 1) & 1) ^ 1;
                mbar_wait(b_tempty + 8 * (st * 2), par_e);
                trace_stamp(p, 1, it, 2);
                tc_fence_after();
                if (!(p.debug & 2)) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ss(tmem + st * 256, ftile_desc_kmajor(sI, k),
                            ftile_desc_kmajor(sJ + slot * kTileBytes, k), idesc, k > 0);
                }
                tc_commit(b_tfull + 8 * (st * 2));            // group 0 can start while group 1's tile runs
                mbar_wait(b_tempty + 8 * (st * 2 + 1), par_e);
                tc_fence_after();
                if (two && !(p.debug & 2)) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_ss(tmem + st * 256 + 128, ftile_desc_kmajor(sI + kTileBytes, k),
                                ftile_desc_kmajor(sJ + slot * kTileBytes, k), idesc, k > 0);
                }
                tc_commit(b_empty + 8 * slot);
                tc_commit(b_tfull + 8 * (st * 2 + 1));
                if (kSweep != SWEEP_C && last) tc_commit(b_iempty);
                trace_stamp(p, 1, it, 3);
                ++it;
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue
        const int g = warp >> 2;                           // group: row block 2U+g of the pair
        const int q = warp & 3;
        const int r = q * 32 + lane;                       // row inside the block == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        TileIter<kSweep == SWEEP_C> iter(p, p.partS, sRange);
        int U, J, curU = -1, it = 0;
        bool last, valid = false;
        float acc0[4], acc1[4], acc2[4], acc3[4];
        float cshift = 0.f, ra = 0.f, rb = 0.f, rden = 1.f;
        int yi = -1, gi = -1, Iloc = 0;
        int2 rI = make_int2(INT_MAX, -1);

        auto reset = [&]() {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc0[u] = (kSweep == SWEEP_A) ? -FLT_MAX : 0.f;
                acc1[u] = acc2[u] = acc3[u] = 0.f;
            }
        };
        auto flush = [&]() {
            if (!valid) return;
            if (kSweep == SWEEP_A) {
                const int seg = blockIdx.x - p.partS.first_cta(curU);
                float mx = fmaxf(fmaxf(acc0[0], acc0[1]), fmaxf(acc0[2], acc0[3]));
                p.pA[(static_cast<size_t>(Iloc) * p.maxsegS + seg) * 128 + r] =
                    make_float4(mx, (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]),
                                (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]), 0.f);
            } else if (kSweep == SWEEP_B) {
                const int seg = blockIdx.x - p.partS.first_cta(curU);
                p.pB[(static_cast<size_t>(Iloc) * p.maxsegS + seg) * 128 + r] =
                    make_float2((acc0[0] + acc0[1]) + (acc0[2] + acc0[3]),
                                (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]));
            } else {
                p.pC[(static_cast<size_t>(Iloc) * p.splitc + (blockIdx.x % p.splitc)) * 128 + r] =
                    make_float4((acc0[0] + acc0[1]) + (acc0[2] + acc0[3]),
                                (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]),
                                (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]),
                                (acc3[0] + acc3[1]) + (acc3[2] + acc3[3]));
            }
        };
        auto begin_segment = [&](int nU) {
            curU = nU;
            Iloc = 2 * nU + g;
            valid = Iloc < p.nI;
            if (!valid) return;
            gi = (p.rb0 + Iloc) * 128 + r;
            yi = p.y[gi];
            rI = sRange[p.rb0 + Iloc];
            if (kSweep == SWEEP_A) {
                cshift = p.sqnorm[gi];
            } else {
                RowA ra_ = combine_A(p, Iloc, r);
                ra = ra_.a;
                rb = ra_.b;
                if (kSweep == SWEEP_C) rden = combine_B(p, Iloc, r).x;
            }
            reset();
        };

        if (kSweep == SWEEP_C) begin_segment(blockIdx.x / p.splitc);   // its slot is always written
        while (iter.next(U, J, last)) {
            if (U != curU) {
                if (curU >= 0) flush();
                begin_segment(U);
            }
            const int st = it & 1;
            if (threadIdx.x == 0) trace_stamp(p, 2, it, 0);
            mbar_wait(b_yfull + 8 * (it & 7), (it >> 3) & 1);   // labels of the column block have landed
            mbar_wait(b_tfull + 8 * (st * 2 + g), (it >> 1) & 1);
            if (threadIdx.x == 0) trace_stamp(p, 2, it, 1);
            tc_fence_after();
            const int32_t* ys = sYg + (it & 7) * 128;
            const int col0 = J * 128;
            const int2 rJ = sRange[J];
            const bool all_valid = sNv[J] == 128;
            const uint32_t taddr = tmem + st * 256 + g * 128 + lane_off;

            if (!valid || (p.debug & 1)) {
                // odd tail: this group has no row block; just release the stage
            } else if (kSweep == SWEEP_A) {
                if (all_valid) {
                    for_each_chunk<4>(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(v[j]) - cshift;
                            acc0[j & 3] = fmaxf(acc0[j & 3], x);
                            acc1[j & 3] += x;
                            acc2[j & 3] = fmaf(x, x, acc2[j & 3]);
                        }
                    });
                } else {
                    for_each_chunk<4>(taddr, [&](int c0, const uint32_t (&v)[32]) {
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
                                                                        : (p.rb0 + Iloc) != J);
                if (fast) {
                    for_each_chunk<4>(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = fmaf(__uint_as_float(v[j]), ra, rb);
                            float e = DCL_EX2_FWD(t, j);
                            acc0[j & 3] += e;
                            acc1[j & 3] = fmaf(e, t, acc1[j & 3]);
                        }
                    });
                } else {
                    for_each_chunk<4>(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = fmaf(__uint_as_float(v[j]), ra, rb);
                            float e = DCL_EX2_FWD(t, j);
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
            } else if (ranges_overlap(rI, rJ)) {
                for_each_chunk<4>(taddr, [&](int c0, const uint32_t (&v)[32]) {
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
            if (lane == 0) mbar_arrive(b_tempty + 8 * (st * 2 + g));
            if (threadIdx.x == 0) trace_stamp(p, 2, it, 2);
            ++it;
        }
        if (curU >= 0) flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarp) tmem_dealloc<kTmemCols>(tmem);
}

// =============================================================================================
// finalize: partials -> per-row loss and backward constants (one thread per local row); the last
// block to finish adds the per-block losses in fixed order
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
    __shared__ bool is_last;
    red[r] = rl;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (r < s) red[r] += red[r + s];
        __syncthreads();
    }
    if (r == 0) {
        p.blockloss[I] = red[0];
        __threadfence();
        is_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && r == 0) {
        __threadfence();
        float s = 0.f;
        const volatile float* bl = p.blockloss;
        for (int i = 0; i < p.nI; ++i) s += bl[i];
        *p.loss_sum = s;
    }
}

// =============================================================================================
// Backward: per tile  S = F_I F_J^T  ->  G (bf16, TMEM, aliasing S)  ->  dF_I += G F_J
// =============================================================================================
template <int kMode>
__global__ void __launch_bounds__(kThreads, 1) k_backward(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sI = base + SmemBwd::kI;
    const uint32_t sJ = base + SmemBwd::kJ;
    const uint32_t sCA = base + SmemBwd::kCA;
    const uint32_t sCB = base + SmemBwd::kCB;
    const float4* gCA = reinterpret_cast<const float4*>(gen + SmemBwd::kCA);
    const float4* gCB = reinterpret_cast<const float4*>(gen + SmemBwd::kCB);
    int2* sRange = reinterpret_cast<int2*>(gen + SmemBwd::kRange);
    const uint32_t bar = base + SmemBwd::kBar;
    constexpr int kSlots = SmemBwd::kSlots, kStages = SmemBwd::kStages;
    const uint32_t b_full = bar, b_empty = bar + 40, b_tfull = bar + 80, b_pfull = bar + 104,
                   b_dfull = bar + 128, b_dempty = bar + 136, b_ifull = bar + 144, b_iempty = bar + 152;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemBwd::kTmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == kProducerWarp) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
    if (threadIdx.x == kMmaWarp * 32) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 1);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(b_tfull + 8 * s, 1);
            mbar_init(b_pfull + 8 * s, 4);      // the 4 warps of one epilogue group
        }
        mbar_init(b_dfull, 1);
        mbar_init(b_dempty, 8);
        mbar_init(b_ifull, 1);
        mbar_init(b_iempty, 1);
        mbar_fence_init();
    }
    {
        // only the label ranges are needed here; reuse the helper with a scratch count array
        // placed over the (not yet used) first colB slot
        int* scratch_nv = reinterpret_cast<int*>(gen + SmemBwd::kCB);
        compute_block_info(p.y, p.nJ, sRange, scratch_nv);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tD = tmem + 384;          // dF accumulator; S/G stage st lives at tmem + st*128

    if (warp == kProducerWarp) {
        if (lane == 0) {
            TileIter<false> iter(p, p.partD, sRange);
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
                const int slot = it % kSlots;
                trace_stamp(p, 0, it, 0);
                mbar_wait(b_empty + 8 * slot, ((it / kSlots) & 1) ^ 1);
                trace_stamp(p, 0, it, 1);
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
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            TileIter<false> iter(p, p.partD, sRange);
            const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
            const uint32_t idesc_d = umma_idesc_bf16(128, 128, 0, 1);   // B = F_J, MN-major
            int I, J, curI = -1, it = 0, seg = 0;
            bool last;
            // Issue order S(0) S(1) S(2) dF(0) S(3) dF(1) S(4) ...: S runs two tiles ahead of the
            // G.F_J product, so while one epilogue group turns S(it) into G(it) the other group's
            // S(it+1) is already complete and S(it+2) is in flight.  Stage it%3 is reused by
            // S(it+3), which is issued after dF(it) (in-order tensor pipe => G(it) is consumed).
            struct Pending { bool first, last; int seg; };
            Pending pend[2] = {{false, false, 0}, {false, false, 0}};
            int n_issued_dF = 0;
            auto issue_dF = [&](int pit) {
                const Pending& q = pend[pit & 1];
                const int pslot = pit % kSlots, pst = pit % kStages;
                trace_stamp(p, 1, pit, 2);
                mbar_wait(b_pfull + 8 * pst, (pit / kStages) & 1);
                trace_stamp(p, 1, pit, 3);
                if (q.first && q.seg > 0) mbar_wait(b_dempty, (q.seg - 1) & 1);
                tc_fence_after();
                if (!(p.debug & 4)) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ts(tD, tmem + pst * 128 + k * 8, ftile_desc_mnmajor(sJ + pslot * kTileBytes, k),
                            idesc_d, (!q.first) || k > 0);
                }
                tc_commit(b_empty + 8 * pslot);
                if (q.last) tc_commit(b_dfull);
            };
            while (iter.next(I, J, last)) {
                const bool first = (I != curI);
                if (first) {
                    mbar_wait(b_ifull, seg & 1);
                    curI = I;
                }
                const int slot = it % kSlots, st = it % kStages;
                trace_stamp(p, 1, it, 0);
                mbar_wait(b_full + 8 * slot, (it / kSlots) & 1);
                trace_stamp(p, 1, it, 1);
                tc_fence_after();
                if (!(p.debug & 2)) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ss(tmem + st * 128, ftile_desc_kmajor(sI, k),
                            ftile_desc_kmajor(sJ + slot * kTileBytes, k), idesc_s, k > 0);
                }
                tc_commit(b_tfull + 8 * st);
                if (last) tc_commit(b_iempty);
                if (it >= 2) { issue_dF(it - 2); n_issued_dF = it - 1; }
                pend[it & 1] = Pending{first, last, seg};
                if (last) ++seg;
                ++it;
            }
            for (int pit = n_issued_dF; pit < it; ++pit) issue_dF(pit);
        }
    } else {
        const int g = warp >> 2;                     // group g handles tiles with (it & 1) == g
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        TileIter<false> iter(p, p.partD, sRange);
        int I, J, curI = -1, it = 0, seg = 0;
        bool last;
        float4 rA = make_float4(0.f, 0.f, 0.f, 0.f), rB = make_float4(0.f, 1.f, 0.f, 0.f);
        int yi = -1, gi = -1;
        int2 rI = make_int2(INT_MAX, -1);
        while (iter.next(I, J, last)) {
            if (I != curI) {
                curI = I;
                gi = (p.rb0 + I) * 128 + r;
                rA = p.colA[gi];
                rB = p.colB[gi];
                yi = __float_as_int(rB.z);
                rI = sRange[p.rb0 + I];
            }
            if ((it & 1) == g) {
                const int slot = it % kSlots, st = it % kStages;
                if (threadIdx.x == g * 128) trace_stamp(p, 2, it, 0);
                mbar_wait(b_full + 8 * slot, (it / kSlots) & 1);
                mbar_wait(b_tfull + 8 * st, (it / kStages) & 1);
                if (threadIdx.x == g * 128) trace_stamp(p, 2, it, 1);
                tc_fence_after();
                const float4* cA = gCA + slot * 128;
                const float4* cB = gCB + slot * 128;
                const int col0 = J * 128;
                const uint32_t tS = tmem + st * 128 + lane_off;
                // fast tile: every pair is a plain "denominator" pair in both directions (no
                // positives, no self pair); padding needs no mask because padded F rows are zero
                const bool fast = !ranges_overlap(rI, sRange[J]) &&
                                  (kMode == DCL_MODE_PIXEL || (p.rb0 + I) != J);
                if (p.debug & 1) {
                    // diagnostics: no G is produced
                } else if (fast) {
                    for_each_chunk<4>(tS, [&](int c0, const uint32_t (&v)[32]) {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float gg[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const float s = __uint_as_float(v[j + u]);
                                const float4 ck = cA[c0 + j + u];
                                const float ti = fmaf(s, rA.x, rA.y);
                                const float tk = fmaf(s, ck.x, ck.y);
                                float acc = ti * rA.z;
                                acc = fmaf(tk, ck.z, acc);
                                acc = fmaf(DCL_EX2_BWD(ti, j + u), rA.w, acc);
                                acc = fmaf(DCL_EX2_BWD(tk, j + u + 4), ck.w, acc);
                                gg[u] = acc;
                            }
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(gg[0], gg[1]);
                            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
                        }
                        tmem_st16(tS + (c0 >> 1), pk);
                    });
                } else {
                    for_each_chunk<4>(tS, [&](int c0, const uint32_t (&v)[32]) {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float gg[2];
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
                                gg[u] = acc;
                            }
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(gg[0], gg[1]);
                            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
                        }
                        tmem_st16(tS + (c0 >> 1), pk);
                    });
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_pfull + 8 * st);
                if (threadIdx.x == g * 128) trace_stamp(p, 2, it, 2);
            }
            ++it;
            if (last) {
                // both groups drain half of the finished dF_I partial (64 columns each)
                mbar_wait(b_dfull, seg & 1);
                tc_fence_after();
                const int sidx = blockIdx.x - p.partD.first_cta(I);
                float* out = p.pD + ((static_cast<size_t>(I) * p.maxsegD + sidx) * 128 + r) * 128 + g * 64;
                for_each_chunk<2>(tD + lane_off + g * 64, [&](int c0, const uint32_t (&v)[32]) {
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
    if (warp == kProducerWarp) tmem_dealloc<kTmemCols>(tmem);
}

// dF[row] = sum over segments of the partials
__global__ void __launch_bounds__(256) k_reduce_dF(const Params p, float* __restrict__ dF) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);     // one warp per row
    const int lane = threadIdx.x & 31;
    if (row >= p.nI * 128) return;
    const int I = row >> 7, r = row & 127;
    const int ns = p.partD.nseg(I);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ns; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(
            p.pD + ((static_cast<size_t>(I) * p.maxsegD + s) * 128 + r) * 128 + lane * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dF + static_cast<size_t>(row) * 128 + lane * 4) = acc;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Layout {
    Part partS, partD;
    int nP, maxsegS, maxsegD, splitc;
    size_t off_pA, off_pB, off_pC, off_pD, off_bl, off_ticket, bytes;
};

static int g_debug_flags = 0;
static long long* g_trace = nullptr;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static void make_part(Part& part, int& maxseg, int nU, int nJ, int ctas) {
    part.total = static_cast<long long>(nU) * nJ;
    part.nJ = nJ;
    part.G = static_cast<int>(part.total < ctas ? part.total : ctas);
    const long long q = part.total / part.G;                  // >= 1 tiles per CTA
    maxseg = static_cast<int>((nJ + q - 1) / q + 1);
}

static Layout make_layout(int nI, int nJ) {
    Layout L;
    const int ctas = sm_count();
    L.nP = (nI + 1) / 2;
    make_part(L.partS, L.maxsegS, L.nP, nJ, ctas);
    make_part(L.partD, L.maxsegD, nI, nJ, ctas);
    int sc = (ctas + L.nP - 1) / L.nP;
    L.splitc = sc < 1 ? 1 : (sc > 8 ? 8 : sc);
    size_t o = 0;
    L.off_pA = o;     o = align_up(o + sizeof(float4) * static_cast<size_t>(nI) * L.maxsegS * 128, 256);
    L.off_pB = o;     o = align_up(o + sizeof(float2) * static_cast<size_t>(nI) * L.maxsegS * 128, 256);
    L.off_pC = o;     o = align_up(o + sizeof(float4) * static_cast<size_t>(nI) * L.splitc * 128, 256);
    L.off_bl = o;     o = align_up(o + sizeof(float) * nI, 256);
    L.off_ticket = o; o = align_up(o + sizeof(unsigned int), 256);
    L.off_pD = o;     o = align_up(o + sizeof(float) * static_cast<size_t>(nI) * L.maxsegD * 128 * 128, 256);
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
    p.nJ = nJ; p.rb0 = rb0; p.nI = nI; p.nP = L.nP; p.n_valid = n_valid; p.mode = mode;
    p.T = T; p.Tb = Tb;
    p.partS = L.partS;
    p.partD = L.partD;
    p.maxsegS = L.maxsegS;
    p.maxsegD = L.maxsegD;
    p.splitc = L.splitc;
    p.pA = reinterpret_cast<float4*>(w + L.off_pA);
    p.pB = reinterpret_cast<float2*>(w + L.off_pB);
    p.pC = reinterpret_cast<float4*>(w + L.off_pC);
    p.pD = reinterpret_cast<float*>(w + L.off_pD);
    p.blockloss = reinterpret_cast<float*>(w + L.off_bl);
    p.ticket = reinterpret_cast<unsigned int*>(w + L.off_ticket);
    p.debug = g_debug_flags;
    p.trace = g_trace;
    return p;
}

template <class K>
static int set_smem(K kernel, int bytes) {
    DCL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
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
    k_sweep<SWEEP_A, kMode><<<L.partS.G, kThreads, SmemSweep::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_sweep<A>");
    k_sweep<SWEEP_B, kMode><<<L.partS.G, kThreads, SmemSweep::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_sweep<B>");
    k_sweep<SWEEP_C, kMode><<<L.nP * L.splitc, kThreads, SmemSweep::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_sweep<C>");
    k_finalize<kMode><<<p.nI, 128, 0, st>>>(p);
    DCL_LAUNCH_CHECK("k_finalize");
    return 0;
}

template <int kMode>
static int run_bwd(Params p, const Layout& L, float* dF, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (int e = set_smem(k_backward<kMode>, SmemBwd::kBytes)) return e;
        configured = true;
    }
    k_backward<kMode><<<L.partD.G, kThreads, SmemBwd::kBytes, st>>>(p);
    DCL_LAUNCH_CHECK("k_backward");
    k_reduce_dF<<<(p.nI * 128 + 7) / 8, 256, 0, st>>>(p, dF);
    DCL_LAUNCH_CHECK("k_reduce_dF");
    return 0;
}

}  // namespace dcl

using namespace dcl;

// Diagnostics only: component-isolation switches for profiling (results are invalid when non-zero).
extern "C" int dcl_debug_flags(int flags) {
    const int old = g_debug_flags;
    g_debug_flags = flags;
    return old;
}

// Diagnostics only: device buffer of 3*32*4 int64 that CTA 0 of k_backward fills with clock64 stamps.
extern "C" int dcl_debug_trace(void* device_buffer) {
    g_trace = static_cast<long long*>(device_buffer);
    return 0;
}

extern "C" size_t dcl_contrast_workspace_bytes(int nI, int nJ) {
    if (nI <= 0 || nJ <= 0) return 0;
    return make_layout(nI, nJ).bytes;
}

static int check_args(const void* tiles, const int32_t* y, int nJ, int rb0, int nI, int mode, void* ws,
                      size_t ws_bytes, const Layout& L) {
    if (!tiles || !y || !ws) return fail(DCL_ERR_ARG, "null pointer argument");
    if (nJ <= 0 || nI <= 0 || rb0 < 0 || rb0 + nI > nJ)
        return fail(DCL_ERR_ARG, "bad block range rb0=%d nI=%d nJ=%d", rb0, nI, nJ);
    if (nJ > kMaxBlocks)
        return fail(DCL_ERR_ARG, "contrast set of %d rows exceeds the supported %d", nJ * 128, kMaxBlocks * 128);
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
