// N x N row-normalised contrast (reference utils/loss.py:339-389 pixel term, :175-204 image term)
// as tcgen05/TMEM tile sweeps; the N x N matrix only ever exists as 128x128 fp32 tiles in TMEM.
//
// Math, in raw dot-product units s_ij = f_i . f_j (a_ij = s_ij / T):
//   l_ij = (s_ij - m_i) kappa_i,  m_i = max_j s_ij (detached),  kappa_i = 1 / max(|s_i - m_i|_2, T*1e-12)
//   pixel: lp_ij = l_ij - log(e^{l_ij} + Den_i), Den_i = sum_{y_k != y_i} e^{l_ik};  image: lp_ij = l_ij - log sum_{k != i} e^{l_ik}
//   loss = mean_i [ -(T/T_b) mean_{j in pos(i)} lp_ij ];  dS_ik = kappa_i (g_ik - l_ik R_i);  dF = (dS + dS^T) F
//
// Two forward pipelines share the kernels below:
//   * pixel term ("v3", the hot path).  e^{l} on a row's logit range [-L_i, 0] (L_i <= 1 because |l_i|_2 = 1) is a
//     per-row economised Chebyshev polynomial of degree n = 1..4, so  Den_i = sum_neg E = sum_j e_j P_j  and
//     sum_neg E x = sum_j e_j P_{j+1}  with the shifted power sums P_j = sum_neg x^j, x = s - c_i (c_i = |f_i|^2, the
//     diagonal: the shift keeps every sum well conditioned; the tensor pipe applies it, as a ninth K step of the
//     product with -c_i split into three bf16 against a ones operand).  ONE sweep (SWEEP_P) produces, per row, the exact
//     maximum and P_0..P_2 of the different-class columns plus Q_0..Q_2 of the same-class columns (packed FFMA2, no
//     MUFU, no per-row constants at all in the epilogue); kappa_i comes from P + Q, and with n = 1 (a few thousand anchors or
//     more) everything else follows per row in k_rows.  Rows whose range needs n > 1 trigger a second sweep
//     (SWEEP_H: x^3..x^{n+1}) that otherwise exits at once; the positive-pair terms use a first-order series in
//     E/Den and fall back to the exact sweep C only when some row has fewer than ~170 negatives.
//   * image term and the diagnostic legacy path: sweep A (max, sums), sweep B (Den), as in round 1.
//   Both continue with sweep C (tiles that can hold positives), finalize, and the fused backward
//         G_ik = dS_ik + dS_ki = rp_i(s_ik) + rp_k(s_ik)   on pairs of different classes,
//     rp = q E(s) + p (a s + b) folded into one polynomial per row (k_bwd_prep), G written as bf16 over S in TMEM,
//     dF_I += G F_J with A from TMEM.  Tiles whose label ranges overlap take the exact masked path.
//
// Kernel shape: one persistent CTA per SM, 352 threads: warps 0..7 = two epilogue groups of four warps (warp w owns
// TMEM lanes 32*(w%4).., one thread per anchor row), warp 8 = TMA producer lane + TMEM allocator, warps 9 and 10 =
// two MMA issuer lanes.  tcgen05.mma issue blocks its thread while the tensor pipe drains, and every mbarrier
// round trip costs ~150-200 clocks, so a single issuer thread leaves the tensor pipe idle half of the time
// (profiles/r01f_*): with two issuers one thread's waits overlap the other's MMAs.
//   sweeps  : a CTA owns a PAIR of row blocks (M = 256): every F_J tile fetched from L2 feeds two 128x128 S tiles,
//             one per epilogue group.  The row blocks are the A operand in TMEM (2 x 64 columns; they arrive through
//             the TMA ring and tcgen05.cp issued by the MMA warps), three 128-column accumulators rotate over the jobs.
//   backward: a CTA owns one row block; issuer 0 produces S tiles (3 TMEM stages), issuer 1 the G.F_J products; the
//             two epilogue groups alternate column tiles.
// Work = the flattened list of (row unit, column block) tiles cut into gridDim.x equal contiguous ranges; a range
// that crosses a row-unit boundary flushes a deterministic partial ("segment") instead of using atomics, so results
// are bit-reproducible.
#include <cfloat>
#include <type_traits>
#include <utility>
#include <cuda_bf16.h>
#include "dcl_common.cuh"
#include "dcl_ptx.cuh"

namespace dcl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 352;
constexpr int kTmemCols = 512;
constexpr int kProducerWarp = 8;
constexpr int kIssuerWarp0 = 9;       // issuer g is warp 9 + g
constexpr int kMaxBlocks = 1024;      // column blocks whose info is cached in shared memory
constexpr int kCoefPairFloats = 12;   // backward column coefficients of one column pair: 3 x float4

// ---------------------------------------------------------------------------------------------
// Partition of nU row units x nJ column blocks (the flattened tile list, unit-major) over G CTAs.
//   flat     : G equal contiguous ranges; a range may cross unit boundaries (the CTA then works on two or more
//              units one after the other, each ending in a pipeline drain, an operand reload and a segment partial).
//   exclusive: when there are at least as many CTAs as units, every unit owns `base` or `base + 1` whole CTAs
//              (the first `extra` units get one more) that split its nJ column blocks evenly; no CTA crosses a
//              unit.  Chosen by make_part when its longest range beats the flat range plus one unit switch.
struct Part {
    long long total;
    int G, nJ;
    int excl, base, extra;
    __host__ __device__ long long begin(int c) const {
        if (!excl) return static_cast<long long>(c) * total / G;
        const int wide = extra * (base + 1);             // CTAs of the units that own base + 1
        int u, k, parts;
        if (c < wide) { u = c / (base + 1); k = c - u * (base + 1); parts = base + 1; }
        else { const int d = c - wide; u = extra + d / base; k = d - (u - extra) * base; parts = base; }
        return static_cast<long long>(u) * nJ + static_cast<long long>(k) * nJ / parts;
    }
    // CTA whose range contains flat tile index x (flat mode)
    __host__ __device__ int cta_of(long long x) const {
        return static_cast<int>(((x + 1) * G + total - 1) / total - 1);
    }
    __host__ __device__ int first_cta(int U) const {
        if (excl) return U < extra ? U * (base + 1) : extra * (base + 1) + (U - extra) * base;
        return cta_of(static_cast<long long>(U) * nJ);
    }
    __host__ __device__ int nseg(int U) const {
        if (excl) return U < extra ? base + 1 : base;
        return cta_of(static_cast<long long>(U + 1) * nJ - 1) - first_cta(U) + 1;
    }
};

struct Params {
    const uint8_t* tiles;    // [nJ] F-tiles
    const int32_t* y;        // [nJ*128]
    const float* sqnorm;     // [nJ*128]
    int nJ, rb0, nI, nP, n_valid, mode, ctas;
    int jstride;             // column-block visiting stride (coprime with nJ)
    int use_series;          // v3 forward: positive-pair sums come from the series unless iscal[4] says otherwise
    float T, Tb;
    Part partS;              // sweeps: units = pairs of row blocks
    Part partD;              // backward: units = row blocks
    int maxsegS, maxsegD;
    int splitc;              // sweep C: CTAs per row-block pair
    // block info (all nJ column blocks)
    int4* binfo;             // (ymin, ymax, nvalid, backward: largest polynomial degree); (INT_MAX, -1, 0) when no row is valid
    float2* bnorm;           // (max, min) of |f|^2 over the block's valid rows
    // legacy sweeps A / B
    float4* pA;              // [nI][maxsegS][128] (max(s-c), S1, S2, -)
    float2* pB;              // [nI][maxsegS][128] (Den, Bt)
    // v3 power-sum forward
    int* iscal;              // [0] largest polynomial degree of the local rows,
                             // [4] != 0: some row has too few negatives for the positive-pair series (sweep C runs)
                             // [5] != 0: some row was not finished by k_rows (k_combine2 and k_finalize run)
    float4* pF;              // [nI][maxsegS][2][128] sweep P partials: (max x, P0, P1, P2) (Q0, Q1, Q2, -)
    float4* pH;              // [nI][maxsegS][2][128] sweep H partials: (P3, P4, P5, -) (Q3, Q4, Q5, -)
    float* rowM;             // [nI*128][8] (P0, P1, P2, Q0, Q1, Q2, d = max x, L)
    float4* rowPos;          // [nI*128] positive-pair sums from the series: (P, sum lp, sum inv, sum inv*l)
    int* rowflag;            // [nI*128] != 0: k_rows left the row to k_combine2 / sweep C / k_finalize
    // per-row state shared by sweep C / finalize
    float4* rowS;            // [nI*128] (a, b, kappa, m)   t = a s + b = l log2(e)
    float4* rowD;            // [nI*128] (Den, Bt = sum_den E t, L, 0)
    float4* pC;              // [nI][splitc][128] (P, sum lp | sum l, sum inv, sum inv*l)
    float* pD;               // [nI][maxsegD][128][128] dF partials
    float* dF;               // [nI*128][128] backward result
    unsigned int* dticket;   // [nI] backward: segments of a row block that have landed
    float4* colA;            // [nJ*128] (a, b, p, q)
    float4* colB;            // [nJ*128] (wn, Den, y bits, L)
    float* coefR;            // [nJ*128][16] backward row polynomials (n0..n4 - - - p0..p4 - - -), see k_bwd_prep
    float* coefP;            // [nJ*64][12] different-class polynomial, pair-interleaved for the column side (+ labels)
    float* coefPP;           // [nJ*64][12] same-class polynomial, pair-interleaved
    float* bden;             // [nJ] smallest Den of each block's valid rows
    float* rowloss;          // [nJ*128]
    float* blockloss;        // [nI]
    float* loss_sum;
    float* loss_out;         // optional: the loss itself (sum / n_valid) once more, where the caller wants it
    unsigned int* ticket;    // last-block counters: [0] k_rows, [1] k_finalize
    long long* trace;        // diagnostics only (dcl_debug_trace)
    long long* cta_times;    // diagnostics only (dcl_debug_cta_times): [kernel slot][256 CTAs][4] globaltimer / clock64 at entry, exit
    int debug;               // diagnostics only (dcl_debug_flags)
};

// ---------------------------------------------------------------------------------------------
// small math helpers
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

// One lane of a fully converged warp.  The single-lane pipeline roles run with ALL lanes executing the (warp-uniform)
// control flow and only the asynchronous instructions predicated on this: ptxas then keeps descriptors, barrier
// addresses and loop state in uniform registers.  Under a divergent `if (lane == 0)` every tcgen05.mma / bulk copy
// is preceded by a ~16-instruction ELECT / R2UR.BROADCAST / BRA.U.ANY uniformisation loop, which made MMA *issue*
// (~110 clk per MMA next to busy epilogue warps) the bottleneck of the first versions.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Programmatic dependent launch: every kernel of a forward / backward chain lets its successor start at once
// (launch_dependents at the top) and waits for its predecessor's results (wait) before the first access to global
// memory that a predecessor writes or still reads; the successor's launch latency and set-up then overlap this
// kernel.  EVERY thread passes pdl_wait() before it exits: a grid that completed without waiting would release its
// own successor while the grid before it still runs.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// The same walk as a real loop (one chunk body in the instruction stream instead of four): the pipelined kernels
// hold several alternative epilogue bodies, and unrolled they overflow the instruction cache.  The TMEM load
// latency of a chunk is covered by the other epilogue warp of the scheduler.
template <class Fn>
__device__ __forceinline__ void for_each_chunk_loop(uint32_t taddr, Fn&& fn) {
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait_on(v);
        fn(c0, v);
    }
}
// diagnostics: stamp (role, tile, event) for CTA 0's first 32 tiles; roles: 0 producer, 1/2 issuers, 3/4 epilogue groups
__device__ __forceinline__ void trace_stamp(const Params& p, int role, int it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 32) p.trace[(role * 32 + it) * 8 + ev] = clock64();
}

__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// kernel slot: 0 sweep P, 1 backward, 2 k_rows, 3 k_bwd_prep; which = 0 entry, 1 exit (one thread per CTA)
__device__ __forceinline__ void cta_time(const Params& p, int slot, int which) {
    if (p.cta_times && blockIdx.x < 256) {
        long long* o = p.cta_times + (static_cast<size_t>(slot) * 256 + blockIdx.x) * 4 + which * 2;
        o[0] = globaltimer_ns();
        o[1] = clock64();
    }
}

__device__ __forceinline__ bool ranges_overlap(int2 a, int2 b) { return a.x <= b.y && b.x <= a.y; }

// =============================================================================================
// Per-row exponential polynomial.  For l = kappa (s - m) in [-L, 0]:
//   e^l = e^{-r} e^z,  z = l + r in [-r, r],  r = L/2,   e^z ~ truncated Chebyshev series (modified Bessel I_k(r)),
// re-expanded in s.  Truncation error ~ 2 I_{deg+1}(r): deg 1 for L <= 0.03 (5.6e-5: the logits of a few thousand
// anchors span so little that exp is linear to that accuracy), 2 for L <= 1/8 (1.1e-5), 3 for L <= 1/2 (2e-5), else 4
// (1.6e-5 at L = 1).
// =============================================================================================
__device__ __forceinline__ int poly_degree_for(double L) { return L <= 0.03 ? 1 : (L <= 0.125 ? 2 : (L <= 0.5 ? 3 : 4)); }

__device__ inline void exp_poly_in_s(double kappa, double m, double L, int deg, double (&d)[5]) {
    double r = 0.5 * L;
    if (r < 1e-8) r = 1e-8;
    // a_k = (2 - [k == 0]) I_k(r),  I_k(r) = (r/2)^k / k! * sum_j (r^2/4)^j / (j! (k+1)..(k+j)): with r <= 1/2 the
    // j-th term is below 0.0625^j / j!^2, five terms reach 1e-12.  Reciprocals are compile-time constants: this
    // runs once per row on the critical path of k_rows / k_bwd_prep, where fp64 divisions dominate the latency.
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const double h = 0.5 * r, h2 = h * h;
    double hk = 1.0;
#pragma unroll
    for (int k = 0; k <= 4; ++k) {
        if (k <= deg) {
            if (k > 0) hk *= h * (1.0 / k);          // h^k / k!
            double sum = 1.0;
#pragma unroll
            for (int j = 4; j >= 1; --j) sum = 1.0 + sum * h2 * (1.0 / (static_cast<double>(j) * (j + k)));   // Horner in h^2
            a[k] = (k == 0 ? 1.0 : 2.0) * hk * sum;
        }
    }
    // Chebyshev -> monomials in x = z / r
    double cx[5] = {0, 0, 0, 0, 0};
    cx[0] = a[0];
    cx[1] = a[1];
    if (deg >= 2) { cx[0] -= a[2]; cx[2] = 2.0 * a[2]; }
    if (deg >= 3) { cx[1] -= 3.0 * a[3]; cx[3] = 4.0 * a[3]; }
    if (deg >= 4) { cx[0] += a[4]; cx[2] -= 8.0 * a[4]; cx[4] = 8.0 * a[4]; }
    const double er = exp(-r), ir = 1.0 / r;
    double cz[5], irn = er;
#pragma unroll
    for (int n = 0; n <= 4; ++n) { cz[n] = (n <= deg) ? irn * cx[n] : 0.0; irn *= ir; }
    // z = al s + be
    const double al = kappa, be = r - kappa * m;
    const double binom[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};
    double alp[5], bep[5];
    alp[0] = bep[0] = 1.0;
#pragma unroll
    for (int n = 1; n <= 4; ++n) { alp[n] = alp[n - 1] * al; bep[n] = bep[n - 1] * be; }
#pragma unroll
    for (int j = 0; j <= 4; ++j) {
        double s = 0.0;
#pragma unroll
        for (int n = j; n <= 4; ++n) s += cz[n] * binom[n][j] * alp[j] * bep[n - j];      // cz[n] = 0 for n > deg
        d[j] = s;
    }
}

// kappa and logit-range bound of a row from its squared norm sum_k (s_ik - m)^2.  A row whose every s_ik equals the
// maximum to 1e-6 relative (all embeddings identical) has l = 0 in the reference (0 / eps); it is mapped to kappa = 0.
__device__ __forceinline__ void row_scale(double nrm2, double m, double c, double cmax, float T, double& kappa, double& L) {
    if (!(nrm2 > 0.0)) nrm2 = 0.0;
    const double nrm = sqrt(nrm2);
    if (nrm <= 1e-6 * fabs(m) || nrm <= static_cast<double>(T) * 1e-12) {
        kappa = 0.0;
        L = 0.0;
        return;
    }
    kappa = 1.0 / nrm;
    // s_ik >= -sqrt(c_i c_k) (Cauchy-Schwarz) and |l_i|_2 = 1
    L = fmin(1.0, kappa * (m + sqrt(c * cmax)) * (1.0 + 1e-6));
    if (L < 0.0) L = 0.0;
}

// =============================================================================================
// Block info
// =============================================================================================
// (ymin, ymax, nvalid) and (cmax, cmin) of one 128-row block; threads 0..127 hold one row each, the result is valid
// in thread 0 after the second barrier.  Must be called by all threads of the CTA (it uses __syncthreads).
struct BlockStat {
    int lo, hi, n;
    float cx, cn;
};
__device__ __forceinline__ BlockStat block_stat(int yv, float c, bool participates, int* si, float* sf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int lo = yv >= 0 ? yv : INT_MAX, hi = yv >= 0 ? yv : -1, n = yv >= 0;
    float cx = yv >= 0 ? c : 0.f, cn = yv >= 0 ? c : FLT_MAX;
    if (participates) {
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            n += __shfl_xor_sync(0xffffffffu, n, o);
            cx = fmaxf(cx, __shfl_xor_sync(0xffffffffu, cx, o));
            cn = fminf(cn, __shfl_xor_sync(0xffffffffu, cn, o));
        }
        if (lane == 0) { si[warp] = lo; si[4 + warp] = hi; si[8 + warp] = n; sf[warp] = cx; sf[4 + warp] = cn; }
    }
    __syncthreads();
    BlockStat b;
    b.lo = min(min(si[0], si[1]), min(si[2], si[3]));
    b.hi = max(max(si[4], si[5]), max(si[6], si[7]));
    b.n = si[8] + si[9] + si[10] + si[11];
    b.cx = fmaxf(fmaxf(sf[0], sf[1]), fmaxf(sf[2], sf[3]));
    b.cn = fminf(fminf(sf[4], sf[5]), fminf(sf[6], sf[7]));
    __syncthreads();
    return b;
}

// =============================================================================================
// shared-memory carve-up (bytes from a 1024-aligned base)
// =============================================================================================
struct SmemSweep {
    static constexpr int kSlots = 6;                      // F_J ring (the row blocks live in TMEM)
    static constexpr int kJ = 0;
    static constexpr int kInfo = kSlots * kTileBytes;     // int4[kMaxBlocks]
    static constexpr int kLab = kInfo + 16 * kMaxBlocks;  // 8 epilogue warps x 128 labels of the current masked column block
    static constexpr int kBar = kLab + 8 * 512;           // full[6] empty[6] tfull[3] tempty[3] afull[2] turn[2] aseen
    static constexpr int kShiftA = kBar + 256;            // sweep P: -c_i of the two row blocks as a K = 16 operand (2 x 4 KB)
    static constexpr int kShiftB = kShiftA + 2 * 4096;    // sweep P: the matching ones rows (256 B, every row group aliases it)
    static constexpr int kTmem = kShiftB + 256;
    static constexpr int kBytes = kTmem + 16 + 1024;      // + alignment slack
};
struct SmemBwd {
    static constexpr int kSlots = 5;                      // F_J tile + its column polynomial block
    static constexpr int kStages = 3;                     // TMEM S/G stages (G aliases its S)
    static constexpr int kCoefBytes = 64 * kCoefPairFloats * 4;   // 3072
    static constexpr int kI = 0;
    static constexpr int kJ = kTileBytes;
    static constexpr int kCP = (1 + kSlots) * kTileBytes;
    static constexpr int kCQ = kCP + kSlots * kCoefBytes;         // same-class column polynomials (tiles that can hold such pairs)
    static constexpr int kRange = kCQ + kSlots * kCoefBytes;      // uint16[kMaxBlocks]: label range of a block, lo | hi << 8
    static constexpr int kDenOk = kRange + 2 * kMaxBlocks;        // uint8[kMaxBlocks]: every Den of the block >= kMinDenSeries
    static constexpr int kBar = kDenOk + kMaxBlocks;      // full[5] empty[5] tfull[3] pfull[3] sfree[3] dfull dempty ifull iempty
    static constexpr int kTmem = kBar + 256;
    static constexpr int kBytes = kTmem + 16 + 1024;
};

enum { SWEEP_A = 0, SWEEP_B = 1, SWEEP_C = 2, SWEEP_P = 3, SWEEP_H = 4 };

// Tile sequence of one CTA: (row unit U, column block J).  Flat mode: contiguous range of the flattened U-major
// list.  Relevant mode (sweep C): CTA = (pair, split s) walks
// the column blocks whose label range overlaps either row block of the pair and keeps every splitc-th one.  All
// warp roles run an identical copy of this iterator, which keeps their barrier phases in step.
template <bool kRelevantOnly>
struct TileIter {
    int left;                 // flat mode: tiles remaining
    int nJ, U, J, r, s, splitc;
    int jp, jstride;          // flat mode: permuted column block (J * jstride) % nJ, kept incrementally
    int2 r0, r1;
    const int4* info;
    __device__ TileIter(const Params& p, const Part& part, const int4* sInfo) {
        nJ = p.nJ;
        info = sInfo;
        left = 0;
        U = J = r = s = 0;
        splitc = 1;
        r0 = r1 = make_int2(INT_MAX, -1);
        if (kRelevantOnly) {
            U = blockIdx.x / p.splitc;
            s = blockIdx.x % p.splitc;
            splitc = p.splitc;
            const int4 a = sInfo[p.rb0 + 2 * U];
            r0 = make_int2(a.x, a.y);
            if (2 * U + 1 < p.nI) {
                const int4 b = sInfo[p.rb0 + 2 * U + 1];
                r1 = make_int2(b.x, b.y);
            }
        } else if (part.total > 0) {
            const long long t0 = part.begin(blockIdx.x), t1 = part.begin(blockIdx.x + 1);
            left = static_cast<int>(t1 - t0);
            U = static_cast<int>(t0 / nJ);            // the only divisions: once per CTA and role
            J = static_cast<int>(t0 - static_cast<long long>(U) * nJ);
        }
        // Column blocks are visited in a stride-permuted order: the tiles that hold same-class pairs are
        // consecutive in J (rows are class-sorted) and cost 2-3x a plain tile; the permutation spreads them over
        // the CTAs that share a row unit instead of piling them onto one.
        jstride = p.jstride;
        jp = static_cast<int>((static_cast<long long>(J) * jstride) % nJ);
    }
    __device__ bool next(int& oU, int& oJ, bool& last_of_seg) {
        if (kRelevantOnly) {
            while (J < nJ) {
                const int j = J++;
                const int4 q = info[j];
                const int2 rj = make_int2(q.x, q.y);
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
            oJ = jp;
            --left;
            jp += jstride;
            if (jp >= nJ) jp -= nJ;
            if (++J == nJ) { J = 0; jp = 0; ++U; }
            last_of_seg = (left == 0) || (J == 0);
            return true;
        }
    }
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// SWEEP_P fast-tile body for 32 columns (16 packed pairs), x = s - c_i: row max, sum x, sum x^2
template <bool kShiftedByMma = false>
__device__ __forceinline__ void psweep_chunk(const uint32_t (&v)[32], f32x2 negc, f32x2 (&x1)[4], f32x2 (&x2)[4], float (&mx)[4]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2 x = kShiftedByMma ? pack2u(v[2 * j], v[2 * j + 1]) : fadd2(pack2u(v[2 * j], v[2 * j + 1]), negc);
        float lo, hi;
        unpack2(x, lo, hi);
        mx[j & 3] = fmaxf(mx[j & 3], fmaxf(lo, hi));
        x1[j & 3] = fadd2(x1[j & 3], x);
        x2[j & 3] = ffma2(x, x, x2[j & 3]);
    }
}
// SWEEP_P body for a tile that may hold same-class pairs, labels of its 128 columns staged in `wy` (shared):
// sums over every valid column (a1, a2; the count comes from the block info) and over the different-class ones
// (n0..n2).  The same-class sums follow as all - different - self.
template <bool kAllValid>
__device__ __forceinline__ void psweep_chunk_masked(const uint32_t (&v)[32], const int* __restrict__ wy, int yi, f32x2 negc,
                                                    f32x2& a0, f32x2& a1, f32x2& a2, f32x2& n0, f32x2& n1, f32x2& n2,
                                                    float (&mx)[4]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int2 yy = *reinterpret_cast<const int2*>(wy + 2 * j);
        const f32x2 x = fadd2(pack2u(v[2 * j], v[2 * j + 1]), negc);
        const f32x2 xx = fmul2(x, x);
        float lo, hi;
        unpack2(x, lo, hi);
        const bool v0 = kAllValid || yy.x >= 0, v1 = kAllValid || yy.y >= 0;
        const f32x2 wn = pack2((v0 && yy.x != yi) ? 1.f : 0.f, (v1 && yy.y != yi) ? 1.f : 0.f);
        if (kAllValid) {
            mx[j & 3] = fmaxf(mx[j & 3], fmaxf(lo, hi));
            a1 = fadd2(a1, x);
            a2 = fadd2(a2, xx);
        } else {
            const f32x2 wv = pack2(v0 ? 1.f : 0.f, v1 ? 1.f : 0.f);
            mx[j & 3] = fmaxf(mx[j & 3], fmaxf(v0 ? lo : -FLT_MAX, v1 ? hi : -FLT_MAX));
            a0 = fadd2(a0, wv);
            a1 = ffma2(wv, x, a1);
            a2 = ffma2(wv, xx, a2);
        }
        n0 = fadd2(n0, wn);
        n1 = ffma2(wn, x, n1);
        n2 = ffma2(wn, xx, n2);
    }
}
// SWEEP_H fast-tile body: sum x^3 .. x^{deg+1}
template <int kDeg>
__device__ __forceinline__ void hsweep_chunk(const uint32_t (&v)[32], f32x2 negc, f32x2 (&h3)[4], f32x2 (&h4)[4], f32x2 (&h5)[4]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2 x = fadd2(pack2u(v[2 * j], v[2 * j + 1]), negc);
        const f32x2 xx = fmul2(x, x);
        h3[j & 3] = ffma2(xx, x, h3[j & 3]);
        if (kDeg >= 3) h4[j & 3] = ffma2(xx, xx, h4[j & 3]);
        if (kDeg >= 4) h5[j & 3] = ffma2(fmul2(xx, x), xx, h5[j & 3]);
    }
}

// =============================================================================================
// Sweeps A / B / C / P / H
// =============================================================================================
template <int kSweep, int kMode>
__global__ void __launch_bounds__(kThreads, 1) k_sweep(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    pdl_launch();
    // conditional sweeps leave before any set-up when their trigger (written by k_rows) is clear; the others do
    // their set-up first and wait for the predecessor just before the first global read
    constexpr bool kConditional = kSweep == SWEEP_H || (kSweep == SWEEP_C && kMode == DCL_MODE_PIXEL);
    if (kConditional) {
        pdl_wait();
        if (kSweep == SWEEP_H && p.iscal[0] <= 1) return;
        if (kSweep == SWEEP_C && p.use_series && p.iscal[4] == 0) return;
    }
    if (threadIdx.x == 0) trace_stamp(p, 0, 0, 7);          // kernel entry
    if (kSweep == SWEEP_P && threadIdx.x == 0) cta_time(p, 0, 0);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sJ = base + SmemSweep::kJ;
    int4* sInfo = reinterpret_cast<int4*>(gen + SmemSweep::kInfo);
    int* sLab = reinterpret_cast<int*>(gen + SmemSweep::kLab);
    const uint32_t bar = base + SmemSweep::kBar;
    constexpr int kSlots = SmemSweep::kSlots;
    // tfull / tempty are per accumulator buffer (job % 3); afull per row block of the pair; turn per issuer
    const uint32_t b_full = bar, b_empty = bar + 48, b_tfull = bar + 96, b_tempty = bar + 120,
                   b_afull = bar + 144, b_turn = bar + 160, b_aseen = bar + 176;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemSweep::kTmem);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

    if (warp == kProducerWarp) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
    if (threadIdx.x == kIssuerWarp0 * 32) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 1);
        }
        for (int s = 0; s < 3; ++s) {
            mbar_init(b_tfull + 8 * s, 1);
            mbar_init(b_tempty + 8 * s, 4);     // the 4 warps of one epilogue group
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(b_afull + 8 * s, 4);
            mbar_init(b_turn + 8 * s, 1);
        }
        mbar_init(b_aseen, 2);                  // both issuers have observed the current row blocks
        mbar_fence_init();
    }
    // Sweep P folds the row shift x = s - c_i into the product: a ninth K = 16 step with A' = (-c_hi, -c_mid, -c_lo, 0..)
    // per row (c_i split into three bf16, exact to 2^-24) and B' = (1, 1, 1, 0..) per column, both K-major without
    // swizzle in shared memory.  That takes the subtraction (one of four FMA-pipe instructions per element pair) out
    // of the epilogue, which is what bounds the sweep, for 1/8 more tensor work.
    const uint32_t sShiftA = base + SmemSweep::kShiftA, sShiftB = base + SmemSweep::kShiftB;
    if (kSweep == SWEEP_P && threadIdx.x < 16) {
        // two core matrices (k 0..7, k 8..15) of 8 rows x 16 bytes; every 8-row group aliases them (stride 0)
        const uint32_t one = 0x3f80u;                 // bf16 1.0
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < 8) v = make_uint4(one | (one << 16), one, 0u, 0u);
        reinterpret_cast<uint4*>(gen + SmemSweep::kShiftB)[threadIdx.x] = v;
        fence_proxy_async_smem();
    }
    if (!kConditional) pdl_wait();
    for (int j = threadIdx.x; j < p.nJ; j += kThreads) sInfo[j] = p.binfo[j];
    const Part& part = p.partS;
    const int deg = (kSweep == SWEEP_H) ? p.iscal[0] : 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) trace_stamp(p, 0, 2, 7);          // set-up done
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // TMEM map: row block g of the pair as the bf16 A operand at columns [64 g, 64 g + 64); three 128-column fp32
    // accumulator buffers at 128, 256, 384.  Job j = 2 * tile + g uses buffer j % 3.
    const uint32_t tAcc = tmem + 128;

    if (warp == kProducerWarp) {
        // ---------------------------------------------------------------------- TMA producer
        // Ring entries in order: at every unit start the unit's row blocks (the issuer copies them on into TMEM with
        // tcgen05.cp and frees the slots at once), then one entry per column block.  `pos` counts entries.
        TileIter<kSweep == SWEEP_C> iter(p, part, sInfo);
        int U, J, curU = -1, it = 0, pos = 0;
        bool last;
        auto push = [&](int block) {
            const int slot = pos % kSlots;
            mbar_wait(b_empty + 8 * slot, ((pos / kSlots) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(b_full + 8 * slot, kTileBytes);
                tma_bulk_g2s(sJ + slot * kTileBytes, p.tiles + static_cast<size_t>(block) * kTileBytes,
                             kTileBytes, b_full + 8 * slot);
            }
            __syncwarp();
            ++pos;
        };
        while (iter.next(U, J, last)) {
            if (U != curU) {
                curU = U;
                push(p.rb0 + 2 * U);
                if (2 * U + 1 < p.nI) push(p.rb0 + 2 * U + 1);
            }
            if (lane == 0) trace_stamp(p, 0, it, 0);
            push(J);
            if (lane == 0) trace_stamp(p, 0, it, 1);
            ++it;
        }
    } else if (warp >= kIssuerWarp0) {
        {
            // ------------------------------------------------------------------ MMA issuer s (whole warp, see elect_one)
            // tcgen05.mma issue blocks its thread while the tensor pipe drains and every barrier wait costs
            // 150-200 clocks, so two issuers alternate bursts: issuer s owns the tiles with it % 2 == s (16 MMAs:
            // both row blocks of the pair) and does its waits for the next burst while the other one's burst
            // executes.  The turn is handed over two MMAs before the end of a burst (bursts touch different
            // accumulators, so their tails may interleave).  A comes from TMEM (TS form): with both operands in
            // shared memory a 128x128x16 MMA needs the full 128 B/clk of shared-memory bandwidth and the TMA
            // writes of the next tiles push it to ~110 clk per MMA (profiles/r01f_*).
            const int s = warp - kIssuerWarp0;
            TileIter<kSweep == SWEEP_C> iter(p, part, sInfo);
            const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
            uint64_t dJ0[8];                         // built once; a slot only shifts the start address
#pragma unroll
            for (int k = 0; k < 8; ++k) dJ0[k] = ftile_desc_kmajor(sJ, k);
            // shift operands (sweep P): core matrices 128 B apart along K; 8-row groups 256 B apart (A') / aliased (B')
            const uint64_t dShiftA0 = umma_smem_desc_noswizzle(sShiftA, 128, 256);
            const uint64_t dShiftA1 = umma_smem_desc_noswizzle(sShiftA + 4096, 128, 256);
            const uint64_t dShiftB = umma_smem_desc_noswizzle(sShiftB, 128, 0);
            int U, J, curU = -1, it = 0, seg = 0, turn = 0, pos = 0, posA = 0;
            bool last, two = false, fresh = false;
            while (iter.next(U, J, last)) {
                if (U != curU) {
                    posA = pos;                          // ring entries of the unit's row blocks
                    pos += (2 * U + 1 < p.nI) ? 2 : 1;
                    fresh = true;                        // the issuer that owns this tile copies them into TMEM
                    // BOTH issuers wait for every unit's row blocks (a waiter that skipped a phase would alias
                    // the 1-bit parity) and acknowledge, so the epilogue never runs two phases ahead of one
                    two = 2 * U + 1 < p.nI;
                    curU = U;
                    mbar_wait(b_afull, seg & 1);
                    mbar_wait(b_afull + 8, seg & 1);
                    if (elect_one()) mbar_arrive(b_aseen);
                    __syncwarp();
                    ++seg;
                }
                const int slot = pos % kSlots, slot_phase = (pos / kSlots) & 1;
                ++pos;
                const bool copyA = fresh;
                fresh = false;
                if ((it & 1) == s) {
                    const int j0 = 2 * it, j1 = 2 * it + 1;
                    const int b0 = j0 % 3, b1 = j1 % 3;
                    if (lane == 0) trace_stamp(p, 1 + s, it, 0);
                    if (copyA) {
                        mbar_wait(b_full + 8 * (posA % kSlots), (posA / kSlots) & 1);
                        if (two) mbar_wait(b_full + 8 * ((posA + 1) % kSlots), ((posA + 1) / kSlots) & 1);
                    }
                    mbar_wait(b_full + 8 * slot, slot_phase);
                    if (lane == 0) trace_stamp(p, 1 + s, it, 1);
                    mbar_wait(b_tempty + 8 * b0, ((j0 / 3) & 1) ^ 1);
                    if (lane == 0) trace_stamp(p, 1 + s, it, 2);
                    mbar_wait(b_turn + 8 * s, (turn & 1) ^ (s == 0 ? 1 : 0));
                    ++turn;
                    if (lane == 0) trace_stamp(p, 1 + s, it, 4);
                    tc_fence_after();
                    const uint64_t soff = static_cast<uint64_t>(slot * (kTileBytes >> 4));
                    const bool run = !(p.debug & 2);
                    if (copyA && elect_one()) {
                        // Row blocks: shared memory -> TMEM as the A operand, eight 128 x 32-byte copies each with the
                        // K-slice descriptors the MMA would use.  The tensor pipe runs cp and mma in issue order and
                        // the other issuer handed the turn over only after its last MMA of the previous unit, so no
                        // MMA that still reads the old blocks is behind these.
                        for (int gg = 0; gg < (two ? 2 : 1); ++gg) {
                            const int sa = (posA + gg) % kSlots;
                            const uint64_t aoff = static_cast<uint64_t>(sa * (kTileBytes >> 4));
#pragma unroll
                            for (int k = 0; k < 8; ++k) tmem_cp_128x256b(tmem + gg * 64 + k * 8, dJ0[k] + aoff);
                            tc_commit(b_empty + 8 * sa);
                        }
                    }
                    __syncwarp();
                    if (elect_one()) {
                        if (run) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) umma_ts(tAcc + b0 * 128, tmem + k * 8, dJ0[k] + soff, idesc, k > 0);
                            if (kSweep == SWEEP_P) umma_ss(tAcc + b0 * 128, dShiftA0, dShiftB, idesc, 1u);
                        }
                        tc_commit(b_tfull + 8 * b0);
                    }
                    __syncwarp();
                    // the second buffer was released by the job three before it (same tile parity, other group):
                    // waiting for it only now lets that epilogue overlap the MMAs above
                    mbar_wait(b_tempty + 8 * b1, ((j1 / 3) & 1) ^ 1);
                    if (lane == 0) trace_stamp(p, 1 + s, it, 5);
                    tc_fence_after();
                    if (elect_one()) {
                        if (run && two) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) umma_ts(tAcc + b1 * 128, tmem + 64 + k * 8, dJ0[k] + soff, idesc, k > 0);
                        }
                        // six MMAs (~400 clk, the other issuer's wake-up) early - unless the next tile starts a new
                        // unit: its row-block copies must queue behind every MMA of this one
                        if (!last) mbar_arrive(b_turn + 8 * (s ^ 1));
                        if (run && two) {
#pragma unroll
                            for (int k = 2; k < 8; ++k) umma_ts(tAcc + b1 * 128, tmem + 64 + k * 8, dJ0[k] + soff, idesc, true);
                            if (kSweep == SWEEP_P) umma_ss(tAcc + b1 * 128, dShiftA1, dShiftB, idesc, 1u);
                        }
                        tc_commit(b_tfull + 8 * b1);
                        tc_commit(b_empty + 8 * slot);
                        if (last) mbar_arrive(b_turn + 8 * (s ^ 1));
                    }
                    __syncwarp();
                    if (lane == 0) trace_stamp(p, 1 + s, it, 3);
                }
                ++it;
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue
        const int g = warp >> 2;                           // group: row block 2U+g of the pair
        const int q = warp & 3;
        const int r = q * 32 + lane;                       // row inside the block == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        TileIter<kSweep == SWEEP_C> iter(p, part, sInfo);
        if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + (warp >> 2), 31, 3);                  // iterator built
        int U, J, curU = -1, it = 0;
        bool last, valid = false;
        float acc0[4], acc1[4], acc2[4], acc3[4];
        f32x2 q3[4], q4[4], q5[4];                         // SWEEP_P: packed sums of x, x^2 (q3, q4); SWEEP_H: x^3, x^4, x^5
        float mx4[4];
        float mQ0 = 0.f, mQ1 = 0.f, mQ2 = 0.f, mN0 = 0.f, mN1 = 0.f, mN2 = 0.f;   // masked-tile sums: positives / negatives
        f32x2 negc = 0ull;
        f32x2 mA0 = 0ull, mA1 = 0ull, mA2 = 0ull, mP0 = 0ull, mP1 = 0ull, mP2 = 0ull;   // SWEEP_P masked tiles (packed)
        int* wy = sLab + warp * 128;
        float cshift = 0.f, ra = 0.f, rb = 0.f, rden = 1.f;
        int yi = -1, gi = -1, Iloc = 0, lrow = 0, nunits = 0;
        int wlo = INT_MAX, whi = -1;                        // label range of this warp's 32 valid rows
        int2 rI = make_int2(INT_MAX, -1);

        auto reset = [&]() {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc0[u] = (kSweep == SWEEP_A) ? -FLT_MAX : 0.f;
                acc1[u] = acc2[u] = acc3[u] = 0.f;
                q3[u] = q4[u] = q5[u] = 0ull;
                mx4[u] = -FLT_MAX;
            }
            mQ0 = mQ1 = mQ2 = mN0 = mN1 = mN2 = 0.f;
            mA0 = mA1 = mA2 = mP0 = mP1 = mP2 = 0ull;
        };
        auto flush = [&]() {
            if (!valid) return;
            const int seg = blockIdx.x - part.first_cta(curU);
            if (kSweep == SWEEP_A) {
                float m4 = fmaxf(fmaxf(acc0[0], acc0[1]), fmaxf(acc0[2], acc0[3]));
                p.pA[(static_cast<size_t>(Iloc) * p.maxsegS + seg) * 128 + r] =
                    make_float4(m4, (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]),
                                (acc2[0] + acc2[1]) + (acc2[2] + acc2[3]), 0.f);
            } else if (kSweep == SWEEP_B) {
                p.pB[(static_cast<size_t>(Iloc) * p.maxsegS + seg) * 128 + r] =
                    make_float2((acc0[0] + acc0[1]) + (acc0[2] + acc0[3]),
                                (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]));
            } else if (kSweep == SWEEP_P) {
                // mN0 counts the 128 columns of every unmasked tile, mQ0 the valid columns of the masked ones (minus the
                // row's own column); masked tiles: same-class = all valid - different-class
                const float n0 = sum2(mP0), n1 = sum2(mP1), n2 = sum2(mP2);
                float4* o = p.pF + ((static_cast<size_t>(Iloc) * p.maxsegS + seg) * 2) * 128 + r;
                o[0] = make_float4(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), mN0 + n0,
                                   (sum2(q3[0]) + sum2(q3[1])) + (sum2(q3[2]) + sum2(q3[3])) + n1,
                                   (sum2(q4[0]) + sum2(q4[1])) + (sum2(q4[2]) + sum2(q4[3])) + n2);
                o[128] = make_float4((mQ0 + sum2(mA0)) - n0, sum2(mA1) - n1, sum2(mA2) - n2, 0.f);
            } else if (kSweep == SWEEP_H) {
                float4* o = p.pH + ((static_cast<size_t>(Iloc) * p.maxsegS + seg) * 2) * 128 + r;
                o[0] = make_float4((sum2(q3[0]) + sum2(q3[1])) + (sum2(q3[2]) + sum2(q3[3])) + mN0,
                                   (sum2(q4[0]) + sum2(q4[1])) + (sum2(q4[2]) + sum2(q4[3])) + mN1,
                                   (sum2(q5[0]) + sum2(q5[1])) + (sum2(q5[2]) + sum2(q5[3])) + mN2, 0.f);
                o[128] = make_float4(mQ0, mQ1, mQ2, 0.f);
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
            // The row block itself reaches TMEM through the ring (TMA) and tcgen05.cp, issued by the MMA warps; this
            // group only needs its rows' labels and shifts.  Every MMA that read the previous unit's shift operand
            // has completed: the group has consumed the accumulators of all its earlier jobs.
            if (valid) {
                gi = (p.rb0 + Iloc) * 128 + r;
                const int yload = __ldg(p.y + gi);
                const float cload = (kSweep == SWEEP_A || kSweep == SWEEP_P || kSweep == SWEEP_H) ? __ldg(p.sqnorm + gi) : 0.f;
                yi = yload;
                cshift = cload;
                if (kSweep == SWEEP_P) {
                    // -c_i as three bf16 (hi + mid + lo reproduces the fp32 value to 2^-24): row r of group g's A'
                    const __nv_bfloat16 hi = __float2bfloat16_rn(cload);
                    const float r1 = cload - __bfloat162float(hi);
                    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
                    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
                    const uint32_t nh = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) ^ 0x8000u;
                    const uint32_t nm = static_cast<uint32_t>(__bfloat16_as_ushort(mid)) ^ 0x8000u;
                    const uint32_t nl = static_cast<uint32_t>(__bfloat16_as_ushort(lo)) ^ 0x8000u;
                    uint8_t* arow = gen + SmemSweep::kShiftA + g * 4096 + (r >> 3) * 256 + (r & 7) * 16;
                    *reinterpret_cast<uint4*>(arow) = make_uint4(nh | (nm << 16), nl, 0u, 0u);
                    *reinterpret_cast<uint4*>(arow + 128) = make_uint4(0u, 0u, 0u, 0u);
                    fence_proxy_async_smem();
                }
            }
            tc_fence_before();
            if (nunits > 0) mbar_wait(b_aseen, (nunits - 1) & 1);   // both issuers are past the previous unit's wait
            ++nunits;
            __syncwarp();
            if (lane == 0) mbar_arrive(b_afull + 8 * g);
            if (!valid) return;
            lrow = Iloc * 128 + r;
            const int4 bi = sInfo[p.rb0 + Iloc];
            rI = make_int2(bi.x, bi.y);
            wlo = __reduce_min_sync(0xffffffffu, yi >= 0 ? yi : INT_MAX);
            whi = __reduce_max_sync(0xffffffffu, yi);
            if (kSweep == SWEEP_A) {
                // cshift loaded above
            } else if (kSweep == SWEEP_P) {
                negc = 0ull;                            // the shift x = s - c_i is part of the product (ninth K step)
            } else if (kSweep == SWEEP_H) {
                negc = pack2(-cshift, -cshift);         // the only per-row constant: the shift x = s - c_i
            } else {
                const float4 rs = p.rowS[lrow];
                ra = rs.x;
                rb = rs.y;
                if (kSweep == SWEEP_C) rden = p.rowD[lrow].x;
            }
            reset();
        };

        if (kSweep == SWEEP_C) begin_segment(blockIdx.x / p.splitc);   // its slot is always written
        while (iter.next(U, J, last)) {
            if (U != curU) {
                if (curU >= 0) flush();
                begin_segment(U);
            }
            const int job = 2 * it + g, buf = job % 3;
            if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 0);
            mbar_wait(b_tfull + 8 * buf, (job / 3) & 1);
            if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 1);
            tc_fence_after();
            const int col0 = J * 128;
            const int4 bj = sInfo[J];
            const int2 rJ = make_int2(bj.x, bj.y);
            const bool all_valid = bj.z == 128;
            const uint32_t taddr = tAcc + buf * 128 + lane_off;
            const int32_t* yJ = p.y + col0;

            if (!valid || (p.debug & 1)) {
                // odd tail: this group has no row block; just release the stage
            } else if (kSweep == SWEEP_P) {
                const bool fast = all_valid && !ranges_overlap(rI, rJ);
                if (fast) {
                    for_each_chunk<4>(taddr, [&](int, const uint32_t (&v)[32]) { psweep_chunk<true>(v, negc, q3, q4, mx4); });
                    mN0 += 128.f;
                } else {
                    // masked tile: stage the column labels once per warp, then packed masked sums
                    __syncwarp();
                    reinterpret_cast<int4*>(wy)[lane] = __ldg(reinterpret_cast<const int4*>(yJ) + lane);
                    __syncwarp();
                    if (all_valid) {
                        // rows are class-sorted, so most 32-row x 32-column pieces of a mixed tile are uniform: no
                        // shared class (plain sums, counted as different-class) or one class on both sides (plain
                        // sums into the all-columns accumulators); only pieces on a class boundary go element-wise
                        for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
                            const int yl = wy[c0 + lane];
                            const int cmin = __reduce_min_sync(0xffffffffu, yl), cmax = __reduce_max_sync(0xffffffffu, yl);
                            if (cmax < wlo || cmin > whi) {
                                psweep_chunk<true>(v, negc, q3, q4, mx4);
                                mN0 += 32.f;
                            } else if (cmin == cmax && wlo == whi && cmin == wlo) {
                                f32x2 t1[4] = {0ull, 0ull, 0ull, 0ull}, t2[4] = {0ull, 0ull, 0ull, 0ull};
                                psweep_chunk<true>(v, negc, t1, t2, mx4);
                                mA1 = fadd2(mA1, fadd2(fadd2(t1[0], t1[1]), fadd2(t1[2], t1[3])));
                                mA2 = fadd2(mA2, fadd2(fadd2(t2[0], t2[1]), fadd2(t2[2], t2[3])));
                                mQ0 += 32.f;
                            } else {
                                psweep_chunk_masked<true>(v, wy + c0, yi, negc, mA0, mA1, mA2, mP0, mP1, mP2, mx4);
                                mQ0 += 32.f;
                            }
                        });
                    } else {
                        for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
                            psweep_chunk_masked<false>(v, wy + c0, yi, negc, mA0, mA1, mA2, mP0, mP1, mP2, mx4);
                        });
                    }
                    if (J == p.rb0 + Iloc) mQ0 -= 1.f;          // the row's own column (x = 0) is not a positive
                }
            } else if (kSweep == SWEEP_H) {
                const bool fast = all_valid && !ranges_overlap(rI, rJ);
                if (fast) {
                    if (deg == 2)      for_each_chunk_loop(taddr, [&](int, const uint32_t (&v)[32]) { hsweep_chunk<2>(v, negc, q3, q4, q5); });
                    else if (deg == 3) for_each_chunk_loop(taddr, [&](int, const uint32_t (&v)[32]) { hsweep_chunk<3>(v, negc, q3, q4, q5); });
                    else               for_each_chunk_loop(taddr, [&](int, const uint32_t (&v)[32]) { hsweep_chunk<4>(v, negc, q3, q4, q5); });
                } else {
                    for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = __uint_as_float(v[j]) - cshift;
                            const int yj = __ldg(yJ + c0 + j);
                            const float xx = x * x, x3 = xx * x;
                            if (yj >= 0) {
                                if (yj != yi) {
                                    mN0 += x3;
                                    mN1 = fmaf(xx, xx, mN1);
                                    mN2 = fmaf(x3, xx, mN2);
                                } else if (col0 + c0 + j != gi) {
                                    mQ0 += x3;
                                    mQ1 = fmaf(xx, xx, mQ1);
                                    mQ2 = fmaf(x3, xx, mQ2);
                                }
                            }
                        }
                    });
                }
            } else if (kSweep == SWEEP_A) {
                if (all_valid) {
                    for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(v[j]) - cshift;
                            acc0[j & 3] = fmaxf(acc0[j & 3], x);
                            acc1[j & 3] += x;
                            acc2[j & 3] = fmaf(x, x, acc2[j & 3]);
                        }
                    });
                } else {
                    for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(v[j]) - cshift;
                            if (__ldg(yJ + c0 + j) >= 0) {
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
                    for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = fmaf(__uint_as_float(v[j]), ra, rb);
                            float e = ex2f(t);
                            acc0[j & 3] += e;
                            acc1[j & 3] = fmaf(e, t, acc1[j & 3]);
                        }
                    });
                } else {
                    for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = fmaf(__uint_as_float(v[j]), ra, rb);
                            float e = ex2f(t);
                            const int yj = __ldg(yJ + c0 + j);
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
                for_each_chunk_loop(taddr, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int yj = __ldg(yJ + c0 + j);
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
            if (lane == 0) mbar_arrive(b_tempty + 8 * buf);
            if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 2);
            ++it;
        }
        if (curU >= 0) flush();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) trace_stamp(p, 0, 1, 7);          // all roles done
    if (kSweep == SWEEP_P && threadIdx.x == 0) cta_time(p, 0, 1);
    if (warp == kProducerWarp) tmem_dealloc<kTmemCols>(tmem);
}

// =============================================================================================
// legacy combines: sweep A partials -> rowS, sweep B partials -> rowD
// =============================================================================================
__global__ void __launch_bounds__(128) k_combine_A(const Params p) {
    pdl_launch();
    pdl_wait();
    const int I = blockIdx.x, r = threadIdx.x, lrow = I * 128 + r;
    if (I == 0 && r == 0) { p.ticket[0] = 0u; p.ticket[1] = 0u; }
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
    const float kappa = 1.0f / rT, smax = c + mx, a = kappa * kLog2e;
    p.rowS[lrow] = make_float4(a, -smax * a, kappa, smax);
}
__global__ void __launch_bounds__(128) k_combine_B(const Params p) {
    pdl_launch();
    pdl_wait();
    const int I = blockIdx.x, r = threadIdx.x, lrow = I * 128 + r;
    const int ns = p.partS.nseg(I >> 1);
    float den = 0.f, bt = 0.f;
    for (int s = 0; s < ns; ++s) {
        float2 v = p.pB[(static_cast<size_t>(I) * p.maxsegS + s) * 128 + r];
        den += v.x;
        bt += v.y;
    }
    p.rowD[lrow] = make_float4(den, bt, 1.0f, 0.f);
}

// =============================================================================================
// v3: per row, shifted power sums + exact maximum -> kappa, logit range, polynomial degree, polynomial, Den, Bt and
//     the positive-pair sums (k_rows; k_combine2 for the rows that need the higher power sums)
// =============================================================================================
constexpr float kMinNegSeries = 174.0f;     // E >= 1/e, so this many negatives guarantee Den >= kMinDenSeries

// Per-row polynomial sums: Den, Bt and the positive-pair sums of the series from the power sums P_j (different
// class) and Q_j (same class) of x = s - c, j = 0..deg+1
__device__ __forceinline__ void row_poly_sums(const double (&P)[6], const double (&Q)[6], double kappa, double a, double d,
                                              double L, float4& rowD, float4& rowPos) {
    // E as a polynomial in x = s - c:  l = kappa (x - d)
    double e[5] = {1.0, 0.0, 0.0, 0.0, 0.0};
    if (kappa > 0.0) exp_poly_in_s(kappa, d, L, poly_degree_for(L), e);
    double den = 0.0, sex = 0.0, pe = 0.0, pex = 0.0;
    for (int j = 0; j < 5; ++j) {
        den += e[j] * P[j]; sex += e[j] * P[j + 1];
        pe += e[j] * Q[j];  pex += e[j] * Q[j + 1];
    }
    // Bt = sum_neg E t,  t = log2(e) kappa (x - d)
    rowD = make_float4(static_cast<float>(den), static_cast<float>(a * (sex - d * den)), static_cast<float>(L), 0.f);
    // positive pairs, first-order series in E/Den (second-order term with sum E^2 ~ (sum E)^2 / P):
    //   lp = l - log(E + Den) ~ l - log Den - E/Den + E^2/(2 Den^2),   1/(E + Den) ~ 1/Den - E/Den^2 + E^2/Den^3
    const double Pn = Q[0], sl = kappa * (Q[1] - d * Q[0]), sel = kappa * (pex - d * pe);
    const double pe2 = Pn > 0.0 ? pe * pe / Pn : 0.0;
    const double id = 1.0 / den;
    const double SL = sl - Pn * log(den) - pe * id + 0.5 * pe2 * id * id;
    const double SI = Pn * id - pe * id * id + pe2 * id * id * id;
    const double SIL = sl * id - sel * id * id;
    rowPos = make_float4(static_cast<float>(Pn), static_cast<float>(SL), static_cast<float>(SI), static_cast<float>(SIL));
}

// Row loss and the packed backward constants from the row's scale (rs), denominator sums (db) and positive-pair
// sums (P, SL, SI, SIL)
template <int kMode>
__device__ __forceinline__ float finalize_row(const Params& p, const float4 rs, const float4 db, float P, float SL, float SI,
                                              float SIL, int yi, float4& cA, float4& cB) {
    const float kappa = rs.z;
    const float ratio = p.T / p.Tb;
    const float c = ratio / static_cast<float>(p.n_valid);
    const float w = -c / P;                       // P == 0 -> NaN, as in the reference (loss.py:383)
    const float Bl = db.y * kLn2;                 // sum_den E*l
    float rl, Q, R, wn;
    if (kMode == DCL_MODE_PIXEL) {
        rl = -ratio * SL / P;
        Q = w * SI;
        R = w * db.x * SIL - Q * Bl;
        wn = kappa * w * db.x;
    } else {
        rl = -ratio * (SL - P * logf(db.x)) / P;
        Q = -c / db.x;
        R = w * SL - Q * Bl;
        wn = kappa * w;
    }
    cA = make_float4(rs.x, rs.y, -kappa * R * kLn2, -kappa * Q);
    cB = make_float4(wn, db.x, __int_as_float(yi), db.z);
    return rl;
}

// Deterministic sum of the per-row losses: block sum in a fixed tree, then the last block to arrive adds the block
// sums in index order
__device__ __forceinline__ float block_sum128(float v, float* red4) {
    // fixed tree: xor-shuffles inside each warp, then the four warp sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red4[threadIdx.x >> 5] = v;
    __syncthreads();
    return (red4[0] + red4[1]) + (red4[2] + red4[3]);
}
__device__ __forceinline__ void block_loss_sum(const Params& p, float rl, unsigned int* ticket) {
    __shared__ float red[4], red2[4];
    __shared__ bool is_last;
    const int I = blockIdx.x, r = threadIdx.x;
    const float bs = block_sum128(rl, red);
    if (r == 0) {
        p.blockloss[I] = bs;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float s = 0.f;
        for (int i = r; i < p.nI; i += 128) s += __ldcg(p.blockloss + i);     // fixed order per thread, fixed tree after
        s = block_sum128(s, red2);
        if (r == 0) {
            p.loss_sum[0] = s;
            p.loss_sum[1] = s / static_cast<float>(p.n_valid);      // the loss itself when the rows are not sharded
            if (p.loss_out) p.loss_out[0] = s / static_cast<float>(p.n_valid);
        }
    }
}

// k_rows (after sweep P, one thread per local row): segment partials -> power sums and the exact maximum ->
// kappa, the logit range L and the polynomial degree.  A row of degree 1 with enough negatives for the series --
// every row of a problem with a few thousand anchors or more -- is finished here: polynomial sums, loss and
// backward constants.  Any other row raises the triggers of the conditional kernels (sweep H for degree > 1,
// k_combine2, sweep C for the exact positive-pair sums, k_finalize), which otherwise leave at once.
__global__ void __launch_bounds__(128) k_rows(const Params p) {
    pdl_launch();
    if (threadIdx.x == 0) cta_time(p, 2, 0);
    pdl_wait();
    if (threadIdx.x == 0) trace_stamp(p, 0, 0, 6);
    const int I = blockIdx.x, r = threadIdx.x, lrow = I * 128 + r, gi = (p.rb0 + I) * 128 + r;
    // all loads of the row first (segment partials in batches of four), the block-wide reduction behind them
    const int ns = p.partS.nseg(I >> 1);
    const float4* pf = p.pF + static_cast<size_t>(I) * p.maxsegS * 2 * 128 + r;
    float mx = -FLT_MAX;
    double P[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, Q[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    float4 a4[4], b4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (u < ns) { a4[u] = __ldcg(pf + u * 256); b4[u] = __ldcg(pf + u * 256 + 128); }
    const int yi = p.y[gi];
    const float cself = p.sqnorm[gi];
    // largest |f|^2 of the contrast set (per-block maxima from k_blockinfo)
    __shared__ float scm[4];
    float cm = 0.f;
    for (int j = r; j < p.nJ; j += 128) cm = fmaxf(cm, p.bnorm[j].x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if ((r & 31) == 0) scm[r >> 5] = cm;
    __syncthreads();
    const double cmax = fmaxf(fmaxf(scm[0], scm[1]), fmaxf(scm[2], scm[3]));
    if (threadIdx.x == 0) trace_stamp(p, 0, 1, 6);
    for (int s0 = 0; s0 < ns; s0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (s0 + u < ns) {
                const float4 a = a4[u], b = b4[u];
                mx = fmaxf(mx, a.x);
                P[0] += a.y; P[1] += a.z; P[2] += a.w;
                Q[0] += b.x; Q[1] += b.y; Q[2] += b.z;
            }
        if (s0 + 4 < ns) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (s0 + 4 + u < ns) { a4[u] = __ldcg(pf + (s0 + 4 + u) * 256); b4[u] = __ldcg(pf + (s0 + 4 + u) * 256 + 128); }
        }
    }
    if (threadIdx.x == 0) trace_stamp(p, 0, 2, 6);
    float4* rm = reinterpret_cast<float4*>(p.rowM + static_cast<size_t>(lrow) * 8);
    float rl = 0.f;
    float4 cA = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cB = make_float4(0.f, 1.f, __int_as_float(-1), 0.f);
    int pending = 0;
    if (yi < 0) {
        rm[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        rm[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        p.rowS[lrow] = make_float4(0.f, 0.f, 0.f, 0.f);
        p.rowD[lrow] = make_float4(1.f, 0.f, 0.f, 0.f);
        p.rowPos[lrow] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        // sum over every valid column of (s - m)^2 = sum (x - d)^2, d = max x (the row itself has x = 0)
        const double d = mx, n = p.n_valid, c = cself, m = c + d;
        const double nrm2 = (P[2] + Q[2]) - 2.0 * d * (P[1] + Q[1]) + n * d * d;
        double kappa, L;
        row_scale(nrm2, m, c, cmax, p.T, kappa, L);
        if (threadIdx.x == 0) trace_stamp(p, 0, 3, 6);
        const int deg = kappa > 0.0 ? poly_degree_for(L) : 1;
        const double a = kappa * static_cast<double>(kLog2e);
        const float4 rs = make_float4(static_cast<float>(a), static_cast<float>(-m * a), static_cast<float>(kappa), static_cast<float>(m));
        const float4 m0 = make_float4(static_cast<float>(P[0]), static_cast<float>(P[1]), static_cast<float>(P[2]), static_cast<float>(Q[0]));
        const float4 m1 = make_float4(static_cast<float>(Q[1]), static_cast<float>(Q[2]), mx, static_cast<float>(L));
        rm[0] = m0;
        rm[1] = m1;
        p.rowS[lrow] = rs;
        const bool few = P[0] < kMinNegSeries;
        pending = (!p.use_series || deg > 1 || few) ? 1 : 0;
        if (pending) {
            p.iscal[5] = 1;                          // same value from every such row
            if (deg > 1) atomicMax(&p.iscal[0], deg);
            if (few) p.iscal[4] = 1;
        } else {
            // the float-rounded values k_combine2 would read back, so both routes give identical bits
            const double Pf[6] = {m0.x, m0.y, m0.z, 0.0, 0.0, 0.0}, Qf[6] = {m0.w, m1.x, m1.y, 0.0, 0.0, 0.0};
            float4 db, pos;
            row_poly_sums(Pf, Qf, rs.z, rs.x, m1.z, m1.w, db, pos);
            if (threadIdx.x == 0) trace_stamp(p, 0, 4, 6);
            p.rowD[lrow] = db;
            rl = finalize_row<DCL_MODE_PIXEL>(p, rs, db, pos.x, pos.y, pos.z, pos.w, yi, cA, cB);
        }
    }
    p.rowflag[lrow] = pending;
    if (!pending) {
        p.colA[gi] = cA;
        p.colB[gi] = cB;
        p.rowloss[gi] = rl;
    }
    if (threadIdx.x == 0) trace_stamp(p, 0, 5, 6);
    block_loss_sum(p, rl, p.ticket);               // provisional when rows are pending: k_finalize then redoes it
    if (threadIdx.x == 0) trace_stamp(p, 0, 6, 6);
    if (threadIdx.x == 0) cta_time(p, 2, 1);
}

// k_combine2 (conditional): the rows k_rows left pending, with the higher power sums of sweep H when it ran
__global__ void __launch_bounds__(128) k_combine2(const Params p) {
    pdl_launch();
    pdl_wait();
    if (p.iscal[5] == 0) return;
    const int I = blockIdx.x, r = threadIdx.x, lrow = I * 128 + r;
    if (p.rowflag[lrow] == 0) return;
    const float4* rm = reinterpret_cast<const float4*>(p.rowM + static_cast<size_t>(lrow) * 8);
    const float4 m0 = rm[0], m1 = rm[1];
    double P[6] = {m0.x, m0.y, m0.z, 0.0, 0.0, 0.0}, Q[6] = {m0.w, m1.x, m1.y, 0.0, 0.0, 0.0};
    if (p.iscal[0] > 1) {
        const int ns = p.partS.nseg(I >> 1);
        for (int s = 0; s < ns; ++s) {
            const float4* o = p.pH + ((static_cast<size_t>(I) * p.maxsegS + s) * 2) * 128 + r;
            const float4 a = o[0], b = o[128];
            P[3] += a.x; P[4] += a.y; P[5] += a.z;
            Q[3] += b.x; Q[4] += b.y; Q[5] += b.z;
        }
    }
    const float4 rs = p.rowS[lrow];                 // (a, b, kappa, m)
    float4 db, pos;
    row_poly_sums(P, Q, rs.z, rs.x, m1.z, m1.w, db, pos);
    p.rowD[lrow] = db;
    p.rowPos[lrow] = pos;
}

// =============================================================================================
// finalize: per-row loss and backward constants (one thread per local row); the last block to finish adds the
// per-block losses in fixed order.  In the v3 pipeline (p.use_series) it only runs when k_rows left rows pending
// and only redoes those; the legacy pipelines finish every row here.
// =============================================================================================
template <int kMode>
__global__ void __launch_bounds__(128) k_finalize(const Params p) {
    pdl_launch();
    pdl_wait();
    if (kMode == DCL_MODE_PIXEL && p.use_series && p.iscal[5] == 0) return;
    const int I = blockIdx.x, r = threadIdx.x, lrow = I * 128 + r;
    const int gi = (p.rb0 + I) * 128 + r;
    const int yi = p.y[gi];
    float rl = 0.f;
    if (kMode == DCL_MODE_PIXEL && p.use_series && p.rowflag[lrow] == 0) {
        rl = p.rowloss[gi];                           // finished by k_rows
    } else {
        float4 cA = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 cB = make_float4(0.f, 1.f, __int_as_float(-1), 0.f);
        if (yi >= 0) {
            float P = 0.f, SL = 0.f, SI = 0.f, SIL = 0.f;
            if (kMode == DCL_MODE_PIXEL && p.use_series && p.iscal[4] == 0) {
                const float4 v = p.rowPos[lrow];          // series sums from k_combine2 (sweep C did not run)
                P = v.x; SL = v.y; SI = v.z; SIL = v.w;
            } else {
                for (int s = 0; s < p.splitc; ++s) {
                    float4 v = p.pC[(static_cast<size_t>(I) * p.splitc + s) * 128 + r];
                    P += v.x; SL += v.y; SI += v.z; SIL += v.w;
                }
            }
            rl = finalize_row<kMode>(p, p.rowS[lrow], p.rowD[lrow], P, SL, SI, SIL, yi, cA, cB);
        }
        p.colA[gi] = cA;
        p.colB[gi] = cB;
        p.rowloss[gi] = rl;
    }
    block_loss_sum(p, rl, p.ticket + 1);
}

// =============================================================================================
// Backward
// =============================================================================================
// Per row (all nJ*128 of them) two polynomials in s = s_ik, in row order (coefR: n0..n4 - - - p0..p4 - - -) and
// pair-interleaved for the column side (coefP / coefPP: [c0k c0k' c1k c1k'] [c2k c2k' c3k c3k'] [c4k c4k' * *]):
//   pairs of different classes:  rpN(s) = q E(s) + p (a s + b)
//   pairs of the same class   :  rpP(s) = p (a s + b) + wn / (E + Den) ~ p (a s + b) + wn/Den - (wn/Den^2) E(s)
// (first-order series in E/Den <= 1/Den; tiles use it only where every Den >= kMinDenSeries, see bden).  The two
// spare floats of a coefP pair block hold the labels of its two columns.
constexpr float kMinDenSeries = 64.0f;      // series error (E/Den)^2 <= 2.5e-4 of the positive-pair term
__global__ void __launch_bounds__(128) k_bwd_prep(const Params p) {
    pdl_launch();
    pdl_wait();
    const int row = blockIdx.x * 128 + threadIdx.x;
    const float4 cA = p.colA[row];
    const float4 cB = p.colB[row];
    const int yv = __float_as_int(cB.z);
    double dn[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, dp[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int deg = 1;
    float den = FLT_MAX;
    if (yv >= 0) {
        const double a = cA.x, b = cA.y;
        double e[5] = {1.0, 0.0, 0.0, 0.0, 0.0};
        if (a > 0.0) {
            const double kappa = a / static_cast<double>(kLog2e), m = -b / a, L = cB.w;
            deg = poly_degree_for(L);
            exp_poly_in_s(kappa, m, L, deg, e);
        }
        den = cB.y;
        const double wn = cB.x, D = cB.y;
        const double qP = D >= 1.0 ? -wn / (D * D) : 0.0, cP = D >= 1.0 ? wn / D : 0.0;
        for (int j = 0; j < 5; ++j) { dn[j] = static_cast<double>(cA.w) * e[j]; dp[j] = qP * e[j]; }
        dn[1] += static_cast<double>(cA.z) * a;
        dn[0] += static_cast<double>(cA.z) * b;
        dp[1] += static_cast<double>(cA.z) * a;
        dp[0] += static_cast<double>(cA.z) * b + cP;
    }
    float4* o = reinterpret_cast<float4*>(p.coefR + static_cast<size_t>(row) * 16);
    o[0] = make_float4(static_cast<float>(dn[0]), static_cast<float>(dn[1]), static_cast<float>(dn[2]), static_cast<float>(dn[3]));
    o[1] = make_float4(static_cast<float>(dn[4]), 0.f, 0.f, 0.f);
    o[2] = make_float4(static_cast<float>(dp[0]), static_cast<float>(dp[1]), static_cast<float>(dp[2]), static_cast<float>(dp[3]));
    o[3] = make_float4(static_cast<float>(dp[4]), 0.f, 0.f, 0.f);
    float* cn = p.coefP + static_cast<size_t>(row >> 1) * kCoefPairFloats + (row & 1);
    float* cq = p.coefPP + static_cast<size_t>(row >> 1) * kCoefPairFloats + (row & 1);
#pragma unroll
    for (int j = 0; j < 5; ++j) { cn[2 * j] = static_cast<float>(dn[j]); cq[2 * j] = static_cast<float>(dp[j]); }
    cn[10] = __int_as_float(yv);
    cq[10] = 0.f;
    // block info of the backward (this kernel is the first of its chain): label range, valid rows, largest degree
    __shared__ int si[12];
    __shared__ float sf[8];
    const BlockStat bs = block_stat(p.y[row], 0.f, true, si, sf);
    const int d4 = __syncthreads_or(deg == 4), d3 = __syncthreads_or(deg == 3), d2 = __syncthreads_or(deg == 2);
    if (threadIdx.x == 0) {
        p.binfo[blockIdx.x] = make_int4(bs.lo, bs.hi, bs.n, d4 ? 4 : (d3 ? 3 : (d2 ? 2 : 1)));
        if (blockIdx.x < p.nI) p.dticket[blockIdx.x] = 0u;
    }
    // smallest Den of the block's valid rows
    for (int o2 = 16; o2 > 0; o2 >>= 1) den = fminf(den, __shfl_xor_sync(0xffffffffu, den, o2));
    __shared__ float sden[4];
    if ((threadIdx.x & 31) == 0) sden[threadIdx.x >> 5] = den;
    __syncthreads();
    if (threadIdx.x == 0) p.bden[blockIdx.x] = fminf(fminf(sden[0], sden[1]), fminf(sden[2], sden[3]));
}

// one column pair of a Horner chain on the column coefficients: h = ((c4 s + c3) s + c2) s + c1 (degree-dependent)
template <int kDeg>
__device__ __forceinline__ f32x2 col_horner(const float4* __restrict__ cp, const float4& q0, f32x2 s) {
    if (kDeg == 1) return pack2(q0.z, q0.w);
    const float4 q1 = cp[1];
    if (kDeg == 2) return ffma2(pack2(q1.x, q1.y), s, pack2(q0.z, q0.w));
    if (kDeg == 3) {
        f32x2 h = ffma2(pack2(q1.z, q1.w), s, pack2(q1.x, q1.y));
        return ffma2(h, s, pack2(q0.z, q0.w));
    }
    const float4 q2 = cp[2];
    f32x2 h = ffma2(pack2(q2.x, q2.y), s, pack2(q1.z, q1.w));
    h = ffma2(h, s, pack2(q1.x, q1.y));
    return ffma2(h, s, pack2(q0.z, q0.w));
}
template <int kDeg>
__device__ __forceinline__ f32x2 row_horner(const f32x2 (&rr)[5], f32x2 s) {
    if (kDeg == 1) return ffma2(rr[1], s, rr[0]);
    f32x2 g = ffma2(rr[kDeg], s, rr[kDeg - 1]);
#pragma unroll
    for (int d = kDeg - 2; d >= 0; --d) g = ffma2(g, s, rr[d]);
    return g;
}
__device__ __forceinline__ uint32_t pack_bf16x2(f32x2 v) {
    float lo, hi;
    unpack2(v, lo, hi);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&b2);
}

// G for 32 columns of a tile without same-class pairs: rpN_i(s) + rpN_k(s), packed bf16 into pk[16]
template <int kDeg>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&v)[32], const f32x2 (&rr)[5], const float4* __restrict__ cp,
                                          uint32_t (&pk)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2 s = pack2u(v[2 * j], v[2 * j + 1]);
        const float4 q0 = cp[3 * j];
        f32x2 g = ffma2(col_horner<kDeg>(cp + 3 * j, q0, s), s, row_horner<kDeg>(rr, s));
        g = fadd2(g, pack2(q0.x, q0.y));
        pk[j] = pack_bf16x2(g);
    }
}
// The same for a tile that mixes classes: both polynomials, selected per element by label equality (`cq`: the
// same-class column polynomials, staged in shared memory like `cp`).
// `self` is the tile-local column of the row's own pair (0..31 inside this chunk, anything else = none): that
// element is neither a denominator nor a positive pair, only the linear term survives: 2 p (a s + b) = fma(s, l1, l0).
template <int kDeg>
__device__ __forceinline__ void bwd_chunk_masked(const uint32_t (&v)[32], const f32x2 (&rn)[5], const f32x2 (&rq)[5],
                                                 const float4* __restrict__ cp, const float4* __restrict__ cq, int yi,
                                                 int self, float l1, float l0, uint32_t (&pk)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2 s = pack2u(v[2 * j], v[2 * j + 1]);
        const float4 n0 = cp[3 * j], n2 = cp[3 * j + 2];
        f32x2 gn = ffma2(col_horner<kDeg>(cp + 3 * j, n0, s), s, row_horner<kDeg>(rn, s));
        gn = fadd2(gn, pack2(n0.x, n0.y));
        const float4 u0 = cq[3 * j];
        f32x2 gp = ffma2(col_horner<kDeg>(cq + 3 * j, u0, s), s, row_horner<kDeg>(rq, s));
        gp = fadd2(gp, pack2(u0.x, u0.y));
        float nl, nh, pl, ph;
        unpack2(gn, nl, nh);
        unpack2(gp, pl, ph);
        float lo = (__float_as_int(n2.z) == yi) ? pl : nl;
        float hi = (__float_as_int(n2.w) == yi) ? ph : nh;
        if (self == 2 * j) lo = fmaf(__uint_as_float(v[2 * j]), l1, l0);
        if (self == 2 * j + 1) hi = fmaf(__uint_as_float(v[2 * j + 1]), l1, l0);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(lo, hi);
        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
    }
}

// dF_I = sum of a row block's segment partials in segment order (the last CTA of the block to arrive; eight epilogue
// warps).  kNs > 0: all segments of eight rows per warp are in flight at once; kNs == 0: any count, one segment at a time.
template <int kNs>
__device__ __forceinline__ void reduce_segments(const float* __restrict__ part, float* __restrict__ dst, int ns, int warp,
                                                int lane) {
#pragma unroll 1
    for (int r0 = warp; r0 < 128; r0 += 64) {
        float4 acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kNs > 0) {
            float4 v[kNs > 0 ? kNs : 1][8];
#pragma unroll
            for (int sg = 0; sg < kNs; ++sg)
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[sg][u] = __ldcg(reinterpret_cast<const float4*>(part + (static_cast<size_t>(sg) * 128 + r0 + 8 * u) * 128 + lane * 4));
#pragma unroll
            for (int sg = 0; sg < kNs; ++sg)
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    acc[u].x += v[sg][u].x; acc[u].y += v[sg][u].y; acc[u].z += v[sg][u].z; acc[u].w += v[sg][u].w;
                }
        } else {
#pragma unroll 1
            for (int sg = 0; sg < ns; ++sg) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[u] = __ldcg(reinterpret_cast<const float4*>(part + (static_cast<size_t>(sg) * 128 + r0 + 8 * u) * 128 + lane * 4));
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            *reinterpret_cast<float4*>(dst + (static_cast<size_t>(r0) + 8 * u) * 128 + lane * 4) = acc[u];
    }
}

template <int kMode>
__global__ void __launch_bounds__(kThreads, 1) k_backward(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    pdl_launch();
    if (threadIdx.x == 0) trace_stamp(p, 0, 0, 7);          // kernel entry
    if (threadIdx.x == 0) cta_time(p, 1, 0);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sI = base + SmemBwd::kI;
    const uint32_t sJ = base + SmemBwd::kJ;
    const uint32_t sCP = base + SmemBwd::kCP;
    const float4* gCP = reinterpret_cast<const float4*>(gen + SmemBwd::kCP);
    const uint32_t sCQ = base + SmemBwd::kCQ;
    const float4* gCQ = reinterpret_cast<const float4*>(gen + SmemBwd::kCQ);
    // label range of every block, packed (labels are 0..255; an empty block is (255, 0))
    uint16_t* sRangeRaw = reinterpret_cast<uint16_t*>(gen + SmemBwd::kRange);
    auto block_range = [&](int j) { const int v = sRangeRaw[j]; return make_int2(v & 255, v >> 8); };
    uint8_t* sDenOk = gen + SmemBwd::kDenOk;
    const uint32_t bar = base + SmemBwd::kBar;
    constexpr int kSlots = SmemBwd::kSlots, kStages = SmemBwd::kStages;
    const uint32_t b_full = bar, b_empty = bar + 40, b_tfull = bar + 80, b_pfull = bar + 104, b_sfree = bar + 128,
                   b_dfull = bar + 152, b_dempty = bar + 160, b_ifull = bar + 168, b_iempty = bar + 176, b_turn = bar + 184;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemBwd::kTmem);
    volatile unsigned int* sTicket = reinterpret_cast<volatile unsigned int*>(gen + SmemBwd::kTmem + 8);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

    if (warp == kProducerWarp) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
    if (threadIdx.x == kIssuerWarp0 * 32) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 1);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(b_tfull + 8 * s, 1);
            mbar_init(b_pfull + 8 * s, 4);      // the 4 warps of one epilogue group
            mbar_init(b_sfree + 8 * s, 1);
        }
        mbar_init(b_dfull, 1);
        mbar_init(b_dempty, 8);
        mbar_init(b_ifull, 1);
        mbar_init(b_iempty, 1);
        mbar_init(b_turn, 1);
        mbar_init(b_turn + 8, 1);
        mbar_fence_init();
    }
    pdl_wait();
    int bdeg = 1;                                // largest polynomial degree of any row (per block from k_bwd_prep)
    for (int j = threadIdx.x; j < p.nJ; j += kThreads) {
        const int4 q = p.binfo[j];
        // labels outside 0..255 cannot be packed: such a block overlaps everything (the element-wise path is always right)
        if (kMode == DCL_MODE_PIXEL) sDenOk[j] = p.bden[j] >= kMinDenSeries;
        sRangeRaw[j] = static_cast<uint16_t>(q.z > 0 ? ((q.x >= 0 && q.y <= 255) ? (q.x | (q.y << 8)) : 0xff00) : 255);
        bdeg = max(bdeg, q.w);
    }
    tc_fence_before();
    const int d4 = __syncthreads_or(bdeg == 4), d3 = __syncthreads_or(bdeg == 3), d2 = __syncthreads_or(bdeg == 2);
    tc_fence_after();
    const int deg = (kMode == DCL_MODE_PIXEL) ? (d4 ? 4 : (d3 ? 3 : (d2 ? 2 : 1))) : 0;
    if (threadIdx.x == 0) trace_stamp(p, 0, 2, 7);          // set-up done
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tD = tmem + 384;          // dF accumulator; S/G stage st lives at tmem + st*128

    if (warp == kProducerWarp) {
        // whole warp runs the (uniform) control flow; the bulk copies are issued by one elected lane (see elect_one)
        TileIter<false> iter(p, p.partD, nullptr);
        int I, J, curI = -1, it = 0, seg = 0;
        bool last;
        while (iter.next(I, J, last)) {
            if (I != curI) {
                mbar_wait(b_iempty, (seg & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(b_ifull, kTileBytes);
                    tma_bulk_g2s(sI, p.tiles + static_cast<size_t>(p.rb0 + I) * kTileBytes, kTileBytes, b_ifull);
                }
                __syncwarp();
                curI = I;
                ++seg;
            }
            const int slot = it % kSlots;
            if (lane == 0) trace_stamp(p, 0, it, 0);
            mbar_wait(b_empty + 8 * slot, ((it / kSlots) & 1) ^ 1);
            if (lane == 0) trace_stamp(p, 0, it, 1);
            const bool mixed = kMode == DCL_MODE_PIXEL && ranges_overlap(block_range(p.rb0 + I), block_range(J));
            if (elect_one()) {
                if (kMode == DCL_MODE_PIXEL) {
                    mbar_arrive_expect_tx(b_full + 8 * slot, kTileBytes + (mixed ? 2 : 1) * SmemBwd::kCoefBytes);
                    tma_bulk_g2s(sCP + slot * SmemBwd::kCoefBytes,
                                 p.coefP + static_cast<size_t>(J) * 64 * kCoefPairFloats, SmemBwd::kCoefBytes,
                                 b_full + 8 * slot);
                    if (mixed)
                        tma_bulk_g2s(sCQ + slot * SmemBwd::kCoefBytes,
                                     p.coefPP + static_cast<size_t>(J) * 64 * kCoefPairFloats, SmemBwd::kCoefBytes,
                                     b_full + 8 * slot);
                } else {
                    mbar_arrive_expect_tx(b_full + 8 * slot, kTileBytes);
                }
                tma_bulk_g2s(sJ + slot * kTileBytes, p.tiles + static_cast<size_t>(J) * kTileBytes,
                             kTileBytes, b_full + 8 * slot);
            }
            __syncwarp();
            ++it;
        }
    } else if (warp >= kIssuerWarp0) {
        {
            // ------------------------------------------------------------------ MMA issuers (whole warp, see elect_one)
            // Burst k = [dF(k-3) = G(k-3) F_J(k-3), then S(k) = F_I F_J(k)^T], issued by issuer k % 2 while the other
            // one waits on the barriers of burst k+1.  S(k) reuses the TMEM stage of tile k-3: the tensor pipe
            // executes in issue order, so G(k-3) has been consumed when S(k) overwrites it - no completion round
            // trip (commit -> mbarrier -> wake-up, ~200 clk per tile) sits between them.  S runs three tiles ahead
            // of dF, which leaves each epilogue two bursts (~2000 clk) for its tile.
            const int s = warp - kIssuerWarp0;
            TileIter<false> iter(p, p.partD, nullptr);
            const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
            const uint32_t idesc_d = umma_idesc_bf16(128, 128, 0, 1);   // B = F_J, MN-major
            uint64_t dI[8], dJk[8], dJm[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                dI[k] = ftile_desc_kmajor(sI, k);
                dJk[k] = ftile_desc_kmajor(sJ, k);
                dJm[k] = ftile_desc_mnmajor(sJ, k);
            }
            // flags of the last four tiles (bit 0 first of its segment, bit 1 last) and their segment index
            int tflags[4] = {0, 0, 0, 0}, tseg[4] = {0, 0, 0, 0};
            int I, J, curI = -1, ntiles = -1, units = 0, segidx = 0, turn = 0;
            bool last = false;
            for (int k = 0;; ++k) {
                bool have = false, first = false;
                if (ntiles < 0) {
                    have = iter.next(I, J, last);
                    if (!have) ntiles = k;
                }
                if (have) {
                    first = (I != curI);
                    if (first) {
                        // both issuers observe every row block landing (either may issue its S tiles)
                        mbar_wait(b_ifull, units & 1);
                        curI = I;
                        ++units;
                    }
                    tflags[k & 3] = (first ? 1 : 0) | (last ? 2 : 0);
                    tseg[k & 3] = segidx;
                    if (last) ++segidx;
                }
                const int kd = k - 3;
                const bool have_d = kd >= 0 && (ntiles < 0 || kd < ntiles);
                if (ntiles >= 0 && k >= ntiles + 3) break;     // bursts 0 .. ntiles+2 (empty ones still pass the turn on)
                if ((k & 1) != s) continue;
                const int slot = k % kSlots, st = k % kStages;
                if (lane == 0) trace_stamp(p, 1 + s, k, 0);
                if (have) mbar_wait(b_full + 8 * slot, (k / kSlots) & 1);
                if (lane == 0) trace_stamp(p, 1 + s, k, 1);
                int dflags = 0;
                if (have_d) {
                    dflags = tflags[kd & 3];
                    mbar_wait(b_pfull + 8 * (kd % kStages), (kd / kStages) & 1);            // G(k-3) is in TMEM
                    if ((dflags & 1) && tseg[kd & 3] > 0) mbar_wait(b_dempty, (tseg[kd & 3] - 1) & 1);   // accumulator drained
                }
                if (lane == 0) trace_stamp(p, 1 + s, k, 2);
                mbar_wait(b_turn + 8 * s, (turn & 1) ^ (s == 0 ? 1 : 0));
                ++turn;
                if (lane == 0) trace_stamp(p, 1 + s, k, 4);
                tc_fence_after();
                if (elect_one()) {
                    if (have_d) {
                        const uint64_t poff = static_cast<uint64_t>((kd % kSlots) * (kTileBytes >> 4));
                        if (!(p.debug & 4)) {
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_ts(tD, tmem + (kd % kStages) * 128 + kk * 8, dJm[kk] + poff, idesc_d,
                                        !(dflags & 1) || kk > 0);
                        }
                        tc_commit(b_empty + 8 * (kd % kSlots));
                        if (dflags & 2) tc_commit(b_dfull);
                    }
                    if (have) {
                        const uint64_t soff = static_cast<uint64_t>(slot * (kTileBytes >> 4));
                        const bool run = !(p.debug & 2);
                        if (run) {
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) umma_ss(tmem + st * 128, dI[kk], dJk[kk] + soff, idesc_s, kk > 0);
                        }
                        // hand the turn over six MMAs (~400 clk: the other issuer's wake-up) early; its dF MMAs
                        // touch another stage and the accumulator, so interleaving with the rest of S(k) is safe
                        mbar_arrive(b_turn + 8 * (s ^ 1));
                        if (run) {
#pragma unroll
                            for (int kk = 2; kk < 8; ++kk) umma_ss(tmem + st * 128, dI[kk], dJk[kk] + soff, idesc_s, true);
                        }
                        tc_commit(b_tfull + 8 * st);
                        if (last) tc_commit(b_iempty);
                    } else {
                        mbar_arrive(b_turn + 8 * (s ^ 1));
                    }
                }
                __syncwarp();
                if (lane == 0) trace_stamp(p, 1 + s, k, 3);
            }
        }
    } else {
        const int g = warp >> 2;                     // group g handles tiles with (it & 1) == g
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        TileIter<false> iter(p, p.partD, nullptr);
        int I, J, curI = -1, it = 0, seg = 0;
        bool last;
        float4 rA = make_float4(0.f, 0.f, 0.f, 0.f), rB = make_float4(0.f, 1.f, 0.f, 0.f);
        f32x2 rn[5] = {0ull, 0ull, 0ull, 0ull, 0ull}, rq[5] = {0ull, 0ull, 0ull, 0ull, 0ull};
        int yi = -1, gi = -1;
        int wlo = INT_MAX, whi = -1;                 // label range of this warp's 32 valid rows
        int2 rI = make_int2(INT_MAX, -1);
        bool denI = false;                           // every Den of the row block is large enough for the series
        while (iter.next(I, J, last)) {
            if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 3);      // loop top (every tile, own or not)
            if (I != curI) {
                curI = I;
                gi = (p.rb0 + I) * 128 + r;
                rA = p.colA[gi];
                rB = p.colB[gi];
                yi = __float_as_int(rB.z);
                wlo = __reduce_min_sync(0xffffffffu, yi >= 0 ? yi : INT_MAX);
                whi = __reduce_max_sync(0xffffffffu, yi);
                rI = block_range(p.rb0 + I);
                if (kMode == DCL_MODE_PIXEL) {
                    const float4* cr = reinterpret_cast<const float4*>(p.coefR + static_cast<size_t>(gi) * 16);
                    const float4 n0 = cr[0], n1 = cr[1], q0 = cr[2], q1 = cr[3];
                    rn[0] = pack2(n0.x, n0.x); rn[1] = pack2(n0.y, n0.y); rn[2] = pack2(n0.z, n0.z);
                    rn[3] = pack2(n0.w, n0.w); rn[4] = pack2(n1.x, n1.x);
                    rq[0] = pack2(q0.x, q0.x); rq[1] = pack2(q0.y, q0.y); rq[2] = pack2(q0.z, q0.z);
                    rq[3] = pack2(q0.w, q0.w); rq[4] = pack2(q1.x, q1.x);
                    denI = sDenOk[p.rb0 + I] != 0;
                }
            }
            if ((it & 1) == g) {
                const int slot = it % kSlots, st = it % kStages;
                if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 0);
                mbar_wait(b_full + 8 * slot, (it / kSlots) & 1);       // column polynomials have landed
                mbar_wait(b_tfull + 8 * st, (it / kStages) & 1);
                if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 1);
                tc_fence_after();
                const float4* cp = gCP + slot * (SmemBwd::kCoefBytes / 16);
                const int col0 = J * 128;
                const uint32_t tS = tmem + st * 128 + lane_off;
                // Tile classes (pixel term): no same-class pair -> one polynomial per side; mixed classes and every
                // Den large -> two polynomials selected per element (+ the self pair on the diagonal tile);
                // otherwise (tiny denominators, image term) the exact masked form.  Padding needs no mask: padded F
                // rows are zero and padded columns carry zero coefficients.
                const bool overlap = ranges_overlap(rI, block_range(J));
                const bool fast = kMode == DCL_MODE_PIXEL && !overlap;
                const bool series = kMode == DCL_MODE_PIXEL && overlap && denI && sDenOk[J];
                if (p.debug & 1) {
                    // diagnostics: no G is produced
                } else if (fast) {
                    auto body = [&](auto degc) {
                        for_each_chunk_loop(tS, [&](int c0, const uint32_t (&v)[32]) {
                            uint32_t pk[16];
                            bwd_chunk<decltype(degc)::value>(v, rn, cp + (c0 >> 1) * 3, pk);
                            tmem_st16(tS + (c0 >> 1), pk);
                        });
                    };
                    // the two common degrees walk the row with the next chunk's TMEM load in flight
                    auto body_pf = [&](auto degc) {
                        for_each_chunk<4>(tS, [&](int c0, const uint32_t (&v)[32]) {
                            uint32_t pk[16];
                            bwd_chunk<decltype(degc)::value>(v, rn, cp + (c0 >> 1) * 3, pk);
                            tmem_st16(tS + (c0 >> 1), pk);
                        });
                    };
                    // degree 3 runs the degree-4 code (its x^4 coefficients are zero): one instantiation less
                    if (deg == 1) body_pf(std::integral_constant<int, 1>{});
                    else if (deg == 2) body_pf(std::integral_constant<int, 2>{});
                    else body(std::integral_constant<int, 4>{});
                } else if (series) {
                    const float4* cq = gCQ + slot * (SmemBwd::kCoefBytes / 16);
                    const int selfcol = ((p.rb0 + I) == J) ? r : -1000;
                    const float l1 = 2.f * rA.z * rA.x, l0 = 2.f * rA.z * rA.y;
                    // rows are class-sorted, so most 32-row x 32-column pieces of a mixed tile are uniform: no shared
                    // class -> the different-class polynomials alone; one class on both sides (and not the piece
                    // with the rows' own columns) -> the same-class polynomials alone; element-wise only on a boundary
                    const bool diag = (p.rb0 + I) == J;
                    auto body = [&](auto degc, auto walk) {
                        constexpr int kD = decltype(degc)::value;
                        walk(tS, [&](int c0, const uint32_t (&v)[32]) {
                            uint32_t pk[16];
                            const float4* cpc = cp + (c0 >> 1) * 3;
                            const int yl = __float_as_int(reinterpret_cast<const float*>(cpc + 3 * (lane >> 1) + 2)[2 + (lane & 1)]);
                            const int cmin = __reduce_min_sync(0xffffffffu, yl), cmax = __reduce_max_sync(0xffffffffu, yl);
                            if (cmax < wlo || cmin > whi) {
                                bwd_chunk<kD>(v, rn, cpc, pk);
                            } else if (cmin == cmax && wlo == whi && cmin == wlo && !(diag && c0 == q * 32)) {
                                bwd_chunk<kD>(v, rq, cq + (c0 >> 1) * 3, pk);
                            } else {
                                bwd_chunk_masked<kD>(v, rn, rq, cpc, cq + (c0 >> 1) * 3, yi, selfcol - c0, l1, l0, pk);
                            }
                            tmem_st16(tS + (c0 >> 1), pk);
                        });
                    };
                    auto walk1 = [](uint32_t t, auto&& fn) { for_each_chunk_loop(t, fn); };
                    if (deg == 1) body(std::integral_constant<int, 1>{}, walk1);
                    else if (deg == 2) body(std::integral_constant<int, 2>{}, walk1);
                    else body(std::integral_constant<int, 4>{}, walk1);
                } else {
                    const float4* cA = p.colA + static_cast<size_t>(J) * 128;
                    const float4* cB = p.colB + static_cast<size_t>(J) * 128;
                    for_each_chunk_loop(tS, [&](int c0, const uint32_t (&v)[32]) {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float gg[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const float s = __uint_as_float(v[j + u]);
                                const float4 ck = __ldg(cA + c0 + j + u);
                                const float4 dk = __ldg(cB + c0 + j + u);
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
                if ((threadIdx.x & 127) == 0) trace_stamp(p, 3 + g, it, 2);
            }
            ++it;
            if (last) {
                // both groups drain half of the finished dF_I partial (64 columns each)
                mbar_wait(b_dfull, seg & 1);
                tc_fence_after();
                // a row block swept by this CTA alone goes straight to dF; otherwise a segment partial, and the last
                // of the block's CTAs to arrive adds the partials in segment order (bit-reproducible, no atomics on data)
                const int sidx = blockIdx.x - p.partD.first_cta(I), ns = p.partD.nseg(I);
                float* out = (ns == 1 ? p.dF + (static_cast<size_t>(I) * 128 + r) * 128
                                      : p.pD + ((static_cast<size_t>(I) * p.maxsegD + sidx) * 128 + r) * 128) + g * 64;
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
                if (ns > 1) {
                    __threadfence();
                    asm volatile("bar.sync 1, 256;" ::: "memory");        // the eight epilogue warps
                    if (threadIdx.x == 0) *sTicket = atomicAdd(p.dticket + I, 1u);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (*sTicket == static_cast<unsigned int>(ns - 1)) {
                        __threadfence();
                        const float* part = p.pD + static_cast<size_t>(I) * p.maxsegD * 128 * 128;
                        // warp w: rows w, w+8, ..; a warp reads 512 contiguous bytes per (row, segment).  Two or three
                        // segments (the usual case: a row block shared by 2-3 CTAs) are loaded in one batch so that the
                        // L2 round trips of all segments overlap; the sums keep the segment order either way.
                        float* dst = p.dF + static_cast<size_t>(I) * 128 * 128;
                        if (ns == 2) reduce_segments<2>(part, dst, ns, warp, lane);
                        else if (ns == 3) reduce_segments<3>(part, dst, ns, warp, lane);
                        else reduce_segments<0>(part, dst, ns, warp, lane);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) cta_time(p, 1, 1);
    if (threadIdx.x == 0) trace_stamp(p, 0, 1, 7);          // all roles done
    if (warp == kProducerWarp) tmem_dealloc<kTmemCols>(tmem);
}

// block info (first kernel of every forward / backward call; also resets the device-side scalars)
__global__ void __launch_bounds__(128) k_blockinfo(const Params p) {
    pdl_launch();
    pdl_wait();     // first kernel of a chain: launched without the attribute, returns at once
    __shared__ int si[12];
    __shared__ float sf[8];
    const int J = blockIdx.x, t = threadIdx.x;
    const int yv = p.y[J * 128 + t];
    // |f|^2 of row t straight from the F-tile (the two 128-byte panel rows of the row, any chunk order): `sqnorm`
    // only has to be valid for the caller's own rows, so a sharded caller need not exchange it
    float c = 0.f;
    if (yv >= 0) {
        const uint8_t* trow = p.tiles + static_cast<size_t>(J) * kTileBytes + t * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 x = __ldg(reinterpret_cast<const uint4*>(trow + h * kHalfBytes) + q);
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
                    c = fmaf(lo, lo, c);
                    c = fmaf(hi, hi, c);
                }
            }
    }
    const BlockStat b = block_stat(yv, c, true, si, sf);
    if (t == 0) {
        p.binfo[J] = make_int4(b.lo, b.hi, b.n, 0);
        p.bnorm[J] = make_float2(b.cx, b.cn);
        if (J < p.nI) p.dticket[J] = 0u;
        if (J == 0) { p.iscal[0] = 1; p.iscal[4] = 0; p.iscal[5] = 0; p.ticket[0] = 0u; p.ticket[1] = 0u; }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Layout {
    Part partS, partD;
    int nP, maxsegS, maxsegD, splitc, ctas;
    size_t off_binfo, off_bnorm, off_pA, off_pB, off_iscal, off_pF, off_pH, off_rowM, off_rowPos, off_rowflag, off_rowS, off_rowD,
        off_pC, off_bl, off_ticket, off_dticket, off_coefR, off_coefP, off_coefPP, off_bden, off_pD, bytes;
};

static int g_debug_flags = 0;
static long long* g_trace = nullptr;
static long long* g_cta_times = nullptr;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// `switch_tiles`: what one unit switch inside a CTA costs, in tiles (measured: ~4 in the sweeps, ~12 in the backward,
// profiles/r01j_cta_times.log)
static void make_part(Part& part, int& maxseg, int nU, int nJ, int ctas, int switch_tiles) {
    part.total = static_cast<long long>(nU) * nJ;
    part.nJ = nJ;
    part.G = static_cast<int>(part.total < ctas ? part.total : ctas);
    part.excl = 0;
    part.base = 1;
    part.extra = 0;
    const long long q = part.total / part.G;                  // >= 1 tiles per CTA
    maxseg = static_cast<int>((nJ + q - 1) / q + 1);
    if (nU <= part.G) {
        const int base = part.G / nU, extra = part.G % nU;
        const long long excl_longest = (nJ + base - 1) / base;                      // a unit that owns `base` CTAs
        const long long flat_longest = (part.total + part.G - 1) / part.G + (part.G % nU == 0 ? 0 : switch_tiles);
        if (excl_longest <= flat_longest) {
            part.excl = 1;
            part.base = base;
            part.extra = extra;
            maxseg = base + (extra ? 1 : 0);
        }
    }
}

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
// stride ~ 0.38 nJ, coprime with nJ: consecutive visits land far apart, every block is visited once per period
static int coprime_stride(int nJ) {
    if (nJ <= 2) return 1;
    int s = static_cast<int>(nJ * 0.381966f + 0.5f);
    if (s < 1) s = 1;
    while (gcd_int(s, nJ) != 1) ++s;
    return s % nJ == 0 ? 1 : s % nJ;
}

static Layout make_layout(int nI, int nJ) {
    Layout L;
    const int ctas = sm_count();
    L.ctas = ctas;
    L.nP = (nI + 1) / 2;
    make_part(L.partS, L.maxsegS, L.nP, nJ, ctas, 4);
    make_part(L.partD, L.maxsegD, nI, nJ, ctas, 12);
    int sc = (ctas + L.nP - 1) / L.nP;
    L.splitc = sc < 1 ? 1 : (sc > 8 ? 8 : sc);
    const size_t rows = static_cast<size_t>(nI) * 128, allrows = static_cast<size_t>(nJ) * 128;
    size_t o = 0;
    auto take = [&](size_t& off, size_t bytes) { off = o; o = align_up(o + bytes, 256); };
    take(L.off_binfo, sizeof(int4) * nJ);
    take(L.off_bnorm, sizeof(float2) * nJ);
    take(L.off_pA, sizeof(float4) * rows * L.maxsegS * 2);    // legacy A partials; the v3 sweep-P partials (pF) share it
    take(L.off_pB, sizeof(float4) * rows * L.maxsegS * 2);    // legacy B partials; the v3 sweep-H partials (pH) share it
    take(L.off_iscal, sizeof(int) * 8);
    L.off_pF = L.off_pA;
    L.off_pH = L.off_pB;
    take(L.off_rowM, sizeof(float) * 8 * rows);
    take(L.off_rowPos, sizeof(float4) * rows);
    take(L.off_rowflag, sizeof(int) * rows);
    take(L.off_rowS, sizeof(float4) * rows);
    take(L.off_rowD, sizeof(float4) * rows);
    take(L.off_pC, sizeof(float4) * rows * L.splitc);
    take(L.off_bl, sizeof(float) * nI);
    take(L.off_ticket, sizeof(unsigned int) * 4);
    take(L.off_dticket, sizeof(unsigned int) * nI);
    take(L.off_coefR, sizeof(float) * 16 * allrows);
    take(L.off_coefP, sizeof(float) * kCoefPairFloats * (allrows / 2));
    take(L.off_coefPP, sizeof(float) * kCoefPairFloats * (allrows / 2));
    take(L.off_bden, sizeof(float) * nJ);
    take(L.off_pD, sizeof(float) * rows * L.maxsegD * 128);
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
    p.nJ = nJ; p.rb0 = rb0; p.nI = nI; p.nP = L.nP; p.n_valid = n_valid; p.mode = mode; p.ctas = L.ctas;
    p.jstride = coprime_stride(nJ);
    p.T = T; p.Tb = Tb;
    p.partS = L.partS;
    p.partD = L.partD;
    p.maxsegS = L.maxsegS;
    p.maxsegD = L.maxsegD;
    p.splitc = L.splitc;
    p.binfo = reinterpret_cast<int4*>(w + L.off_binfo);
    p.bnorm = reinterpret_cast<float2*>(w + L.off_bnorm);
    p.pA = reinterpret_cast<float4*>(w + L.off_pA);
    p.pB = reinterpret_cast<float2*>(w + L.off_pB);
    p.iscal = reinterpret_cast<int*>(w + L.off_iscal);
    p.pF = reinterpret_cast<float4*>(w + L.off_pF);
    p.pH = reinterpret_cast<float4*>(w + L.off_pH);
    p.rowM = reinterpret_cast<float*>(w + L.off_rowM);
    p.rowPos = reinterpret_cast<float4*>(w + L.off_rowPos);
    p.rowflag = reinterpret_cast<int*>(w + L.off_rowflag);
    p.rowS = reinterpret_cast<float4*>(w + L.off_rowS);
    p.rowD = reinterpret_cast<float4*>(w + L.off_rowD);
    p.pC = reinterpret_cast<float4*>(w + L.off_pC);
    p.pD = reinterpret_cast<float*>(w + L.off_pD);
    p.coefR = reinterpret_cast<float*>(w + L.off_coefR);
    p.coefP = reinterpret_cast<float*>(w + L.off_coefP);
    p.coefPP = reinterpret_cast<float*>(w + L.off_coefPP);
    p.bden = reinterpret_cast<float*>(w + L.off_bden);
    p.blockloss = reinterpret_cast<float*>(w + L.off_bl);
    p.ticket = reinterpret_cast<unsigned int*>(w + L.off_ticket);
    p.dticket = reinterpret_cast<unsigned int*>(w + L.off_dticket);
    p.debug = g_debug_flags;
    p.trace = g_trace;
    p.cta_times = g_cta_times;
    return p;
}

// Launch on `st`; `dependent` marks the kernel as a programmatic dependent of the launch before it on the stream
// (see pdl_wait): the first kernel of a call is launched plainly and so is ordered after all earlier work.
static bool g_use_pdl = true;
template <class... KArgs, class... Args>
static cudaError_t launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                          bool dependent, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (dependent && g_use_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define DCL_LAUNCH(name, ...)                                                                              \
    do {                                                                                                   \
        cudaError_t e_ = launch(__VA_ARGS__);                                                              \
        if (e_ != cudaSuccess)                                                                             \
            return ::dcl::fail(static_cast<int>(e_), "launch %s: %s", name, cudaGetErrorString(e_));       \
    } while (0)

template <class K>
static int set_smem(K kernel, int bytes) {
    DCL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}
// The shared-memory opt-in is a per-device (per-context) attribute: one flag per device ordinal and kernel family.
struct Configured {
    bool done[64] = {false};
    bool& here() {
        const int dev = current_device();
        return done[(dev >= 0 && dev < 64) ? dev : 0];
    }
};

// v3 pixel forward: block info, one power-sum sweep, per-row combination; the higher-moment sweep and the exact
// positive-pair sweep C are launched too but leave at once unless k_rows raised their triggers
static int run_fwd_v3(Params p, const Layout& L, cudaStream_t st) {
    static Configured cfgd;
    bool& configured = cfgd.here();
    if (!configured) {
        if (int e = set_smem(k_sweep<SWEEP_P, DCL_MODE_PIXEL>, SmemSweep::kBytes)) return e;
        if (int e = set_smem(k_sweep<SWEEP_H, DCL_MODE_PIXEL>, SmemSweep::kBytes)) return e;
        if (int e = set_smem(k_sweep<SWEEP_C, DCL_MODE_PIXEL>, SmemSweep::kBytes)) return e;
        configured = true;
    }
    p.use_series = 1;
    DCL_LAUNCH("k_blockinfo", k_blockinfo, p.nJ, 128, 0, st, false, p);
    DCL_LAUNCH("k_sweep<P>", k_sweep<SWEEP_P, DCL_MODE_PIXEL>, L.partS.G, kThreads, SmemSweep::kBytes, st, true, p);
    DCL_LAUNCH("k_rows", k_rows, p.nI, 128, 0, st, true, p);
    DCL_LAUNCH("k_sweep<H>", k_sweep<SWEEP_H, DCL_MODE_PIXEL>, L.partS.G, kThreads, SmemSweep::kBytes, st, true, p);
    DCL_LAUNCH("k_combine2", k_combine2, p.nI, 128, 0, st, true, p);
    DCL_LAUNCH("k_sweep<C>", k_sweep<SWEEP_C, DCL_MODE_PIXEL>, L.nP * L.splitc, kThreads, SmemSweep::kBytes, st, true, p);
    DCL_LAUNCH("k_finalize", k_finalize<DCL_MODE_PIXEL>, p.nI, 128, 0, st, true, p);
    return 0;
}

template <int kMode>
static int run_fwd_legacy(Params p, const Layout& L, cudaStream_t st) {
    static Configured cfgd;
    bool& configured = cfgd.here();
    if (!configured) {
        if (int e = set_smem(k_sweep<SWEEP_A, kMode>, SmemSweep::kBytes)) return e;
        if (int e = set_smem(k_sweep<SWEEP_B, kMode>, SmemSweep::kBytes)) return e;
        if (int e = set_smem(k_sweep<SWEEP_C, kMode>, SmemSweep::kBytes)) return e;
        configured = true;
    }
    DCL_LAUNCH("k_blockinfo", k_blockinfo, p.nJ, 128, 0, st, false, p);
    DCL_LAUNCH("k_sweep<A>", k_sweep<SWEEP_A, kMode>, L.partS.G, kThreads, SmemSweep::kBytes, st, true, p);
    DCL_LAUNCH("k_combine_A", k_combine_A, p.nI, 128, 0, st, true, p);
    DCL_LAUNCH("k_sweep<B>", k_sweep<SWEEP_B, kMode>, L.partS.G, kThreads, SmemSweep::kBytes, st, true, p);
    DCL_LAUNCH("k_combine_B", k_combine_B, p.nI, 128, 0, st, true, p);
    DCL_LAUNCH("k_sweep<C>", k_sweep<SWEEP_C, kMode>, L.nP * L.splitc, kThreads, SmemSweep::kBytes, st, true, p);
    DCL_LAUNCH("k_finalize", k_finalize<kMode>, p.nI, 128, 0, st, true, p);
    return 0;
}

// `chained`: the call directly follows this module's forward of the same problem on `st`, so the first kernel may be
// a programmatic dependent of the forward's last one instead of waiting for the stream to drain
template <int kMode>
static int run_bwd(Params p, const Layout& L, float* dF, bool chained, cudaStream_t st) {
    static Configured cfgd;
    bool& configured = cfgd.here();
    if (!configured) {
        if (int e = set_smem(k_backward<kMode>, SmemBwd::kBytes)) return e;
        configured = true;
    }
    // first kernel of the chain (launched plainly): block info, and for the pixel term the row polynomials with it
    if (kMode == DCL_MODE_PIXEL) DCL_LAUNCH("k_bwd_prep", k_bwd_prep, p.nJ, 128, 0, st, chained, p);
    else DCL_LAUNCH("k_blockinfo", k_blockinfo, p.nJ, 128, 0, st, chained, p);
    p.dF = dF;
    DCL_LAUNCH("k_backward", k_backward<kMode>, L.partD.G, kThreads, SmemBwd::kBytes, st, true, p);
    return 0;
}

}  // namespace dcl

using namespace dcl;

// Diagnostics only: component-isolation switches for profiling (results are invalid when bits 0..2 are set).
// Bit 3 (8) selects the legacy three-sweep forward for the pixel term; bit 4 (16) launches every kernel plainly
// (no programmatic dependent launch).
extern "C" int dcl_debug_flags(int flags) {
    const int old = g_debug_flags;
    g_debug_flags = flags;
    g_use_pdl = !(flags & 16);
    return old;
}

// Diagnostics only: device buffer of 5*32*8 int64 that CTA 0 of the pipelined kernels fills with clock64 stamps.
// Diagnostics only: device buffer of 4*256*4 int64; every CTA of the sweep-P, backward and k_rows kernels records
// (globaltimer, clock64) at entry and exit.
extern "C" int dcl_debug_cta_times(void* device_buffer) {
    g_cta_times = static_cast<long long*>(device_buffer);
    return 0;
}

extern "C" int dcl_debug_trace(void* device_buffer) {
    g_trace = static_cast<long long*>(device_buffer);
    return 0;
}

// kernels launched by one forward (backward != 0: one backward) call for `mode` with the current debug flags
extern "C" int dcl_contrast_launches(int mode, int backward) {
    if (backward) return 2;
    return 7;
}

// Diagnostics / tests (host only): the tile partition a sweep (backward == 0: units = pairs of row blocks) or the
// backward (units = row blocks) would use for nI local row blocks, nJ column blocks and `ctas` CTAs.
//   begin [ctas + 1] out : first flat tile index (unit * nJ + k) of every CTA, begin[G] = units * nJ
//   unit_first, unit_nseg [units] out : first CTA and number of CTAs of every unit
//   meta [4] out : G, exclusive mode (0/1), maxseg (segments the workspace is sized for), units
extern "C" int dcl_debug_partition(int nI, int nJ, int ctas, int backward, long long* begin, int* unit_first,
                                   int* unit_nseg, int* meta) {
    if (nI <= 0 || nJ <= 0 || ctas <= 0 || !begin || !unit_first || !unit_nseg || !meta)
        return fail(DCL_ERR_ARG, "bad argument");
    Part part;
    int maxseg = 0;
    const int units = backward ? nI : (nI + 1) / 2;
    make_part(part, maxseg, units, nJ, ctas, backward ? 12 : 4);
    for (int c = 0; c <= part.G; ++c) begin[c] = part.begin(c);
    for (int u = 0; u < units; ++u) { unit_first[u] = part.first_cta(u); unit_nseg[u] = part.nseg(u); }
    meta[0] = part.G; meta[1] = part.excl; meta[2] = maxseg; meta[3] = units;
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

int dcl::contrast_fwd_ex(const void* tiles, const int32_t* y, const float* sqnorm, int nJ, int rb0, int nI, int n_valid,
                         int mode, float temperature, float base_temperature, void* workspace, size_t workspace_bytes,
                         float* colA, float* colB, float* rowloss, float* loss_sum, float* loss_out, void* stream) {
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
    p.loss_out = loss_out;
    if (mode == DCL_MODE_PIXEL)
        return (g_debug_flags & 8) ? run_fwd_legacy<DCL_MODE_PIXEL>(p, L, as_stream(stream))
                                   : run_fwd_v3(p, L, as_stream(stream));
    return run_fwd_legacy<DCL_MODE_SUPCON>(p, L, as_stream(stream));
}

extern "C" int dcl_contrast_fwd(const void* tiles, const int32_t* y, const float* sqnorm, int nJ, int rb0,
                                int nI, int n_valid, int mode, float temperature,
                                float base_temperature, void* workspace, size_t workspace_bytes,
                                float* colA, float* colB, float* rowloss, float* loss_sum,
                                void* stream) {
    return contrast_fwd_ex(tiles, y, sqnorm, nJ, rb0, nI, n_valid, mode, temperature, base_temperature, workspace,
                           workspace_bytes, colA, colB, rowloss, loss_sum, nullptr, stream);
}

int dcl::contrast_bwd_ex(const void* tiles, const int32_t* y, const float* colA, const float* colB, int nJ, int rb0,
                         int nI, int mode, void* workspace, size_t workspace_bytes, float* dF, bool chained, void* stream) {
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
    return mode == DCL_MODE_PIXEL ? run_bwd<DCL_MODE_PIXEL>(p, L, dF, chained, as_stream(stream))
                                  : run_bwd<DCL_MODE_SUPCON>(p, L, dF, chained, as_stream(stream));
}

extern "C" int dcl_contrast_bwd(const void* tiles, const int32_t* y, const float* colA, const float* colB,
                                int nJ, int rb0, int nI, int mode, void* workspace,
                                size_t workspace_bytes, float* dF, void* stream) {
    return contrast_bwd_ex(tiles, y, colA, colB, nJ, rb0, nI, mode, workspace, workspace_bytes, dF, false, stream);
}
