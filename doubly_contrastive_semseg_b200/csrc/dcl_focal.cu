// Segmentation-loss neighbour of the contrastive path: BoundaryAwareFocalLoss, reference
// utils/loss.py:27-80, forward AND gradient in one pass over the label-resolution pixels, without ever forming the
// up-sampled logits (the reference materialises [B,19,H,W] = 1.27 GB at batch 8 @ 1024x2048 and walks it ~10 times).
//
//   z = U x            x [B,C,h,w] pre-upsample logits, U = F.interpolate(bilinear, align_corners=False) (loss.py:5,41-42)
//   p = softmax_c(z),  k_i = class weight x EDT weight x exp(gamma (1 - p_{i,t_i}))   (focal factor detached, loss.py:61-70)
//   loss = - sum_i k_i log p_{i,t_i} / N,  N = #(EDT weight > 0)                       (loss.py:45,71)
//   d loss / d x = U^T [ k_i (p_{i,c} - delta_{c,t_i}) ] / N
//
// Every label pixel is evaluated exactly once, nothing is accumulated with atomics and the result is bit-reproducible.
// A lane owns ONE low-resolution column c of a strip of low-resolution rows and evaluates the run of label columns X
// whose FIRST bilinear tap is c (the first tap is monotone in X, so the runs tile the row).  A pixel adds (1 - lx) of
// its term to column c -- the lane's own registers -- and lx to column c + 1: those parts are summed over the run and
// handed to the next lane with one shuffle per class and label row.  Lane 0 of a warp re-evaluates the run of the
// column before the warp's first one just for that hand-over (31 owned columns per warp).  Along Y the lane walks the
// label rows whose first tap lies in its strip, holding the two low-resolution rows a label row touches in registers
// (the taps are monotone in Y: a row is complete when the walk leaves it and is stored once); what it has collected for
// the first row of the NEXT strip goes to a spill row that k_focal_spill adds afterwards (fixed order: two terms).
// Per pixel: the class logits by two FMAs from the four taps (interpolated along Y once per label row), softmax with
// ex2, the target-class logit re-read from the taps (no dynamic register indexing), and three FMAs per class into the
// accumulators.  Work is dealt in warp tasks (column group, strip, image); the strip height is chosen so that the
// tasks fill the machine in whole waves.
#include <cfloat>
#include <cstdlib>
#include "dcl_common.cuh"
#include "dcl_ptx.cuh"

namespace dcl {

constexpr int kFocalThreads = 128;
constexpr float kFocalLog2e = 1.4426950408889634f;
constexpr float kFocalLn2 = 0.6931471805599453f;
enum { FOCAL_FULL = 0, FOCAL_PLAIN = 1, FOCAL_NO_CLASS_WEIGHTS = 2, FOCAL_NO_EDT = 3 };

// source tap of output index d: src = max((d + 0.5) * scale - 0.5, 0), i0 = floor(src), lambda = src - i0 (ATen
// area_pixel_compute_source_index, align_corners = false, float32)
__device__ __forceinline__ void bilinear_tap(int d, float scale, int in_size, int& i0, int& i1, float& lam) {
    float src = (static_cast<float>(d) + 0.5f) * scale - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = static_cast<int>(src);
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    lam = src - static_cast<float>(i0);
}
__device__ __forceinline__ int tap0(int d, float scale, int in_size) {
    int i0, i1;
    float lam;
    bilinear_tap(d, scale, in_size, i0, i1, lam);
    return i0;
}
// first output index whose first tap is >= t (out_size if none); the tap is monotone in the index
__device__ __forceinline__ int first_with_tap_ge(int t, float scale, int in_size, int out_size) {
    if (t <= 0) return 0;
    int d = static_cast<int>((static_cast<float>(t) + 0.5f) / scale - 0.5f) - 2;
    d = d < 0 ? 0 : (d > out_size ? out_size : d);
    while (d > 0 && tap0(d - 1, scale, in_size) >= t) --d;
    while (d < out_size && tap0(d, scale, in_size) < t) ++d;
    return d;
}


struct FocalParams {
    const float* logits;       // [B,C,h,w]
    long long* target;         // [B,H,W], ignore_id rewritten to 0
    const float* alpha;        // [B,H,W] EDT weight
    const float* weight;       // [C] class weights (unused in the plain / no_class_weights modes)
    float* dlogits;            // [B,C,h,w] sum_i k_i (p_ic - delta_ic) U^T, NOT yet divided by N
    float* spill;              // [B][nstrips][C][w] part of the first row of the next strip
    double* partial;           // [blocks][2] loss sum (before the division), EDT count
    int B, C, h, w, H, W, ignore_id, mode, strip, nstrips, ncg;
    float gamma, scale_h, scale_w;
};

constexpr int kFocalCols = 31;      // owned columns per warp (lane 0 is the hand-over lane)
constexpr int kFocalRun = 4;        // label pixels of a run whose loads are issued together
constexpr int kFocalStage = 33;     // staged low-resolution columns per warp: the 32 lanes' own and one to the right

// (packed fp32 pairs, FFMA2 / FADD2, from dcl_ptx.cuh: the kernel is bound by issue slots, two classes share one)
// z - (a == b ? s : 0) as a compare and a predicated subtract (the compiler turns the plain form, unrolled over the
// classes, into a jump table on the pixel's label: one divergent branch per pixel)
__device__ __forceinline__ float sub_if_eq(float z, float s, int a, int b) {
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\tmov.f32 %0, %1;\n\t@p sub.f32 %0, %1, %4;\n\t}"
        : "=&f"(r) : "f"(z), "r"(a), "r"(b), "f"(s));
    return r;
}

template <int kC, int kMinBlocks>
__global__ void __launch_bounds__(kFocalThreads, kMinBlocks) k_focal(const FocalParams p) {
    constexpr int kP = (kC + 1) / 2;                              // class pairs; a padding class has logit -1e30
    __shared__ float stage[kFocalThreads / 32][2][kC][kFocalStage];
    const int lane = threadIdx.x & 31;
    const long long task = static_cast<long long>(blockIdx.x) * (kFocalThreads / 32) + (threadIdx.x >> 5);
    const int C = (kC == 32) ? p.C : kC;
    double loss_sum = 0.0;
    float cnt = 0.f;
    if (task < static_cast<long long>(p.ncg) * p.nstrips * p.B) {            // warp-uniform
        float (*st)[kC][kFocalStage] = stage[threadIdx.x >> 5];              // st[row y0 / y1][class][column]
        const int cg = static_cast<int>(task % p.ncg);
        const int k = static_cast<int>((task / p.ncg) % p.nstrips);
        const int b = static_cast<int>(task / (static_cast<long long>(p.ncg) * p.nstrips));
        const int c = cg * kFocalCols + lane - 1;
        const bool active = c >= 0 && c < p.w;
        const bool owner = active && lane > 0;
        const bool last_col = active && c == p.w - 1;
        const int hw = p.h * p.w;
        const float* L = p.logits + static_cast<size_t>(b) * C * hw;
        float* G = p.dlogits + static_cast<size_t>(b) * C * hw;
        long long* T = p.target + static_cast<size_t>(b) * p.H * p.W;
        const float* A = p.alpha + static_cast<size_t>(b) * p.H * p.W;
        // staged columns: index j <-> column clamp(c0 - 1 + j); lane l evaluates with columns l and l + 1
        const int gcol = min(max(c, 0), p.w - 1);
        const int gcol32 = min(cg * kFocalCols + 31, p.w - 1);
        const int ys0 = k * p.strip, ys1 = min(ys0 + p.strip, p.h);
        const int Ylo = first_with_tap_ge(ys0, p.scale_h, p.h, p.H);
        const int Yhi = first_with_tap_ge(ys1, p.scale_h, p.h, p.H);            // exclusive
        int Xlo = 0, Xhi = 0;
        if (active) {
            Xlo = first_with_tap_ge(c, p.scale_w, p.w, p.W);
            Xhi = first_with_tap_ge(c + 1, p.scale_w, p.w, p.W);                // exclusive
        }
        const bool use_w = p.mode == FOCAL_FULL || p.mode == FOCAL_NO_EDT;
        const bool use_a = p.mode == FOCAL_FULL || p.mode == FOCAL_NO_CLASS_WEIGHTS;
        f32x2 accA[kP], accB[kP];
#pragma unroll
        for (int j = 0; j < kP; ++j) accA[j] = accB[j] = 0ull;
        int ra = ys0;                            // accA collects low-res row ra, accB row ra + 1
        int staged = -1;                         // low-res row pair in shared memory
        auto flush = [&]() {
#pragma unroll
            for (int j = 0; j < kP; ++j) {
                float lo, hi;
                unpack2(accA[j], lo, hi);
                if (owner) {
                    float* g = G + static_cast<size_t>(ra) * p.w + c;
                    if (2 * j < C) g[(2 * j) * hw] = lo;
                    if (2 * j + 1 < C) g[(2 * j + 1) * hw] = hi;
                }
                accA[j] = accB[j];
                accB[j] = 0ull;
            }
            ++ra;
        };
        for (int Y = Ylo; Y < Yhi; ++Y) {
            int y0, y1;
            float ly;
            bilinear_tap(Y, p.scale_h, p.h, y0, y1, ly);
            while (y0 > ra) flush();
            if (y0 != staged) {                  // warp-uniform: the two low-res rows of this run of label rows
                __syncwarp();
                const float* s0 = L + y0 * p.w;
                const float* s1 = L + y1 * p.w;
#pragma unroll
                for (int q = 0; q < kC; ++q)
                    if (q < C) {
                        st[0][q][lane] = __ldg(s0 + q * hw + gcol);
                        st[1][q][lane] = __ldg(s1 + q * hw + gcol);
                        if (lane == 0) {
                            st[0][q][32] = __ldg(s0 + q * hw + gcol32);
                            st[1][q][32] = __ldg(s1 + q * hw + gcol32);
                        }
                    }
                staged = y0;
                __syncwarp();
            }
            // row Y feeds low-res rows y0 (weight 1 - ly) and y1 (weight ly; y1 == y0 at the bottom edge)
            const float w0 = (y1 == y0) ? 1.f : 1.f - ly, w1 = (y1 == y0) ? 0.f : ly;
            // columns c and c + 1 interpolated to row Y, in log2 units: z_q(X) = r_q + lx d_q
            f32x2 r2[kP], d2[kP], SR[kP];
#pragma unroll
            for (int j = 0; j < kP; ++j) {
                float rr[2], dd[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = 2 * j + e;
                    rr[e] = -1e30f;
                    dd[e] = 0.f;
                    if (q < kC && q < C) {
                        const float t0 = st[0][q][lane], t1 = st[0][q][lane + 1];
                        const float u0 = st[1][q][lane], u1 = st[1][q][lane + 1];
                        const float v0 = t0 + ly * (u0 - t0), v1 = t1 + ly * (u1 - t1);
                        rr[e] = v0 * kFocalLog2e;
                        dd[e] = (v1 - v0) * kFocalLog2e;
                    }
                }
                r2[j] = pack2(rr[0], rr[1]);
                d2[j] = pack2(dd[0], dd[1]);
                SR[j] = 0ull;
            }
            for (int Xb = Xlo; Xb < Xhi; Xb += kFocalRun) {
                // loads of the whole batch first: label, EDT weight, then the four taps of the target class
                long long tt[kFocalRun];
                float aa[kFocalRun], zt[kFocalRun], lxs[kFocalRun];
#pragma unroll
                for (int u = 0; u < kFocalRun; ++u) {
                    const size_t pi = static_cast<size_t>(Y) * p.W + min(Xb + u, Xhi - 1);
                    tt[u] = T[pi];
                    aa[u] = __ldg(A + pi);
                }
#pragma unroll
                for (int u = 0; u < kFocalRun; ++u) {
                    const int X = min(Xb + u, Xhi - 1);
                    int x0, x1;
                    bilinear_tap(X, p.scale_w, p.w, x0, x1, lxs[u]);
                    if (tt[u] == p.ignore_id) {
                        tt[u] = 0;
                        if (owner && Xb + u < Xhi) T[static_cast<size_t>(Y) * p.W + X] = 0;   // the reference rewrites its argument (loss.py:43)
                    }
                    const int tc = min(max(static_cast<int>(tt[u]), 0), C - 1);
                    const float t0 = st[0][tc][lane], t1 = st[0][tc][lane + 1];
                    const float u0 = st[1][tc][lane], u1 = st[1][tc][lane + 1];
                    const float v0 = t0 + ly * (u0 - t0), v1 = t1 + ly * (u1 - t1);
                    zt[u] = fmaf(lxs[u], (v1 - v0) * kFocalLog2e, v0 * kFocalLog2e);
                }
#pragma unroll
                for (int u = 0; u < kFocalRun; ++u) {
                    if (Xb + u < Xhi) {
                        const float lx = lxs[u];
                        const int tc = static_cast<int>(tt[u]);
                        const f32x2 lx2 = pack2(lx, lx);
                        f32x2 z2[kP];
                        float m0 = -FLT_MAX, m1 = -FLT_MAX;
#pragma unroll
                        for (int j = 0; j < kP; ++j) {
                            z2[j] = ffma2(lx2, d2[j], r2[j]);
                            float lo, hi;
                            unpack2(z2[j], lo, hi);
                            m0 = fmaxf(m0, lo);
                            m1 = fmaxf(m1, hi);
                        }
                        const float m = fmaxf(m0, m1);
                        const f32x2 negm = pack2(-m, -m);
                        f32x2 s2 = 0ull;
#pragma unroll
                        for (int j = 0; j < kP; ++j) {
                            float lo, hi;
                            unpack2(fadd2(z2[j], negm), lo, hi);
                            z2[j] = pack2(exp2f(lo), exp2f(hi));
                            s2 = fadd2(s2, z2[j]);
                        }
                        float slo, shi;
                        unpack2(s2, slo, shi);
                        const float s = slo + shi;
                        const float logpt = ((zt[u] - m) - __log2f(s)) * kFocalLn2;
                        const float pt = __expf(logpt);
                        const float focal = __expf(p.gamma * (1.f - pt));
                        const float wt = use_w ? __ldg(p.weight + min(max(tc, 0), C - 1)) : 1.f;
                        const float kf = wt * (use_a ? aa[u] : 1.f) * focal;
                        if (owner) {
                            loss_sum -= static_cast<double>(kf * logpt);
                            cnt += aa[u] > 0.f ? 1.f : 0.f;
                        }
                        // k (p_q - delta_qt) = (k / s) (e_q - delta_qt s)
                        const float kq = kf / s;
                        const float qa = (1.f - lx) * w0 * kq, qb = (1.f - lx) * w1 * kq, qr = lx * kq;
                        const f32x2 qa2 = pack2(qa, qa), qb2 = pack2(qb, qb), qr2 = pack2(qr, qr);
#pragma unroll
                        for (int j = 0; j < kP; ++j) {
                            float lo, hi;
                            unpack2(z2[j], lo, hi);
                            lo = sub_if_eq(lo, s, 2 * j, tc);
                            hi = sub_if_eq(hi, s, 2 * j + 1, tc);
                            const f32x2 e2 = pack2(lo, hi);
                            accA[j] = ffma2(qa2, e2, accA[j]);
                            accB[j] = ffma2(qb2, e2, accB[j]);
                            SR[j] = ffma2(qr2, e2, SR[j]);
                        }
                    }
                }
            }
            // the lx parts belong to column c + 1: one lane up (the last column has both taps on itself)
            const f32x2 w02 = pack2(w0, w0), w12 = pack2(w1, w1);
#pragma unroll
            for (int j = 0; j < kP; ++j) {
                float lo, hi;
                unpack2(SR[j], lo, hi);
                float rlo = __shfl_up_sync(0xffffffffu, lo, 1), rhi = __shfl_up_sync(0xffffffffu, hi, 1);
                if (lane == 0) rlo = rhi = 0.f;
                if (last_col) { rlo += lo; rhi += hi; }
                const f32x2 recv = pack2(rlo, rhi);
                accA[j] = ffma2(w02, recv, accA[j]);
                accB[j] = ffma2(w12, recv, accB[j]);
            }
        }
        while (ra < ys1) flush();
        if (owner && ys1 < p.h) {               // accA now holds the strip's part of row ys1
            float* S = p.spill + (static_cast<size_t>(b) * p.nstrips + k) * C * p.w + c;
#pragma unroll
            for (int j = 0; j < kP; ++j) {
                float lo, hi;
                unpack2(accA[j], lo, hi);
                if (2 * j < C) S[(2 * j) * p.w] = lo;
                if (2 * j + 1 < C) S[(2 * j + 1) * p.w] = hi;
            }
        }
    }
    // deterministic block partial: fixed shuffle tree, then the warps in order
    __shared__ double sl[kFocalThreads / 32];
    __shared__ float sc[kFocalThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        loss_sum += __shfl_xor_sync(0xffffffffu, loss_sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = loss_sum; sc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0, n = 0.0;
        for (int i = 0; i < kFocalThreads / 32; ++i) { l += sl[i]; n += sc[i]; }
        p.partial[2 * static_cast<size_t>(blockIdx.x)] = l;
        p.partial[2 * static_cast<size_t>(blockIdx.x) + 1] = n;
    }
}

// first row of every strip but the first += what the strip above collected for it
__global__ void __launch_bounds__(256) k_focal_spill(const FocalParams p) {
    const size_t n = static_cast<size_t>(p.B) * (p.nstrips - 1) * p.C * p.w;
    const size_t hw = static_cast<size_t>(p.h) * p.w;
    for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256) {
        const int x = static_cast<int>(i % p.w);
        const int q = static_cast<int>((i / p.w) % p.C);
        const int k = static_cast<int>((i / (static_cast<size_t>(p.w) * p.C)) % (p.nstrips - 1));
        const int b = static_cast<int>(i / (static_cast<size_t>(p.w) * p.C * (p.nstrips - 1)));
        const float v = p.spill[((static_cast<size_t>(b) * p.nstrips + k) * p.C + q) * p.w + x];
        p.dlogits[(static_cast<size_t>(b) * p.C + q) * hw + static_cast<size_t>(k + 1) * p.strip * p.w + x] += v;
    }
}

// partials in block order -> out[0] = loss (0 when N == 0), out[1] = N
__global__ void __launch_bounds__(256) k_focal_reduce(const double* __restrict__ partial, int blocks, float* __restrict__ out) {
    __shared__ double sl[256], sn[256];
    double l = 0.0, n = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 256) { l += partial[2 * i]; n += partial[2 * i + 1]; }
    sl[threadIdx.x] = l;
    sn[threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sl[threadIdx.x] += sl[threadIdx.x + o]; sn[threadIdx.x] += sn[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sn[0] > 0.0 ? static_cast<float>(sl[0] / sn[0]) : 0.f;
        out[1] = static_cast<float>(sn[0]);
    }
}

// d logits = unscaled * (*grad_out) / N   (N == 0: zeros)
__global__ void __launch_bounds__(256)
k_focal_scale(const float* __restrict__ unscaled, const float* __restrict__ loss_n, const float* __restrict__ grad_out,
              float* __restrict__ out, size_t n) {
    const float N = loss_n[1];
    const float f = N > 0.f ? __ldg(grad_out) / N : 0.f;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = unscaled[i] * f;
}

}  // namespace dcl

using namespace dcl;

// Strip height: warp tasks = column groups x strips x images; a task's time is proportional to the strip height, the
// machine runs `slots` warps at a time, so the cost of a height is waves x height.  Heights below 4 rows are not
// considered (every strip but the last writes one spill row).
static int focal_strip(int ncg, int h, int B, long long slots) {
    int best = h;
    long long best_cost = -1;
    for (int s = h; s >= (h < 4 ? h : 4); --s) {
        const long long tasks = static_cast<long long>(ncg) * ((h + s - 1) / s) * B;
        const long long cost = ((tasks + slots - 1) / slots) * s;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    return best;
}

static size_t focal_spill_floats(int B, int h, int w) {
    const int min_strip = h < 4 ? h : 4;
    return static_cast<size_t>(B) * ((h + min_strip - 1) / min_strip) * 32 * w;      // C <= 32
}
static size_t focal_max_blocks(int B, int h, int w) {
    const int min_strip = h < 4 ? h : 4;
    const size_t tasks = static_cast<size_t>((w + kFocalCols - 1) / kFocalCols) * ((h + min_strip - 1) / min_strip) * B;
    return (tasks + kFocalThreads / 32 - 1) / (kFocalThreads / 32);
}

extern "C" size_t dcl_focal_workspace_bytes(int B, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    const size_t partial = (sizeof(double) * 2 * focal_max_blocks(B, h, w) + 255) / 256 * 256;
    return partial + sizeof(float) * focal_spill_floats(B, h, w);
}

extern "C" int dcl_focal_fwd(const float* logits, int64_t* target, const float* alpha, const float* weight, int B, int C,
                             int h, int w, int H, int W, int ignore_id, float gamma, int mode, float* dlogits_unscaled,
                             float* loss_n, void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!logits || !target || !alpha || !dlogits_unscaled || !loss_n || !workspace) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || C <= 0 || C > 32 || h <= 0 || w <= 0 || H < h || W < w || B > 65535)
        return fail(DCL_ERR_ARG, "bad shape B=%d C=%d h=%d w=%d H=%d W=%d (C <= 32, H >= h, W >= w)", B, C, h, w, H, W);
    if (static_cast<long long>(C) * h * w >= (1LL << 31) || static_cast<long long>(H) * W >= (1LL << 31))
        return fail(DCL_ERR_ARG, "image too large for 32-bit offsets (C*h*w and H*W must be below 2^31)");
    if (mode < FOCAL_FULL || mode > FOCAL_NO_EDT) return fail(DCL_ERR_ARG, "bad mode %d", mode);
    if ((mode == FOCAL_FULL || mode == FOCAL_NO_EDT) && !weight) return fail(DCL_ERR_ARG, "class weights are required in this mode");
    if (workspace_bytes < dcl_focal_workspace_bytes(B, h, w)) return fail(DCL_ERR_WORKSPACE, "workspace too small");
    FocalParams p{};
    p.logits = logits; p.target = reinterpret_cast<long long*>(target); p.alpha = alpha; p.weight = weight;
    p.dlogits = dlogits_unscaled;
    p.partial = static_cast<double*>(workspace);
    p.spill = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) +
                                       (sizeof(double) * 2 * focal_max_blocks(B, h, w) + 255) / 256 * 256);
    p.B = B; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W; p.ignore_id = ignore_id; p.mode = mode; p.gamma = gamma;
    p.scale_h = static_cast<float>(h) / static_cast<float>(H);       // ATen: (float)input_size / output_size
    p.scale_w = static_cast<float>(w) / static_cast<float>(W);
    // resident warps: blocks per SM of the kernel that runs, from the occupancy calculator (once per device and kernel)
    // kernel variant: 19 classes at 3 blocks per SM (168 registers, no spills; DCL_FOCAL_BLOCKS=4 selects the 128-register
    // build, diagnostics), any other class count through the 32-class build
    static const int want4 = [] { const char* e = std::getenv("DCL_FOCAL_BLOCKS"); return (e && std::atoi(e) == 4) ? 1 : 0; }();
    void (*kern)(FocalParams) = C != 19 ? k_focal<32, 1> : (want4 ? k_focal<19, 4> : k_focal<19, 3>);
    const int ki = C != 19 ? 0 : 1 + want4;
    static int occ_cache[64][3] = {};
    const int dev = current_device(), di = (dev >= 0 && dev < 64) ? dev : 0;
    if (occ_cache[di][ki] == 0) {
        int occ = 0;
        DCL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFocalThreads, 0));
        occ_cache[di][ki] = occ > 0 ? occ : 1;
    }
    p.ncg = (w + kFocalCols - 1) / kFocalCols;
    const long long slots = static_cast<long long>(sm_count()) * occ_cache[di][ki] * (kFocalThreads / 32);
    p.strip = focal_strip(p.ncg, h, B, slots);
    p.nstrips = (h + p.strip - 1) / p.strip;
    const long long tasks = static_cast<long long>(p.ncg) * p.nstrips * B;
    const long long blocks = (tasks + kFocalThreads / 32 - 1) / (kFocalThreads / 32);
    kern<<<static_cast<unsigned>(blocks), kFocalThreads, 0, as_stream(stream)>>>(p);
    DCL_LAUNCH_CHECK("k_focal");
    if (p.nstrips > 1) {
        const size_t n = static_cast<size_t>(B) * (p.nstrips - 1) * C * w;
        size_t sb = (n + 255) / 256;
        if (sb > 148 * 8) sb = 148 * 8;
        k_focal_spill<<<static_cast<unsigned>(sb), 256, 0, as_stream(stream)>>>(p);
        DCL_LAUNCH_CHECK("k_focal_spill");
    }
    k_focal_reduce<<<1, 256, 0, as_stream(stream)>>>(p.partial, static_cast<int>(blocks), loss_n);
    DCL_LAUNCH_CHECK("k_focal_reduce");
    return 0;
}

extern "C" int dcl_focal_bwd(const float* dlogits_unscaled, const float* loss_n, const float* grad_out, float* dlogits,
                             size_t n, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!dlogits_unscaled || !loss_n || !grad_out || !dlogits) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_focal_scale<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(dlogits_unscaled, loss_n, grad_out, dlogits, n);
    DCL_LAUNCH_CHECK("k_focal_scale");
    return 0;
}
