// Segmentation-loss neighbour of the contrastive path: BoundaryAwareFocalLoss, reference
// utils/loss.py:27-80, forward AND gradient in one pass over the label-resolution pixels, without ever forming the
// up-sampled logits (the reference materialises [B,19,H,W] = 1.27 GB at batch 8 @ 1024x2048 and walks it ~10 times).
//
//   z = U x            x [B,C,h,w] pre-upsample logits, U = F.interpolate(bilinear, align_corners=False) (loss.py:5,41-42)
//   p = softmax_c(z),  k_i = class weight x EDT weight x exp(gamma (1 - p_{i,t_i}))   (focal factor detached, loss.py:61-70)
//   loss = - sum_i k_i log p_{i,t_i} / N,  N = #(EDT weight > 0)                       (loss.py:45,71)
//   d loss / d x = U^T [ k_i (p_{i,c} - delta_{c,t_i}) ] / N
//
// Owner-computes, no atomics, bit-reproducible: a thread owns ONE low-resolution column x of a strip of low-resolution
// rows and walks the label-resolution rows Y that touch the strip.  For every Y it evaluates the pixels X whose
// bilinear taps include column x (each pixel has two column taps, so every pixel is evaluated by two threads), reduces
// them along X with the column weights, and adds the result to the two low-resolution rows Y touches, which it holds in
// registers (the row taps are monotone in Y, so a row is complete when the walk leaves it and is stored exactly once).
// The pixel's loss, the EDT count and the reference's in-place `target[target == ignore_id] = 0` (loss.py:43) are done
// by the thread that owns the pixel's first tap.  MUFU-bound: 19 ex2 per evaluated pixel.
#include <cfloat>
#include "dcl_common.cuh"

namespace dcl {

constexpr int kFocalThreads = 128;
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
    double* partial;           // [blocks][2] loss sum (before the division), EDT count
    int B, C, h, w, H, W, ignore_id, mode, strip;
    float gamma, scale_h, scale_w;
};

template <int kC>
__global__ void __launch_bounds__(kFocalThreads) k_focal(const FocalParams p) {
    const int x = blockIdx.x * kFocalThreads + threadIdx.x;
    const int ys0 = blockIdx.y * p.strip, ys1 = min(ys0 + p.strip, p.h);
    const int b = blockIdx.z;
    const int C = (kC == 32) ? p.C : kC;
    double loss_sum = 0.0;
    float cnt = 0.f;
    if (x < p.w) {
        const size_t hw = static_cast<size_t>(p.h) * p.w;
        const float* L = p.logits + static_cast<size_t>(b) * C * hw;
        float* G = p.dlogits + static_cast<size_t>(b) * C * hw;
        long long* T = p.target + static_cast<size_t>(b) * p.H * p.W;
        const float* A = p.alpha + static_cast<size_t>(b) * p.H * p.W;
        // pixels X whose taps include column x: first tap in {x-1, x}
        const int Xlo = first_with_tap_ge(x - 1, p.scale_w, p.w, p.W);
        const int Xhi = first_with_tap_ge(x + 1, p.scale_w, p.w, p.W);          // exclusive
        const int Ylo = first_with_tap_ge(ys0 - 1, p.scale_h, p.h, p.H);
        const int Yhi = first_with_tap_ge(ys1, p.scale_h, p.h, p.H);            // exclusive
        const int xm = max(x - 1, 0), xp = min(x + 1, p.w - 1);
        float accA[kC], accB[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) accA[c] = accB[c] = 0.f;
        int ra = ys0 - 1;                        // accA collects low-res row ra, accB row ra + 1
        auto flush = [&]() {
            if (ra >= ys0 && ra < ys1) {
#pragma unroll
                for (int c = 0; c < kC; ++c)
                    if (c < C) G[static_cast<size_t>(c) * hw + static_cast<size_t>(ra) * p.w + x] = accA[c];
            }
#pragma unroll
            for (int c = 0; c < kC; ++c) { accA[c] = accB[c]; accB[c] = 0.f; }
            ++ra;
        };
        for (int Y = Ylo; Y < Yhi; ++Y) {
            int y0, y1;
            float ly;
            bilinear_tap(Y, p.scale_h, p.h, y0, y1, ly);
            while (y0 > ra) flush();
            if (y0 < ys0 && (y1 == y0 || ly == 0.f)) continue;      // the row above the strip with no weight on the strip
            // logits of columns x-1, x, x+1 interpolated to row Y
            float r0[kC], r1[kC], r2[kC];
            const float* Ly0 = L + static_cast<size_t>(y0) * p.w;
            const float* Ly1 = L + static_cast<size_t>(y1) * p.w;
#pragma unroll
            for (int c = 0; c < kC; ++c)
                if (c < C) {
                    const float* a0 = Ly0 + static_cast<size_t>(c) * hw;
                    const float* a1 = Ly1 + static_cast<size_t>(c) * hw;
                    const float t0 = __ldg(a0 + xm), t1 = __ldg(a0 + x), t2 = __ldg(a0 + xp);
                    const float u0 = __ldg(a1 + xm), u1 = __ldg(a1 + x), u2 = __ldg(a1 + xp);
                    r0[c] = t0 + ly * (u0 - t0);
                    r1[c] = t1 + ly * (u1 - t1);
                    r2[c] = t2 + ly * (u2 - t2);
                }
            const bool core_row = y0 >= ys0 && y0 < ys1;
            float rowg[kC];
#pragma unroll
            for (int c = 0; c < kC; ++c) rowg[c] = 0.f;
            for (int X = Xlo; X < Xhi; ++X) {
                int x0, x1;
                float lx;
                bilinear_tap(X, p.scale_w, p.w, x0, x1, lx);
                const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
                const bool owner = core_row && x0 == x;
                if (wx == 0.f && !owner) continue;
                const size_t pi = static_cast<size_t>(Y) * p.W + X;
                long long t = T[pi];
                if (t == p.ignore_id) {
                    t = 0;
                    if (owner) T[pi] = 0;                 // the reference rewrites its argument (loss.py:43)
                }
                const float a = __ldg(A + pi);
                const int tc = static_cast<int>(t);
                // z_c = (1 - lx) row[x0] + lx row[x1]; taps are x-1 / x / x+1 (clamped taps coincide with x)
                float z[kC];
                float m = -FLT_MAX;
#pragma unroll
                for (int c = 0; c < kC; ++c)
                    if (c < C) {
                        const float v0 = (x0 == x) ? r1[c] : ((x0 < x) ? r0[c] : r2[c]);
                        const float v1 = (x1 == x) ? r1[c] : ((x1 < x) ? r0[c] : r2[c]);
                        z[c] = v0 + lx * (v1 - v0);
                        m = fmaxf(m, z[c]);
                    }
                float s = 0.f, zt = 0.f;
#pragma unroll
                for (int c = 0; c < kC; ++c)
                    if (c < C) {
                        const float e = __expf(z[c] - m);
                        s += e;
                        if (c == tc) zt = z[c];
                        z[c] = e;
                    }
                const float inv = 1.f / s;
                const float logpt = (zt - m) - __logf(s);
                const float pt = __expf(logpt);
                const float focal = __expf(p.gamma * (1.f - pt));
                const float wt = (p.mode == FOCAL_FULL || p.mode == FOCAL_NO_EDT) ? __ldg(p.weight + min(max(tc, 0), C - 1)) : 1.f;
                const float k = p.mode == FOCAL_PLAIN ? focal
                              : p.mode == FOCAL_NO_CLASS_WEIGHTS ? a * focal
                              : p.mode == FOCAL_NO_EDT ? wt * focal : wt * a * focal;
                if (owner) {
                    loss_sum -= static_cast<double>(k * logpt);
                    cnt += a > 0.f ? 1.f : 0.f;
                }
                if (wx != 0.f) {
                    const float kw = k * wx;
#pragma unroll
                    for (int c = 0; c < kC; ++c)
                        if (c < C) rowg[c] += kw * (z[c] * inv - (c == tc ? 1.f : 0.f));
                }
            }
            // row Y feeds low-res rows y0 (weight 1 - ly) and y1 (weight ly; y1 == y0 at the bottom edge)
            const float w0 = (y1 == y0) ? 1.f : 1.f - ly, w1 = (y1 == y0) ? 0.f : ly;
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                accA[c] += w0 * rowg[c];
                accB[c] += w1 * rowg[c];
            }
        }
        while (ra < ys1) flush();
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
        const size_t blk = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partial[2 * blk] = l;
        p.partial[2 * blk + 1] = n;
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

extern "C" size_t dcl_focal_workspace_bytes(int B, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    const size_t bx = (static_cast<size_t>(w) + kFocalThreads - 1) / kFocalThreads;
    return sizeof(double) * 2 * bx * static_cast<size_t>(h) * B;      // one partial per block at the smallest strip (1 row)
}

extern "C" int dcl_focal_fwd(const float* logits, int64_t* target, const float* alpha, const float* weight, int B, int C,
                             int h, int w, int H, int W, int ignore_id, float gamma, int mode, float* dlogits_unscaled,
                             float* loss_n, void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!logits || !target || !alpha || !dlogits_unscaled || !loss_n || !workspace) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || C <= 0 || C > 32 || h <= 0 || w <= 0 || H < h || W < w || B > 65535)
        return fail(DCL_ERR_ARG, "bad shape B=%d C=%d h=%d w=%d H=%d W=%d (C <= 32, H >= h, W >= w)", B, C, h, w, H, W);
    if (mode < FOCAL_FULL || mode > FOCAL_NO_EDT) return fail(DCL_ERR_ARG, "bad mode %d", mode);
    if ((mode == FOCAL_FULL || mode == FOCAL_NO_EDT) && !weight) return fail(DCL_ERR_ARG, "class weights are required in this mode");
    FocalParams p{};
    p.logits = logits; p.target = reinterpret_cast<long long*>(target); p.alpha = alpha; p.weight = weight;
    p.dlogits = dlogits_unscaled;
    p.partial = static_cast<double*>(workspace);
    p.B = B; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W; p.ignore_id = ignore_id; p.mode = mode; p.gamma = gamma;
    p.scale_h = static_cast<float>(h) / static_cast<float>(H);       // ATen: (float)input_size / output_size
    p.scale_w = static_cast<float>(w) / static_cast<float>(W);
    // strips: enough blocks to fill the machine a few times over, at least 4 low-res rows each (one extra label row
    // group per strip is evaluated twice)
    const int bx = (w + kFocalThreads - 1) / kFocalThreads;
    int strip = h;
    const long long want = 6LL * sm_count();                          // a few 128-thread blocks per SM
    while (strip > 4 && static_cast<long long>(bx) * ((h + strip - 1) / strip) * B < want) strip = (strip + 1) / 2;
    p.strip = strip;
    const int by = (h + strip - 1) / strip;
    const size_t blocks = static_cast<size_t>(bx) * by * B;
    if (workspace_bytes < sizeof(double) * 2 * blocks) return fail(DCL_ERR_WORKSPACE, "workspace too small");
    if (by > 65535) return fail(DCL_ERR_ARG, "too many strips");
    const dim3 grid(bx, by, B);
    if (C == 19) k_focal<19><<<grid, kFocalThreads, 0, as_stream(stream)>>>(p);
    else k_focal<32><<<grid, kFocalThreads, 0, as_stream(stream)>>>(p);
    DCL_LAUNCH_CHECK("k_focal");
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
