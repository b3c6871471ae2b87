// HBM-bound front and back ends of the doubly contrastive loss:
//   classify : argmax over class logits + legacy-nearest label down-sampling + per-chunk
//              (label, hard/easy) histograms + per-image totals    (reference loss.py:396-408, :278-312)
//   select   : "rank-th pixel of (image, label, hard/easy) in raster order" -> pixel id  (loss.py:308-331)
//   gather   : NCHW embeddings at the selected pixels -> bf16 F-tiles (+ |f|^2)          (loss.py:333, :409-410)
//   scatter  : anchor-row gradients back into a dense NCHW gradient                      (autograd of :333)
//   gap      : global average pool forward / backward for the image-level term          (loss.py:104,115)
// All index work is integer-exact with respect to the reference; only `gather` rounds (to bf16).
#include <cstdlib>
#include <cuda_bf16.h>
#include "dcl_common.cuh"
#include "dcl_ptx.cuh"

namespace dcl {

constexpr int kChunk = DCL_CHUNK_PIXELS;   // 2048 pixels per CTA
constexpr int kBins = DCL_HIST_BINS;       // 512

// legacy 'nearest' source index: min(floor(dst * float(in/out)), in-1)  (ATen upsample_nearest)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    int s = static_cast<int>(floorf(static_cast<float>(dst) * scale));
    return s < in_size - 1 ? s : in_size - 1;
}

// ---- classify: one 2048-pixel chunk per CTA, kPx pixels per thread (4 -> 512 threads, 8 -> 256 threads).  The kernel is
// HBM-bound: what matters is bytes in flight per SM, so all label loads and a batch of class planes are issued before
// anything is compared.  Four pixels per thread keep the thread under 64 registers (two 512-thread CTAs = 32 warps
// per SM instead of 16) and make a warp's plane loads contiguous (512 B per instruction).
template <int kPx, int kBatch>
__device__ __forceinline__ void argmax_batch(const float* __restrict__ pl, size_t hw, int c0, int cn, float (&best)[kPx],
                                             int (&arg)[kPx]) {
    float4 v[kBatch][kPx / 4];
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
        if (u < cn) {
            const float4* p = reinterpret_cast<const float4*>(pl + static_cast<size_t>(c0 + u) * hw);
#pragma unroll
            for (int q = 0; q < kPx / 4; ++q) v[u][q] = __ldcs(p + q);
        }
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
        if (u < cn) {
#pragma unroll
            for (int q = 0; q < kPx / 4; ++q) {
                const float x[4] = {v[u][q].x, v[u][q].y, v[u][q].z, v[u][q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // first-index argmax: strict '>' while scanning classes upward; NaN wins once (torch.max)
                    float& b = best[q * 4 + e];
                    const bool take = (c0 + u == 0) || (x[e] > b) || (x[e] != x[e] && b == b);
                    if (take) { b = x[e]; arg[q * 4 + e] = c0 + u; }
                }
            }
        }
}

template <int kPx>
__global__ void __launch_bounds__(kChunk / kPx, kPx == 4 ? 2 : 1)
k_classify(const int64_t* __restrict__ labels, const float* __restrict__ predict, int H, int W, int h,
           int w, int C, float scale_h, float scale_w, uint16_t* __restrict__ code,
           int32_t* __restrict__ chunk_hist, int32_t* __restrict__ counts, int n_chunks) {
    constexpr int kBatch = kPx == 4 ? 5 : 4;
    __shared__ int hist[kBins];
    const int b = blockIdx.y, chunk = blockIdx.x, hw = h * w;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int p0 = chunk * kChunk + threadIdx.x * kPx;
    const float* pl = predict + static_cast<size_t>(b) * C * hw;
    const int64_t* lb = labels + static_cast<size_t>(b) * H * W;

    // labels first: independent loads (one DRAM sector each at a 4x down-sampling)
    long long lab[kPx];
#pragma unroll
    for (int e = 0; e < kPx; ++e) {
        const int p = p0 + e;
        lab[e] = -1;
        if (p < hw) {
            const int yy = p / w, xx = p - yy * w;
            lab[e] = __ldcs(lb + static_cast<size_t>(nearest_src(yy, scale_h, H)) * W + nearest_src(xx, scale_w, W));
        }
    }
    float best[kPx];
    int arg[kPx];
#pragma unroll
    for (int e = 0; e < kPx; ++e) { best[e] = 0.f; arg[e] = 0; }
    const bool vec = ((hw & 3) == 0) && (p0 + kPx - 1 < hw);
    if (vec) {
        const float* q = pl + p0;
        int c = 0;
        for (; c + kBatch <= C; c += kBatch) argmax_batch<kPx, kBatch>(q, hw, c, kBatch, best, arg);
        if (c < C) argmax_batch<kPx, kBatch>(q, hw, c, C - c, best, arg);
    } else {
        for (int e = 0; e < kPx; ++e) {
            if (p0 + e >= hw) break;
            for (int c = 0; c < C; ++c) {
                const float v = __ldg(pl + static_cast<size_t>(c) * hw + p0 + e);
                const bool take = (c == 0) || (v > best[e]) || (v != v && best[e] == best[e]);
                if (take) { best[e] = v; arg[e] = c; }
            }
        }
    }
    // codes + histogram.  Labels are spatially coherent: the pixels of a thread that share its first pixel's label are
    // counted in two registers (hard / easy) and warp-aggregated with one match; the others (class boundaries) go
    // one by one.
    uint16_t out[kPx];
    const int lab0 = (lab[0] >= 0 && lab[0] <= 255) ? static_cast<int>(lab[0]) : -1;
    int n_hard0 = 0, n_easy0 = 0;
#pragma unroll
    for (int e = 0; e < kPx; ++e) {
        out[e] = 0xFFFF;
        if (lab[e] >= 0 && lab[e] <= 255) {
            const int l = static_cast<int>(lab[e]);
            const int easy = (arg[e] == l) ? 1 : 0;
            out[e] = static_cast<uint16_t>(l | (easy << 8));
            if (l == lab0) { n_hard0 += 1 - easy; n_easy0 += easy; }
            else atomicAdd(&hist[l * 2 + easy], 1);
        }
    }
    {
        const unsigned peers = __match_any_sync(0xffffffffu, lab0);
        const int sh = __reduce_add_sync(peers, n_hard0), se = __reduce_add_sync(peers, n_easy0);
        if (lab0 >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) {
            if (sh) atomicAdd(&hist[lab0 * 2], sh);
            if (se) atomicAdd(&hist[lab0 * 2 + 1], se);
        }
    }
    if (vec && (hw & 7) == 0) {
        uint32_t wds[kPx / 2];
#pragma unroll
        for (int e = 0; e < kPx / 2; ++e) wds[e] = out[2 * e] | (static_cast<uint32_t>(out[2 * e + 1]) << 16);
        uint16_t* dstc = code + static_cast<size_t>(b) * hw + p0;
        if (kPx == 8) *reinterpret_cast<uint4*>(dstc) = make_uint4(wds[0], wds[1], wds[kPx / 2 - 2], wds[kPx / 2 - 1]);
        else *reinterpret_cast<uint2*>(dstc) = make_uint2(wds[0], wds[1]);
    } else {
        for (int e = 0; e < kPx; ++e)
            if (p0 + e < hw) code[static_cast<size_t>(b) * hw + p0 + e] = out[e];
    }
    __syncthreads();
    // per-chunk histogram, bin-major ([image][bin][chunk]: select reads one bin's chunks as a contiguous run), and the
    // per-image totals (the host's count table)
    int32_t* dst = chunk_hist + static_cast<size_t>(b) * kBins * n_chunks + chunk;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
        const int v = hist[i];
        dst[static_cast<size_t>(i) * n_chunks] = v;
        if (v) atomicAdd(counts + b * kBins + i, v);
    }
}

// one warp per request: chunk from a warp scan over the bin's per-chunk counts, then the whole 2048-pixel chunk of
// codes in one round trip (8 x 16 bytes per lane, lane l owns pixels [64 l, 64 l + 64) of the chunk: raster order)
__global__ void __launch_bounds__(256)
k_select(const uint16_t* __restrict__ code, const int32_t* __restrict__ chunk_hist, int hw, int n_chunks,
         const int4* __restrict__ req, int N, int32_t* __restrict__ pix, int32_t* __restrict__ rowof) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const int4 rq = req[n];
    if (rq.x < 0) {
        if (lane == 0) pix[n] = -1;
        return;
    }
    const int b = rq.x, bin = rq.y * 2 + rq.z;
    int rem = rq.w;
    const uint16_t want = static_cast<uint16_t>(rq.y | (rq.z << 8));
    const int32_t* hst = chunk_hist + (static_cast<size_t>(b) * kBins + bin) * n_chunks;
    int chunk = -1;
    for (int c0 = 0; c0 < n_chunks && chunk < 0; c0 += 32) {
        const int c = c0 + lane;
        const int cnt = (c < n_chunks) ? __ldg(hst + c) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (rem < total) {
            const unsigned m = __ballot_sync(0xffffffffu, rem < incl);     // first lane whose inclusive count exceeds rem
            const int src = __ffs(m) - 1;
            chunk = c0 + src;
            rem -= __shfl_sync(0xffffffffu, incl - cnt, src);
        } else {
            rem -= total;
        }
    }
    if (chunk < 0) {                      // rank beyond the bin's population: cannot happen for a valid plan
        if (lane == 0) pix[n] = -1;
        return;
    }
    const uint16_t* cc = code + static_cast<size_t>(b) * hw;
    const int pbeg = chunk * kChunk, pend = min(pbeg + kChunk, hw);
    const int mine = pbeg + lane * 64;
    unsigned long long mask = 0ull;
    if (((hw & 7) == 0) && mine + 63 < pend) {
        uint4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(reinterpret_cast<const uint4*>(cc + mine) + q);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t wds[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint16_t cv = static_cast<uint16_t>(wds[e >> 1] >> ((e & 1) * 16));
                mask |= static_cast<unsigned long long>(cv == want ? 1u : 0u) << (q * 8 + e);
            }
        }
    } else {
        for (int e = 0; e < 64; ++e)
            if (mine + e < pend && cc[mine + e] == want) mask |= 1ull << e;
    }
    const int cnt = __popcll(mask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned m = __ballot_sync(0xffffffffu, rem < incl);
    int found = -1;
    if (m) {
        const int src = __ffs(m) - 1;
        int pos = -1;
        if (lane == src) {
            int k = rem - (incl - cnt);                // k-th set bit of this lane's mask
            unsigned long long mm = mask;
            for (int i = 0; i < k; ++i) mm &= mm - 1;
            pos = mine + __ffsll(static_cast<long long>(mm)) - 1;
        }
        found = __shfl_sync(0xffffffffu, pos, src);
    }
    if (lane == 0) {
        pix[n] = found >= 0 ? b * hw + found : -1;
        if (rowof && found >= 0) rowof[static_cast<size_t>(b) * hw + found] = n;     // inverse map for the chunk-wise kernels
    }
}

// one warp per anchor row; lane l owns channels 4l..4l+3
template <bool kFromPixels>
__global__ void __launch_bounds__(256)
k_gather(const float* __restrict__ src, int hw, const int32_t* __restrict__ pix, int n_rows, int n_pad,
         uint8_t* __restrict__ tiles, float* __restrict__ sqnorm) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= n_pad) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (kFromPixels) {
        const int pid = pix[n];
        if (pid >= 0) {
            const int b = pid / hw, p = pid - b * hw;
            const float* f = src + (static_cast<size_t>(b) * kDim + lane * 4) * hw + p;
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __ldg(f + static_cast<size_t>(e) * hw);
        }
    } else if (n < n_rows) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(n) * kDim + lane * 4));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
    const float r0 = __low2float(lo), r1 = __high2float(lo), r2 = __low2float(hi), r3 = __high2float(hi);
    float sq = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    uint8_t* t = tiles + static_cast<size_t>(n >> 7) * kTileBytes + ftile_offset(n & 127, lane * 4);
    *reinterpret_cast<uint2*>(t) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    if (lane == 0) sqnorm[n] = sq;
}

template <bool kAdd>
__global__ void __launch_bounds__(256)
k_scatter(const float* __restrict__ dF, const int32_t* __restrict__ pix, int n_rows,
          const float* __restrict__ grad_out, float* __restrict__ dfeats, int hw) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= n_rows) return;
    const int pid = pix[n];
    if (pid < 0) return;
    const float g = __ldg(grad_out);
    const float4 v = __ldg(reinterpret_cast<const float4*>(dF + static_cast<size_t>(n) * kDim + lane * 4));
    const int b = pid / hw, p = pid - b * hw;
    float* o = dfeats + (static_cast<size_t>(b) * kDim + lane * 4) * hw + p;
    if (kAdd) {
        // sampled pixels are distinct: every (pixel, channel) element has one writer
        const float a0 = o[0], a1 = o[static_cast<size_t>(hw)], a2 = o[static_cast<size_t>(hw) * 2],
                    a3 = o[static_cast<size_t>(hw) * 3];
        o[0] = fmaf(v.x, g, a0);
        o[static_cast<size_t>(hw)] = fmaf(v.y, g, a1);
        o[static_cast<size_t>(hw) * 2] = fmaf(v.z, g, a2);
        o[static_cast<size_t>(hw) * 3] = fmaf(v.w, g, a3);
    } else {
        o[0] = v.x * g;
        o[static_cast<size_t>(hw)] = v.y * g;
        o[static_cast<size_t>(hw) * 2] = v.z * g;
        o[static_cast<size_t>(hw) * 3] = v.w * g;
    }
}

__global__ void __launch_bounds__(256)
k_unpack(const float* __restrict__ dF, int n, const float* __restrict__ grad_out, float* __restrict__ dZ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (kDim / 4)) return;
    const float g = __ldg(grad_out);
    float4 v = __ldg(reinterpret_cast<const float4*>(dF) + i);
    v.x *= g; v.y *= g; v.z *= g; v.w *= g;
    reinterpret_cast<float4*>(dZ)[i] = v;
}

// ------------------------------------------------------------------------------- pixel-ordered gather / scatter
// The sampled pixels of an image are a sparse subset of its h*w positions, and the embedding tensor is NCHW: one
// anchor row touches 128 channel planes at the same pixel offset.  Walking the anchors row by row (one warp per row)
// costs a DRAM line per (row, channel) in an order without any locality (26 us for 8192 rows, 8x that at 65536).
// k_select therefore also records `rowof[image*h*w + pixel] = row`, and the kernels below work per 2048-pixel chunk:
// a CTA reads its chunk of the map (8 KB, coalesced), compacts the sampled pixels in ascending order in shared
// memory, and then moves 32 of them at a time, channel plane by channel plane - the 32 lanes of a warp touch
// ascending addresses inside one 8 KB window of one plane, so lines are shared between neighbouring anchors when the
// sampling is dense and DRAM pages stay open when it is not.
constexpr int kGsThreads = 256;

// compact the sampled pixels of chunk (b, chunk): list[i] = (pixel offset inside the image, row); returns the count
__device__ __forceinline__ int chunk_samples(const int32_t* __restrict__ rowof, int b, int chunk, int hw, int2* list,
                                             int* warp_tot) {
    const int p0 = chunk * kChunk + threadIdx.x * 8;
    int r[8];
    const int32_t* src = rowof + static_cast<size_t>(b) * hw + p0;
    if (((hw & 3) == 0) && p0 + 7 < hw) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(src)), c = __ldg(reinterpret_cast<const int4*>(src) + 1);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = c.x; r[5] = c.y; r[6] = c.z; r[7] = c.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = (p0 + e < hw) ? __ldg(src + e) : -1;
    }
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) cnt += r[e] >= 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kGsThreads / 32; ++w) {
        const int t = warp_tot[w];
        if (w < warp) base += t;
        total += t;
    }
    int o = base + incl - cnt;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (r[e] >= 0) list[o++] = make_int2(p0 + e, r[e]);
    __syncthreads();
    return total;
}

// gather: feats [B,128,hw] f32 at the sampled pixels -> bf16 F-tiles + |f|^2 (rows that no pixel maps to, i.e. the
// padding rows, are cleared by k_gather_pad)
__global__ void __launch_bounds__(kGsThreads)
k_gather_px(const float* __restrict__ feats, int hw, const int32_t* __restrict__ rowof, uint8_t* __restrict__ tiles,
            float* __restrict__ sqnorm) {
    __shared__ int2 list[kChunk];
    __shared__ int warp_tot[kGsThreads / 32];
    __shared__ float tile[32][kDim + 1];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int n = chunk_samples(rowof, b, chunk, hw, list, warp_tot);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* plane0 = feats + static_cast<size_t>(b) * kDim * hw;
    for (int g0 = 0; g0 < n; g0 += 32) {
        const int m = min(32, n - g0);
        const int2 mine = list[g0 + min(lane, m - 1)];
        // warp w: channels 16 w .. 16 w + 15, sixteen independent loads in flight per lane
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldg(plane0 + static_cast<size_t>(warp * 16 + u) * hw + mine.x);
#pragma unroll
        for (int u = 0; u < 16; ++u) tile[lane][warp * 16 + u] = v[u];
        __syncthreads();
        // warp w writes rows w, w + 8, ..: lane l owns channels 4 l .. 4 l + 3 (256 contiguous bytes per row)
        for (int i = warp; i < m; i += kGsThreads / 32) {
            const int row = list[g0 + i].y;
            const float a0 = tile[i][lane * 4], a1 = tile[i][lane * 4 + 1], a2 = tile[i][lane * 4 + 2], a3 = tile[i][lane * 4 + 3];
            __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1);
            __nv_bfloat162 hi = __floats2bfloat162_rn(a2, a3);
            const float r0 = __low2float(lo), r1 = __high2float(lo), r2 = __low2float(hi), r3 = __high2float(hi);
            float sq = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            uint8_t* t = tiles + static_cast<size_t>(row >> 7) * kTileBytes + ftile_offset(row & 127, lane * 4);
            *reinterpret_cast<uint2*>(t) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            if (lane == 0) sqnorm[row] = sq;
        }
        __syncthreads();
    }
}

// padding rows (pix < 0): zero features, zero norm.  One warp per row; rows with a pixel are left alone.
__global__ void __launch_bounds__(256)
k_gather_pad(const int32_t* __restrict__ pix, int n_pad, uint8_t* __restrict__ tiles, float* __restrict__ sqnorm) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= n_pad || pix[n] >= 0) return;
    uint8_t* t = tiles + static_cast<size_t>(n >> 7) * kTileBytes + ftile_offset(n & 127, lane * 4);
    *reinterpret_cast<uint2*>(t) = make_uint2(0u, 0u);
    if (lane == 0) sqnorm[n] = 0.f;
}

// scatter: dfeats [B,128,hw] at the sampled pixels = (kAdd ? old : 0) + dF[row] * g, chunk by chunk, plane by plane
template <bool kAdd>
__global__ void __launch_bounds__(kGsThreads)
k_scatter_px(const float* __restrict__ dF, const int32_t* __restrict__ rowof, const float* __restrict__ grad_out,
             float* __restrict__ dfeats, int hw) {
    __shared__ int2 list[kChunk];
    __shared__ int warp_tot[kGsThreads / 32];
    __shared__ float tile[32][kDim + 1];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int n = chunk_samples(rowof, b, chunk, hw, list, warp_tot);
    if (n == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float g = __ldg(grad_out);
    float* plane0 = dfeats + static_cast<size_t>(b) * kDim * hw;
    for (int g0 = 0; g0 < n; g0 += 32) {
        const int m = min(32, n - g0);
        for (int i = warp; i < m; i += kGsThreads / 32) {
            const int row = list[g0 + i].y;
            const float4 v = __ldg(reinterpret_cast<const float4*>(dF + static_cast<size_t>(row) * kDim) + lane);
            tile[i][lane * 4] = v.x; tile[i][lane * 4 + 1] = v.y;
            tile[i][lane * 4 + 2] = v.z; tile[i][lane * 4 + 3] = v.w;
        }
        __syncthreads();
        if (lane < m) {
            const int p = list[g0 + lane].x;
            float* o = plane0 + static_cast<size_t>(warp * 16) * hw + p;
            if (kAdd) {
                float old[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) old[u] = o[static_cast<size_t>(u) * hw];
#pragma unroll
                for (int u = 0; u < 16; ++u) o[static_cast<size_t>(u) * hw] = fmaf(tile[lane][warp * 16 + u], g, old[u]);
            } else {
#pragma unroll
                for (int u = 0; u < 16; ++u) o[static_cast<size_t>(u) * hw] = tile[lane][warp * 16 + u] * g;
            }
        }
        __syncthreads();
    }
}

// zero-fill that can share the SMs with the persistent tensor-core kernels (32 registers, no shared memory): issued
// on a second stream it clears the dense gradient buffer underneath the N x N sweeps, which leave HBM idle
__global__ void __launch_bounds__(128)
k_zero_fill(float4* __restrict__ dst, size_t n16) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(dst + i, z);
}

// ------------------------------------------------------------------------------- global avg pool
__global__ void __launch_bounds__(256)
k_gap_fwd(const float* __restrict__ x, int hw, float* __restrict__ pooled) {
    const float* row = x + static_cast<size_t>(blockIdx.x) * hw;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if ((hw & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const int n4 = hw >> 2;
        int i = threadIdx.x;
        for (; i + 3 * 256 < n4; i += 4 * 256) {
            const float4 a = __ldg(r4 + i), b = __ldg(r4 + i + 256), c = __ldg(r4 + i + 512),
                         d = __ldg(r4 + i + 768);
            s[0] += (a.x + a.y) + (a.z + a.w);
            s[1] += (b.x + b.y) + (b.z + b.w);
            s[2] += (c.x + c.y) + (c.z + c.w);
            s[3] += (d.x + d.y) + (d.z + d.w);
        }
        for (; i < n4; i += 256) {
            const float4 a = __ldg(r4 + i);
            s[0] += (a.x + a.y) + (a.z + a.w);
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += 256) s[0] += __ldg(row + i);
    }
    float t = (s[0] + s[1]) + (s[2] + s[3]);
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < 8; ++i) tot += red[i];
        pooled[blockIdx.x] = tot / static_cast<float>(hw);
    }
}

template <bool kAccumulate>
__global__ void __launch_bounds__(256)
k_gap_bwd(const float* __restrict__ g, int hw, float* __restrict__ dx) {
    const float v = __ldg(g + blockIdx.x) / static_cast<float>(hw);
    float* row = dx + static_cast<size_t>(blockIdx.x) * hw;
    if ((hw & 3) == 0) {
        float4* r4 = reinterpret_cast<float4*>(row);
        const int n4 = hw >> 2;
        for (int i = threadIdx.x; i < n4; i += 256) {
            if (kAccumulate) {
                float4 o = r4[i];
                o.x += v; o.y += v; o.z += v; o.w += v;
                r4[i] = o;
            } else {
                r4[i] = make_float4(v, v, v, v);
            }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += 256) row[i] = kAccumulate ? row[i] + v : v;
    }
}

}  // namespace dcl

using namespace dcl;

extern "C" int dcl_sample_classify(const int64_t* labels, const float* predict, int B, int H, int W, int h,
                                   int w, int C_cls, uint16_t* code, int32_t* chunk_hist, int32_t* counts,
                                   void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!labels || !predict || !code || !chunk_hist || !counts) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || C_cls <= 0)
        return fail(DCL_ERR_ARG, "bad shape B=%d H=%d W=%d h=%d w=%d C=%d", B, H, W, h, w, C_cls);
    if (B > 65535) return fail(DCL_ERR_ARG, "B > 65535");
    const int hw = h * w;
    const int n_chunks = (hw + kChunk - 1) / kChunk;
    // float32 scale exactly as ATen computes it: (float)in / out
    const float sh = static_cast<float>(H) / static_cast<float>(h);
    const float sw = static_cast<float>(W) / static_cast<float>(w);
    // the per-image totals are accumulated with atomics: clear them first (2 KB per image)
    DCL_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * kBins * B, as_stream(stream)));
    // pixels per thread: 4 by default (DCL_CLASSIFY_PX=8 selects the 256-thread variant, diagnostics)
    static const int px = [] { const char* e = std::getenv("DCL_CLASSIFY_PX"); return (e && std::atoi(e) == 8) ? 8 : 4; }();
    if (px == 8)
        k_classify<8><<<dim3(n_chunks, B), kChunk / 8, 0, as_stream(stream)>>>(labels, predict, H, W, h, w, C_cls, sh, sw, code,
                                                                               chunk_hist, counts, n_chunks);
    else
        k_classify<4><<<dim3(n_chunks, B), kChunk / 4, 0, as_stream(stream)>>>(labels, predict, H, W, h, w, C_cls, sh, sw, code,
                                                                               chunk_hist, counts, n_chunks);
    DCL_LAUNCH_CHECK("k_classify");
    return 0;
}

extern "C" int dcl_sample_select(const uint16_t* code, const int32_t* chunk_hist, int B, int hw,
                                 const int32_t* req, int N, int32_t* pix, int32_t* rowof, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!code || !chunk_hist || !req || !pix) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || hw <= 0 || N < 0) return fail(DCL_ERR_ARG, "bad shape");
    if (reinterpret_cast<uintptr_t>(req) % 16) return fail(DCL_ERR_ARG, "req must be 16-byte aligned");
    if (rowof) DCL_CUDA(cudaMemsetAsync(rowof, 0xFF, sizeof(int32_t) * static_cast<size_t>(B) * hw, as_stream(stream)));
    if (N == 0) return 0;
    const int n_chunks = (hw + kChunk - 1) / kChunk;
    k_select<<<(N + 7) / 8, 256, 0, as_stream(stream)>>>(code, chunk_hist, hw, n_chunks,
                                                         reinterpret_cast<const int4*>(req), N, pix, rowof);
    DCL_LAUNCH_CHECK("k_select");
    return 0;
}

extern "C" int dcl_gather_tiles(const float* feats, int B, int hw, const int32_t* pix, int n_pad, void* tiles,
                                float* sqnorm, const int32_t* rowof, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!feats || !pix || !tiles || !sqnorm) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || hw <= 0 || n_pad <= 0 || n_pad % 128) return fail(DCL_ERR_ARG, "n_pad must be a positive multiple of 128");
    if (rowof) {
        if (B > 65535) return fail(DCL_ERR_ARG, "B > 65535");
        const int n_chunks = (hw + kChunk - 1) / kChunk;
        k_gather_pad<<<(n_pad + 7) / 8, 256, 0, as_stream(stream)>>>(pix, n_pad, static_cast<uint8_t*>(tiles), sqnorm);
        DCL_LAUNCH_CHECK("k_gather_pad");
        k_gather_px<<<dim3(n_chunks, B), kGsThreads, 0, as_stream(stream)>>>(feats, hw, rowof, static_cast<uint8_t*>(tiles), sqnorm);
        DCL_LAUNCH_CHECK("k_gather_px");
        return 0;
    }
    k_gather<true><<<(n_pad + 7) / 8, 256, 0, as_stream(stream)>>>(feats, hw, pix, n_pad, n_pad,
                                                                   static_cast<uint8_t*>(tiles), sqnorm);
    DCL_LAUNCH_CHECK("k_gather<pixels>");
    return 0;
}

extern "C" int dcl_pack_rows(const float* Z, int n, int n_pad, void* tiles, float* sqnorm, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!Z || !tiles || !sqnorm) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n <= 0 || n_pad < n || n_pad % 128) return fail(DCL_ERR_ARG, "n_pad must be a multiple of 128 and >= n");
    if (reinterpret_cast<uintptr_t>(Z) % 16) return fail(DCL_ERR_ARG, "Z must be 16-byte aligned");
    k_gather<false><<<(n_pad + 7) / 8, 256, 0, as_stream(stream)>>>(Z, 0, nullptr, n, n_pad,
                                                                    static_cast<uint8_t*>(tiles), sqnorm);
    DCL_LAUNCH_CHECK("k_gather<rows>");
    return 0;
}

extern "C" int dcl_zero_fill(void* dst, size_t bytes, int persistent, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!dst || bytes % 16 || reinterpret_cast<uintptr_t>(dst) % 16) return fail(DCL_ERR_ARG, "dst and bytes must be 16-byte aligned");
    if (bytes == 0) return 0;
    const size_t n16 = bytes / 16;
    size_t blocks;
    if (persistent) {
        // two 128-thread blocks per SM that walk the whole buffer: they never have blocks waiting for a slot, so a
        // kernel issued on another stream meanwhile starts at once next to them
        blocks = static_cast<size_t>(sm_count()) * 2;
        if (blocks * 128 > n16) blocks = (n16 + 127) / 128;
    } else {
        blocks = (n16 + 4095) / 4096;                     // 64 KB per block
        if (blocks > 1u << 20) blocks = 1u << 20;
    }
    k_zero_fill<<<static_cast<unsigned>(blocks), 128, 0, as_stream(stream)>>>(static_cast<float4*>(dst), n16);
    DCL_LAUNCH_CHECK("k_zero_fill");
    return 0;
}

extern "C" int dcl_scatter_grad(const float* dF, const int32_t* pix, int n_rows, const float* grad_out,
                                float* dfeats, int B, int hw, int zero_fill, const int32_t* rowof, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!dF || !pix || !grad_out || !dfeats) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n_rows < 0 || B <= 0 || hw <= 0) return fail(DCL_ERR_ARG, "bad shape");
    if (zero_fill == 1)
        DCL_CUDA(cudaMemsetAsync(dfeats, 0, static_cast<size_t>(B) * kDim * hw * sizeof(float), as_stream(stream)));
    if (n_rows == 0) return 0;
    if (rowof) {
        if (B > 65535) return fail(DCL_ERR_ARG, "B > 65535");
        const int n_chunks = (hw + kChunk - 1) / kChunk;
        if (zero_fill == 2) k_scatter_px<true><<<dim3(n_chunks, B), kGsThreads, 0, as_stream(stream)>>>(dF, rowof, grad_out, dfeats, hw);
        else k_scatter_px<false><<<dim3(n_chunks, B), kGsThreads, 0, as_stream(stream)>>>(dF, rowof, grad_out, dfeats, hw);
        DCL_LAUNCH_CHECK("k_scatter_px");
        return 0;
    }
    if (zero_fill == 2) k_scatter<true><<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(dF, pix, n_rows, grad_out, dfeats, hw);
    else k_scatter<false><<<(n_rows + 7) / 8, 256, 0, as_stream(stream)>>>(dF, pix, n_rows, grad_out, dfeats, hw);
    DCL_LAUNCH_CHECK("k_scatter");
    return 0;
}

// ------------------------------------------------------------------------------- sharded exchange
// send = [colA of the local rows | colB of the local rows | (local loss sum, 0, 0, 0)] as float4
__global__ void __launch_bounds__(256)
k_shard_pack(const float4* __restrict__ colA, const float4* __restrict__ colB, const float* __restrict__ loss_sum,
             int row0, int n_pad, float4* __restrict__ send) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) {
        send[i] = colA[row0 + i];
        send[n_pad + i] = colB[row0 + i];
    }
    if (i == 0) send[2 * n_pad] = make_float4(loss_sum[0], 0.f, 0.f, 0.f);
}
// recv = `world` such messages; colA / colB of every rank's rows land in place, loss = sum of the partials / n_global
__global__ void __launch_bounds__(256)
k_shard_unpack(const float4* __restrict__ recv, int world, int n_pad, float4* __restrict__ colA,
               float4* __restrict__ colB, int n_global, float* __restrict__ loss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int msg = 2 * n_pad + 1;
    if (i < world * n_pad) {
        const int r = i / n_pad, k = i - r * n_pad;
        colA[i] = recv[static_cast<size_t>(r) * msg + k];
        colB[i] = recv[static_cast<size_t>(r) * msg + n_pad + k];
    }
    if (i == 0) {
        float sum = 0.f;
        for (int r = 0; r < world; ++r) sum += recv[static_cast<size_t>(r) * msg + 2 * n_pad].x;      // rank order
        loss[0] = sum / static_cast<float>(n_global);
    }
}

extern "C" int dcl_shard_pack(const float* colA, const float* colB, const float* loss_sum, int rank, int n_pad,
                              float* send, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!colA || !colB || !loss_sum || !send || n_pad <= 0 || rank < 0) return fail(DCL_ERR_ARG, "bad argument");
    k_shard_pack<<<(n_pad + 255) / 256, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(colA), reinterpret_cast<const float4*>(colB), loss_sum, rank * n_pad, n_pad,
        reinterpret_cast<float4*>(send));
    DCL_LAUNCH_CHECK("k_shard_pack");
    return 0;
}

extern "C" int dcl_shard_unpack(const float* recv, int world, int n_pad, float* colA, float* colB, int n_global,
                                float* loss, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!recv || !colA || !colB || !loss || n_pad <= 0 || world <= 0 || n_global <= 0) return fail(DCL_ERR_ARG, "bad argument");
    k_shard_unpack<<<(world * n_pad + 255) / 256, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(recv), world, n_pad, reinterpret_cast<float4*>(colA),
        reinterpret_cast<float4*>(colB), n_global, loss);
    DCL_LAUNCH_CHECK("k_shard_unpack");
    return 0;
}

extern "C" int dcl_unpack_rows(const float* dF, int n, const float* grad_out, float* dZ, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!dF || !grad_out || !dZ || n <= 0) return fail(DCL_ERR_ARG, "bad argument");
    const int total = n * (kDim / 4);
    k_unpack<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(dF, n, grad_out, dZ);
    DCL_LAUNCH_CHECK("k_unpack");
    return 0;
}

extern "C" int dcl_gap_fwd(const float* x, int R, int hw, float* pooled, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!x || !pooled || R <= 0 || hw <= 0) return fail(DCL_ERR_ARG, "bad argument");
    if ((hw & 3) == 0 && reinterpret_cast<uintptr_t>(x) % 16) return fail(DCL_ERR_ARG, "x must be 16-byte aligned");
    k_gap_fwd<<<R, 256, 0, as_stream(stream)>>>(x, hw, pooled);
    DCL_LAUNCH_CHECK("k_gap_fwd");
    return 0;
}

extern "C" int dcl_gap_bwd(const float* g, int R, int hw, float* dx, int accumulate, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!g || !dx || R <= 0 || hw <= 0) return fail(DCL_ERR_ARG, "bad argument");
    if ((hw & 3) == 0 && reinterpret_cast<uintptr_t>(dx) % 16) return fail(DCL_ERR_ARG, "dx must be 16-byte aligned");
    if (accumulate) k_gap_bwd<true><<<R, 256, 0, as_stream(stream)>>>(g, hw, dx);
    else k_gap_bwd<false><<<R, 256, 0, as_stream(stream)>>>(g, hw, dx);
    DCL_LAUNCH_CHECK("k_gap_bwd");
    return 0;
}

// Dense gradient of the doubly step: dfeats = pooled-gradient broadcast (+ anchor gradients at the
// sampled pixels of the first B_pix images).  The dense tensor is written ONCE, by the streaming broadcast kernel
// (k_gap_bwd, at the HBM write roofline); the anchor rows are then added by the pixel-ordered scatter, which touches
// N x 128 elements only (1.03x the tensor size in DRAM traffic at the cfg3 shapes).  A single kernel that merged the
// two (per-chunk sample lists inside the broadcast loop) reached 58 % of the roofline where the broadcast alone runs
// at 100 %: 450 us against 313 + 35 us (profiles/r02_ncu_sampler_ncu_summary.csv), so the two stay separate launches.
extern "C" int dcl_dense_grad(const float* dF, const int32_t* rowof, int B_pix, const float* grad_out, const float* gap_g,
                              float* dfeats, int B_all, int hw, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!dF || !rowof || !grad_out || !gap_g || !dfeats) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B_pix < 0 || B_all < B_pix || B_all <= 0 || B_all > 65535 || hw <= 0) return fail(DCL_ERR_ARG, "bad shape");
    if (int e = dcl_gap_bwd(gap_g, B_all * kDim, hw, dfeats, 0, stream)) return e;
    if (B_pix == 0) return 0;
    const int n_chunks = (hw + kChunk - 1) / kChunk;
    k_scatter_px<true><<<dim3(n_chunks, B_pix), kGsThreads, 0, as_stream(stream)>>>(dF, rowof, grad_out, dfeats, hw);
    DCL_LAUNCH_CHECK("k_scatter_px");
    return 0;
}
