// HBM-bound front and back ends of the doubly contrastive loss:
//   classify : argmax over class logits + legacy-nearest label down-sampling + per-chunk
//              (label, hard/easy) histograms                      (reference loss.py:396-408, :278-312)
//   select   : "rank-th pixel of (image, label, hard/easy) in raster order" -> pixel id  (loss.py:308-331)
//   gather   : NCHW embeddings at the selected pixels -> bf16 F-tiles (+ |f|^2)          (loss.py:333, :409-410)
//   scatter  : anchor-row gradients back into a dense NCHW gradient                      (autograd of :333)
//   gap      : global average pool forward / backward for the image-level term          (loss.py:104,115)
// All index work is integer-exact with respect to the reference; only `gather` rounds (to bf16).
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

__global__ void __launch_bounds__(512)
k_classify(const int64_t* __restrict__ labels, const float* __restrict__ predict, int H, int W, int h,
           int w, int C, float scale_h, float scale_w, uint16_t* __restrict__ code,
           int32_t* __restrict__ chunk_hist, int n_chunks) {
    __shared__ int hist[kBins];
    const int b = blockIdx.y, chunk = blockIdx.x, hw = h * w;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int p0 = chunk * kChunk + threadIdx.x * 4;
    const float* pl = predict + static_cast<size_t>(b) * C * hw;
    const int64_t* lb = labels + static_cast<size_t>(b) * H * W;

    float best[4];
    int arg[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { best[e] = 0.f; arg[e] = 0; }
    const bool vec = ((hw & 3) == 0) && (p0 + 3 < hw);
    if (vec) {
        // first-index argmax: strict '>' while scanning classes upward; NaN wins once (torch.max)
        for (int c = 0; c < C; ++c) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(pl + static_cast<size_t>(c) * hw + p0));
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool take = (c == 0) || (v[e] > best[e]) || (v[e] != v[e] && best[e] == best[e]);
                if (take) { best[e] = v[e]; arg[e] = c; }
            }
        }
    } else {
        for (int e = 0; e < 4; ++e) {
            if (p0 + e >= hw) break;
            for (int c = 0; c < C; ++c) {
                const float v = __ldg(pl + static_cast<size_t>(c) * hw + p0 + e);
                const bool take = (c == 0) || (v > best[e]) || (v != v && best[e] == best[e]);
                if (take) { best[e] = v; arg[e] = c; }
            }
        }
    }
    uint16_t out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int p = p0 + e;
        int bin = -1;
        out[e] = 0xFFFF;
        if (p < hw) {
            const int yy = p / w, xx = p - yy * w;
            const long long lab = lb[static_cast<size_t>(nearest_src(yy, scale_h, H)) * W +
                                     nearest_src(xx, scale_w, W)];
            if (lab >= 0 && lab <= 255) {
                const int easy = (arg[e] == static_cast<int>(lab)) ? 1 : 0;
                out[e] = static_cast<uint16_t>(lab | (easy << 8));
                bin = static_cast<int>(lab) * 2 + easy;
            }
        }
        // warp-aggregated shared-memory histogram (labels are spatially coherent: heavy collisions)
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[bin], __popc(peers));
    }
    if (vec) {
        *reinterpret_cast<uint2*>(code + static_cast<size_t>(b) * hw + p0) =
            make_uint2(out[0] | (static_cast<uint32_t>(out[1]) << 16),
                       out[2] | (static_cast<uint32_t>(out[3]) << 16));
    } else {
        for (int e = 0; e < 4; ++e)
            if (p0 + e < hw) code[static_cast<size_t>(b) * hw + p0 + e] = out[e];
    }
    __syncthreads();
    int32_t* dst = chunk_hist + (static_cast<size_t>(b) * n_chunks + chunk) * kBins;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) dst[i] = hist[i];
}

// per (image, bin): exclusive prefix over chunks in place, totals to counts
__global__ void __launch_bounds__(kBins)
k_chunk_prefix(int32_t* __restrict__ chunk_hist, int32_t* __restrict__ counts, int n_chunks) {
    const int b = blockIdx.x, bin = threadIdx.x;
    int32_t* base = chunk_hist + static_cast<size_t>(b) * n_chunks * kBins + bin;
    int run = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int v = base[static_cast<size_t>(c) * kBins];
        base[static_cast<size_t>(c) * kBins] = run;
        run += v;
    }
    counts[b * kBins + bin] = run;
}

// one warp per request
__global__ void __launch_bounds__(256)
k_select(const uint16_t* __restrict__ code, const int32_t* __restrict__ chunk_prefix, int hw, int n_chunks,
         const int4* __restrict__ req, int N, int32_t* __restrict__ pix) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const int4 rq = req[n];
    if (rq.x < 0) {
        if (lane == 0) pix[n] = -1;
        return;
    }
    const int b = rq.x, bin = rq.y * 2 + rq.z, rank = rq.w;
    const uint16_t want = static_cast<uint16_t>(rq.y | (rq.z << 8));
    // chunk = last one whose exclusive prefix is <= rank
    const int32_t* pre = chunk_prefix + static_cast<size_t>(b) * n_chunks * kBins + bin;
    int chunk = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
        const int c = c0 + lane;
        const bool le = (c < n_chunks) && (pre[static_cast<size_t>(c) * kBins] <= rank);
        const unsigned m = __ballot_sync(0xffffffffu, le);
        if (m) chunk = c0 + 31 - __clz(m);
        if (m != 0xffffffffu) break;
    }
    int rem = rank - pre[static_cast<size_t>(chunk) * kBins];
    // scan the chunk, 8 codes per lane per step
    const uint16_t* cc = code + static_cast<size_t>(b) * hw;
    const int pbeg = chunk * kChunk;
    const int pend = min(pbeg + kChunk, hw);
    int found = -1;
    for (int p0 = pbeg; p0 < pend && found < 0; p0 += 256) {
        const int mine = p0 + lane * 8;
        unsigned mask = 0;
        if (((hw & 7) == 0) && mine + 7 < pend) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(cc + mine));
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint16_t cv = static_cast<uint16_t>(wds[e >> 1] >> ((e & 1) * 16));
                mask |= (cv == want ? 1u : 0u) << e;
            }
        } else {
            for (int e = 0; e < 8; ++e)
                if (mine + e < pend && cc[mine + e] == want) mask |= 1u << e;
        }
        const int cnt = __popc(mask);
        int incl = cnt;                                   // inclusive warp scan
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (rem < total) {
            const int excl = incl - cnt;
            const bool here = (rem >= excl) && (rem < incl);
            int pos = -1;
            if (here) {
                int k = rem - excl;                        // k-th set bit of mask
                unsigned m = mask;
                for (int i = 0; i < k; ++i) m &= m - 1;
                pos = mine + __ffs(m) - 1;
            }
            const unsigned who = __ballot_sync(0xffffffffu, here);
            found = __shfl_sync(0xffffffffu, pos, __ffs(who) - 1);
        } else {
            rem -= total;
        }
    }
    if (lane == 0) pix[n] = found >= 0 ? b * hw + found : -1;
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
    k_classify<<<dim3(n_chunks, B), 512, 0, as_stream(stream)>>>(labels, predict, H, W, h, w, C_cls, sh, sw,
                                                                 code, chunk_hist, n_chunks);
    DCL_LAUNCH_CHECK("k_classify");
    k_chunk_prefix<<<B, kBins, 0, as_stream(stream)>>>(chunk_hist, counts, n_chunks);
    DCL_LAUNCH_CHECK("k_chunk_prefix");
    return 0;
}

extern "C" int dcl_sample_select(const uint16_t* code, const int32_t* chunk_hist, int B, int hw,
                                 const int32_t* req, int N, int32_t* pix, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!code || !chunk_hist || !req || !pix) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || hw <= 0 || N < 0) return fail(DCL_ERR_ARG, "bad shape");
    if (reinterpret_cast<uintptr_t>(req) % 16) return fail(DCL_ERR_ARG, "req must be 16-byte aligned");
    if (N == 0) return 0;
    const int n_chunks = (hw + kChunk - 1) / kChunk;
    k_select<<<(N + 7) / 8, 256, 0, as_stream(stream)>>>(code, chunk_hist, hw, n_chunks,
                                                         reinterpret_cast<const int4*>(req), N, pix);
    DCL_LAUNCH_CHECK("k_select");
    return 0;
}

extern "C" int dcl_gather_tiles(const float* feats, int B, int hw, const int32_t* pix, int n_pad, void* tiles,
                                float* sqnorm, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!feats || !pix || !tiles || !sqnorm) return fail(DCL_ERR_ARG, "null pointer argument");
    if (B <= 0 || hw <= 0 || n_pad <= 0 || n_pad % 128) return fail(DCL_ERR_ARG, "n_pad must be a positive multiple of 128");
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

extern "C" int dcl_scatter_grad(const float* dF, const int32_t* pix, int n_rows, const float* grad_out,
                                float* dfeats, int B, int hw, int zero_fill, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!dF || !pix || !grad_out || !dfeats) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n_rows < 0 || B <= 0 || hw <= 0) return fail(DCL_ERR_ARG, "bad shape");
    if (zero_fill == 1)
        DCL_CUDA(cudaMemsetAsync(dfeats, 0, static_cast<size_t>(B) * kDim * hw * sizeof(float), as_stream(stream)));
    if (n_rows == 0) return 0;
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
