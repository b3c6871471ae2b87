// One-call forms of the pixel term's forward and backward (host glue only: every stage is one of the entry points
// of dcl_sampler.cu / dcl_host_rng.cpp / dcl_contrast.cu).  At the headline size the step is bound by the host's
// issue time, and most of that was the interpreter walking from one entry point to the next; here the whole chain
//   classify -> count table D2H -> (zero-fill || host plan) -> requests H2D -> select -> gather -> N x N forward
// is issued from C with a single wait (the count table), reference utils/loss.py:391-415 -> :250-389.
#include <cstdint>
#include <cuda_runtime.h>
#include "dcl_common.cuh"

using namespace dcl;

namespace {
cudaEvent_t count_event() {
    static thread_local cudaEvent_t ev = nullptr;
    static thread_local int ev_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (ev && ev_dev != dev) { cudaEventDestroy(ev); ev = nullptr; }
    if (!ev) {
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
        ev_dev = dev;
    }
    return ev;
}
}  // namespace

// First half: everything that does not depend on the host plan.  Issued as early as possible so that the caller's
// remaining preparation (allocations, generator state, descriptor) overlaps the classification on the GPU.
extern "C" int dcl_pixel_begin(const int64_t* labels, const float* predict, int B, int H, int W, int h, int w, int C_cls,
                               uint16_t* code, int32_t* chunk_hist, int32_t* counts_dev, int32_t* counts_host,
                               void* zero_fill, size_t zero_fill_bytes, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!counts_host) return fail(DCL_ERR_ARG, "null pointer argument");
    cudaStream_t st = as_stream(stream);
    cudaEvent_t ev = count_event();
    if (!ev) return fail(DCL_ERR_ARG, "cannot create a CUDA event");
    if (int e = dcl_sample_classify(labels, predict, B, H, W, h, w, C_cls, code, chunk_hist, counts_dev, stream)) return e;
    DCL_CUDA(cudaMemcpyAsync(counts_host, counts_dev, sizeof(int32_t) * 512 * B, cudaMemcpyDeviceToHost, st));
    DCL_CUDA(cudaEventRecord(ev, st));
    if (zero_fill && zero_fill_bytes)                  // runs on the GPU while the host plans
        DCL_CUDA(cudaMemsetAsync(zero_fill, 0, zero_fill_bytes, st));
    return 0;
}

extern "C" int dcl_pixel_fwd(const dcl_pixel_step_t* s, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!s) return fail(DCL_ERR_ARG, "null step descriptor");
    if (!s->labels || !s->predict || !s->feats || !s->code || !s->chunk_hist || !s->counts_dev || !s->counts_host ||
        !s->stage_host || !s->stage_dev || !s->info || !s->pix || !s->tiles || !s->sqnorm || !s->colA || !s->colB ||
        !s->rowloss || !s->loss_sum || !s->workspace)
        return fail(DCL_ERR_ARG, "null pointer in step descriptor");
    const int cap_min = (s->max_samples + DCL_TILE_ROWS - 1) / DCL_TILE_ROWS * DCL_TILE_ROWS;
    if (s->B <= 0 || s->h <= 0 || s->w <= 0 || s->cap < cap_min || s->cap < DCL_TILE_ROWS || s->cap % DCL_TILE_ROWS)
        return fail(DCL_ERR_ARG, "bad step shape (B=%d h=%d w=%d cap=%d max_samples=%d)", s->B, s->h, s->w, s->cap, s->max_samples);
    cudaStream_t st = as_stream(stream);
    cudaEvent_t ev = count_event();
    if (!ev) return fail(DCL_ERR_ARG, "cannot create a CUDA event");
    const int hw = s->h * s->w;
    if (!s->begun) {
        if (int e = dcl_pixel_begin(s->labels, s->predict, s->B, s->H, s->W, s->h, s->w, s->C_cls, s->code, s->chunk_hist,
                                    s->counts_dev, s->counts_host, s->zero_fill, s->zero_fill_bytes, stream))
            return e;
    }
    DCL_CUDA(cudaEventSynchronize(ev));                // the one unavoidable wait: the count table
    int32_t* req = s->stage_host;
    int32_t* y = s->stage_host + static_cast<size_t>(s->cap) * 4;
    const int rc = dcl_host_plan_rows(s->counts_host, s->B, s->ignore_label, s->max_samples, s->max_views,
                                      s->torch_rng_state, s->state_bytes, s->info, s->image, s->cls, s->num_hard,
                                      s->num_easy, s->keep_hard, s->ranks, req, y, s->ref_row, s->anchor);
    if (rc != 0) return rc;                            // 1: no class qualifies, 2: the reference's unreachable branch
    const int n_view = s->info[1], n = s->info[2], n_pad = s->info[3];
    if (n_view <= 0) return 3;                         // max_samples // total_classes == 0 (reference fails in torch.cat)
    if (n_pad > s->cap) return fail(DCL_ERR_ARG, "plan needs %d rows, capacity is %d", n_pad, s->cap);
    int32_t* req_dev = s->stage_dev;
    int32_t* y_dev = s->stage_dev + static_cast<size_t>(s->cap) * 4;
    DCL_CUDA(cudaMemcpyAsync(req_dev, req, sizeof(int32_t) * 4 * n_pad, cudaMemcpyHostToDevice, st));
    DCL_CUDA(cudaMemcpyAsync(y_dev, y, sizeof(int32_t) * n_pad, cudaMemcpyHostToDevice, st));
    if (int e = dcl_sample_select(s->code, s->chunk_hist, s->B, hw, req_dev, n_pad, s->pix, stream)) return e;
    if (int e = dcl_gather_tiles(s->feats, s->B, hw, s->pix, n_pad, s->tiles, s->sqnorm, stream)) return e;
    const int nJ = n_pad / DCL_TILE_ROWS;
    if (s->ev_begin) DCL_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(s->ev_begin), st));
    if (int e = dcl_contrast_fwd(s->tiles, y_dev, s->sqnorm, nJ, 0, nJ, n, DCL_MODE_PIXEL, s->temperature,
                                 s->base_temperature, s->workspace, s->workspace_bytes, s->colA, s->colB, s->rowloss,
                                 s->loss_sum, stream))
        return e;
    if (s->ev_end) DCL_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(s->ev_end), st));
    return 0;
}

extern "C" int dcl_pixel_bwd(const void* tiles, const int32_t* y, const float* colA, const float* colB, int n_pad,
                             void* workspace, size_t workspace_bytes, float* dF, const int32_t* pix,
                             const float* grad_out, float* dfeats, int B, int hw, int zero_fill, void* ev_begin,
                             void* ev_end, void* stream) {
    cudaStream_t st = as_stream(stream);
    const int nJ = n_pad / DCL_TILE_ROWS;
    if (n_pad <= 0 || n_pad % DCL_TILE_ROWS) return fail(DCL_ERR_ARG, "n_pad must be a positive multiple of %d", DCL_TILE_ROWS);
    if (ev_begin) DCL_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_begin), st));
    if (int e = dcl_contrast_bwd(tiles, y, colA, colB, nJ, 0, nJ, DCL_MODE_PIXEL, workspace, workspace_bytes, dF, stream))
        return e;
    if (ev_end) DCL_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_end), st));
    return dcl_scatter_grad(dF, pix, n_pad, grad_out, dfeats, B, hw, zero_fill, stream);
}
