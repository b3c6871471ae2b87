// PixelContrastLoss.forward (+ its gradient) as ONE host call, single GPU or one rank of a sharded job
// (reference utils/loss.py:391-415 -> :250-389; the sharding has no reference counterpart):
//
//   classify -> [all-gather of the count tables] -> count table D2H          (side stream: zero-fill of d feats)
//   -> the one host wait -> plan on the host (anchor list, n_view, split rule; O(#anchors)) -> permutations:
//        k_plan on the GPU from the mirrored generator stream, or the host replay when the generator was touched
//   -> select -> gather -> [all-gather of the F-tiles] -> N x N forward -> [all-gather of the row constants]
//   -> N x N backward (dF of the local rows, for an upstream gradient of 1)
//
// The backward of the similarity is issued here, eagerly: the gradient is linear in the upstream scalar, so the
// autograd backward only has to scale dF while scattering it (dcl_step_bwd).  The GPU therefore always has the
// whole chain queued while the host goes through the autograd hand-over, which at the headline size used to cost
// more than the kernels themselves.
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <cuda_runtime.h>
#include "dcl_common.cuh"
#include "dcl_plan.h"

using namespace dcl;

namespace dcl {
int comm_all_gather(void* comm, const void* send, void* recv, size_t bytes, cudaStream_t st);
struct P2p;
uint8_t* p2p_begin(P2p* p, int kind);
int p2p_exchange(P2p* p, int kind, size_t bytes_per_rank, bool use_copy_engine, cudaStream_t st);
}

namespace {

constexpr int kMaxDevices = 64;

struct PerDevice {
    cudaEvent_t ev_counts = nullptr;      // count table has reached the host
    cudaEvent_t ev_fork = nullptr;        // main stream position at the start of a step (side stream waits for it)
    cudaEvent_t ev_zero = nullptr;        // zero-fill done (main stream waits for it before the step ends)
    // device mirror of the host generator look-ahead (raw mt19937 state blocks)
    uint32_t* d_ring = nullptr;
    uint64_t ring_blocks = 0, epoch = ~0ull, up_hi = 0;
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_up = nullptr, ev_need = nullptr, ev_used = nullptr;
    bool used_recorded = false;
    uint64_t last_window = 0;
    // dcl_step_sim_timing: the library's own events around the contrast forward / backward of every step
    cudaEvent_t* sim_ev = nullptr;        // [kSimRing][4]
    long long sim_steps = 0;
    bool sim_on = false;
};
constexpr int kSimRing = 2048;
PerDevice g_dev[kMaxDevices];
std::mutex g_mu;
// diagnostics: host nanoseconds of the last dcl_step_fwd at the end of each of its sections (dcl_step_timing)
long long g_step_ns[12] = {0};
inline long long now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Where the dense gradient buffer is cleared on the second stream: 0 = right after classify, while the count table
// is with the host (the GPU is idle then, but the fill also runs into k_plan / select and slows them); 1 = after the
// gather, underneath the N x N sweeps (they lose ~10 us to it, which only pays when they are long).  Measured
// (profiles/r02y_*): 8192 anchors 0.396 ms early vs 0.406 late; 65536 anchors 2.956 ms early vs 2.885 late.
int fill_placement(int max_samples) {
    static const int v = [] { const char* e = std::getenv("DCL_FILL_PLACEMENT"); return e ? std::atoi(e) : -1; }();
    if (v >= 0) return v;
    return max_samples >= 16384 ? 1 : 0;
}

int device_slot(PerDevice*& pd) {
    int dev = 0;
    DCL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(DCL_ERR_ARG, "device ordinal %d out of range", dev);
    pd = &g_dev[dev];
    if (!pd->ev_counts) {
        DCL_CUDA(cudaEventCreateWithFlags(&pd->ev_counts, cudaEventDisableTiming));
        DCL_CUDA(cudaEventCreateWithFlags(&pd->ev_fork, cudaEventDisableTiming));
        DCL_CUDA(cudaEventCreateWithFlags(&pd->ev_zero, cudaEventDisableTiming));
        DCL_CUDA(cudaEventCreateWithFlags(&pd->ev_up, cudaEventDisableTiming));
        DCL_CUDA(cudaEventCreateWithFlags(&pd->ev_need, cudaEventDisableTiming));
        DCL_CUDA(cudaEventCreateWithFlags(&pd->ev_used, cudaEventDisableTiming));
        DCL_CUDA(cudaStreamCreateWithFlags(&pd->copy, cudaStreamNonBlocking));
    }
    return 0;
}

// copy stream blocks [up_hi, target) of the host look-ahead ring into the device mirror (copy stream)
int upload_blocks(PerDevice& pd, const DevicePlan& dp, uint64_t target) {
    if (target <= pd.up_hi) return 0;
    if (pd.used_recorded) DCL_CUDA(cudaStreamWaitEvent(pd.copy, pd.ev_used, 0));
    uint64_t b = pd.up_hi;
    while (b < target) {
        const uint64_t slot = b % pd.ring_blocks;
        uint64_t n = target - b;
        if (n > pd.ring_blocks - slot) n = pd.ring_blocks - slot;
        DCL_CUDA(cudaMemcpyAsync(pd.d_ring + slot * kMtWords, dp.host_ring + slot * kMtWords,
                                 sizeof(uint32_t) * kMtWords * n, cudaMemcpyHostToDevice, pd.copy));
        b += n;
    }
    pd.up_hi = target;
    return 0;
}

// A guess at the NEXT step's window, pushed behind what is already mirrored: the next plan starts where this one
// ended (every rank consumes the whole stream), and this rank's permutations will sit at about the same offset into
// it.  Called once everything of the current step has been issued: the copy then shares neither the copy engine with
// the step's own small copies nor anything else with its critical path, and nothing waits for it until the next
// step's k_plan.  (The blocks of other ranks' permutations in between are copied too: 4 MB per step at most.)
int mirror_prefetch(PerDevice& pd, const DevicePlan& dp) {
    const uint64_t span = dp.last_block - dp.plan_first_block + 1;         // plan start .. end of this rank's window
    uint64_t target = dp.plan_end_block + span + span / 4 + 16;
    if (target > dp.produced) target = dp.produced;
    if (target > dp.first_block + pd.ring_blocks - 8) target = dp.first_block + pd.ring_blocks - 8;
    if (int e = upload_blocks(pd, dp, target)) return e;
    DCL_CUDA(cudaEventRecord(pd.ev_up, pd.copy));
    return 0;
}

// Make stream blocks [first, last] of the look-ahead readable on the device before anything later on `st`, and
// push the blocks the next step will probably need behind them (the upload then overlaps this step's kernels).
// Block b lives in slot b % ring_blocks on both sides.  A slot is overwritten only by block b + ring_blocks, i.e.
// never by anything the current window needs; kernels of earlier steps that may still read the old content are
// ordered before the copy through ev_used.
int mirror_window(PerDevice& pd, const DevicePlan& dp, cudaStream_t st) {
    if (!pd.d_ring) {
        pd.ring_blocks = dp.ring_blocks;
        DCL_CUDA(cudaMalloc(&pd.d_ring, sizeof(uint32_t) * kMtWords * pd.ring_blocks));
    }
    if (pd.ring_blocks != dp.ring_blocks) return fail(DCL_ERR_ARG, "generator ring size changed");
    if (pd.epoch != dp.epoch || dp.first_block > pd.up_hi || dp.first_block + pd.ring_blocks <= pd.up_hi) {
        pd.epoch = dp.epoch;                              // stream restarted (or a gap): nothing uploaded is usable
        pd.up_hi = dp.first_block;
    }
    const uint64_t window = dp.last_block - dp.first_block + 1;
    if (window > pd.ring_blocks - 16) return fail(DCL_ERR_ARG, "plan window of %llu blocks exceeds the generator ring",
                                                  static_cast<unsigned long long>(window));
    // (1) what this step reads: normally already there (pushed by the previous step), so the event below only marks
    //     the completion of that earlier copy; `st` waits for it
    if (int e = upload_blocks(pd, dp, dp.last_block + 1)) return e;
    DCL_CUDA(cudaEventRecord(pd.ev_need, pd.copy));
    DCL_CUDA(cudaStreamWaitEvent(st, pd.ev_need, 0));
    pd.last_window = window;
    return 0;
}

}  // namespace

extern "C" size_t dcl_step_plan_bytes(int B_local, int world) {
    if (B_local <= 0 || world <= 0) return 0;
    const size_t A_cap = static_cast<size_t>(B_local) * world * 256;
    // [ycnt world][yoff world][ycls A_cap] | [anchors B_local*256] | [yanchor A_cap] (host only)
    size_t o = sizeof(int32_t) * (2 * static_cast<size_t>(world) + A_cap);
    o = (o + 63) / 64 * 64;
    o += sizeof(PlanAnchor) * static_cast<size_t>(B_local) * 256;
    o = (o + 63) / 64 * 64;
    o += sizeof(int32_t) * A_cap;
    return o;
}

extern "C" int dcl_step_begin(const dcl_step_t* s, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!s) return fail(DCL_ERR_ARG, "null step descriptor");
    if (!s->labels || !s->predict || !s->code || !s->chunk_hist || !s->counts_dev || !s->counts_host)
        return fail(DCL_ERR_ARG, "null pointer in step descriptor");
    if (s->world <= 0 || s->rank < 0 || s->rank >= s->world || (s->world > 1 && !s->comm && !s->p2p))
        return fail(DCL_ERR_ARG, "bad sharding (world=%d rank=%d)", s->world, s->rank);
    std::lock_guard<std::mutex> lock(g_mu);
    PerDevice* pd = nullptr;
    if (int e = device_slot(pd)) return e;
    cudaStream_t st = as_stream(stream);
    const size_t table = static_cast<size_t>(512) * s->B;                       // ints per rank
    P2p* p2p = s->world > 1 ? static_cast<P2p*>(s->p2p) : nullptr;
    int32_t* counts_all = p2p ? reinterpret_cast<int32_t*>(p2p_begin(p2p, 0)) : s->counts_dev;
    int32_t* mine = counts_all + table * s->rank;
    if (int e = dcl_sample_classify(s->labels, s->predict, s->B, s->H, s->W, s->h, s->w, s->C_cls, s->code, s->chunk_hist,
                                    mine, stream))
        return e;
    if (s->zero_fill && s->zero_fill_bytes && s->side_stream && fill_placement(s->max_samples) == 0) {
        // The dense gradient buffer is cleared on the second stream from here on: while the count table travels to
        // the host, the host plans and the plan travels back, the GPU has nothing else to do.  The fill runs as a
        // few persistent blocks per SM (dcl_zero_fill_persistent), so kernels issued on `st` in the meantime always
        // find room next to it.  The buffer was allocated on `st`: the side stream first waits for st's position.
        cudaStream_t side = as_stream(s->side_stream);
        DCL_CUDA(cudaEventRecord(pd->ev_fork, st));
        DCL_CUDA(cudaStreamWaitEvent(side, pd->ev_fork, 0));
        if (int e = dcl_zero_fill(s->zero_fill, s->zero_fill_bytes, 1, s->side_stream)) return e;
        DCL_CUDA(cudaEventRecord(pd->ev_zero, side));
    }
    if (p2p) {
        if (int e = p2p_exchange(p2p, 0, table * sizeof(int32_t), false, st)) return e;
    } else if (s->world > 1) {
        if (int e = comm_all_gather(s->comm, mine, s->counts_dev, table * sizeof(int32_t), st)) return e;
    }
    DCL_CUDA(cudaMemcpyAsync(s->counts_host, counts_all, sizeof(int32_t) * table * s->world, cudaMemcpyDeviceToHost, st));
    DCL_CUDA(cudaEventRecord(pd->ev_counts, st));
    if (s->zero_fill && s->zero_fill_bytes && !s->side_stream)
        DCL_CUDA(cudaMemsetAsync(s->zero_fill, 0, s->zero_fill_bytes, st));    // still overlaps the host's plan
    return 0;
}

extern "C" int dcl_step_fwd(const dcl_step_t* s, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!s) return fail(DCL_ERR_ARG, "null step descriptor");
    if (!s->feats || !s->req_dev || !s->y_dev || !s->pix || !s->rowof || !s->plan_dev || !s->plan_host || !s->stage_host || !s->tiles ||
        !s->sqnorm || !s->colA || !s->colB || !s->rowloss || !s->loss_sum || !s->loss || !s->workspace || !s->info ||
        !s->image || !s->cls || !s->num_hard || !s->num_easy || !s->keep_hard || !s->ranks || !s->torch_rng_state)
        return fail(DCL_ERR_ARG, "null pointer in step descriptor");
    if (s->world > 1 && !s->p2p && (!s->xchg_send || !s->xchg_recv)) return fail(DCL_ERR_ARG, "sharded step without exchange buffers");
    const int cap_min = (s->max_samples + DCL_TILE_ROWS - 1) / DCL_TILE_ROWS * DCL_TILE_ROWS;
    if (s->B <= 0 || s->h <= 0 || s->w <= 0 || s->cap < cap_min || s->cap < DCL_TILE_ROWS || s->cap % DCL_TILE_ROWS)
        return fail(DCL_ERR_ARG, "bad step shape (B=%d h=%d w=%d cap=%d max_samples=%d)", s->B, s->h, s->w, s->cap, s->max_samples);
    if (s->plan_bytes < dcl_step_plan_bytes(s->B, s->world)) return fail(DCL_ERR_ARG, "plan buffer too small");
    cudaStream_t st = as_stream(stream);
    const long long t0 = now_ns();
    if (!s->begun)
        if (int e = dcl_step_begin(s, stream)) return e;
    g_step_ns[0] = now_ns() - t0;                            // classify, count table copy, zero-fill issued
    std::lock_guard<std::mutex> lock(g_mu);
    PerDevice* pd = nullptr;
    if (int e = device_slot(pd)) return e;
    const int world = s->world, rank = s->rank, hw = s->h * s->w;
    const int Bg = s->B * world;
    DCL_CUDA(cudaEventSynchronize(pd->ev_counts));           // the one unavoidable wait: the count table
    // the previous step's upload of generator blocks has left the host ring (it completed before that step's
    // k_plan ran, which precedes this step's classify on the same stream; a caller that switched streams waits here)
    DCL_CUDA(cudaEventSynchronize(pd->ev_up));
    g_step_ns[1] = now_ns() - t0;                            // + wait for the count table
    // labels outside 0..255 cannot be coded (the reference would treat them as classes of their own): refuse
    // instead of silently dropping them.  Every pixel lands in exactly one of the 512 bins otherwise.
    for (int b = 0; b < Bg; ++b) {
        long long tot = 0;
        const int32_t* c = s->counts_host + static_cast<size_t>(b) * 512;
        for (int i = 0; i < 512; ++i) tot += c[i];
        if (tot != hw) return fail(DCL_ERR_LABEL, "image %d has %lld pixels whose label is outside 0..255", b, hw - tot);
    }
    // plan buffer: [ycnt | yoff | ycls] [anchors] [yanchor]
    const size_t A_cap = static_cast<size_t>(Bg) * 256;
    uint8_t* ph = static_cast<uint8_t*>(s->plan_host);
    uint8_t* pdv = static_cast<uint8_t*>(s->plan_dev);
    size_t o_anchor = sizeof(int32_t) * (2 * static_cast<size_t>(world) + A_cap);
    o_anchor = (o_anchor + 63) / 64 * 64;
    size_t o_yanchor = o_anchor + sizeof(PlanAnchor) * static_cast<size_t>(s->B) * 256;
    o_yanchor = (o_yanchor + 63) / 64 * 64;
    DevicePlan dp{};
    dp.allow_device = s->device_plan ? 1 : 0;
    dp.ycnt = reinterpret_cast<int32_t*>(ph);
    dp.yoff = dp.ycnt + world;
    dp.ycls = dp.yoff + world;
    dp.anchors = reinterpret_cast<PlanAnchor*>(ph + o_anchor);
    dp.yanchor = reinterpret_cast<int32_t*>(ph + o_yanchor);
    int32_t* req_h = s->stage_host;
    int32_t* y_h = s->stage_host + static_cast<size_t>(s->cap) * 4;
    int32_t inf[6] = {0, 0, 0, 0, 0, 0};
    const int rc = plan_rows(s->counts_host, s->B, world, rank, s->ignore_label, s->max_samples, s->max_views,
                             s->torch_rng_state, s->state_bytes, inf, s->image, s->cls, s->num_hard, s->num_easy,
                             s->keep_hard, s->ranks, req_h, y_h, nullptr, nullptr, &dp);
    for (int i = 0; i < 5; ++i) s->info[i] = inf[i];
    s->info[5] = dp.taken;
    g_step_ns[2] = now_ns() - t0;                            // + host plan
    if (rc != 0) return rc;                                  // 1: no class qualifies, 2: the reference's unreachable branch
    const int A = inf[0], n_view = inf[1], n_pad = inf[3], n_global = inf[4];
    if (n_view <= 0) return 3;                               // max_samples // total_classes == 0 (reference fails in torch.cat)
    if (n_pad > s->cap) return fail(DCL_ERR_ARG, "plan needs %d rows per rank, capacity is %d", n_pad, s->cap);
    if (dp.taken) {
        if (int e = mirror_window(*pd, dp, st)) return e;
        g_step_ns[3] = now_ns() - t0;                        // + generator blocks queued for upload
        // one copy: [ycnt | yoff | ycls[A]] and, moved right behind it, the local anchors
        size_t o_pack = sizeof(int32_t) * (2 * static_cast<size_t>(world) + A);
        o_pack = (o_pack + 15) / 16 * 16;
        if (dp.n_local_anchors && o_pack != o_anchor)
            std::memmove(ph + o_pack, ph + o_anchor, sizeof(PlanAnchor) * dp.n_local_anchors);
        DCL_CUDA(cudaMemcpyAsync(pdv, ph, o_pack + sizeof(PlanAnchor) * dp.n_local_anchors, cudaMemcpyHostToDevice, st));
        const int32_t* d_ycnt = reinterpret_cast<const int32_t*>(pdv);
        if (int e = launch_plan(pd->d_ring, pd->ring_blocks, reinterpret_cast<const PlanAnchor*>(pdv + o_pack),
                                dp.n_local_anchors, d_ycnt + 2 * world, d_ycnt, d_ycnt + world, world, rank, n_view,
                                n_pad, s->req_dev, s->y_dev, stream))
            return e;
        DCL_CUDA(cudaEventRecord(pd->ev_used, st));
        pd->used_recorded = true;
    } else {
        DCL_CUDA(cudaMemcpyAsync(s->req_dev, req_h, sizeof(int32_t) * 4 * n_pad, cudaMemcpyHostToDevice, st));
        DCL_CUDA(cudaMemcpyAsync(s->y_dev, y_h, sizeof(int32_t) * static_cast<size_t>(world) * n_pad, cudaMemcpyHostToDevice, st));
    }
    g_step_ns[4] = now_ns() - t0;                            // + plan kernel / row copies issued
    if (int e = dcl_sample_select(s->code, s->chunk_hist, s->B, hw, s->req_dev, n_pad, s->pix, s->rowof, stream)) return e;
    const size_t tile_bytes = static_cast<size_t>(n_pad) * DCL_DIM * 2;
    P2p* p2p = world > 1 ? static_cast<P2p*>(s->p2p) : nullptr;
    // The contrast set (megabytes) goes through NCCL's all-gather when a communicator is at hand: on eight GPUs its
    // ring moves 16.8 MB in 66 us, where pushing every block to seven peers took 90 us plus the peers' skew
    // (profiles/r02w_*).  tile_exchange (diagnostics): 0 = NCCL, 1 = push kernel, 2 = peer copies by the copy engines.
    static const int tile_exchange = [] { const char* e = std::getenv("DCL_TILE_EXCHANGE"); return e ? std::atoi(e) : 0; }();
    const bool tiles_p2p = p2p && (!s->comm || tile_exchange != 0);
    uint8_t* tiles_all = tiles_p2p ? p2p_begin(p2p, 1) : static_cast<uint8_t*>(s->tiles);
    uint8_t* tl = tiles_all + tile_bytes * rank;
    if (int e = dcl_gather_tiles(s->feats, s->B, hw, s->pix, n_pad, tl, s->sqnorm + static_cast<size_t>(rank) * n_pad,
                                 s->rowof, stream))
        return e;
    if (s->zero_fill && s->zero_fill_bytes && s->side_stream && fill_placement(s->max_samples) == 1) {
        // large problems: clear the gradient buffer underneath the N x N sweeps instead (see fill_placement)
        cudaStream_t side = as_stream(s->side_stream);
        DCL_CUDA(cudaEventRecord(pd->ev_fork, st));
        DCL_CUDA(cudaStreamWaitEvent(side, pd->ev_fork, 0));
        if (int e = dcl_zero_fill(s->zero_fill, s->zero_fill_bytes, 1, s->side_stream)) return e;
        DCL_CUDA(cudaEventRecord(pd->ev_zero, side));
    }
    if (tiles_p2p) {
        if (int e = p2p_exchange(p2p, 1, tile_bytes, tile_exchange == 2, st)) return e;
    } else if (world > 1) {
        if (int e = comm_all_gather(s->comm, tl, s->tiles, tile_bytes, st)) return e;
    }
    const int nI = n_pad / DCL_TILE_ROWS, nJ = nI * world, rb0 = nI * rank;
    // similarity timing: the caller's events, else (dcl_step_sim_timing) the next slot of the library's ring
    cudaEvent_t ev_sim[4] = {static_cast<cudaEvent_t>(s->ev_fwd_begin), static_cast<cudaEvent_t>(s->ev_fwd_end),
                             static_cast<cudaEvent_t>(s->ev_bwd_begin), static_cast<cudaEvent_t>(s->ev_bwd_end)};
    if (pd->sim_on && !ev_sim[0] && !ev_sim[1] && !ev_sim[2] && !ev_sim[3]) {
        cudaEvent_t* slot = pd->sim_ev + 4 * (pd->sim_steps % kSimRing);
        for (int i = 0; i < 4; ++i) ev_sim[i] = slot[i];
        ++pd->sim_steps;
        if (!s->dF) {                                        // no backward in this step: an empty interval
            DCL_CUDA(cudaEventRecord(ev_sim[2], st));
            DCL_CUDA(cudaEventRecord(ev_sim[3], st));
        }
    }
    if (ev_sim[0]) DCL_CUDA(cudaEventRecord(ev_sim[0], st));
    // one GPU: the forward's last block writes the loss where the caller wants it and the backward is chained onto
    // the forward (no copy and no stream drain between the two)
    if (int e = contrast_fwd_ex(tiles_all, s->y_dev, s->sqnorm, nJ, rb0, nI, n_global, DCL_MODE_PIXEL, s->temperature,
                                s->base_temperature, s->workspace, s->workspace_bytes, s->colA, s->colB, s->rowloss,
                                s->loss_sum, world == 1 ? s->loss : nullptr, stream))
        return e;
    if (ev_sim[1]) DCL_CUDA(cudaEventRecord(ev_sim[1], st));
    g_step_ns[5] = now_ns() - t0;                            // + select, gather, forward issued
    if (world > 1) {
        // every rank's row constants (32 B per row: the dS_ki terms of the backward) and loss partial, one message
        const size_t msg = sizeof(float) * 4 * (2 * static_cast<size_t>(n_pad) + 1);
        if (p2p) {
            float* recv = reinterpret_cast<float*>(p2p_begin(p2p, 2));
            float* send = recv + (msg / sizeof(float)) * rank;          // the rank's message is packed in place
            if (int e = dcl_shard_pack(s->colA, s->colB, s->loss_sum, rank, n_pad, send, stream)) return e;
            if (int e = p2p_exchange(p2p, 2, msg, false, st)) return e;
            if (int e = dcl_shard_unpack(recv, world, n_pad, s->colA, s->colB, n_global, s->loss, stream)) return e;
        } else {
            if (int e = dcl_shard_pack(s->colA, s->colB, s->loss_sum, rank, n_pad, s->xchg_send, stream)) return e;
            if (int e = comm_all_gather(s->comm, s->xchg_send, s->xchg_recv, msg, st)) return e;
            if (int e = dcl_shard_unpack(s->xchg_recv, world, n_pad, s->colA, s->colB, n_global, s->loss, stream)) return e;
        }
    }
    if (s->dF) {
        if (ev_sim[2]) DCL_CUDA(cudaEventRecord(ev_sim[2], st));
        // one GPU: the backward's first kernel is a programmatic dependent of the forward's last one (with an event
        // recorded in between the launch simply orders normally)
        const bool chained = world == 1;
        if (int e = contrast_bwd_ex(tiles_all, s->y_dev, s->colA, s->colB, nJ, rb0, nI, DCL_MODE_PIXEL, s->workspace,
                                    s->workspace_bytes, s->dF, chained, stream))
            return e;
        if (ev_sim[3]) DCL_CUDA(cudaEventRecord(ev_sim[3], st));
    }
    if (s->zero_fill && s->zero_fill_bytes && s->side_stream)
        DCL_CUDA(cudaStreamWaitEvent(st, pd->ev_zero, 0));   // the cleared buffer is ordered before anything later on st
    if (dp.taken)
        if (int e = mirror_prefetch(*pd, dp)) return e;
    g_step_ns[6] = now_ns() - t0;                            // + exchange, backward, generator prefetch issued
    return 0;
}

// Measurement aid (bench.py): with timing enabled every dcl_step_fwd on the current device records events of the
// library's own around its contrast forward and backward (a ring of kSimRing steps; creating and recording events
// from Python costs the host more than a small step can hide).  Enabling resets the ring.
extern "C" int dcl_step_sim_timing(int enable) {
    if (int e = dcl_check_device()) return e;
    std::lock_guard<std::mutex> lock(g_mu);
    PerDevice* pd = nullptr;
    if (int e = device_slot(pd)) return e;
    if (enable && !pd->sim_ev) {
        pd->sim_ev = new cudaEvent_t[4 * kSimRing];
        for (int i = 0; i < 4 * kSimRing; ++i) DCL_CUDA(cudaEventCreate(&pd->sim_ev[i]));
    }
    const int old = pd->sim_on ? 1 : 0;
    pd->sim_on = enable != 0;
    if (enable) pd->sim_steps = 0;
    return old;
}

// Sums of the recorded intervals in milliseconds over the last min(steps, kSimRing) steps since timing was enabled;
// the stream must have been synchronized.  *steps = how many steps the sums cover.
extern "C" int dcl_step_sim_elapsed(double* fwd_ms, double* bwd_ms, long long* steps) {
    if (!fwd_ms || !bwd_ms || !steps) return fail(DCL_ERR_ARG, "null pointer argument");
    std::lock_guard<std::mutex> lock(g_mu);
    PerDevice* pd = nullptr;
    if (int e = device_slot(pd)) return e;
    *fwd_ms = *bwd_ms = 0.0;
    *steps = 0;
    if (!pd->sim_ev) return 0;
    const long long n = pd->sim_steps < kSimRing ? pd->sim_steps : kSimRing;
    for (long long i = 0; i < n; ++i) {
        const cudaEvent_t* slot = pd->sim_ev + 4 * ((pd->sim_steps - 1 - i) % kSimRing);
        float a = 0.f, b = 0.f;
        DCL_CUDA(cudaEventElapsedTime(&a, slot[0], slot[1]));
        DCL_CUDA(cudaEventElapsedTime(&b, slot[2], slot[3]));
        *fwd_ms += a;
        *bwd_ms += b;
    }
    *steps = n;
    return 0;
}

// Diagnostics: cumulative host nanoseconds of the last dcl_step_fwd at the end of each section: out[8] = classify etc.
// issued, + count-table wait, + host plan, + generator upload queued, + plan kernel issued, + select / gather /
// forward issued, + exchange / backward issued, -.
extern "C" int dcl_step_timing(long long* out) {
    if (!out) return fail(DCL_ERR_ARG, "null pointer argument");
    for (int i = 0; i < 8; ++i) out[i] = g_step_ns[i];
    return 0;
}

// Autograd backward of the step: d feats = upstream * scatter(dF) [+ the image-level term's pooled gradient
// broadcast over the pixels, written in the same pass: `gap_g` [gap_rows] = d loss / d pooled, one value per
// (image, channel) row of d feats; the dense tensor is then written once and needs no zero-fill].
extern "C" int dcl_step_bwd(const float* dF, const int32_t* pix, const int32_t* rowof, int n_pad, const float* grad_out,
                            float* dfeats, int B, int hw, int zero_fill, const float* gap_g, int gap_rows, void* stream) {
    if (n_pad <= 0 || n_pad % DCL_TILE_ROWS) return fail(DCL_ERR_ARG, "n_pad must be a positive multiple of %d", DCL_TILE_ROWS);
    if (gap_g) {
        if (gap_rows < B * DCL_DIM || gap_rows % DCL_DIM) return fail(DCL_ERR_ARG, "pooled gradient covers %d rows, the pixel term %d", gap_rows, B * DCL_DIM);
        if (rowof) return dcl_dense_grad(dF, rowof, B, grad_out, gap_g, dfeats, gap_rows / DCL_DIM, hw, stream);
        if (int e = dcl_gap_bwd(gap_g, gap_rows, hw, dfeats, 0, stream)) return e;
        return dcl_scatter_grad(dF, pix, n_pad, grad_out, dfeats, B, hw, 2, nullptr, stream);
    }
    return dcl_scatter_grad(dF, pix, n_pad, grad_out, dfeats, B, hw, zero_fill, rowof, stream);
}
