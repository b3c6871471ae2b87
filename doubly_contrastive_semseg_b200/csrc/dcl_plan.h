// Internal interface between the host half of the sampler (dcl_host_rng.cpp), the device permutation kernel
// (dcl_plan.cu) and the one-call steps (dcl_step.cu).  Not part of the C ABI.
#pragma once
#include <cstddef>
#include <cstdint>

namespace dcl {

constexpr int kMtWords = 624;                 // words of one regenerated mt19937 state block
constexpr int kMaxDeviceViews = 1024;         // largest n_view k_plan replays on the device (shared memory per warp)

// One local anchor = one (image, class) pair of this rank: its two torch.randperm calls (hard first, then easy,
// utils/loss.py:327-330) as positions in the generator's output stream.  48 bytes, read by k_plan.
struct PlanAnchor {
    uint64_t g_hard;                          // draw index (counted from word 0 of stream block 0) of the first draw
    uint64_t g_easy;
    int32_t num_hard, num_easy, keep_hard, keep_easy;
    int32_t row0;                             // first local row; the n_view views are contiguous, hard then easy
    int32_t image;                            // image index relative to the rank's first image
    int32_t cls;
    int32_t reserved;
};

// Device-plan request / result of plan_rows.  When `taken` comes back 1 the host did NOT draw the permutations:
// it only advanced the generator, and the descriptors below tell k_plan where every permutation starts in the
// look-ahead stream.  taken == 0: the host drew everything itself (generator touched since the last plan, stream
// not running yet, plan longer than the ring, n_view too large) and req / y hold the finished rows.
struct DevicePlan {
    // in
    int allow_device;                         // 0: fill the class-sorted order only, draw on the host
    PlanAnchor* anchors;                      // [>= Bl*256] host (pinned)
    int32_t* ycls;                            // [>= world*Bl*256] class of the o-th class-sorted anchor of rank r at ycls[yoff[r] + o]
    int32_t* yanchor;                         // same indexing: its anchor id a (reference order)
    int32_t* ycnt;                            // [world] anchors per rank
    int32_t* yoff;                            // [world]
    // out
    int taken;
    int n_local_anchors;
    uint64_t epoch;                           // restart count of the stream (the device mirror resets when it changes)
    uint64_t first_block, last_block;         // stream blocks the permutations read
    uint64_t plan_first_block, plan_end_block;   // blocks where the whole plan (every rank's draws) starts and ends
    const uint32_t* host_ring;                // look-ahead ring: block b lives at host_ring + (b % ring_blocks) * 624
    uint64_t ring_blocks;
    uint64_t produced;                        // blocks [.., produced) exist on the host right now (prefetch hint)
};

int plan_rows(const int32_t* counts, int Bl, int world, int rank, int ignore_label, int max_samples, int max_views,
              void* torch_rng_state, size_t state_bytes, int32_t* info, int64_t* image, int64_t* cls,
              int64_t* num_hard, int64_t* num_easy, int64_t* keep_hard, int64_t* ranks, int32_t* req, int32_t* y_all,
              int64_t* ref_row, int64_t* anchor, DevicePlan* dp);

// Launch k_plan on `stream`: permutations of the local anchors -> req rows; labels of every rank block -> y_all.
//   d_ring [ring_blocks*624] raw (untempered) state words; anchors/ycls/ycnt/yoff: DEVICE copies of the host arrays
int launch_plan(const uint32_t* d_ring, uint64_t ring_blocks, const PlanAnchor* anchors, int n_local_anchors,
                const int32_t* ycls, const int32_t* ycnt, const int32_t* yoff, int world, int rank, int n_view,
                int n_pad, int32_t* req, int32_t* y_all, void* stream);

}  // namespace dcl
