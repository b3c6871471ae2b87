// Device half of the hard-anchor sampler's plan (reference utils/loss.py:297-335): the kept prefixes of the
// torch.randperm calls, replayed on the GPU from a mirror of the CPU generator's output stream, and the
// class-sorted row layout (labels of every rank block, row requests of the local block).
//
// torch.randperm(n) on the CPU generator is Fisher-Yates:  r = arange(n); for i < n-1: z = mt19937() % (n - i);
// swap(r[i], r[i + z]).  Entry i is final after step i and n - 1 draws are consumed whatever the outcome, so the
// host can place every permutation in the generator's output stream without drawing it (dcl_host_rng.cpp,
// device mode of plan_rows) and this kernel only replays the first `keep` steps of each one.  The stream itself
// (regenerated mt19937 state blocks, produced ahead of time by the look-ahead thread) is uploaded by dcl_step.cu;
// the tempering of a state word into an output happens here.  Results are bit-identical with the host replay.
#include "dcl_common.cuh"
#include "dcl_plan.h"

namespace dcl {

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// Blocks [0, n_perm): one warp per permutation (anchor a = block / 2, hard = even, easy = odd).
//   phase 1 (32 lanes): j_i = i + temper(stream[g + i]) % (n - i) for the kept steps
//   phase 2 (lane 0)  : sparse Fisher-Yates - positions < k in a dense array, the touched positions >= k in an
//                       open-addressing table (at most one new entry per step), all in shared memory
//   phase 3 (32 lanes): row requests (image, label, easy, rank) for dcl_sample_select
// Blocks [n_perm, ..): labels of every rank's row block (-1 = padding) and the padding requests of the local block.
// Dynamic shared memory: 3 * n_view ints + tab_slots * 2 ints.
__global__ void __launch_bounds__(256)
k_plan(const uint32_t* __restrict__ ring, unsigned long long ring_blocks, const PlanAnchor* __restrict__ anchors,
       int n_perm, const int32_t* __restrict__ ycls, const int32_t* __restrict__ ycnt,
       const int32_t* __restrict__ yoff, int world, int rank, int n_view, int n_pad, int tab_slots,
       int4* __restrict__ req, int32_t* __restrict__ y_all) {
    extern __shared__ int32_t sm[];
    if (static_cast<int>(blockIdx.x) < n_perm) {
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x;
        const PlanAnchor pa = anchors[blockIdx.x >> 1];
        const bool easy = blockIdx.x & 1;
        const int n = easy ? pa.num_easy : pa.num_hard;
        const int k = easy ? pa.keep_easy : pa.keep_hard;
        const unsigned long long g = easy ? pa.g_easy : pa.g_hard;
        const int row = pa.row0 + (easy ? pa.keep_hard : 0);
        if (k <= 0) return;
        int32_t* jbuf = sm;
        int32_t* front = sm + n_view;
        int32_t* out = sm + 2 * n_view;
        int2* tab = reinterpret_cast<int2*>(sm + 3 * n_view + ((3 * n_view) & 1));
        for (int i = lane; i < k; i += 32) {
            int j = i;
            if (i < n - 1) {
                const unsigned long long gi = g + static_cast<unsigned long long>(i);
                const unsigned long long blk = gi / kMtWords;
                const uint32_t word = static_cast<uint32_t>(gi - blk * kMtWords);
                const uint32_t x = mt_temper(__ldg(ring + (blk % ring_blocks) * kMtWords + word));
                j = i + static_cast<int>(x % static_cast<uint32_t>(n - i));
            }
            jbuf[i] = j;
            front[i] = i;
        }
        for (int t = lane; t < tab_slots; t += 32) tab[t] = make_int2(-1, 0);
        __syncwarp();
        if (lane == 0) {
            // The chain is latency-bound (shared-memory round trips), so the operands of step i + 1 - its target j, the
            // value at position i + 1 and the table slot j hashes to - are fetched while step i completes, and patched
            // in the rare case that step i has just written one of them.
            const uint32_t mask = static_cast<uint32_t>(tab_slots - 1);
            auto slot_of = [&](int j) { return (static_cast<uint32_t>(j) * 0x9E3779B1u >> 12) & mask; };
            int j_n = jbuf[0], v_n = front[0];
            uint32_t h_n = j_n >= k ? slot_of(j_n) : 0u;
            int2 t_n = j_n >= k ? tab[h_n] : make_int2(-1, 0);
            for (int i = 0; i < k; ++i) {
                const int j = j_n, vi = v_n;
                uint32_t h = h_n;
                int2 t = t_n;
                const bool more = i + 1 < k;
                if (more) {
                    j_n = jbuf[i + 1];
                    v_n = front[i + 1];
                    h_n = j_n >= k ? slot_of(j_n) : 0u;
                    t_n = j_n >= k ? tab[h_n] : make_int2(-1, 0);
                }
                int o;
                if (j < k) {
                    o = (j == i) ? vi : front[j];
                    front[j] = vi;
                    if (more && j == i + 1) v_n = vi;                 // the value just moved to the next position
                } else {
                    while (t.x != -1 && t.x != j) {
                        h = (h + 1) & mask;
                        t = tab[h];
                    }
                    o = (t.x == j) ? t.y : j;
                    tab[h] = make_int2(j, vi);
                    if (more && j_n >= k && h_n == h) t_n = tab[h_n];   // the next probe starts on the slot just written
                }
                out[i] = o;
            }
        }
        __syncwarp();
        for (int i = lane; i < k; i += 32) req[row + i] = make_int4(pa.image, pa.cls, easy ? 1 : 0, out[i]);
        return;
    }
    // ---- labels / padding
    const int per = blockDim.x;
    const long long idx = static_cast<long long>(blockIdx.x - n_perm) * per + threadIdx.x;
    if (idx >= static_cast<long long>(world) * n_pad) return;
    const int r = static_cast<int>(idx / n_pad), i = static_cast<int>(idx - static_cast<long long>(r) * n_pad);
    const int o = i / n_view;
    const bool valid = o < ycnt[r];
    y_all[idx] = valid ? ycls[yoff[r] + o] : -1;
    if (!valid && r == rank) req[i] = make_int4(-1, -1, -1, -1);
}

int launch_plan(const uint32_t* d_ring, uint64_t ring_blocks, const PlanAnchor* anchors, int n_local_anchors,
                const int32_t* ycls, const int32_t* ycnt, const int32_t* yoff, int world, int rank, int n_view,
                int n_pad, int32_t* req, int32_t* y_all, void* stream) {
    if (!d_ring || !anchors || !ycls || !ycnt || !yoff || !req || !y_all) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n_view <= 0 || n_view > kMaxDeviceViews || n_pad <= 0 || world <= 0) return fail(DCL_ERR_ARG, "bad plan shape");
    int tab_slots = 16;
    while (tab_slots < 2 * n_view + 4) tab_slots <<= 1;
    const size_t smem = sizeof(int32_t) * (3 * static_cast<size_t>(n_view) + 1 + 2 * static_cast<size_t>(tab_slots));
    const int n_perm = 2 * n_local_anchors;
    const long long fill = static_cast<long long>(world) * n_pad;
    const unsigned grid = static_cast<unsigned>(n_perm + (fill + 255) / 256);
    k_plan<<<grid, 256, smem, as_stream(stream)>>>(d_ring, ring_blocks, anchors, n_perm, ycls, ycnt, yoff, world, rank,
                                                   n_view, n_pad, tab_slots, reinterpret_cast<int4*>(req), y_all);
    DCL_LAUNCH_CHECK("k_plan");
    return 0;
}

}  // namespace dcl
