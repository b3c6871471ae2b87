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

// Blocks [0, n_perm): one 128-thread block per permutation (anchor a = block / 2, hard = even, easy = odd), no serial
// chain.  With j_i = i + draw_i % (n - i) the target of step i, Fisher-Yates' swap sequence means
//   val(p, t) = value at position p before step t = V(q) for q = the latest step < t with j_q == p, else p
//   V(i)      = val(i, i)   (the value step i moves to position j_i),     out[i] = val(j_i, i)
// and since a writer q of position i has q < i, V(i) is the root of the chain i -> q -> ... (pointer jumping).  The
// "latest earlier step with target p" queries are predecessor look-ups in the steps sorted by (target, step):
//   phase 1: targets from the mirrored generator stream (tempering here), keys (j_i << 32 | i)
//   phase 2: bitonic sort of the keys in shared memory; position of every step in the sorted order
//   phase 3: parent(i) by binary search for (i, i), log2(k) rounds of pointer jumping
//   phase 4: out[i] from the sorted neighbour of (j_i, i); row requests (image, label, easy, rank) for select
// Blocks [n_perm, ..): labels of every rank's row block (-1 = padding) and the padding requests of the local block.
// Dynamic shared memory: 20 bytes x (n_view rounded up to a power of two).
constexpr int kPlanThreads = 128;
__global__ void __launch_bounds__(kPlanThreads)
k_plan(const uint32_t* __restrict__ ring, unsigned long long ring_blocks, const PlanAnchor* __restrict__ anchors,
       int n_perm, const int32_t* __restrict__ ycls, const int32_t* __restrict__ ycnt,
       const int32_t* __restrict__ yoff, int world, int rank, int n_view, int n_pad, int k2max,
       int4* __restrict__ req, int32_t* __restrict__ y_all) {
    extern __shared__ unsigned long long sm64[];
    const int tid = threadIdx.x;
    if (static_cast<int>(blockIdx.x) < n_perm) {
        const PlanAnchor pa = anchors[blockIdx.x >> 1];
        const bool easy = blockIdx.x & 1;
        const int n = easy ? pa.num_easy : pa.num_hard;
        const int k = easy ? pa.keep_easy : pa.keep_hard;
        const unsigned long long g = easy ? pa.g_easy : pa.g_hard;
        const int row = pa.row0 + (easy ? pa.keep_hard : 0);
        if (k <= 0) return;                                         // block-uniform
        int K2 = 32;
        while (K2 < k) K2 <<= 1;
        unsigned long long* key = sm64;                             // [k2max]
        int* pos = reinterpret_cast<int*>(sm64 + k2max);            // [k2max] sorted index of step i
        int* P = pos + k2max;                                       // [k2max] parent -> root
        int* jb = P + k2max;                                        // [k2max] target of step i
        for (int i = tid; i < K2; i += kPlanThreads) {
            unsigned long long kv = ~0ull;
            if (i < k) {
                int j = i;
                if (i < n - 1) {
                    const unsigned long long gi = g + static_cast<unsigned long long>(i);
                    const unsigned long long blk = gi / kMtWords;
                    const uint32_t word = static_cast<uint32_t>(gi - blk * kMtWords);
                    const uint32_t x = mt_temper(__ldg(ring + (blk % ring_blocks) * kMtWords + word));
                    j = i + static_cast<int>(x % static_cast<uint32_t>(n - i));
                }
                jb[i] = j;
                kv = (static_cast<unsigned long long>(static_cast<uint32_t>(j)) << 32) | static_cast<uint32_t>(i);
            }
            key[i] = kv;
        }
        __syncthreads();
        for (int size = 2; size <= K2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (K2 >> 1); t += kPlanThreads) {
                    const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const unsigned long long a = key[lo], b = key[hi];
                    if ((a > b) == up) { key[lo] = b; key[hi] = a; }
                }
                __syncthreads();
            }
        for (int s = tid; s < k; s += kPlanThreads) pos[static_cast<int>(key[s] & 0xffffffffu)] = s;
        for (int i = tid; i < k; i += kPlanThreads) {
            // latest earlier step that targeted position i: the predecessor of (i, i) in the sorted keys
            const unsigned long long want = (static_cast<unsigned long long>(static_cast<uint32_t>(i)) << 32) | static_cast<uint32_t>(i);
            int lo = 0, hi = k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (key[mid] < want) lo = mid + 1; else hi = mid;
            }
            int par = i;
            if (lo > 0 && static_cast<int>(key[lo - 1] >> 32) == i) par = static_cast<int>(key[lo - 1] & 0xffffffffu);
            P[i] = par;
        }
        __syncthreads();
        for (int span = 1; span < k; span <<= 1) {                  // chains only run towards smaller steps
            int nxt[(kMaxDeviceViews + kPlanThreads - 1) / kPlanThreads];
            int c = 0;
            for (int i = tid; i < k; i += kPlanThreads) nxt[c++] = P[P[i]];
            __syncthreads();
            c = 0;
            for (int i = tid; i < k; i += kPlanThreads) P[i] = nxt[c++];
            __syncthreads();
        }
        for (int i = tid; i < k; i += kPlanThreads) {
            const int j = jb[i], s = pos[i];
            int o = j;
            if (s > 0 && static_cast<int>(key[s - 1] >> 32) == j) o = P[static_cast<int>(key[s - 1] & 0xffffffffu)];
            req[row + i] = make_int4(pa.image, pa.cls, easy ? 1 : 0, o);
        }
        return;
    }
    // ---- labels / padding
    const long long idx = static_cast<long long>(blockIdx.x - n_perm) * kPlanThreads + tid;
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
    int k2max = 32;
    while (k2max < n_view) k2max <<= 1;
    const size_t smem = static_cast<size_t>(k2max) * (sizeof(unsigned long long) + 3 * sizeof(int32_t));
    const int n_perm = 2 * n_local_anchors;
    const long long fill = static_cast<long long>(world) * n_pad;
    const unsigned grid = static_cast<unsigned>(n_perm + (fill + kPlanThreads - 1) / kPlanThreads);
    k_plan<<<grid, kPlanThreads, smem, as_stream(stream)>>>(d_ring, ring_blocks, anchors, n_perm, ycls, ycnt, yoff, world,
                                                            rank, n_view, n_pad, k2max, reinterpret_cast<int4*>(req), y_all);
    DCL_LAUNCH_CHECK("k_plan");
    return 0;
}

}  // namespace dcl
