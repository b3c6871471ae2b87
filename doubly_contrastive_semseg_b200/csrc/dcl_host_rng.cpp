// Host-side replay of torch's CPU generator (ATen mt19937) for the hard-anchor sampler.
//
// The reference draws `torch.randperm(num_hard)[:keep_hard]` then `torch.randperm(num_easy)[:keep_easy]`
// for every (image, class) anchor on the GLOBAL CPU generator (utils/loss.py:327-330).  Calling
// torch.randperm A*2 times from Python costs milliseconds; this replays the identical stream in C
// directly on the serialized generator state (torch.get_rng_state()), producing only the prefixes
// that are used while still advancing the generator by the full n-1 draws of every permutation.
//
// Serialized state (ATen CPUGeneratorImplStateLegacy, 5056 bytes, verified at import time by
// loss._verify_host_rng against torch.randperm): u64 seed | i32 left | i32 seeded | u64 next |
// u64 state[624] | normal-distribution cache.
// randperm(n) for n < 2^32/20: r = arange(n); for i < n-1: z = u32() % (n-i); swap(r[i], r[i+z]).
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include "dcl_common.cuh"

namespace {

constexpr int kN = 624, kM = 397;

struct Mt {
    uint32_t s[kN];
    int left;
    uint32_t next;
};

inline __attribute__((always_inline)) uint32_t mix(uint32_t u, uint32_t v) {
    uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((0u - (v & 1u)) & 0x9908b0dfu);       // branch-free: vectorises
}

// State regeneration.  Nearly all of the sampler's host time is spent here (one regeneration per 624 draws, and
// every pixel of the batch is one draw), so the loops are written to vectorise: element i reads i+1 and i+397
// (not yet rewritten) in the first loop and i-227 (rewritten 227 iterations earlier) in the second, i.e. no
// dependence closer than the vector width.
inline __attribute__((always_inline)) void twist_body(uint32_t* __restrict__ s) {
#pragma GCC ivdep
    for (int i = 0; i < kN - kM; ++i) s[i] = s[i + kM] ^ mix(s[i], s[i + 1]);
    // split the second range so that the vectorised part never reads an element written inside the same vector
#pragma GCC ivdep
    for (int i = kN - kM; i < kN - 1; ++i) s[i] = s[i + kM - kN] ^ mix(s[i], s[i + 1]);
    s[kN - 1] = s[kM - 1] ^ mix(s[kN - 1], s[0]);
}
__attribute__((target("avx512f"))) void twist_avx512(uint32_t* s) { twist_body(s); }
__attribute__((target("avx2"))) void twist_avx2(uint32_t* s) { twist_body(s); }
void twist_generic(uint32_t* s) { twist_body(s); }

void twist(Mt& m) {
    static const int isa = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    if (isa == 2) twist_avx512(m.s);
    else if (isa == 1) twist_avx2(m.s);
    else twist_generic(m.s);
    m.left = kN;
    m.next = 0;
}

inline uint32_t draw(Mt& m) {
    if (--m.left == 0) twist(m);
    uint32_t y = m.s[m.next++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// advance the generator by `cnt` draws without producing outputs
inline void skip(Mt& m, int64_t cnt) {
    while (cnt > 0) {
        if (m.left <= 1) {          // next draw would regenerate
            --m.left;
            twist(m);
            m.next++;
            --cnt;
            continue;
        }
        int64_t take = m.left - 1;  // draws available before the next regenerate
        if (take > cnt) take = cnt;
        m.left -= static_cast<int>(take);
        m.next += static_cast<uint32_t>(take);
        cnt -= take;
    }
}

// Sparse Fisher-Yates state: positions [0, k) live in a dense array, the (at most k) touched positions >= k in a
// small open-addressing map with key and value side by side.
struct Sparse {
    struct Slot { int32_t key, val; };
    std::vector<Slot> tab;
    std::vector<uint32_t> used;        // slots filled by the current permutation (cleared on reset)
    std::vector<int32_t> front;
    uint32_t mask = 0;
    void reset(int64_t k) {
        size_t cap = 16;
        while (cap < static_cast<size_t>(2 * k + 4)) cap <<= 1;
        if (tab.size() != cap) {
            tab.assign(cap, Slot{-1, 0});
            mask = static_cast<uint32_t>(cap - 1);
        } else {
            for (uint32_t h : used) tab[h].key = -1;
        }
        used.clear();
        front.resize(static_cast<size_t>(k));
        for (int64_t i = 0; i < k; ++i) front[i] = static_cast<int32_t>(i);
    }
    uint32_t slot(int32_t k) const {
        uint32_t h = (static_cast<uint32_t>(k) * 0x9E3779B1u >> 12) & mask;
        while (tab[h].key != -1 && tab[h].key != k) h = (h + 1) & mask;
        return h;
    }
};

void randperm_prefix(Mt& m, int64_t n, int64_t k, int64_t* out, Sparse& sp) {
    if (n <= 1) {
        for (int64_t i = 0; i < k && i < n; ++i) out[i] = i;
        return;
    }
    if (k > n) k = n;
    sp.reset(k);
    int64_t drawn = 0;
    for (int64_t i = 0; i < k; ++i) {
        int64_t j = i;
        if (i < n - 1) {
            j = i + static_cast<int64_t>(draw(m) % static_cast<uint32_t>(n - i));   // n < 2^31: 32-bit modulo, same value
            ++drawn;
        }
        const int32_t vi = sp.front[i];
        if (j < k) {
            out[i] = sp.front[j];
            sp.front[j] = vi;
        } else {
            const uint32_t h = sp.slot(static_cast<int32_t>(j));
            Sparse::Slot& t = sp.tab[h];
            if (t.key == static_cast<int32_t>(j)) {
                out[i] = t.val;
            } else {
                out[i] = j;
                t.key = static_cast<int32_t>(j);
                sp.used.push_back(h);
            }
            t.val = vi;
        }
    }
    skip(m, (n - 1) - drawn);
}

int load_state(void* torch_rng_state, size_t state_bytes, Mt& m) {
    if (!torch_rng_state || state_bytes < 24 + 8 * static_cast<size_t>(kN))
        return dcl::fail(DCL_ERR_ARG, "rng state buffer too small (%zu bytes)", state_bytes);
    const uint8_t* raw = static_cast<const uint8_t*>(torch_rng_state);
    int32_t left;
    uint64_t next;
    std::memcpy(&left, raw + 8, 4);
    std::memcpy(&next, raw + 16, 8);
    const uint64_t* st = reinterpret_cast<const uint64_t*>(raw + 24);
    for (int i = 0; i < kN; ++i) m.s[i] = static_cast<uint32_t>(st[i]);
    m.left = left;
    m.next = static_cast<uint32_t>(next);
    if (m.left < 0 || m.left > kN || m.next > static_cast<uint32_t>(kN))
        return dcl::fail(DCL_ERR_ARG, "unexpected generator state (left=%d next=%u)", m.left, m.next);
    return 0;
}
void store_state(void* torch_rng_state, const Mt& m) {
    uint8_t* raw = static_cast<uint8_t*>(torch_rng_state);
    const int32_t left = m.left;
    const uint64_t next = m.next;
    std::memcpy(raw + 8, &left, 4);
    std::memcpy(raw + 16, &next, 8);
    uint64_t* sw = reinterpret_cast<uint64_t*>(raw + 24);
    for (int i = 0; i < kN; ++i) sw[i] = m.s[i];
}

}  // namespace

// Whole host side of the sampler in one call (loss.py:264-337 minus the tensor indexing): class list, n_view,
// split rule, the randperm draws and the class-sorted device row layout.  Shared by the single-process entry point
// and the sharded one: `world` ranks own `Bl` consecutive images each; every rank replays the SAME generator stream
// over the global batch, but only draws the permutations of its own anchors (the others just advance the state).
static int plan_rows(const int32_t* counts, int Bl, int world, int rank, int ignore_label, int max_samples,
                     int max_views, void* torch_rng_state, size_t state_bytes, int32_t* info, int64_t* image,
                     int64_t* cls, int64_t* num_hard, int64_t* num_easy, int64_t* keep_hard, int64_t* ranks,
                     int32_t* req, int32_t* y_all, int64_t* ref_row, int64_t* anchor) {
    if (!counts || !info || Bl <= 0 || world <= 0 || rank < 0 || rank >= world) return dcl::fail(DCL_ERR_ARG, "bad argument");
    const int B = Bl * world;
    int A = 0;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < 256; ++c) {
            const int64_t nh = counts[(b * 256 + c) * 2], ne = counts[(b * 256 + c) * 2 + 1];
            if (c == ignore_label || nh + ne <= max_views) continue;          // loss.py:281-282
            image[A] = b; cls[A] = c; num_hard[A] = nh; num_easy[A] = ne;
            ++A;
        }
    for (int i = 0; i < 6; ++i) info[i] = 0;
    info[0] = A;
    if (A == 0) return 1;
    int n_view = max_samples / A;                                              // loss.py:290-291
    if (n_view > max_views) n_view = max_views;
    info[1] = n_view;
    const double half = n_view / 2.0;                                          // true division, loss.py:314
    for (int a = 0; a < A; ++a) {
        const int64_t nh = num_hard[a], ne = num_easy[a];
        if (nh >= half && ne >= half) keep_hard[a] = n_view / 2;
        else if (nh >= half) keep_hard[a] = n_view - ne;
        else if (ne >= half) keep_hard[a] = nh;
        else {
            info[0] = static_cast<int32_t>(nh); info[1] = static_cast<int32_t>(ne); info[2] = n_view;
            return 2;
        }
    }
    const int lo = rank * Bl, hi = lo + Bl;
    if (n_view > 0) {
        Mt m;
        if (int e = load_state(torch_rng_state, state_bytes, m)) return e;
        Sparse sp;
        for (int a = 0; a < A; ++a) {
            const int64_t kh = keep_hard[a], ke = n_view - kh;
            if (kh < 0 || ke < 0 || kh > num_hard[a] || ke > num_easy[a] || num_hard[a] >= 214748364 ||
                num_easy[a] >= 214748364)
                return dcl::fail(DCL_ERR_ARG, "anchor %d: keep (%lld,%lld) exceeds counts (%lld,%lld)", a,
                                 (long long)kh, (long long)ke, (long long)num_hard[a], (long long)num_easy[a]);
            if (image[a] >= lo && image[a] < hi) {
                randperm_prefix(m, num_hard[a], kh, ranks + static_cast<size_t>(a) * n_view, sp);
                randperm_prefix(m, num_easy[a], ke, ranks + static_cast<size_t>(a) * n_view + kh, sp);
            } else {
                // another rank's anchor: randperm(n) consumes n - 1 draws (none for n <= 1)
                if (num_hard[a] > 1) skip(m, num_hard[a] - 1);
                if (num_easy[a] > 1) skip(m, num_easy[a] - 1);
            }
        }
        store_state(torch_rng_state, m);
    }
    // rows per rank -> common padded block size
    std::vector<int> per_rank(world, 0);
    for (int a = 0; a < A; ++a) per_rank[image[a] / Bl] += n_view;
    int n_max = 0, n_global = 0;
    for (int r = 0; r < world; ++r) { n_max = per_rank[r] > n_max ? per_rank[r] : n_max; n_global += per_rank[r]; }
    int n_pad = (n_max + 127) / 128 * 128;
    if (n_pad < 128) n_pad = 128;
    info[2] = per_rank[rank]; info[3] = n_pad; info[4] = n_global;
    // rows of every rank block: anchors of that rank stably sorted by class (counting sort), views contiguous;
    // labels for all blocks (y_all), requests / bookkeeping for the local block only
    std::vector<int> start(257), order(A);
    for (int r = 0; r < world; ++r) {
        std::fill(start.begin(), start.end(), 0);
        const int rlo = r * Bl, rhi = rlo + Bl;
        int cnt = 0;
        for (int a = 0; a < A; ++a)
            if (image[a] >= rlo && image[a] < rhi) { ++start[cls[a] + 1]; ++cnt; }
        for (int c = 0; c < 256; ++c) start[c + 1] += start[c];
        for (int a = 0; a < A; ++a)
            if (image[a] >= rlo && image[a] < rhi) order[start[cls[a]]++] = a;
        int32_t* yb = y_all + static_cast<size_t>(r) * n_pad;
        int row = 0;
        for (int o = 0; o < cnt; ++o) {
            const int a = order[o];
            const int64_t* rk = ranks + static_cast<size_t>(a) * n_view;
            for (int v = 0; v < n_view; ++v, ++row) {
                yb[row] = static_cast<int32_t>(cls[a]);
                if (r == rank) {
                    req[row * 4 + 0] = static_cast<int32_t>(image[a] - lo);
                    req[row * 4 + 1] = static_cast<int32_t>(cls[a]);
                    req[row * 4 + 2] = v >= keep_hard[a];
                    req[row * 4 + 3] = static_cast<int32_t>(rk[v]);
                    ref_row[row] = static_cast<int64_t>(v) * A + a;
                    anchor[row] = a;
                }
            }
        }
        for (; row < n_pad; ++row) {
            yb[row] = -1;
            if (r == rank) {
                req[row * 4 + 0] = req[row * 4 + 1] = req[row * 4 + 2] = req[row * 4 + 3] = -1;
                ref_row[row] = -1;
                anchor[row] = -1;
            }
        }
    }
    return 0;
}

//   counts [B][256][2] i32 : pixels per (image, label, hard|easy) from dcl_sample_classify
//   info out [6] i32      : A, n_view, n (valid local rows), n_pad, n_global, -
//   anchor arrays out     : image, cls, num_hard, num_easy, keep_hard [A <= B*256] i64; ranks [A*n_view] i64
//   row arrays out        : req [n_pad*4] i32 (image, label, easy, rank; -1 padding), y [n_pad] i32,
//                           ref_row [n_pad] i64 (v*A + a in the reference's order), anchor [n_pad] i64
// Returns 0, 1 when no class qualifies (the reference's `return None, None`, loss.py:287-288), 2 when the
// split rule hits the reference's "this shoud be never touched" branch (info[0..2] = num_hard, num_easy, n_view),
// or a negative dcl_status.
extern "C" int dcl_host_plan_rows(const int32_t* counts, int B, int ignore_label, int max_samples, int max_views,
                                  void* torch_rng_state, size_t state_bytes, int32_t* info, int64_t* image,
                                  int64_t* cls, int64_t* num_hard, int64_t* num_easy, int64_t* keep_hard,
                                  int64_t* ranks, int32_t* req, int32_t* y, int64_t* ref_row, int64_t* anchor) {
    int32_t inf[6];
    const int rc = plan_rows(counts, B, 1, 0, ignore_label, max_samples, max_views, torch_rng_state, state_bytes, inf,
                             image, cls, num_hard, num_easy, keep_hard, ranks, req, y, ref_row, anchor);
    if (info) for (int i = 0; i < 4; ++i) info[i] = inf[i];
    return rc;
}

// Sharded form: counts [world*Bl][256][2] all-gathered, rank-major image order.  y_all [world*n_pad] receives the
// labels of EVERY rank's row block (so they need not be exchanged); req / ref_row / anchor describe the local block;
// `ranks` is filled for the local anchors only.  info [6] as above.
extern "C" int dcl_host_plan_rows_sharded(const int32_t* counts, int Bl, int world, int rank, int ignore_label,
                                          int max_samples, int max_views, void* torch_rng_state, size_t state_bytes,
                                          int32_t* info, int64_t* image, int64_t* cls, int64_t* num_hard,
                                          int64_t* num_easy, int64_t* keep_hard, int64_t* ranks, int32_t* req,
                                          int32_t* y_all, int64_t* ref_row, int64_t* anchor) {
    return plan_rows(counts, Bl, world, rank, ignore_label, max_samples, max_views, torch_rng_state, state_bytes, info,
                     image, cls, num_hard, num_easy, keep_hard, ranks, req, y_all, ref_row, anchor);
}

extern "C" int dcl_host_sample_ranks(void* torch_rng_state, size_t state_bytes, int A, int n_view,
                                     const int64_t* num_hard, const int64_t* num_easy,
                                     const int64_t* keep_hard, int64_t* ranks) {
    if (A < 0 || n_view < 0 || (A > 0 && (!num_hard || !num_easy || !keep_hard || !ranks)))
        return dcl::fail(DCL_ERR_ARG, "bad argument");
    Mt m;
    if (int e = load_state(torch_rng_state, state_bytes, m)) return e;
    Sparse sp;
    for (int a = 0; a < A; ++a) {
        const int64_t kh = keep_hard[a], ke = n_view - kh;
        if (kh < 0 || ke < 0 || kh > num_hard[a] || ke > num_easy[a] || num_hard[a] >= 214748364 ||
            num_easy[a] >= 214748364)
            return dcl::fail(DCL_ERR_ARG, "anchor %d: keep (%lld,%lld) exceeds counts (%lld,%lld)", a,
                             (long long)kh, (long long)ke, (long long)num_hard[a], (long long)num_easy[a]);
        randperm_prefix(m, num_hard[a], kh, ranks + static_cast<size_t>(a) * n_view, sp);
        randperm_prefix(m, num_easy[a], ke, ranks + static_cast<size_t>(a) * n_view + kh, sp);
    }
    store_state(torch_rng_state, m);
    return 0;
}
