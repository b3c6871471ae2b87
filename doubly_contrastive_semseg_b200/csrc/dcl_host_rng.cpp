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
#include <vector>
#include "dcl_common.cuh"

namespace {

constexpr int kN = 624, kM = 397;

struct Mt {
    uint32_t s[kN];
    int left;
    uint32_t next;
};

inline uint32_t mix(uint32_t u, uint32_t v) {
    uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

void twist(Mt& m) {
    uint32_t* s = m.s;
    for (int i = 0; i < kN - kM; ++i) s[i] = s[i + kM] ^ mix(s[i], s[i + 1]);
    for (int i = kN - kM; i < kN - 1; ++i) s[i] = s[i + kM - kN] ^ mix(s[i], s[i + 1]);
    s[kN - 1] = s[kM - 1] ^ mix(s[kN - 1], s[0]);
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

// tiny open-addressing map for the sparse Fisher-Yates state (only touched positions are stored)
struct Sparse {
    std::vector<int64_t> key, val;
    uint64_t mask;
    void reset(int k) {
        size_t cap = 16;
        while (cap < static_cast<size_t>(4 * k + 4)) cap <<= 1;
        key.assign(cap, -1);
        val.resize(cap);
        mask = cap - 1;
    }
    size_t slot(int64_t k) const {
        size_t h = (static_cast<uint64_t>(k) * 0x9E3779B97F4A7C15ull) & mask;
        while (key[h] != -1 && key[h] != k) h = (h + 1) & mask;
        return h;
    }
    int64_t get(int64_t k) const {
        size_t h = slot(k);
        return key[h] == k ? val[h] : k;
    }
    void set(int64_t k, int64_t v) {
        size_t h = slot(k);
        key[h] = k;
        val[h] = v;
    }
};

void randperm_prefix(Mt& m, int64_t n, int64_t k, int64_t* out, Sparse& sp) {
    if (n <= 1) {
        for (int64_t i = 0; i < k && i < n; ++i) out[i] = i;
        return;
    }
    if (k > n) k = n;
    sp.reset(static_cast<int>(k));
    int64_t drawn = 0;
    for (int64_t i = 0; i < k; ++i) {
        int64_t j = i;
        if (i < n - 1) {
            j = i + static_cast<int64_t>(draw(m) % static_cast<uint64_t>(n - i));
            ++drawn;
        }
        const int64_t vi = sp.get(i), vj = sp.get(j);
        sp.set(i, vj);
        sp.set(j, vi);
        out[i] = vj;
    }
    skip(m, (n - 1) - drawn);
}

}  // namespace

extern "C" int dcl_host_sample_ranks(void* torch_rng_state, size_t state_bytes, int A, int n_view,
                                     const int64_t* num_hard, const int64_t* num_easy,
                                     const int64_t* keep_hard, int64_t* ranks) {
    if (!torch_rng_state || state_bytes < 24 + 8 * static_cast<size_t>(kN))
        return dcl::fail(DCL_ERR_ARG, "rng state buffer too small (%zu bytes)", state_bytes);
    if (A < 0 || n_view < 0 || (A > 0 && (!num_hard || !num_easy || !keep_hard || !ranks)))
        return dcl::fail(DCL_ERR_ARG, "bad argument");
    uint8_t* raw = static_cast<uint8_t*>(torch_rng_state);
    Mt m;
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
    left = m.left;
    next = m.next;
    std::memcpy(raw + 8, &left, 4);
    std::memcpy(raw + 16, &next, 8);
    uint64_t* sw = reinterpret_cast<uint64_t*>(raw + 24);
    for (int i = 0; i < kN; ++i) sw[i] = m.s[i];
    return 0;
}
