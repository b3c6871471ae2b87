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
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <immintrin.h>
#include <unistd.h>
#include "dcl_common.cuh"
#include "dcl_plan.h"

namespace {

constexpr int kN = 624, kM = 397;

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
#pragma GCC ivdep
    for (int i = kN - kM; i < kN - 1; ++i) s[i] = s[i + kM - kN] ^ mix(s[i], s[i + 1]);
    s[kN - 1] = s[kM - 1] ^ mix(s[kN - 1], s[0]);
}
__attribute__((target("avx512f"))) void twist_avx512(uint32_t* s) { twist_body(s); }
__attribute__((target("avx2"))) void twist_avx2(uint32_t* s) { twist_body(s); }
void twist_generic(uint32_t* s) { twist_body(s); }
inline void twist_words(uint32_t* s) {
    static const int isa = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    if (isa == 2) twist_avx512(s);
    else if (isa == 1) twist_avx2(s);
    else twist_generic(s);
}

// Publishing a block to the look-ahead ring.  The ring (41 MB) is far larger than the worker's cache share, and a
// plain store to it first reads the line from DRAM: regenerating straight into the ring ran at 0.46 us per block, the
// same arithmetic on a cache-resident block at 0.12 us.  So the worker keeps ONE block in its L1 (in-place twist) and
// streams a copy into the ring slot with non-temporal stores (no read-for-ownership, no cache pollution; the slots
// are 64-byte aligned and 2496 = 39 x 64 bytes long).  The consumers - the GPU's copy engine and, for a host-side
// plan, the main thread - read the ring from memory anyway.
__attribute__((target("avx2"))) void stream_block_avx2(uint32_t* dst, const uint32_t* src) {
    for (int i = 0; i < kN; i += 8)
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)));
    _mm_sfence();
}
inline void stream_block(uint32_t* dst, const uint32_t* src) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) stream_block_avx2(dst, src);
    else std::memcpy(dst, src, sizeof(uint32_t) * kN);
}

// Look-ahead of the generator's state sequence.  The sequence of regenerated state blocks depends only on the
// generator state, not on the data, so a worker thread produces it ahead of time into a ring; the sampler then
// consumes blocks instead of regenerating them on the step's critical path (0.2 ms per step at cfg2).  The stream is
// valid while torch's generator is where the previous plan left it (`expected`); any other use of the generator
// (manual_seed, torch.rand, ...) is detected by comparing the serialized state and the stream restarts from it.
class Lookahead {
public:
    static constexpr uint64_t kRing = 16384;                   // blocks (41 MB): ~10 M draws ahead
    // blocks [floor, produced) are readable.  `seen` is the caller's last reading of `produced`: the worker stores to
    // that counter all the time, so loading it once per block (a plan walks ~1700 blocks) was a cache-line transfer
    // each time - about half of a cfg2 plan.
    const uint32_t* block(uint64_t idx, uint64_t& seen) {
        if (idx >= seen) {
            seen = produced_.load(std::memory_order_acquire);
            if (seen <= idx) {
                // the consumer has caught up with the worker: diagnostics keep the time lost here
                const auto t0 = std::chrono::steady_clock::now();
                while (seen <= idx) {
                    std::this_thread::yield();
                    seen = produced_.load(std::memory_order_acquire);
                }
                const long long ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
                wait_ns_ += ns;
                if (ns > wait_max_ns_) wait_max_ns_ = ns;
            }
        }
        return ring_ + (idx % kRing) * kN;
    }
    // does the serialized torch state equal where the last plan left the generator?  (same process only)
    bool active() const { return active_ && pid_ == getpid(); }
    bool matches(const uint8_t* raw) const {
        return have_expected_ && epid_ == getpid() && std::memcmp(raw + 8, expected_ + 8, 16 + 8 * kN) == 0;
    }
    void expect(const uint8_t* raw) {                          // record where a plan left the generator
        std::memcpy(expected_, raw, sizeof(expected_));
        have_expected_ = true;
        epid_ = getpid();
    }
    uint64_t position() const { return pos_block_; }
    long long wait_ns() const { return wait_ns_; }
    long long wait_max_ns() const { return wait_max_ns_; }
    uint64_t epoch() const { return epoch_; }
    uint64_t produced() const { return produced_.load(std::memory_order_acquire); }
    uint32_t* ring() const { return ring_; }
    // restart from the state words `s` as block 0
    void restart(const uint32_t* s) {
        stop();
        ++epoch_;
        if (!ring_) {
            void* mem = nullptr;
            if (posix_memalign(&mem, 4096, sizeof(uint32_t) * kRing * kN) != 0) mem = nullptr;
            ring_ = static_cast<uint32_t*>(mem);
            // page-locked so the device mirror of the stream (dcl_step.cu) can be filled by asynchronous copies;
            // without a CUDA device (CPU tests) the registration simply fails and nothing depends on it
            if (ring_ && cudaHostRegister(ring_, sizeof(uint32_t) * kRing * kN, cudaHostRegisterPortable) != cudaSuccess)
                cudaGetLastError();
        }
        std::memcpy(ring_, s, sizeof(uint32_t) * kN);
        std::memcpy(last_, s, sizeof(uint32_t) * kN);
        produced_.store(1, std::memory_order_release);
        floor_.store(0, std::memory_order_release);
        quit_ = false;
        pid_ = getpid();
        pos_block_ = 0;
        worker_ = std::thread([this] { run(); });
        active_ = true;
    }
    // the plan ended in block `idx`; remember the serialized state torch now holds.  Blocks from `keep_from` on stay
    // readable (the worker only reuses slots of older blocks): a device-mode plan passes the first block its
    // permutations read, because the upload of that window to the GPU reads the ring after the plan has returned.
    void commit(uint64_t idx, const uint8_t* raw, uint64_t keep_from) {
        pos_block_ = idx;
        expect(raw);
        {
            // under the worker's mutex: a store + notify that slipped between the worker's predicate check and its
            // wait would be lost and the worker would sleep through its time-out while the ring runs dry
            std::lock_guard<std::mutex> g(mu_);
            floor_.store(keep_from < idx ? keep_from : idx, std::memory_order_release);
        }
        cv_.notify_one();
    }
    void stop() {
        if (worker_.joinable() && pid_ == getpid()) {
            { std::lock_guard<std::mutex> g(mu_); quit_ = true; }
            cv_.notify_one();
            worker_.join();
        } else if (worker_.joinable()) {
            worker_.detach();                                  // forked child: the thread does not exist here
        }
        active_ = false;
    }
    ~Lookahead() { stop(); }

private:
    void run() {
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait_for(lk, std::chrono::milliseconds(2), [this] {
                    return quit_ || produced_.load(std::memory_order_relaxed) - floor_.load(std::memory_order_acquire) < kRing - 1;
                });
                if (quit_) return;
            }
            while (produced_.load(std::memory_order_relaxed) - floor_.load(std::memory_order_acquire) < kRing - 1) {
                twist_words(last_);
                const uint64_t idx = produced_.load(std::memory_order_relaxed);
                stream_block(ring() + (idx % kRing) * kN, last_);
                produced_.store(idx + 1, std::memory_order_release);
                if (quit_) return;
            }
        }
    }
    uint32_t* ring_ = nullptr;                 // kRing blocks, page-aligned, page-locked when a device is present
    alignas(64) uint32_t last_[kN];
    std::atomic<uint64_t> produced_{0}, floor_{0};
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread worker_;
    bool quit_ = false, active_ = false, have_expected_ = false;
    pid_t pid_ = 0, epid_ = 0;
    uint64_t pos_block_ = 0, epoch_ = 0;
    long long wait_ns_ = 0, wait_max_ns_ = 0;
    uint8_t expected_[24 + 8 * kN] = {0};
};

struct Mt {
    const uint32_t* s;       // current state block: `own` or a block of the look-ahead ring
    uint32_t own[kN];
    int left;
    uint32_t next;
    Lookahead* la = nullptr;
    uint64_t blk = 0;
    uint64_t la_seen = 0;    // blocks known to be produced (see Lookahead::block)
};

void twist(Mt& m) {
    if (m.la) {
        m.s = m.la->block(++m.blk, m.la_seen);
    } else {
        twist_words(m.own);
        m.s = m.own;
    }
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
    std::vector<double> inv;           // 1 / (n - i) of the current permutation (see randperm_prefix)
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
    // draw % (n - i) without the integer divider (a third of the loop's time): the reciprocals of the k divisors
    // come from one vectorisable loop, the quotient estimate x * (1 / d) is off by at most one (x < 2^32, d < 2^31,
    // 53-bit products), and the remainder is corrected exactly.  Same value as the reference's 32-bit modulo.
    sp.inv.resize(static_cast<size_t>(k));
    double* inv = sp.inv.data();
    for (int64_t i = 0; i < k; ++i) inv[i] = 1.0 / static_cast<double>(n - i);
    int64_t drawn = 0;
    for (int64_t i = 0; i < k; ++i) {
        int64_t j = i;
        if (i < n - 1) {
            const uint32_t x = draw(m);
            const int64_t d = n - i;
            const int64_t q = static_cast<int64_t>(static_cast<double>(x) * inv[i]);
            int64_t r = static_cast<int64_t>(x) - q * d;
            r += (r < 0) ? d : 0;
            r -= (r >= d) ? d : 0;
            j = i + r;
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

// Small persistent pool for the stream mode of plan_rows: with the state blocks coming from the look-ahead ring the
// permutations of different anchors are independent (each starts at a known offset of the stream), so they are
// replayed on several cores.  Helpers sleep on a condition variable between plans.
class Pool {
public:
    // run fn(i) for i in [0, n) on up to `threads` threads (the caller is one of them)
    template <class Fn>
    void run(int n, int threads, Fn&& fn) {
        if (threads > n) threads = n;
        if (threads <= 1) {
            for (int i = 0; i < n; ++i) fn(i);
            return;
        }
        ensure(threads - 1);
        std::function<void(int)> f = fn;
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &f;
            n_ = n;
            next_.store(0, std::memory_order_relaxed);
            pending_ = threads - 1;
            wanted_ = threads - 1;
            ++epoch_;
        }
        cv_.notify_all();
        for (int i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < n;) f(i);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~Pool() { shutdown(); }

private:
    void ensure(int helpers) {
        if (pid_ != getpid()) {                     // forked child: the helper threads do not exist here
            for (auto& t : workers_) t.detach();
            workers_.clear();
            pid_ = getpid();
        }
        while (static_cast<int>(workers_.size()) < helpers) {
            const int id = static_cast<int>(workers_.size());
            workers_.emplace_back([this, id] { loop(id); });
        }
    }
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int)>* f;
            int n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || (epoch_ != seen && id < wanted_); });
                if (quit_) return;
                seen = epoch_;
                f = fn_;
                n = n_;
            }
            for (int i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < n;) (*f)(i);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    void shutdown() {
        if (pid_ == getpid()) {
            { std::lock_guard<std::mutex> g(mu_); quit_ = true; }
            cv_.notify_all();
            for (auto& t : workers_) t.join();
        } else {
            for (auto& t : workers_) t.detach();
        }
        workers_.clear();
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::function<void(int)>* fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, pending_ = 0, wanted_ = 0;
    uint64_t epoch_ = 0;
    bool quit_ = false;
    pid_t pid_ = getpid();
};

// threads for `draws` sampled positions: waking the helpers only pays from a few ten thousand draws on (measured on
// the 16-core B200 host: 8192 draws 126 us alone, 148 us on two threads; 65536 draws 660 us alone, 215 us on eight)
int host_threads(int64_t draws) {
    // DCL_HOST_THREADS=n: exactly n threads whatever the size (tests, experiments)
    static const int forced = [] { const char* e = std::getenv("DCL_HOST_THREADS"); return e ? std::max(1, std::atoi(e)) : 0; }();
    if (forced) return forced;
    static const int cap = static_cast<int>(std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)));
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(cap, draws / 8192)));
}
Pool& pool() {
    static Pool p;
    return p;
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
    for (int i = 0; i < kN; ++i) m.own[i] = static_cast<uint32_t>(st[i]);
    m.s = m.own;
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
static long long g_plan_ns[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // diagnostics: sections of the last plan (dcl_host_plan_timing)
static inline long long now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static long long g_la_stats[4] = {0, 0, 0, 0};
static Lookahead g_la;
static std::mutex g_la_mu;
// Diagnostics: out[2] = nanoseconds the plans have spent waiting for the look-ahead worker (total, longest wait)
extern "C" int dcl_host_lookahead_wait(long long* out) {
    if (!out) return dcl::fail(DCL_ERR_ARG, "null pointer argument");
    out[0] = g_la.wait_ns();
    out[1] = g_la.wait_max_ns();
    return 0;
}   // plans served by the look-ahead stream, inline plans, stream starts, drops
// Diagnostics: counters of the generator look-ahead (see Lookahead): out[4] = stream plans, inline plans, starts, drops.
extern "C" int dcl_host_plan_timing(long long* out) {
    if (!out) return dcl::fail(DCL_ERR_ARG, "null pointer argument");
    for (int i = 0; i < 8; ++i) out[i] = g_plan_ns[i];
    return 0;
}

extern "C" int dcl_host_lookahead_wait(long long* out);
extern "C" int dcl_host_lookahead_stats(long long* out) {
    for (int i = 0; i < 4; ++i) out[i] = g_la_stats[i];
    return 0;
}
int dcl::plan_rows(const int32_t* counts, int Bl, int world, int rank, int ignore_label, int max_samples,
                   int max_views, void* torch_rng_state, size_t state_bytes, int32_t* info, int64_t* image,
                   int64_t* cls, int64_t* num_hard, int64_t* num_easy, int64_t* keep_hard, int64_t* ranks,
                   int32_t* req, int32_t* y_all, int64_t* ref_row, int64_t* anchor, DevicePlan* dp) {
    if (dp) dp->taken = 0;
    if (!counts || !info || Bl <= 0 || world <= 0 || rank < 0 || rank >= world) return dcl::fail(DCL_ERR_ARG, "bad argument");
    const int B = Bl * world;
    const long long t_start = now_ns();
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
    g_plan_ns[0] = now_ns() - t_start;                       // anchor list + split rule
    bool device_mode = false;
    if (n_view > 0) {
        Mt m;
        if (int e = load_state(torch_rng_state, state_bytes, m)) return e;
        // Look-ahead policy: the stream is used while the generator is exactly where the previous plan left it.
        // A mismatch (manual_seed, other draws) drops to the inline replay; the stream is (re)started after a plan
        // that found the generator untouched, so loops that reseed every step never pay for a useless worker.
        Lookahead& la = g_la;
        static const bool la_enabled = [] { const char* e = std::getenv("DCL_HOST_LOOKAHEAD"); return !(e && e[0] == '0'); }();
        std::lock_guard<std::mutex> la_lock(g_la_mu);
        const uint8_t* raw = static_cast<const uint8_t*>(torch_rng_state);
        const bool continuous = la_enabled && la.matches(raw);
        // the ring is only refilled when a plan commits: a plan longer than the ring has to run inline
        int64_t total_draws = 0;
        for (int a = 0; a < A; ++a) total_draws += (num_hard[a] > 1 ? num_hard[a] - 1 : 0) + (num_easy[a] > 1 ? num_easy[a] - 1 : 0);
        const bool fits = total_draws / kN + 4 < static_cast<int64_t>(Lookahead::kRing) - 4;
        if (la.active()) {
            if (continuous && fits) {
                m.la = &la;
                m.blk = la.position();
                m.s = la.block(m.blk, m.la_seen);
                ++g_la_stats[0];
            } else {
                la.stop();
                ++g_la_stats[3];
            }
        }
        if (!m.la) ++g_la_stats[1];
        g_plan_ns[1] = now_ns() - t_start;                   // + generator state, look-ahead attach
        int n_local = 0;
        for (int a = 0; a < A; ++a) {
            const int64_t kh = keep_hard[a], ke = n_view - kh;
            if (kh < 0 || ke < 0 || kh > num_hard[a] || ke > num_easy[a] || num_hard[a] >= 214748364 ||
                num_easy[a] >= 214748364)
                return dcl::fail(DCL_ERR_ARG, "anchor %d: keep (%lld,%lld) exceeds counts (%lld,%lld)", a,
                                 (long long)kh, (long long)ke, (long long)num_hard[a], (long long)num_easy[a]);
            n_local += image[a] >= lo && image[a] < hi;
        }
        // Device mode: with the stream attached every permutation's position in the generator's output is known
        // without drawing it (randperm(n) consumes n - 1 draws whatever its outcome), so the host only records
        // where each local permutation starts and advances the generator; k_plan (dcl_plan.cu) replays the kept
        // prefixes on the GPU from its mirror of the same stream blocks.
        device_mode = dp && dp->allow_device && m.la && n_view <= dcl::kMaxDeviceViews;
        const int threads = (m.la && !device_mode) ? std::min(n_local, host_threads(static_cast<int64_t>(n_local) * n_view)) : 1;
        // draw index of the generator's next output, counted from word 0 of stream block 0
        auto gabs = [](const Mt& g) -> uint64_t {
            return g.left > 1 ? g.blk * static_cast<uint64_t>(kN) + g.next : (g.blk + 1) * static_cast<uint64_t>(kN);
        };
        if (device_mode) {
            int li = 0;
            uint64_t g_first = ~0ull, g_last = 0;
            dp->plan_first_block = gabs(m) / kN;
            for (int a = 0; a < A; ++a) {
                const bool local = image[a] >= lo && image[a] < hi;
                const uint64_t gh = gabs(m);
                if (num_hard[a] > 1) skip(m, num_hard[a] - 1);
                const uint64_t ge = gabs(m);
                if (num_easy[a] > 1) skip(m, num_easy[a] - 1);
                if (local) {
                    dcl::PlanAnchor& pa = dp->anchors[li++];
                    pa.g_hard = gh;
                    pa.g_easy = ge;
                    pa.num_hard = static_cast<int32_t>(num_hard[a]);
                    pa.num_easy = static_cast<int32_t>(num_easy[a]);
                    pa.keep_hard = static_cast<int32_t>(keep_hard[a]);
                    pa.keep_easy = static_cast<int32_t>(n_view - keep_hard[a]);
                    pa.row0 = 0;                               // set with the layout below
                    pa.image = static_cast<int32_t>(image[a] - lo);
                    pa.cls = static_cast<int32_t>(cls[a]);
                    pa.reserved = a;                           // anchor id, resolved to row0 below
                    if (gh < g_first) g_first = gh;
                    const uint64_t end = ge + static_cast<uint64_t>(pa.keep_easy > 0 ? pa.keep_easy : 0);
                    const uint64_t endh = gh + static_cast<uint64_t>(pa.keep_hard > 0 ? pa.keep_hard : 0);
                    if (end > g_last) g_last = end;
                    if (endh > g_last) g_last = endh;
                }
            }
            dp->plan_end_block = gabs(m) / kN;
            dp->n_local_anchors = li;
            dp->first_block = li ? g_first / kN : la.position();
            dp->last_block = li ? g_last / kN : la.position();
            uint64_t seen = 0;
            la.block(dp->last_block, seen);                    // the worker has produced every block the device reads
            dp->epoch = la.epoch();
            dp->host_ring = la.ring();
            dp->ring_blocks = Lookahead::kRing;
            dp->taken = 1;
        } else if (threads > 1) {
            // Stream mode: one cheap pass records the generator at the start of each local anchor (and leaves `m`
            // where the whole plan ends); the permutations themselves then run in parallel.
            std::vector<Mt> at(static_cast<size_t>(n_local));
            std::vector<int> which(static_cast<size_t>(n_local));
            int li = 0;
            for (int a = 0; a < A; ++a) {
                if (image[a] >= lo && image[a] < hi) { at[li] = m; which[li] = a; ++li; }
                if (num_hard[a] > 1) skip(m, num_hard[a] - 1);
                if (num_easy[a] > 1) skip(m, num_easy[a] - 1);
            }
            pool().run(n_local, threads, [&](int i) {
                thread_local Sparse sp;
                Mt mm = at[i];
                const int a = which[i];
                const int64_t kh = keep_hard[a];
                randperm_prefix(mm, num_hard[a], kh, ranks + static_cast<size_t>(a) * n_view, sp);
                randperm_prefix(mm, num_easy[a], n_view - kh, ranks + static_cast<size_t>(a) * n_view + kh, sp);
            });
        } else {
            Sparse sp;
            for (int a = 0; a < A; ++a) {
                const int64_t kh = keep_hard[a], ke = n_view - kh;
                if (image[a] >= lo && image[a] < hi) {
                    randperm_prefix(m, num_hard[a], kh, ranks + static_cast<size_t>(a) * n_view, sp);
                    randperm_prefix(m, num_easy[a], ke, ranks + static_cast<size_t>(a) * n_view + kh, sp);
                } else {
                    // another rank's anchor: randperm(n) consumes n - 1 draws (none for n <= 1)
                    if (num_hard[a] > 1) skip(m, num_hard[a] - 1);
                    if (num_easy[a] > 1) skip(m, num_easy[a] - 1);
                }
            }
        }
        g_plan_ns[2] = now_ns() - t_start;                   // + permutations
        store_state(torch_rng_state, m);
        if (m.la) {
            la.commit(m.blk, raw, device_mode ? dp->first_block : m.blk);
        } else {
            if (continuous && fits) {
                uint32_t words[kN];
                for (int i = 0; i < kN; ++i) words[i] = m.s[i];
                la.restart(words);                 // generator untouched since the last plan: stream from here on
                ++g_la_stats[2];
                la.commit(0, raw, 0);
            } else {
                la.expect(raw);
            }
        }
        if (dp) dp->produced = la.active() ? la.produced() : 0;
    }
    g_plan_ns[3] = now_ns() - t_start;                       // + state write-back, look-ahead commit
    // rows per rank -> common padded block size
    std::vector<int> per_rank(world, 0);
    for (int a = 0; a < A; ++a) per_rank[image[a] / Bl] += n_view;
    int n_max = 0, n_global = 0;
    for (int r = 0; r < world; ++r) { n_max = per_rank[r] > n_max ? per_rank[r] : n_max; n_global += per_rank[r]; }
    int n_pad = (n_max + 127) / 128 * 128;
    if (n_pad < 128) n_pad = 128;
    info[2] = per_rank[rank]; info[3] = n_pad; info[4] = n_global;
    if (!y_all && req) y_all = req + static_cast<size_t>(4) * n_pad;      // compact form: labels right behind the requests
    // rows of every rank block: anchors of that rank stably sorted by class (counting sort), views contiguous;
    // labels for all blocks (y_all), requests / bookkeeping for the local block only
    std::vector<int> start(257), order(A);
    std::vector<int> row0_of(device_mode ? A : 0);
    int yo = 0;
    for (int r = 0; r < world; ++r) {
        std::fill(start.begin(), start.end(), 0);
        const int rlo = r * Bl, rhi = rlo + Bl;
        int cnt = 0;
        for (int a = 0; a < A; ++a)
            if (image[a] >= rlo && image[a] < rhi) { ++start[cls[a] + 1]; ++cnt; }
        for (int c = 0; c < 256; ++c) start[c + 1] += start[c];
        for (int a = 0; a < A; ++a)
            if (image[a] >= rlo && image[a] < rhi) order[start[cls[a]]++] = a;
        if (dp) {
            // the class-sorted anchor order of every rank block (what k_plan needs for the labels, and what the
            // caller needs to map device rows back to the reference's row order)
            dp->ycnt[r] = cnt;
            dp->yoff[r] = yo;
            for (int o = 0; o < cnt; ++o) {
                dp->ycls[yo + o] = static_cast<int32_t>(cls[order[o]]);
                dp->yanchor[yo + o] = order[o];
                if (device_mode && r == rank) row0_of[order[o]] = o * n_view;
            }
            yo += cnt;
        }
        if (device_mode) continue;                           // k_plan writes the rows
        int32_t* yb = y_all + static_cast<size_t>(r) * n_pad;
        auto fill_anchor = [&](int o) {
            const int a = order[o];
            const int64_t* rk = ranks + static_cast<size_t>(a) * n_view;
            int row = o * n_view;
            for (int v = 0; v < n_view; ++v, ++row) {
                yb[row] = static_cast<int32_t>(cls[a]);
                if (r == rank) {
                    req[row * 4 + 0] = static_cast<int32_t>(image[a] - lo);
                    req[row * 4 + 1] = static_cast<int32_t>(cls[a]);
                    req[row * 4 + 2] = v >= keep_hard[a];
                    req[row * 4 + 3] = static_cast<int32_t>(rk[v]);
                    if (ref_row) ref_row[row] = static_cast<int64_t>(v) * A + a;
                    if (anchor) anchor[row] = a;
                }
            }
        };
        const int fill_threads = (r == rank) ? std::min(cnt, host_threads(static_cast<int64_t>(cnt) * n_view)) : 1;
        if (fill_threads > 1) pool().run(cnt, fill_threads, fill_anchor);
        else for (int o = 0; o < cnt; ++o) fill_anchor(o);
        for (int row = cnt * n_view; row < n_pad; ++row) {
            yb[row] = -1;
            if (r == rank) {
                req[row * 4 + 0] = req[row * 4 + 1] = req[row * 4 + 2] = req[row * 4 + 3] = -1;
                if (ref_row) ref_row[row] = -1;
                if (anchor) anchor[row] = -1;
            }
        }
    }
    if (device_mode)
        for (int i = 0; i < dp->n_local_anchors; ++i) {
            dcl::PlanAnchor& pa = dp->anchors[i];
            pa.row0 = row0_of[pa.reserved];
            pa.reserved = 0;
        }
    g_plan_ns[4] = now_ns() - t_start;                       // + row requests / labels
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
    const int rc = dcl::plan_rows(counts, B, 1, 0, ignore_label, max_samples, max_views, torch_rng_state, state_bytes, inf,
                                  image, cls, num_hard, num_easy, keep_hard, ranks, req, y, ref_row, anchor, nullptr);
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
    return dcl::plan_rows(counts, Bl, world, rank, ignore_label, max_samples, max_views, torch_rng_state, state_bytes, info,
                          image, cls, num_hard, num_easy, keep_hard, ranks, req, y_all, ref_row, anchor, nullptr);
}

// Diagnostics / tests (host only): plan_rows with the device-plan descriptors requested, as dcl_step_fwd calls it.
// When meta[0] comes back 1 the host did not draw the permutations: `anchors` (48-byte PlanAnchor records, see
// dcl_plan.h) place them in the look-ahead stream, whose ring (raw mt19937 state blocks, block b at
// ring + (b % ring_blocks) * 624 words) is returned through ring_out.  meta [8]: taken, local anchors, epoch,
// first block, last block, ring blocks, produced, -.
extern "C" int dcl_debug_plan_device(const int32_t* counts, int Bl, int world, int rank, int ignore_label,
                                     int max_samples, int max_views, void* torch_rng_state, size_t state_bytes,
                                     int32_t* info, int64_t* image, int64_t* cls, int64_t* num_hard,
                                     int64_t* num_easy, int64_t* keep_hard, int64_t* ranks, int32_t* req,
                                     int32_t* y_all, void* anchors, int32_t* ycls, int32_t* yanchor, int32_t* ycnt,
                                     int32_t* yoff, long long* meta, const void** ring_out) {
    if (!anchors || !ycls || !yanchor || !ycnt || !yoff || !meta || !ring_out) return dcl::fail(DCL_ERR_ARG, "null pointer argument");
    dcl::DevicePlan dp{};
    dp.allow_device = 1;
    dp.anchors = static_cast<dcl::PlanAnchor*>(anchors);
    dp.ycls = ycls; dp.yanchor = yanchor; dp.ycnt = ycnt; dp.yoff = yoff;
    const int rc = dcl::plan_rows(counts, Bl, world, rank, ignore_label, max_samples, max_views, torch_rng_state,
                                  state_bytes, info, image, cls, num_hard, num_easy, keep_hard, ranks, req, y_all,
                                  nullptr, nullptr, &dp);
    meta[0] = dp.taken; meta[1] = dp.n_local_anchors; meta[2] = static_cast<long long>(dp.epoch);
    meta[3] = static_cast<long long>(dp.first_block); meta[4] = static_cast<long long>(dp.last_block);
    meta[5] = static_cast<long long>(dp.ring_blocks); meta[6] = static_cast<long long>(dp.produced); meta[7] = 0;
    *ring_out = dp.host_ring;
    return rc;
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
