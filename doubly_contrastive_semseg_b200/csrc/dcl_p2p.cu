// Exchange steps of the sharded pixel term over NVLink peer memory (one process per GPU; the reference has no
// multi-GPU path).  The three all-gathers of a step - count tables (2 KB per image), F-tiles (256 B per
// anchor row), row constants (32 B per row) - move little data but sit on the critical path, where an NCCL
// all-gather costs 50-100 us each on eight GPUs (profiles/r02q_sharded_timeline_8gpu.log).  Here every rank owns an
// arena that its peers map through CUDA IPC; a producer kernel leaves the rank's block in the local arena, k_p2p_push
// stores it into every peer's arena (16-byte remote stores through NVSwitch), and k_p2p_sync publishes a sequence
// number in every peer's flag row and waits for all peers' numbers in its own.  Regions are double-buffered by the
// parity of the sequence number: a rank can be at most one exchange of a kind ahead of the slowest peer, and that
// peer's reads of the older parity were ordered before its own signal by its stream.
#include <cstring>
#include "dcl_common.cuh"

namespace dcl {

constexpr int kP2pMaxWorld = 16;
constexpr int kP2pKinds = 3;                       // 0 count tables, 1 F-tiles, 2 row constants

struct PeerPtrs { uint8_t* p[kP2pMaxWorld]; };

struct P2p {
    int world = 0, rank = 0;
    size_t bytes = 0;
    uint8_t* local = nullptr;
    uint8_t* peer[kP2pMaxWorld] = {nullptr};
    size_t off_flags = 0;
    size_t off_region[kP2pKinds][2] = {{0}};
    size_t cap_bytes[kP2pKinds] = {0};             // bytes per rank a region can hold
    unsigned long long seq[kP2pKinds] = {0};
    bool opened = false;
};

__global__ void __launch_bounds__(256)
k_p2p_push(const uint4* __restrict__ src, size_t n16, PeerPtrs dst, int world, int rank) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = src[i];
#pragma unroll 1
        for (int r = 0; r < world; ++r)
            if (r != rank) reinterpret_cast<uint4*>(dst.p[r])[i] = v;
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns64() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// thread r: publish `seq` in peer r's flag row (slot = this rank), then wait until peer r's number in the local row
// has reached `seq`.  The kernel before this one on the stream wrote the data; its completion and the system-scope
// fence order those stores before the flag.  A peer that never arrives becomes a trap after 20 s, not a hung box.
__global__ void __launch_bounds__(32)
k_p2p_sync(PeerPtrs flags_peer, unsigned long long* flags_local, int world, int rank, unsigned long long seq) {
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(flags_peer.p[r]) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(seq) : "memory");
    const unsigned long long* mine = flags_local + r;
    const unsigned long long t0 = globaltimer_ns64();
    unsigned long long v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= seq) break;
        if (globaltimer_ns64() - t0 > 20000000000ull) __trap();
        __nanosleep(200);
    }
}

// ---- used by dcl_step.cu ---------------------------------------------------------------------------------------
// start exchange `kind` of a step: returns the local region the producers fill (this rank's block at
// rank * bytes_per_rank, everybody else's arrives by p2p_exchange)
uint8_t* p2p_begin(P2p* p, int kind) {
    ++p->seq[kind];
    return p->local + p->off_region[kind][p->seq[kind] & 1];
}

int p2p_exchange(P2p* p, int kind, size_t bytes_per_rank, bool use_copy_engine, cudaStream_t st) {
    if (!p->opened) return fail(DCL_ERR_COMM, "peer arena not opened");
    if (bytes_per_rank == 0 || bytes_per_rank % 16 || bytes_per_rank > p->cap_bytes[kind])
        return fail(DCL_ERR_ARG, "exchange of %zu bytes per rank does not fit region %d (%zu, multiples of 16)", bytes_per_rank, kind, p->cap_bytes[kind]);
    const size_t off = p->off_region[kind][p->seq[kind] & 1] + bytes_per_rank * p->rank;
    PeerPtrs dst{}, flags{};
    for (int r = 0; r < p->world; ++r) {
        dst.p[r] = p->peer[r] + off;
        flags.p[r] = p->peer[r] + p->off_flags + sizeof(unsigned long long) * kP2pMaxWorld * kind;
    }
    if (use_copy_engine) {
        // large blocks: one peer copy each (the copy engines keep NVLink busier than 16-byte stores from the SMs)
        for (int i = 1; i < p->world; ++i) {
            const int r = (p->rank + i) % p->world;                  // every rank starts with a different peer
            DCL_CUDA(cudaMemcpyAsync(dst.p[r], p->local + off, bytes_per_rank, cudaMemcpyDeviceToDevice, st));
        }
    } else {
        const size_t n16 = bytes_per_rank / 16;
        size_t blocks = (n16 + 255) / 256;
        const size_t max_blocks = static_cast<size_t>(sm_count()) * 4;
        if (blocks > max_blocks) blocks = max_blocks;
        k_p2p_push<<<static_cast<unsigned>(blocks), 256, 0, st>>>(reinterpret_cast<const uint4*>(p->local + off), n16, dst, p->world, p->rank);
        DCL_LAUNCH_CHECK("k_p2p_push");
    }
    k_p2p_sync<<<1, 32, 0, st>>>(flags, reinterpret_cast<unsigned long long*>(p->local + p->off_flags) + kP2pMaxWorld * kind,
                                 p->world, p->rank, p->seq[kind]);
    DCL_LAUNCH_CHECK("k_p2p_sync");
    return 0;
}

}  // namespace dcl

using namespace dcl;

extern "C" int dcl_p2p_create(int world, int rank, int B_local, int cap, void** p2p_out, void* ipc_handle_out) {
    if (int e = dcl_check_device()) return e;
    if (!p2p_out || !ipc_handle_out) return fail(DCL_ERR_ARG, "null pointer argument");
    if (world < 2 || world > kP2pMaxWorld || rank < 0 || rank >= world || B_local <= 0 || cap <= 0 || cap % DCL_TILE_ROWS)
        return fail(DCL_ERR_ARG, "bad peer arena shape (world=%d rank=%d B=%d cap=%d; world <= %d)", world, rank, B_local, cap, kP2pMaxWorld);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    P2p* p = new P2p();
    p->world = world;
    p->rank = rank;
    p->cap_bytes[0] = sizeof(int32_t) * 512 * static_cast<size_t>(B_local);
    p->cap_bytes[1] = static_cast<size_t>(cap) * DCL_DIM * 2;
    p->cap_bytes[2] = sizeof(float) * 4 * (2 * static_cast<size_t>(cap) + 1);
    size_t o = 4096;                                   // flags: kinds x 16 x u64
    p->off_flags = 0;
    for (int k = 0; k < kP2pKinds; ++k)
        for (int h = 0; h < 2; ++h) {
            p->off_region[k][h] = o;
            o += (p->cap_bytes[k] * world + 4095) / 4096 * 4096;
        }
    p->bytes = o;
    cudaError_t e = cudaMalloc(&p->local, p->bytes);
    if (e != cudaSuccess) { delete p; return fail(static_cast<int>(e), "cudaMalloc of the peer arena (%zu bytes): %s", o, cudaGetErrorString(e)); }
    e = cudaMemset(p->local, 0, 4096);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) { cudaFree(p->local); delete p; return fail(static_cast<int>(e), "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    std::memcpy(ipc_handle_out, &h, sizeof(h));
    p->peer[rank] = p->local;
    *p2p_out = p;
    return 0;
}

extern "C" int dcl_p2p_open(void* p2p, const void* all_handles) {
    P2p* p = static_cast<P2p*>(p2p);
    if (!p || !all_handles) return fail(DCL_ERR_ARG, "null pointer argument");
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const uint8_t*>(all_handles) + sizeof(h) * r, sizeof(h));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(DCL_ERR_COMM, "cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
        }
        p->peer[r] = static_cast<uint8_t*>(ptr);
    }
    p->opened = true;
    return 0;
}

extern "C" int dcl_p2p_destroy(void* p2p) {
    P2p* p = static_cast<P2p*>(p2p);
    if (!p) return 0;
    for (int r = 0; r < p->world; ++r)
        if (r != p->rank && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
    if (p->local) cudaFree(p->local);
    cudaGetLastError();
    delete p;
    return 0;
}
