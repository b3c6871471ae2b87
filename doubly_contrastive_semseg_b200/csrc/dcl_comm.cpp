// NCCL plumbing for the sharded step (one process per GPU): the library talks to NCCL itself, on the caller's
// stream, instead of walking through torch.distributed between kernels.  libnccl.so.2 is not linked: it is taken
// from the process (torch has it loaded) or dlopen'ed by soname, so the library still loads on a box without NCCL
// and single-GPU users never touch it.  Only the five entry points below are used; their ABI is stable over
// NCCL 2.x (ncclUniqueId = 128 opaque bytes passed by value).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include "dcl_common.cuh"

namespace {

struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void** comm, int nranks, NcclId id, int rank);
typedef int (*fn_comm_destroy)(void* comm);
typedef int (*fn_all_gather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t st);
typedef const char* (*fn_error_string)(int);
typedef int (*fn_comm_count)(void* comm, int* count);

struct Nccl {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
    fn_comm_count comm_count = nullptr;
    bool tried = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (n.tried) return n;
    n.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch already mapped, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return n;
    n.get_unique_id = reinterpret_cast<fn_get_unique_id>(dlsym(h, "ncclGetUniqueId"));
    n.comm_init_rank = reinterpret_cast<fn_comm_init_rank>(dlsym(h, "ncclCommInitRank"));
    n.comm_destroy = reinterpret_cast<fn_comm_destroy>(dlsym(h, "ncclCommDestroy"));
    n.all_gather = reinterpret_cast<fn_all_gather>(dlsym(h, "ncclAllGather"));
    n.error_string = reinterpret_cast<fn_error_string>(dlsym(h, "ncclGetErrorString"));
    n.comm_count = reinterpret_cast<fn_comm_count>(dlsym(h, "ncclCommCount"));
    if (n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_gather) n.handle = h;
    return n;
}

int nccl_fail(const char* what, int rc) {
    Nccl& n = nccl();
    return dcl::fail(DCL_ERR_COMM, "%s: NCCL error %d (%s)", what, rc, n.error_string ? n.error_string(rc) : "?");
}

}  // namespace

namespace dcl {
// all-gather of `bytes` per rank on `stream`; in place when send == recv + rank * bytes
int comm_all_gather(void* comm, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    Nccl& n = nccl();
    if (!n.handle) return fail(DCL_ERR_COMM, "libnccl.so.2 is not available in this process");
    if (!comm) return fail(DCL_ERR_ARG, "null communicator");
    const int rc = n.all_gather(send, recv, bytes, /* ncclUint8 */ 1, comm, st);
    if (rc != 0) return nccl_fail("ncclAllGather", rc);
    return 0;
}
}  // namespace dcl

extern "C" int dcl_comm_unique_id(void* out128) {
    Nccl& n = nccl();
    if (!n.handle) return dcl::fail(DCL_ERR_COMM, "libnccl.so.2 is not available in this process");
    if (!out128) return dcl::fail(DCL_ERR_ARG, "null pointer argument");
    NcclId id;
    const int rc = n.get_unique_id(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    std::memcpy(out128, &id, sizeof(id));
    return 0;
}

extern "C" int dcl_comm_init(const void* id128, int world, int rank, void** comm) {
    Nccl& n = nccl();
    if (!n.handle) return dcl::fail(DCL_ERR_COMM, "libnccl.so.2 is not available in this process");
    if (!id128 || !comm || world <= 0 || rank < 0 || rank >= world) return dcl::fail(DCL_ERR_ARG, "bad argument");
    NcclId id;
    std::memcpy(&id, id128, sizeof(id));
    void* c = nullptr;
    const int rc = n.comm_init_rank(&c, world, id, rank);
    if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
    *comm = c;
    return 0;
}

extern "C" int dcl_comm_destroy(void* comm) {
    Nccl& n = nccl();
    if (!n.handle || !comm) return 0;
    const int rc = n.comm_destroy(comm);
    if (rc != 0) return nccl_fail("ncclCommDestroy", rc);
    return 0;
}

extern "C" int dcl_comm_all_gather(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
    return dcl::comm_all_gather(comm, send, recv, bytes_per_rank, dcl::as_stream(stream));
}
