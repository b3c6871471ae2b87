// Library-level entry points: version, error string, device check.
#include <cstring>
#include "dcl_common.cuh"

namespace dcl {

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// Per-device caches (a process may drive several GPUs): indexed by the device ordinal.
constexpr int kMaxDev = 64;

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

int sm_count() {
    static int cached[kMaxDev] = {0};
    const int dev = current_device();
    if (dev >= 0 && dev < kMaxDev && cached[dev] > 0) return cached[dev];
    int n = 0;
    if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;     // B200; only reached when sizing a workspace without a device
    }
    if (dev < kMaxDev) cached[dev] = n;
    return n;
}

}  // namespace dcl

extern "C" int dcl_version(void) { return 100; }

extern "C" const char* dcl_last_error(void) { return dcl::last_error_buf(); }

extern "C" int dcl_check_device(void) {
    static bool ok[dcl::kMaxDev] = {false};
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev >= 0 && dev < dcl::kMaxDev && ok[dev]) return 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return dcl::fail(DCL_ERR_ARCH, "no usable CUDA device: %s", cudaGetErrorString(e));
    }
    if (major != 10)
        return dcl::fail(DCL_ERR_ARCH, "libdcl_b200 is built for sm_100a only; device has compute capability %d.x", major);
    if (dev >= 0 && dev < dcl::kMaxDev) ok[dev] = true;
    return 0;
}
