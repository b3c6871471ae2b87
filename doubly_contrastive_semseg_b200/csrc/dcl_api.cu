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

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;     // B200; only reached when sizing a workspace without a device
    }
    cached = n;
    return cached;
}

}  // namespace dcl

extern "C" int dcl_version(void) { return 100; }

extern "C" const char* dcl_last_error(void) { return dcl::last_error_buf(); }

extern "C" int dcl_check_device(void) {
    static int ok = -1;
    if (ok == 1) return 0;
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return dcl::fail(DCL_ERR_ARCH, "no usable CUDA device: %s", cudaGetErrorString(e));
    }
    if (major != 10)
        return dcl::fail(DCL_ERR_ARCH, "libdcl_b200 is built for sm_100a only; device has compute capability %d.x", major);
    ok = 1;
    return 0;
}
