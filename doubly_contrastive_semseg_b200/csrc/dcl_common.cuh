// Host-side plumbing shared by the translation units of libdcl_b200.so.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/dcl_b200.h"

namespace dcl {

char* last_error_buf();                       // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);     // records message, returns code
int sm_count();                               // SMs of the current device (148 on B200), cached per device
int current_device();                         // ordinal of the current device, -1 without one

#define DCL_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return ::dcl::fail(static_cast<int>(e_), "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define DCL_LAUNCH_CHECK(name)                                                             \
    do {                                                                                   \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess)                                                             \
            return ::dcl::fail(static_cast<int>(e_), "launch %s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// dcl_contrast.cu, for the one-call step (dcl_step.cu): the C-ABI entry points plus `loss_out` (the loss written by the
// forward's last block, no copy afterwards) and `chained` (the backward directly follows that forward on the stream)
int contrast_fwd_ex(const void* tiles, const int32_t* y, const float* sqnorm, int nJ, int rb0, int nI, int n_valid,
                    int mode, float temperature, float base_temperature, void* workspace, size_t workspace_bytes,
                    float* colA, float* colB, float* rowloss, float* loss_sum, float* loss_out, void* stream);
int contrast_bwd_ex(const void* tiles, const int32_t* y, const float* colA, const float* colB, int nJ, int rb0,
                    int nI, int mode, void* workspace, size_t workspace_bytes, float* dF, bool chained, void* stream);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace dcl
