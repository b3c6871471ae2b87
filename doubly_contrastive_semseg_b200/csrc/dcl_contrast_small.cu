// Image-level term at its real size (reference utils/loss.py:161-204): 2B rows, B = images per step, i.e. tens of
// rows.  The tensor-core path quantises the rows to bf16, and this term is the one place where that is not
// harmless: the rows are projections of globally pooled features, nearly identical across images, and the loss only
// sees DIFFERENCES of their dot products (l = normalize(a - max a)), so an operand rounding of 2^-9 can be larger than
// the signal (measured at the cfg3 shapes: 5 % of the gradient's maximum).  A 32 x 32 problem has nothing to gain
// from tensor cores either, so rows <= 128 take this exact fp32 path: one CTA, the n x n matrix in shared memory,
// forward and gradient in one launch (the 7 + 2 launches of the tiled path cost more than the arithmetic here).
//
//   a = Z Z^T / T,  m_i = max_j a_ij (detached),  l = (a - m) / max(|a_i - m_i|_2, 1e-12),  E = exp(l)
//   image: lp_ij = l_ij - log sum_{k != i} E_ik      pixel: lp_ij = l_ij - log(E_ij + sum_{y_k != y_i} E_ik)
//   loss = mean_i [ -(T/T_b) mean_{j in pos(i)} lp_ij ];   dS_ik = (g_ik - l_ik R_i) / (r_i T),  dZ = (dS + dS^T) Z
#include <cfloat>
#include "dcl_common.cuh"

namespace dcl {

constexpr int kSmallMax = 128;          // rows
constexpr int kSmallDim = 128;          // channels
constexpr int kSmallThreads = 512;
constexpr int kSmallLd = kSmallMax + 1; // padded leading dimension of both shared arrays

template <int kMode>
__global__ void __launch_bounds__(kSmallThreads)
k_contrast_small(const float* __restrict__ Z, const int32_t* __restrict__ y, int n, float T, float Tb,
                 float* __restrict__ loss, float* __restrict__ dZ) {
    extern __shared__ float sm[];
    float* sZ = sm;                                  // [n][129]
    float* sS = sm + kSmallMax * kSmallLd;           // [n][129]: a -> l -> dS
    __shared__ float rowloss[kSmallMax];
    __shared__ int sy[kSmallMax];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n * kSmallDim; i += kSmallThreads) sZ[(i >> 7) * kSmallLd + (i & 127)] = Z[i];
    for (int i = tid; i < n; i += kSmallThreads) sy[i] = y[i];
    __syncthreads();
    // a_ij = z_i . z_j / T  (fp32, sequential in d like a plain GEMM)
    const float invT = 1.f / T;
    for (int e = tid; e < n * n; e += kSmallThreads) {
        const int i = e / n, j = e - i * n;
        const float* a = sZ + i * kSmallLd;
        const float* b = sZ + j * kSmallLd;
        float acc = 0.f;
#pragma unroll 8
        for (int d = 0; d < kSmallDim; ++d) acc = fmaf(a[d], b[d], acc);
        sS[i * kSmallLd + j] = acc * invT;
    }
    __syncthreads();
    // one warp per row: statistics, loss term, dS in place
    const float ratio = T / Tb, c = ratio / static_cast<float>(n);
    for (int i = warp; i < n; i += kSmallThreads / 32) {
        float* row = sS + i * kSmallLd;
        const int yi = sy[i];
        float m = -FLT_MAX;
        for (int k = lane; k < n; k += 32) m = fmaxf(m, row[k]);
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float ss = 0.f;
        for (int k = lane; k < n; k += 32) {
            const float u = row[k] - m;
            ss = fmaf(u, u, ss);
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float r = fmaxf(sqrtf(ss), 1e-12f);                  // F.normalize eps
        const float invr = 1.f / r;
        float den = 0.f, P = 0.f;
        for (int k = lane; k < n; k += 32) {
            const float l = (row[k] - m) * invr;
            row[k] = l;
            const bool same = sy[k] == yi;
            const bool in_den = kMode == DCL_MODE_PIXEL ? !same : (k != i);
            if (in_den) den += expf(l);
            if (same && k != i) P += 1.f;
        }
        for (int o = 16; o > 0; o >>= 1) {
            den += __shfl_xor_sync(0xffffffffu, den, o);
            P += __shfl_xor_sync(0xffffffffu, P, o);
        }
        const float w = -c / P;                                     // P == 0 -> NaN like the reference (loss.py:383)
        float slp = 0.f, Q = 0.f;
        for (int k = lane; k < n; k += 32) {
            if (sy[k] == yi && k != i) {
                const float l = row[k];
                if (kMode == DCL_MODE_PIXEL) {
                    const float d = expf(l) + den;
                    slp += l - logf(d);
                    Q += w / d;
                } else {
                    slp += l - logf(den);
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            slp += __shfl_xor_sync(0xffffffffu, slp, o);
            Q += __shfl_xor_sync(0xffffffffu, Q, o);
        }
        if (kMode != DCL_MODE_PIXEL) Q = -c / den;
        if (lane == 0) rowloss[i] = -ratio * slp / P;
        // g_ik and R_i = sum_k g_ik l_ik
        float R = 0.f;
        for (int k = lane; k < n; k += 32) {
            const float l = row[k], E = expf(l);
            const bool same = sy[k] == yi, pos = same && k != i;
            float g;
            if (kMode == DCL_MODE_PIXEL) g = (pos ? w * den / (E + den) : 0.f) - (!same ? E * Q : 0.f);
            else g = (pos ? w : 0.f) - (k != i ? E * Q : 0.f);
            R = fmaf(g, l, R);
        }
        for (int o = 16; o > 0; o >>= 1) R += __shfl_xor_sync(0xffffffffu, R, o);
        const float s = invr * invT;
        for (int k = lane; k < n; k += 32) {
            const float l = row[k], E = expf(l);
            const bool same = sy[k] == yi, pos = same && k != i;
            float g;
            if (kMode == DCL_MODE_PIXEL) g = (pos ? w * den / (E + den) : 0.f) - (!same ? E * Q : 0.f);
            else g = (pos ? w : 0.f) - (k != i ? E * Q : 0.f);
            row[k] = (g - l * R) * s;                               // dS_ik
        }
    }
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < n; ++i) t += rowloss[i];                // fixed order
        loss[0] = t / static_cast<float>(n);
    }
    if (dZ) {
        // dZ_i = sum_k (dS_ik + dS_ki) z_k
        for (int e = tid; e < n * kSmallDim; e += kSmallThreads) {
            const int i = e >> 7, d = e & 127;
            float acc = 0.f;
            for (int k = 0; k < n; ++k) acc = fmaf(sS[i * kSmallLd + k] + sS[k * kSmallLd + i], sZ[k * kSmallLd + d], acc);
            dZ[e] = acc;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// The rest of the image-level head at the same size: SupConLoss.projection (Linear 128 -> 128, ReLU, Linear 128 -> 128;
// reference utils/loss.py:102-106, applied at :120) forward and backward, and the group ids of the label mask
// (:151-159).  With torch these are ~45 launches of a few microseconds of work each (two cuBLAS GEMMs and their
// epilogues forward, four backward, stack / cat / eq / argmax / repeat and autograd's glue): at the cfg3 shapes the
// host spends more time issuing them than the GPU spends on the whole pixel term.  Here: one launch forward, two
// backward.  Plain fp32 FMAs in a fixed order (bit-reproducible); nothing here is large enough for tensor cores.
//   H = relu(X W1^T + b1)   Z = H W2^T + b2                                      (nn.Linear: weight [out, in])
//   dH = (dZ W2) * (H > 0)   dX = dH W1   dW2 = dZ^T H   db2 = sum_i dZ   dW1 = dH^T X   db1 = sum_i dH
// ---------------------------------------------------------------------------------------------
constexpr int kMlpDim = 128;

__device__ __forceinline__ float warp_dot128(const float* __restrict__ wrow, const float4 xv, int lane) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(wrow) + lane);
    float acc = w.x * xv.x;
    acc = fmaf(w.y, xv.y, acc);
    acc = fmaf(w.z, xv.z, acc);
    acc = fmaf(w.w, xv.w, acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// one CTA per row, eight warps x sixteen outputs per layer; a warp reads a weight row as one 512-byte request
__global__ void __launch_bounds__(256) k_mlp_fwd(const float* __restrict__ X, const float* __restrict__ W1,
                                                 const float* __restrict__ b1, const float* __restrict__ W2,
                                                 const float* __restrict__ b2, float* __restrict__ H, float* __restrict__ Z) {
    __shared__ float4 sx[kMlpDim / 4], sh[kMlpDim / 4];
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kMlpDim / 4) sx[threadIdx.x] = reinterpret_cast<const float4*>(X + static_cast<size_t>(i) * kMlpDim)[threadIdx.x];
    __syncthreads();
    const float4 xv = sx[lane];
#pragma unroll 4
    for (int u = 0; u < 16; ++u) {
        const int j = warp * 16 + u;
        const float h = fmaxf(warp_dot128(W1 + static_cast<size_t>(j) * kMlpDim, xv, lane) + __ldg(b1 + j), 0.f);
        if (lane == 0) {
            reinterpret_cast<float*>(sh)[j] = h;
            H[static_cast<size_t>(i) * kMlpDim + j] = h;
        }
    }
    __syncthreads();
    const float4 hv = sh[lane];
#pragma unroll 4
    for (int u = 0; u < 16; ++u) {
        const int j = warp * 16 + u;
        const float z = warp_dot128(W2 + static_cast<size_t>(j) * kMlpDim, hv, lane) + __ldg(b2 + j);
        if (lane == 0) Z[static_cast<size_t>(i) * kMlpDim + j] = z;
    }
}

// one CTA per row, thread k: dH[i][k] and dX[i][k]; the upstream scalar is applied to dZ here
__global__ void __launch_bounds__(kMlpDim) k_mlp_bwd_rows(const float* __restrict__ W1, const float* __restrict__ W2,
                                                          const float* __restrict__ H, const float* __restrict__ dZ,
                                                          const float* __restrict__ grad_out, float* __restrict__ dH,
                                                          float* __restrict__ dX) {
    __shared__ float sdz[kMlpDim], sdh[kMlpDim];
    const int i = blockIdx.x, k = threadIdx.x;
    const float g = __ldg(grad_out);
    sdz[k] = dZ[static_cast<size_t>(i) * kMlpDim + k] * g;
    __syncthreads();
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < kMlpDim; ++j) acc = fmaf(sdz[j], __ldg(W2 + static_cast<size_t>(j) * kMlpDim + k), acc);
    const float dh = H[static_cast<size_t>(i) * kMlpDim + k] > 0.f ? acc : 0.f;
    sdh[k] = dh;
    dH[static_cast<size_t>(i) * kMlpDim + k] = dh;
    __syncthreads();
    acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < kMlpDim; ++j) acc = fmaf(sdh[j], __ldg(W1 + static_cast<size_t>(j) * kMlpDim + k), acc);
    dX[static_cast<size_t>(i) * kMlpDim + k] = acc;
}

// grid (128 output rows j, 2 layers), thread k: dW[j][k] = sum_i a[i][j] b[i][k] in row order; thread 0 also the bias
__global__ void __launch_bounds__(kMlpDim) k_mlp_bwd_weights(const float* __restrict__ X, const float* __restrict__ H,
                                                             const float* __restrict__ dZ, const float* __restrict__ dH,
                                                             const float* __restrict__ grad_out, int n,
                                                             float* __restrict__ dW1, float* __restrict__ db1,
                                                             float* __restrict__ dW2, float* __restrict__ db2) {
    const int j = blockIdx.x, k = threadIdx.x;
    const bool second = blockIdx.y == 1;
    const float* a = second ? dZ : dH;                 // [n][128], column j
    const float* b = second ? H : X;                   // [n][128], column k
    const float g = second ? __ldg(grad_out) : 1.f;    // dH already carries the upstream scalar
    float acc = 0.f, bias = 0.f;
    for (int i = 0; i < n; ++i) {
        const float aij = a[static_cast<size_t>(i) * kMlpDim + j] * g;
        acc = fmaf(aij, b[static_cast<size_t>(i) * kMlpDim + k], acc);
        bias += aij;
    }
    (second ? dW2 : dW1)[static_cast<size_t>(j) * kMlpDim + k] = acc;
    if (k == 0) (second ? db2 : db1)[j] = bias;
}

// y[v * n + i] = first row k with labels[k] == labels[i] (loss.py:157: mask = eq(labels, labels^T) on the raw values;
// labels == nullptr: the identity, loss.py:151), for v < views
template <class T>
__global__ void __launch_bounds__(128) k_group_ids(const T* __restrict__ labels, int n, int views, int32_t* __restrict__ y) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int first = i;
        if (labels) {
            const T li = labels[i];
            for (int k = 0; k < i; ++k)
                if (labels[k] == li) { first = k; break; }
        }
        for (int v = 0; v < views; ++v) y[v * n + i] = first;
    }
}

}  // namespace dcl

using namespace dcl;

extern "C" int dcl_contrast_small_max_rows(void) { return kSmallMax; }

extern "C" int dcl_contrast_small(const float* Z, const int32_t* y, int n, int mode, float temperature,
                                  float base_temperature, float* loss, float* dZ, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!Z || !y || !loss) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n <= 0 || n > kSmallMax) return fail(DCL_ERR_ARG, "n must be 1..%d", kSmallMax);
    if (mode != DCL_MODE_PIXEL && mode != DCL_MODE_SUPCON) return fail(DCL_ERR_ARG, "bad mode %d", mode);
    const size_t smem = sizeof(float) * 2 * kSmallMax * kSmallLd;
    static bool configured[64] = {false};
    const int dev = current_device();
    bool& done = configured[(dev >= 0 && dev < 64) ? dev : 0];
    if (!done) {
        DCL_CUDA(cudaFuncSetAttribute(k_contrast_small<DCL_MODE_PIXEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        DCL_CUDA(cudaFuncSetAttribute(k_contrast_small<DCL_MODE_SUPCON>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        done = true;
    }
    if (mode == DCL_MODE_PIXEL)
        k_contrast_small<DCL_MODE_PIXEL><<<1, kSmallThreads, smem, as_stream(stream)>>>(Z, y, n, temperature, base_temperature, loss, dZ);
    else
        k_contrast_small<DCL_MODE_SUPCON><<<1, kSmallThreads, smem, as_stream(stream)>>>(Z, y, n, temperature, base_temperature, loss, dZ);
    DCL_LAUNCH_CHECK("k_contrast_small");
    return 0;
}

extern "C" int dcl_supcon_mlp_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                                  int n, float* H, float* Z, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!X || !W1 || !b1 || !W2 || !b2 || !H || !Z) return fail(DCL_ERR_ARG, "null pointer argument");
    if (n <= 0) return fail(DCL_ERR_ARG, "n must be positive");
    if (reinterpret_cast<uintptr_t>(X) % 16 || reinterpret_cast<uintptr_t>(W1) % 16 || reinterpret_cast<uintptr_t>(W2) % 16)
        return fail(DCL_ERR_ARG, "X, W1 and W2 must be 16-byte aligned");
    k_mlp_fwd<<<n, 256, 0, as_stream(stream)>>>(X, W1, b1, W2, b2, H, Z);
    DCL_LAUNCH_CHECK("k_mlp_fwd");
    return 0;
}

extern "C" int dcl_supcon_mlp_bwd(const float* X, const float* W1, const float* W2, const float* H, const float* dZ,
                                  const float* grad_out, int n, float* dH, float* dX, float* dW1, float* db1,
                                  float* dW2, float* db2, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!X || !W1 || !W2 || !H || !dZ || !grad_out || !dH || !dX || !dW1 || !db1 || !dW2 || !db2)
        return fail(DCL_ERR_ARG, "null pointer argument");
    if (n <= 0) return fail(DCL_ERR_ARG, "n must be positive");
    k_mlp_bwd_rows<<<n, kMlpDim, 0, as_stream(stream)>>>(W1, W2, H, dZ, grad_out, dH, dX);
    DCL_LAUNCH_CHECK("k_mlp_bwd_rows");
    k_mlp_bwd_weights<<<dim3(kMlpDim, 2), kMlpDim, 0, as_stream(stream)>>>(X, H, dZ, dH, grad_out, n, dW1, db1, dW2, db2);
    DCL_LAUNCH_CHECK("k_mlp_bwd_weights");
    return 0;
}

extern "C" int dcl_group_ids(const void* labels, int elem_bytes, int n, int views, int32_t* y, void* stream) {
    if (int e = dcl_check_device()) return e;
    if (!y || n <= 0 || views <= 0) return fail(DCL_ERR_ARG, "bad argument");
    if (labels && elem_bytes != 4 && elem_bytes != 8) return fail(DCL_ERR_ARG, "labels must be int32 or int64");
    if (elem_bytes == 8)
        k_group_ids<long long><<<1, 128, 0, as_stream(stream)>>>(static_cast<const long long*>(labels), n, views, y);
    else
        k_group_ids<int><<<1, 128, 0, as_stream(stream)>>>(static_cast<const int*>(labels), n, views, y);
    DCL_LAUNCH_CHECK("k_group_ids");
    return 0;
}
