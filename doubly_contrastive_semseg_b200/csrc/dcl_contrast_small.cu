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
