// siren_common.cuh -- shared device/host helpers for the SIREN kernels.
//
// Forward-mode stream algebra of one sine layer (reference: Sine.forward,
// base/networks.py:21-27, differentiated once/twice as base/diff_ops.py does through
// torch.autograd).  With t = omega * z the pre-activation (omega already folded in) and
// s = sin t, c = cos t:
//     value      a0  = s
//     tangent d  a_d = c * t_d
//     trace      a_q = c * t_q - s * sum_d t_d^2                 (ORDER_LAP)
//     hess d<=e  a_q = c * t_q - s * t_d * t_e                   (ORDER_HESS)
#pragma once
#include "insr_platform.h"

struct SirenDims {
    int D, O, H, L;
    float omega;
};

template <int D, int ORDER>
struct StreamCfg {
    static constexpr int ND = (ORDER >= 1) ? D : 0;
    static constexpr int NQ = (ORDER == 2) ? 1 : (ORDER == 3 ? D * (D + 1) / 2 : 0);
    static constexpr int S = 1 + ND + NQ;
};

__host__ __device__ inline int insr_nstreams(int D, int order) {
    const int nd = order >= 1 ? D : 0;
    const int nq = order == 2 ? 1 : (order == 3 ? D * (D + 1) / 2 : 0);
    return 1 + nd + nq;
}

// flat-theta offsets (floats).  layer 0: D->H, layers 1..L: H->H, layer L+1: H->O
__host__ __device__ inline int64_t insr_w_offset(const SirenDims &dm, int layer) {
    if (layer == 0) return 0;
    const int64_t first = (int64_t)dm.H * dm.D + dm.H;
    return first + (int64_t)(layer - 1) * ((int64_t)dm.H * dm.H + dm.H);
}
__host__ __device__ inline int64_t insr_b_offset(const SirenDims &dm, int layer) {
    const int64_t w = insr_w_offset(dm, layer);
    if (layer == 0) return w + (int64_t)dm.H * dm.D;
    if (layer == dm.L + 1) return w + (int64_t)dm.O * dm.H;
    return w + (int64_t)dm.H * dm.H;
}
__host__ __device__ inline int64_t insr_theta_size(const SirenDims &dm) {
    return insr_b_offset(dm, dm.L + 1) + dm.O;
}

// sin/cos of a (possibly large: |t| ~ 30 * |z|) argument.  Two-term Cody-Waite reduction by
// 2*pi keeps the MUFU approximations inside [-pi, pi] where their absolute error is
// ~2^-21.4, so the pair costs 2 MUFU + 4 FP32 ops and is computed ONCE per activation for
// every stream (the reference evaluates sin/cos of the same activations dozens of times per
// iteration across its autograd sweeps).
__device__ __forceinline__ void insr_sincos(float t, float &s, float &c) {
    const float n = rintf(t * 0.15915494309189535f);
    float r = fmaf(n, -6.2831854820251465f, t);
    r = fmaf(n, 1.7484555349182897e-07f, r);
    s = __sinf(r);
    c = __cosf(r);
}

// forward activation of all streams of one neuron; z = pre-activations (omega folded)
template <int D, int ORDER>
__device__ __forceinline__ void insr_sine_fwd(const float *z, float *a) {
    typedef StreamCfg<D, ORDER> C;
    float s, c;
    insr_sincos(z[0], s, c);
    a[0] = s;
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < C::ND; ++d) a[1 + d] = c * z[1 + d];
    if constexpr (ORDER == 2) {
        float quad = 0.f;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < C::ND; ++d) quad = fmaf(z[1 + d], z[1 + d], quad);
        a[1 + C::ND] = c * z[1 + C::ND] - s * quad;
    }
    if constexpr (ORDER == 3) {
        int q = 0;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) {
            INSR_PRAGMA_UNROLL
            for (int e = d; e < D; ++e) {
                a[1 + C::ND + q] = c * z[1 + C::ND + q] - s * (z[1 + d] * z[1 + e]);
                ++q;
            }
        }
    }
}

// adjoint of the activation: zb = (d a / d z)^T ab
template <int D, int ORDER>
__device__ __forceinline__ void insr_sine_bwd(const float *z, const float *ab, float *zb) {
    typedef StreamCfg<D, ORDER> C;
    float s, c;
    insr_sincos(z[0], s, c);
    float zb0 = c * ab[0];
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < C::ND; ++d) {
        zb[1 + d] = c * ab[1 + d];
        zb0 = fmaf(-s * z[1 + d], ab[1 + d], zb0);
    }
    if constexpr (ORDER == 2) {
        const float aq = ab[1 + C::ND];
        float quad = 0.f;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < C::ND; ++d) quad = fmaf(z[1 + d], z[1 + d], quad);
        zb[1 + C::ND] = c * aq;
        zb0 = fmaf(aq, -(s * z[1 + C::ND] + c * quad), zb0);
        const float m = -2.f * s * aq;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < C::ND; ++d) zb[1 + d] = fmaf(m, z[1 + d], zb[1 + d]);
    }
    if constexpr (ORDER == 3) {
        int q = 0;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) {
            INSR_PRAGMA_UNROLL
            for (int e = d; e < D; ++e) {
                const float aq = ab[1 + C::ND + q];
                zb[1 + C::ND + q] = c * aq;
                zb0 = fmaf(aq, -(s * z[1 + C::ND + q] + c * (z[1 + d] * z[1 + e])), zb0);
                const float m = -s * aq;
                if (d == e) {
                    zb[1 + d] = fmaf(2.f * m, z[1 + d], zb[1 + d]);
                } else {
                    zb[1 + d] = fmaf(m, z[1 + e], zb[1 + d]);
                    zb[1 + e] = fmaf(m, z[1 + d], zb[1 + e]);
                }
                ++q;
            }
        }
    }
    zb[0] = zb0;
}

// scatter the output-layer streams out[s] of output o of point n into y / jac / h2
template <int D, int O, int ORDER>
__device__ __forceinline__ void insr_store_outputs(int64_t n, int o, const float *out, float *y,
                                                   float *jac, float *h2) {
    typedef StreamCfg<D, ORDER> C;
    y[n * O + o] = out[0];
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < C::ND; ++d) jac[(n * O + o) * D + d] = out[1 + d];
    if constexpr (ORDER == 2) h2[n * O + o] = out[1 + C::ND];
    if constexpr (ORDER == 3) {
        int q = 0;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) {
            INSR_PRAGMA_UNROLL
            for (int e = d; e < D; ++e) {
                const float v = out[1 + C::ND + q];
                h2[((n * O + o) * D + d) * D + e] = v;
                h2[((n * O + o) * D + e) * D + d] = v;
                ++q;
            }
        }
    }
}

// gather the cotangents of output o of point n into g[s] (NULL pointers = zero)
template <int D, int O, int ORDER>
__device__ __forceinline__ void insr_load_cotangents(int64_t n, int o, const float *gy,
                                                     const float *gjac, const float *gh2, float *g) {
    typedef StreamCfg<D, ORDER> C;
    g[0] = gy ? gy[n * O + o] : 0.f;
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < C::ND; ++d) g[1 + d] = gjac ? gjac[(n * O + o) * D + d] : 0.f;
    if constexpr (ORDER == 2) g[1 + C::ND] = gh2 ? gh2[n * O + o] : 0.f;
    if constexpr (ORDER == 3) {
        int q = 0;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) {
            INSR_PRAGMA_UNROLL
            for (int e = d; e < D; ++e) {
                float v = 0.f;
                if (gh2) {
                    v = gh2[((n * O + o) * D + d) * D + e];
                    if (e != d) v += gh2[((n * O + o) * D + e) * D + d];
                }
                g[1 + C::ND + q] = v;
                ++q;
            }
        }
    }
}
