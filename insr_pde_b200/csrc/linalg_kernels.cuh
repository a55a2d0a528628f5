// linalg_kernels.cuh -- batched 2x2 / 3x3 singular value decomposition and the fused ARAP / volume energy of the
// elasticity closure.
//
// Reference semantics restated (elasticity/model.py:143-149): F = d(net(x) + x)/dx  (N, D, D);
//   U, S, V = torch.svd(F);  E_arap = r_a sum (S - 1)^2;  E_volume = r_v sum (prod(S) - 1)^2
// One thread per matrix, one-sided (Hestenes) Jacobi on the columns of F: rotations J_k from the right until the
// columns are orthogonal, F J_1 J_2 ... = U diag(S), V = J_1 J_2 ...; backward stable (no F^T F squaring), exact after
// one rotation for D = 2, quadratically convergent for D = 3.  d sigma_k / dF = u_k v_k^T, so the adjoint of any
// function of the singular values is  U diag(dE/dsigma) V^T  -- computed in the same thread.
#pragma once
#include "insr_platform.h"
#include <math.h>

template <int D>
__device__ inline void insr_svd_small(const float (&F)[D][D], float (&U)[D][D], float (&S)[D], float (&V)[D][D]) {
    float A[D][D];
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < D; ++i)
        INSR_PRAGMA_UNROLL
        for (int j = 0; j < D; ++j) { A[i][j] = F[i][j]; V[i][j] = (i == j) ? 1.f : 0.f; }
    constexpr int SWEEPS = (D == 2) ? 2 : 6;
    for (int sweep = 0; sweep < SWEEPS; ++sweep) {
        INSR_PRAGMA_UNROLL
        for (int p = 0; p < D - 1; ++p) {
            INSR_PRAGMA_UNROLL
            for (int q = p + 1; q < D; ++q) {
                float alpha = 0.f, beta = 0.f, gamma = 0.f;
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < D; ++i) { alpha = fmaf(A[i][p], A[i][p], alpha); beta = fmaf(A[i][q], A[i][q], beta); gamma = fmaf(A[i][p], A[i][q], gamma); }
                if (fabsf(gamma) > 1e-30f && fabsf(gamma) > 1e-9f * sqrtf(alpha * beta)) {
                    const float zeta = (beta - alpha) / (2.f * gamma);
                    const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
                    const float c = 1.f / sqrtf(fmaf(t, t, 1.f)), s = c * t;
                    INSR_PRAGMA_UNROLL
                    for (int i = 0; i < D; ++i) {
                        const float ap = A[i][p], aq = A[i][q];
                        A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
                        const float vp = V[i][p], vq = V[i][q];
                        V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                    }
                }
            }
        }
    }
    INSR_PRAGMA_UNROLL
    for (int k = 0; k < D; ++k) {
        float n2 = 0.f;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < D; ++i) n2 = fmaf(A[i][k], A[i][k], n2);
        S[k] = sqrtf(n2);
    }
    // descending order (torch.svd), columns of A and V follow
    INSR_PRAGMA_UNROLL
    for (int pass = 0; pass < D - 1; ++pass) {
        INSR_PRAGMA_UNROLL
        for (int k = 0; k < D - 1 - pass; ++k) {
            if (S[k] < S[k + 1]) {
                const float ts = S[k]; S[k] = S[k + 1]; S[k + 1] = ts;
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < D; ++i) {
                    const float ta = A[i][k]; A[i][k] = A[i][k + 1]; A[i][k + 1] = ta;
                    const float tv = V[i][k]; V[i][k] = V[i][k + 1]; V[i][k + 1] = tv;
                }
            }
        }
    }
    // U = A diag(1/S); rank-deficient columns are completed to an orthonormal basis
    const float tol = 1e-7f * S[0];
    INSR_PRAGMA_UNROLL
    for (int k = 0; k < D; ++k) {
        const float inv = (S[k] > tol && S[k] > 0.f) ? 1.f / S[k] : 0.f;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < D; ++i) U[i][k] = A[i][k] * inv;
    }
    if (!(S[0] > 0.f)) {
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < D; ++i)
            INSR_PRAGMA_UNROLL
            for (int j = 0; j < D; ++j) U[i][j] = (i == j) ? 1.f : 0.f;
    } else if constexpr (D == 2) {
        if (!(S[1] > tol)) { U[0][1] = -U[1][0]; U[1][1] = U[0][0]; }
    } else {
        if (!(S[1] > tol)) {                      // rank 1: any unit vector orthogonal to u0
            const float ax = fabsf(U[0][0]), ay = fabsf(U[1][0]), az = fabsf(U[2][0]);
            float e[3] = {0.f, 0.f, 0.f};
            e[(ax <= ay && ax <= az) ? 0 : ((ay <= az) ? 1 : 2)] = 1.f;
            float w[3] = {U[1][0] * e[2] - U[2][0] * e[1], U[2][0] * e[0] - U[0][0] * e[2], U[0][0] * e[1] - U[1][0] * e[0]};
            const float nw = 1.f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            U[0][1] = w[0] * nw; U[1][1] = w[1] * nw; U[2][1] = w[2] * nw;
        }
        if (!(S[2] > tol)) {
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
        }
    }
}

template <int D>
__global__ void k_svd_small(const float *__restrict__ F, int64_t n, float *__restrict__ U, float *__restrict__ S,
                            float *__restrict__ V) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float f[D][D], u[D][D], s[D], v[D][D];
        INSR_PRAGMA_UNROLL
        for (int r = 0; r < D; ++r)
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < D; ++c) f[r][c] = F[i * D * D + r * D + c];
        insr_svd_small<D>(f, u, s, v);
        INSR_PRAGMA_UNROLL
        for (int r = 0; r < D; ++r) {
            S[i * D + r] = s[r];
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < D; ++c) {
                if (U) U[i * D * D + r * D + c] = u[r][c];
                if (V) V[i * D * D + r * D + c] = v[r][c];
            }
        }
    }
}

// energy += r_a sum_k (s_k - 1)^2 + r_v (prod s - 1)^2 over all matrices;  gF = U diag(dE/ds) V^T  (if gF != NULL)
template <int D>
__global__ void k_elastic_energy(const float *__restrict__ F, int64_t n, float ratio_arap, float ratio_volume,
                                 float *__restrict__ energy, float *__restrict__ gF) {
    float e_sum = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float f[D][D], u[D][D], s[D], v[D][D];
        INSR_PRAGMA_UNROLL
        for (int r = 0; r < D; ++r)
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < D; ++c) f[r][c] = F[i * D * D + r * D + c];
        insr_svd_small<D>(f, u, s, v);
        float prod = 1.f, arap = 0.f;
        INSR_PRAGMA_UNROLL
        for (int k = 0; k < D; ++k) { prod *= s[k]; arap = fmaf(s[k] - 1.f, s[k] - 1.f, arap); }
        e_sum += ratio_arap * arap + ratio_volume * (prod - 1.f) * (prod - 1.f);
        if (gF) {
            float ds[D];
            INSR_PRAGMA_UNROLL
            for (int k = 0; k < D; ++k) {
                float others = 1.f;
                INSR_PRAGMA_UNROLL
                for (int j = 0; j < D; ++j) if (j != k) others *= s[j];
                ds[k] = 2.f * ratio_arap * (s[k] - 1.f) + 2.f * ratio_volume * (prod - 1.f) * others;
            }
            INSR_PRAGMA_UNROLL
            for (int r = 0; r < D; ++r)
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < D; ++c) {
                    float acc = 0.f;
                    INSR_PRAGMA_UNROLL
                    for (int k = 0; k < D; ++k) acc = fmaf(u[r][k] * ds[k], v[c][k], acc);
                    gF[i * D * D + r * D + c] = acc;
                }
        }
    }
    INSR_PRAGMA_UNROLL
    for (int m = 16; m >= 1; m >>= 1) e_sum += __shfl_xor_sync(0xffffffffu, e_sum, m);
    if ((threadIdx.x & 31) == 0) atomicAdd(energy, e_sum);
}

// ---------------------------------------------------------------------------------------------------------------------
// Every term of ElasticityModel._solve_deformation (elasticity/model.py:127-189, elasticity/losses.py:6-39) in ONE kernel:
// the loss and its cotangents w.r.t. the trainable field's value y and Jacobian J at every row of the batch
//     rows [0, n)                    interior samples x:  q = y + x,  F = J + I,  qdot = (y - y_prev) / dt,
//                                    qdot_prev = (y_prev - y_pp) / dt   (the x of q, q_prev, q_pp cancels)
//         r_a sum_k (s_k - 1)^2 + r_v (prod s - 1)^2             arap / volume        (model.py:143-149)
//         r_k |qdot - qdot_prev|^2                                kinematics           (:151-153)
//         - dt qdot . f_ext                                       external             (:155-158)
//         - dt r_c [q_last < h] qdot_last (h - q_last)            plane collision      (losses.py:10-20)
//         - dt r_c [|q - c| < R] qdot . (q - c)                   sphere collision, 2-D form (losses.py:22-39; force =
//                                                                 r_c * dist * dir = r_c (q - c))
//     rows [n, n + n_left)           clamped left face:   r_l |y|^2                    (:160-162)
//     rows [n + n_left, n_all)       right face:          r_r |y - off|^2              (:164-171; off carries the sign)
// A ratio of zero switches a term off.  loss: device scalar, ACCUMULATED; gy (n_all, D), gJ (n_all, D, D; NULL when
// r_a = r_v = 0) are overwritten.
struct insr_elastic_terms_k {
    int64_t n, n_left, n_right;
    float dt, r_arap, r_volume, r_kin, r_left, r_right, r_plane, plane_height, r_sphere, radius;
    float ext[3], off[3], center[3];
};

template <int D>
__global__ void k_elastic_terms(insr_elastic_terms_k t, const float *__restrict__ y, const float *__restrict__ J,
                                const float *__restrict__ x, const float *__restrict__ y_prev, const float *__restrict__ y_pp,
                                float *__restrict__ loss, float *__restrict__ gy, float *__restrict__ gJ) {
    const int64_t n_all = t.n + t.n_left + t.n_right;
    const float inv_dt = 1.f / t.dt;
    float e_sum = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += (int64_t)gridDim.x * blockDim.x) {
        float yv[D], g[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) { yv[d] = y[i * D + d]; g[d] = 0.f; }
        if (i < t.n) {
            float qd[D], qdp[D], q[D];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) {
                const float yp = y_prev[i * D + d], ypp = y_pp[i * D + d];
                qd[d] = (yv[d] - yp) * inv_dt;
                qdp[d] = (yp - ypp) * inv_dt;
                q[d] = yv[d] + x[i * D + d];
            }
            if (J) {
                float f[D][D], u[D][D], s[D], v[D][D];
                INSR_PRAGMA_UNROLL
                for (int r = 0; r < D; ++r)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < D; ++c) f[r][c] = J[i * D * D + r * D + c] + (r == c ? 1.f : 0.f);
                insr_svd_small<D>(f, u, s, v);
                float prod = 1.f, arap = 0.f, ds[D];
                INSR_PRAGMA_UNROLL
                for (int k = 0; k < D; ++k) { prod *= s[k]; arap = fmaf(s[k] - 1.f, s[k] - 1.f, arap); }
                e_sum += t.r_arap * arap + t.r_volume * (prod - 1.f) * (prod - 1.f);
                INSR_PRAGMA_UNROLL
                for (int k = 0; k < D; ++k) {
                    float others = 1.f;
                    INSR_PRAGMA_UNROLL
                    for (int j = 0; j < D; ++j) if (j != k) others *= s[j];
                    ds[k] = 2.f * t.r_arap * (s[k] - 1.f) + 2.f * t.r_volume * (prod - 1.f) * others;
                }
                INSR_PRAGMA_UNROLL
                for (int r = 0; r < D; ++r)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < D; ++c) {
                        float acc = 0.f;
                        INSR_PRAGMA_UNROLL
                        for (int k = 0; k < D; ++k) acc = fmaf(u[r][k] * ds[k], v[c][k], acc);
                        gJ[i * D * D + r * D + c] = acc;
                    }
            }
            float dist2 = 0.f, qd_vec = 0.f;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) {
                const float a = qd[d] - qdp[d];
                e_sum = fmaf(t.r_kin * a, a, e_sum);
                g[d] = fmaf(2.f * t.r_kin * inv_dt, a, g[d]);
                e_sum = fmaf(-t.dt * qd[d], t.ext[d], e_sum);
                g[d] -= t.ext[d];
                const float vec = q[d] - t.center[d];
                dist2 = fmaf(vec, vec, dist2);
                qd_vec = fmaf(qd[d], vec, qd_vec);
            }
            if (t.r_plane != 0.f && q[D - 1] < t.plane_height) {
                const float gap = t.plane_height - q[D - 1];
                e_sum -= t.dt * t.r_plane * qd[D - 1] * gap;
                g[D - 1] -= t.dt * t.r_plane * (gap * inv_dt - qd[D - 1]);
            }
            if (t.r_sphere != 0.f && sqrtf(dist2) < t.radius) {
                e_sum -= t.dt * t.r_sphere * qd_vec;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) g[d] -= t.dt * t.r_sphere * ((q[d] - t.center[d]) * inv_dt + qd[d]);
            }
        } else {
            const bool left = i < t.n + t.n_left;
            const float r = left ? t.r_left : t.r_right;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) {
                const float a = yv[d] - (left ? 0.f : t.off[d]);
                e_sum = fmaf(r * a, a, e_sum);
                g[d] = 2.f * r * a;
            }
            if (gJ) {
                INSR_PRAGMA_UNROLL
                for (int k = 0; k < D * D; ++k) gJ[i * D * D + k] = 0.f;
            }
        }
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) gy[i * D + d] = g[d];
    }
    INSR_PRAGMA_UNROLL
    for (int m = 16; m >= 1; m >>= 1) e_sum += __shfl_xor_sync(0xffffffffu, e_sum, m);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, e_sum);
}
