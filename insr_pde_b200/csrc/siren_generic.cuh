// siren_generic.cuh -- the any-shape SIREN kernels (D,O <= 3, H <= 512, any L, any order).
//
// Role: correctness path and the path for widths whose weights do not fit in shared memory
// (H > INSR_FUSED_MAX_H).  One thread owns one collocation point; the per-layer stream
// activations live in a caller-provided workspace laid out [stream][neuron][slot] so that a
// warp's accesses are coalesced; weights are read through the read-only path (every lane of
// a warp reads the same address -> one broadcast transaction).  The parameter gradient is a
// reduction over points, done by separate tiled-reduction kernels over the workspace.
//
// Reference behaviour restated: base/networks.py:50-71 (forward), base/diff_ops.py:33-82
// (derivative streams), base/baseModel.py:77 (backward).
#pragma once
#include "siren_common.cuh"

#define INSR_GEN_JB 4   // output neurons per register block
#define INSR_GEN_THREADS 128

// ws element (s, k) of buffer `buf` for this slot: buf[(s*H + k) * T]
template <int D, int O, int ORDER>
__global__ void __launch_bounds__(INSR_GEN_THREADS)
k_generic_fwd(SirenDims dm, const float *__restrict__ theta, const float *__restrict__ x, int64_t N,
              float *__restrict__ y, float *__restrict__ jac, float *__restrict__ h2,
              float *__restrict__ ws, int T) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= T) return;
    const int H = dm.H, L = dm.L;
    const float w = dm.omega;
    float *bufA = ws + slot;
    float *bufB = ws + (size_t)S * H * T + slot;

    for (int64_t n = slot; n < N; n += T) {
        float xv[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) xv[d] = x[n * D + d];

        // ---- first sine layer: tangents are the columns of W1, second order is zero
        {
            const float *W1 = theta;
            const float *b1 = theta + (size_t)H * D;
            for (int j = 0; j < H; ++j) {
                float z[S], a[S];
                float acc = __ldg(b1 + j);
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) acc = fmaf(__ldg(W1 + j * D + d), xv[d], acc);
                z[0] = w * acc;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < C::ND; ++d) z[1 + d] = w * __ldg(W1 + j * D + d);
                INSR_PRAGMA_UNROLL
                for (int q = 0; q < C::NQ; ++q) z[1 + C::ND + q] = 0.f;
                insr_sine_fwd<D, ORDER>(z, a);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) bufA[(size_t)(s * H + j) * T] = a[s];
            }
        }
        float *in = bufA, *out = bufB;
        // ---- hidden sine layers
        for (int l = 1; l <= L; ++l) {
            const float *W = theta + insr_w_offset(dm, l);
            const float *b = theta + insr_b_offset(dm, l);
            for (int j0 = 0; j0 < H; j0 += INSR_GEN_JB) {
                float acc[INSR_GEN_JB][S];
                INSR_PRAGMA_UNROLL
                for (int jj = 0; jj < INSR_GEN_JB; ++jj)
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) acc[jj][s] = 0.f;
                for (int k = 0; k < H; ++k) {
                    float av[S];
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) av[s] = in[(size_t)(s * H + k) * T];
                    INSR_PRAGMA_UNROLL
                    for (int jj = 0; jj < INSR_GEN_JB; ++jj) {
                        const int j = (j0 + jj < H) ? (j0 + jj) : (H - 1);
                        const float wv = __ldg(W + (size_t)j * H + k);
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s) acc[jj][s] = fmaf(wv, av[s], acc[jj][s]);
                    }
                }
                INSR_PRAGMA_UNROLL
                for (int jj = 0; jj < INSR_GEN_JB; ++jj) {
                    const int j = j0 + jj;
                    if (j < H) {
                        float z[S], a[S];
                        z[0] = w * (acc[jj][0] + __ldg(b + j));
                        INSR_PRAGMA_UNROLL
                        for (int s = 1; s < S; ++s) z[s] = w * acc[jj][s];
                        insr_sine_fwd<D, ORDER>(z, a);
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s) out[(size_t)(s * H + j) * T] = a[s];
                    }
                }
            }
            float *t = in; in = out; out = t;
        }
        // ---- output layer (linear)
        {
            const float *Wo = theta + insr_w_offset(dm, L + 1);
            const float *bo = theta + insr_b_offset(dm, L + 1);
            float acc[O][S];
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) acc[o][s] = 0.f;
            for (int j = 0; j < H; ++j) {
                float av[S];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) av[s] = in[(size_t)(s * H + j) * T];
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o) {
                    const float wv = __ldg(Wo + o * H + j);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) acc[o][s] = fmaf(wv, av[s], acc[o][s]);
                }
            }
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                acc[o][0] += __ldg(bo + o);
                insr_store_outputs<D, O, ORDER>(n, o, acc[o], y, jac, h2);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// backward, phase 1: per point forward-with-tape + reverse sweep.  Workspace per chunk of T
// slots (floats):   Z[(L+1)][S][H][T]  pre-activations, overwritten in place by their
//                                      adjoints zbar during the sweep
//                   A[(L+1)][S][H][T]  post-activations (inputs of the next layer)
//                   G[O][S][T]         output-layer cotangents
// ------------------------------------------------------------------------------------
__host__ __device__ inline size_t insr_gen_bwd_ws_floats(int S, int H, int L, int O, int T) {
    return ((size_t)2 * (L + 1) * S * H + (size_t)O * S) * T;
}

template <int D, int O, int ORDER>
__global__ void __launch_bounds__(INSR_GEN_THREADS)
k_generic_bwd_sweep(SirenDims dm, const float *__restrict__ theta, const float *__restrict__ x,
                    int64_t N, int64_t n0, const float *__restrict__ gy,
                    const float *__restrict__ gjac, const float *__restrict__ gh2,
                    float *__restrict__ gx, float *__restrict__ ws, int T) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = n0 + slot;
    if (slot >= T || n >= N) return;
    const int H = dm.H, L = dm.L;
    const float w = dm.omega;
    const size_t layer_stride = (size_t)S * H * T;
    float *Zt = ws + slot;
    float *At = ws + (size_t)(L + 1) * layer_stride + slot;
    float *Gt = ws + (size_t)2 * (L + 1) * layer_stride + slot;

    float xv[D];
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < D; ++d) xv[d] = x[n * D + d];

    // ---- forward with tape
    {
        const float *W1 = theta;
        const float *b1 = theta + (size_t)H * D;
        for (int j = 0; j < H; ++j) {
            float z[S], a[S];
            float acc = __ldg(b1 + j);
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) acc = fmaf(__ldg(W1 + j * D + d), xv[d], acc);
            z[0] = w * acc;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < C::ND; ++d) z[1 + d] = w * __ldg(W1 + j * D + d);
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < C::NQ; ++q) z[1 + C::ND + q] = 0.f;
            insr_sine_fwd<D, ORDER>(z, a);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                Zt[(size_t)(s * H + j) * T] = z[s];
                At[(size_t)(s * H + j) * T] = a[s];
            }
        }
    }
    for (int l = 1; l <= L; ++l) {
        const float *W = theta + insr_w_offset(dm, l);
        const float *b = theta + insr_b_offset(dm, l);
        const float *in = At + (size_t)(l - 1) * layer_stride;
        float *zo = Zt + (size_t)l * layer_stride;
        float *ao = At + (size_t)l * layer_stride;
        for (int j0 = 0; j0 < H; j0 += INSR_GEN_JB) {
            float acc[INSR_GEN_JB][S];
            INSR_PRAGMA_UNROLL
            for (int jj = 0; jj < INSR_GEN_JB; ++jj)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) acc[jj][s] = 0.f;
            for (int k = 0; k < H; ++k) {
                float av[S];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) av[s] = in[(size_t)(s * H + k) * T];
                INSR_PRAGMA_UNROLL
                for (int jj = 0; jj < INSR_GEN_JB; ++jj) {
                    const int j = (j0 + jj < H) ? (j0 + jj) : (H - 1);
                    const float wv = __ldg(W + (size_t)j * H + k);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) acc[jj][s] = fmaf(wv, av[s], acc[jj][s]);
                }
            }
            INSR_PRAGMA_UNROLL
            for (int jj = 0; jj < INSR_GEN_JB; ++jj) {
                const int j = j0 + jj;
                if (j < H) {
                    float z[S], a[S];
                    z[0] = w * (acc[jj][0] + __ldg(b + j));
                    INSR_PRAGMA_UNROLL
                    for (int s = 1; s < S; ++s) z[s] = w * acc[jj][s];
                    insr_sine_fwd<D, ORDER>(z, a);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) {
                        zo[(size_t)(s * H + j) * T] = z[s];
                        ao[(size_t)(s * H + j) * T] = a[s];
                    }
                }
            }
        }
    }

    // ---- output-layer cotangents
    float g[O][S];
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o) {
        insr_load_cotangents<D, O, ORDER>(n, o, gy, gjac, gh2, g[o]);
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) Gt[(size_t)(o * S + s) * T] = g[o][s];
    }

    // ---- reverse: last sine layer receives Wout^T g
    {
        const float *Wo = theta + insr_w_offset(dm, L + 1);
        float *zl = Zt + (size_t)L * layer_stride;
        for (int k = 0; k < H; ++k) {
            float ab[S], z[S], zb[S];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) ab[s] = 0.f;
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                const float wv = __ldg(Wo + o * H + k);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) ab[s] = fmaf(wv, g[o][s], ab[s]);
            }
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) z[s] = zl[(size_t)(s * H + k) * T];
            insr_sine_bwd<D, ORDER>(z, ab, zb);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) zl[(size_t)(s * H + k) * T] = zb[s];
        }
    }
    // ---- reverse through hidden layers: abar_{l-1} = omega * W_l^T zbar_l, then activation adjoint
    for (int l = L; l >= 1; --l) {
        const float *W = theta + insr_w_offset(dm, l);
        const float *zbl = Zt + (size_t)l * layer_stride;
        float *zprev = Zt + (size_t)(l - 1) * layer_stride;
        for (int k0 = 0; k0 < H; k0 += INSR_GEN_JB) {
            float acc[INSR_GEN_JB][S];
            INSR_PRAGMA_UNROLL
            for (int kk = 0; kk < INSR_GEN_JB; ++kk)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) acc[kk][s] = 0.f;
            for (int j = 0; j < H; ++j) {
                float zv[S];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) zv[s] = zbl[(size_t)(s * H + j) * T];
                INSR_PRAGMA_UNROLL
                for (int kk = 0; kk < INSR_GEN_JB; ++kk) {
                    const int k = (k0 + kk < H) ? (k0 + kk) : (H - 1);
                    const float wv = __ldg(W + (size_t)j * H + k);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) acc[kk][s] = fmaf(wv, zv[s], acc[kk][s]);
                }
            }
            INSR_PRAGMA_UNROLL
            for (int kk = 0; kk < INSR_GEN_JB; ++kk) {
                const int k = k0 + kk;
                if (k < H) {
                    float ab[S], z[S], zb[S];
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) {
                        ab[s] = w * acc[kk][s];
                        z[s] = zprev[(size_t)(s * H + k) * T];
                    }
                    insr_sine_bwd<D, ORDER>(z, ab, zb);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) zprev[(size_t)(s * H + k) * T] = zb[s];
                }
            }
        }
    }
    // ---- gradient w.r.t. the point (exact: every derivative stream funnels into zbar_0 of layer 1)
    if (gx) {
        const float *W1 = theta;
        float acc[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) acc[d] = 0.f;
        for (int j = 0; j < H; ++j) {
            const float zb0 = Zt[(size_t)j * T];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) acc[d] = fmaf(__ldg(W1 + j * D + d), zb0, acc[d]);
        }
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) gx[n * D + d] = w * acc[d];
    }
}

// ------------------------------------------------------------------------------------
// backward, phase 2a: hidden-layer weight gradients.   gW_l[j][k] += omega * sum_{t<nv} sum_s
// zbar_l[s][j][t] * a_{l-1}[s][k][t].   32x32 output tile per CTA, reduction over slots in
// chunks of 32 staged through shared memory; grid.z = L * zsplit.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_generic_wgrad_hidden(SirenDims dm, int S, const float *__restrict__ ws, int T, int nv, int zsplit,
                       float *__restrict__ gtheta) {
    const int H = dm.H, L = dm.L;
    const int l = 1 + blockIdx.z / zsplit;
    const int slice = blockIdx.z % zsplit;
    const int j0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const size_t layer_stride = (size_t)S * H * T;
    const float *Zb = ws + (size_t)l * layer_stride;
    const float *Ap = ws + (size_t)(L + 1) * layer_stride + (size_t)(l - 1) * layer_stride;
    __shared__ float Zs[32][33];
    __shared__ float As[32][33];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;   // 16 x 16 threads, 2 x 2 outputs each
    const int lt = tid % 32, lr = tid / 32;   // loader mapping: 32 slots x 8 rows
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float bsum = 0.f;
    const int chunk = (nv + zsplit - 1) / zsplit;
    const int t_begin = slice * chunk;
    const int t_end = (t_begin + chunk < nv) ? (t_begin + chunk) : nv;
    for (int t0 = t_begin; t0 < t_end; t0 += 32) {
        for (int s = 0; s < S; ++s) {
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < 4; ++i) {
                const int r = lr + 8 * i;
                const int t = t0 + lt;
                const bool tv = t < t_end;
                Zs[r][lt] = (tv && j0 + r < H) ? Zb[(size_t)(s * H + j0 + r) * T + t] : 0.f;
                As[r][lt] = (tv && k0 + r < H) ? Ap[(size_t)(s * H + k0 + r) * T + t] : 0.f;
            }
            __syncthreads();
            INSR_PRAGMA_UNROLL
            for (int tt = 0; tt < 32; ++tt) {
                const float z0 = Zs[ty * 2][tt], z1 = Zs[ty * 2 + 1][tt];
                const float a0 = As[tx * 2][tt], a1 = As[tx * 2 + 1][tt];
                acc[0][0] = fmaf(z0, a0, acc[0][0]);
                acc[0][1] = fmaf(z0, a1, acc[0][1]);
                acc[1][0] = fmaf(z1, a0, acc[1][0]);
                acc[1][1] = fmaf(z1, a1, acc[1][1]);
            }
            if (s == 0 && blockIdx.y == 0 && tid < 32) {
                INSR_PRAGMA_UNROLL
                for (int tt = 0; tt < 32; ++tt) bsum += Zs[tid][tt];
            }
            __syncthreads();
        }
    }
    float *gW = gtheta + insr_w_offset(dm, l);
    float *gb = gtheta + insr_b_offset(dm, l);
    const float w = dm.omega;
    INSR_PRAGMA_UNROLL
    for (int a = 0; a < 2; ++a)
        INSR_PRAGMA_UNROLL
        for (int b = 0; b < 2; ++b) {
            const int j = j0 + ty * 2 + a, k = k0 + tx * 2 + b;
            if (j < H && k < H) atomicAdd(gW + (size_t)j * H + k, w * acc[a][b]);
        }
    if (blockIdx.y == 0 && tid < 32 && j0 + tid < H) atomicAdd(gb + j0 + tid, w * bsum);
}

// ------------------------------------------------------------------------------------
// backward, phase 2b: first- and output-layer gradients (thin: one warp per output element)
//   gW1[j][d] += omega * sum_t ( zbar_0[0][j][t] * x[n0+t][d] + zbar_0[1+d][j][t] )
//   gb1[j]    += omega * sum_t zbar_0[0][j][t]
//   gWo[o][j] += sum_t sum_s G[o][s][t] * a_L[s][j][t]
//   gbo[o]    += sum_t G[o][0][t]
// element index e: [0,H*D) gW1 | [H*D, H*D+H) gb1 | then O*H gWo | then O gbo
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_generic_wgrad_edge(SirenDims dm, int S, int ND, const float *__restrict__ ws, int T, int nv,
                     const float *__restrict__ x, int64_t n0, float *__restrict__ gtheta) {
    const int H = dm.H, L = dm.L, D = dm.D, O = dm.O;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x % 32;
    const int n_elem = H * D + H + O * H + O;
    if (warp >= n_elem) return;                     // warp-uniform
    const size_t layer_stride = (size_t)S * H * T;
    const float *Z0 = ws;
    const float *AL = ws + (size_t)(L + 1) * layer_stride + (size_t)L * layer_stride;
    const float *G = ws + (size_t)2 * (L + 1) * layer_stride;
    float sum = 0.f;
    float *dst;
    float scale = 1.f;
    int e = warp;
    if (e < H * D) {
        const int j = e / D, d = e % D;
        const float *z0 = Z0 + (size_t)j * T;
        const float *zd = (ND > 0) ? Z0 + (size_t)((1 + d) * H + j) * T : nullptr;
        for (int t = lane; t < nv; t += 32) {
            float v = z0[t] * x[(n0 + t) * D + d];
            if (zd) v += zd[t];
            sum += v;
        }
        dst = gtheta + insr_w_offset(dm, 0) + e;
        scale = dm.omega;
    } else if ((e -= H * D) < H) {
        const float *z0 = Z0 + (size_t)e * T;
        for (int t = lane; t < nv; t += 32) sum += z0[t];
        dst = gtheta + insr_b_offset(dm, 0) + e;
        scale = dm.omega;
    } else if ((e -= H) < O * H) {
        const int o = e / H, j = e % H;
        for (int s = 0; s < S; ++s) {
            const float *gp = G + (size_t)(o * S + s) * T;
            const float *ap = AL + (size_t)(s * H + j) * T;
            for (int t = lane; t < nv; t += 32) sum = fmaf(gp[t], ap[t], sum);
        }
        dst = gtheta + insr_w_offset(dm, L + 1) + e;
    } else {
        e -= O * H;
        const float *gp = G + (size_t)(e * S) * T;
        for (int t = lane; t < nv; t += 32) sum += gp[t];
        dst = gtheta + insr_b_offset(dm, L + 1) + e;
    }
    INSR_PRAGMA_UNROLL
    for (int m = 16; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    if (lane == 0) atomicAdd(dst, scale * sum);
}
