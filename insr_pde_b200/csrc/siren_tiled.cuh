// siren_tiled.cuh -- tiled (shared-memory GEMM) SIREN kernels for hidden widths 32 < H <= 512
// (elasticity H = 66 / 68 / 128 and the 64-512 synthetic sweep).
//
// At these widths the weights no longer fit next to the activations in shared memory, and one
// hidden layer is a real GEMM:  Z[(point,stream)][j] = sum_k A[(point,stream)][k] W[j][k]  with
// M = S*N rows.  The family therefore runs LAYER BY LAYER with the stream activations in a
// caller-provided HBM workspace (arithmetic intensity 2*H^2*S / (16*S*H) = H/8 FLOP/B per layer,
// i.e. 8-64 FLOP/B here: still FP32-pipe bound on B200) and fuses everything elementwise into
// the GEMM epilogues:
//   forward  layer: GEMM + bias + sine-stream activation (sin/cos once) -> writes the
//                   pre-activations (tape) and the post-activations (next operand)
//   backward layer: data-gradient GEMM (Zbar W) + activation adjoint in the epilogue; the
//                   weight gradient is a second GEMM with the reduction over (point,stream) rows,
//                   split across CTAs and finished with red.global
// GEMM core: 256 threads, k-slabs of 8 staged in shared memory (transposed so that every
// operand read is a 64/128-bit LDS), register tiles of (TP points x S streams) x 8 columns
// = 8x8 for the common S, i.e. 4 FFMA per float delivered by shared memory (the balance point
// of the LSU pipe measured in profiles/r1_ncu_full_fused_v1.txt).
//
// Workspace layout (floats), all buffers [stream][point][HP] with HP = roundup(H, 8):
//   Zpre[l], Act[l]   l = 0..L      pre- / post-activations of sine layer l   (backward only: Zpre)
//   ZB[2]                           zbar ping-pong (backward)
//   G[O][S][NCp]                    output cotangents (backward)
// Points are processed in chunks of NC so that the workspace stays bounded.
#pragma once
#include "siren_tiled_api.h"
#include "siren_wide_tc.cuh"
#include "siren_mid_api.h"

namespace insr_tiled {

constexpr int BK = 8;       // reduction slab
constexpr int TN = 8;       // output columns per thread
constexpr int NT = 256;     // threads per CTA
constexpr int PANEL = 128;  // max columns per CTA panel

template <int S> struct Tile { static constexpr int TP = (S <= 2) ? 4 : ((S <= 4) ? 2 : 1); };

struct Geo {
    int H, HP;        // HP = roundup(H, 8): row length of every activation buffer
    int CT, RT;       // column threads per panel, row threads
    int BN;           // columns per panel = 8 * CT
    int panels;       // ceil(HP / BN)
    int BP;           // points per CTA tile = TP * RT
    int AST, BST;     // shared-memory row strides (floats)
};

__host__ __device__ inline Geo make_geo(int H, int S, int TP) {
    Geo g;
    g.H = H; g.HP = (H + 7) & ~7;
    g.panels = (g.HP + PANEL - 1) / PANEL;
    const int cols = (g.HP / 8 + g.panels - 1) / g.panels;    // column threads, balanced over panels
    g.CT = cols; g.BN = 8 * cols;
    g.RT = NT / g.CT;
    g.BP = TP * g.RT;
    g.AST = S * g.BP + 4;       // == 4 mod 8: conflict-free transposing stores
    g.BST = g.BN + 4;
    return g;
}
__host__ __device__ inline size_t gemm_smem_floats(const Geo &g) { return (size_t)BK * (g.AST + g.BST); }

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

// ---------------------------------------------------------------------------------------------
// GEMM main loop.  acc[s][t][c] += sum_k A[s][p0+t][k] * B[k][c0+c]
//   A rows come from `Ain` ([S][NCp][HP], k fastest);  B(k, col) is supplied by `bfetch`.
// ---------------------------------------------------------------------------------------------
template <int S, int TP, typename BFetch>
__device__ __forceinline__ void gemm_mainloop(const Geo &g, const float *__restrict__ Ain, int64_t NCp,
                                              int64_t p_tile0, int K, BFetch bfetch, float *smA, float *smB,
                                              int rt, int ct, bool active, float (&acc)[S][TP][TN]) {
    INSR_PRAGMA_UNROLL
    for (int s = 0; s < S; ++s)
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < TP; ++t)
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < TN; ++c) acc[s][t][c] = 0.f;
    const int tid = threadIdx.x;
    // per-thread staging assignments (independent of the slab): <= 4 float4 of A, <= 4 scalars of B
    constexpr int MAXA = 4, MAXB = 4;
    const int a_items = S * g.BP * 2;          // float4 loads per slab: (row, half)
    const int b_items = g.BN * BK;             // scalar loads per slab
    const float *a_src[MAXA];
    int a_dst[MAXA], b_kk[MAXB], b_col[MAXB];
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < MAXA; ++i) {
        const int idx = tid + i * NT;
        const int r = idx >> 1, q = idx & 1;
        const int s = r / g.BP, p = r - s * g.BP;
        a_src[i] = (idx < a_items) ? Ain + ((int64_t)s * NCp + p_tile0 + p) * g.HP + 4 * q : nullptr;
        a_dst[i] = (4 * q) * g.AST + r;
    }
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < MAXB; ++i) {
        const int idx = tid + i * NT;
        b_kk[i] = (idx < b_items) ? idx / g.BN : -1;
        b_col[i] = idx - (idx / g.BN) * g.BN;
    }
    float4 ra[MAXA];
    float rb[MAXB];
    auto gload = [&](int k0) {                 // global -> registers (next slab, overlaps the FFMA loop)
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < MAXA; ++i)
            if (a_src[i]) ra[i] = ld4(a_src[i] + k0);
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < MAXB; ++i)
            if (b_kk[i] >= 0) rb[i] = bfetch(k0 + b_kk[i], b_col[i]);
    };
    auto sstore = [&]() {                      // registers -> shared (A transposed: smA[kk][s*BP + p])
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < MAXA; ++i)
            if (a_src[i]) {
                float *dst = smA + a_dst[i];
                dst[0] = ra[i].x; dst[g.AST] = ra[i].y; dst[2 * g.AST] = ra[i].z; dst[3 * g.AST] = ra[i].w;
            }
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < MAXB; ++i)
            if (b_kk[i] >= 0) smB[b_kk[i] * g.BST + b_col[i]] = rb[i];
    };
    gload(0);
    sstore();
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += BK) {
        const bool more = k0 + BK < K;
        if (more) gload(k0 + BK);
        if (active) {
            const float *ap = smA + TP * rt;
            const float *bp = smB + 4 * ct;          // columns 4ct..4ct+3 and 4CT+4ct..: conflict-free LDS.128
            INSR_PRAGMA_UNROLL
            for (int kk = 0; kk < BK; ++kk) {
                float a[S][TP];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    if (TP == 4) {
                        const float4 v = ld4(ap + kk * g.AST + s * g.BP);
                        a[s][0] = v.x; a[s][1 % TP] = v.y; a[s][2 % TP] = v.z; a[s][3 % TP] = v.w;
                    } else if (TP == 2) {
                        const float2 v = *reinterpret_cast<const float2 *>(ap + kk * g.AST + s * g.BP);
                        a[s][0] = v.x; a[s][1 % TP] = v.y;
                    } else {
                        a[s][0] = ap[kk * g.AST + s * g.BP];
                    }
                }
                const float4 b0 = ld4(bp + kk * g.BST), b1 = ld4(bp + kk * g.BST + 4 * g.CT);
                const float b[TN] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int t = 0; t < TP; ++t)
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < TN; ++c) acc[s][t][c] = fmaf(a[s][t], b[c], acc[s][t][c]);
            }
        }
        __syncthreads();                       // everyone is done reading the current slab
        if (more) {
            sstore();
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// T1: first sine layer (D -> H), elementwise.  One thread per (point, 4 neurons).
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER>
__global__ void __launch_bounds__(NT) k_tiled_layer0(SirenDims dm, int HP, const float *__restrict__ theta,
                                                     const float *__restrict__ x, int64_t n0, int nv, int64_t NCp,
                                                     float *__restrict__ Zpre, float *__restrict__ Act) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    const int quads = HP / 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= NCp * quads) return;
    const int64_t p = gid / quads;
    const int j4 = (int)(gid - p * quads) * 4;
    float xv[D];
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < D; ++d) xv[d] = (p < nv) ? x[(n0 + p) * D + d] : 0.f;
    const float *W1 = theta, *b1 = theta + (size_t)dm.H * D;
    float zo[S][4], ao[S][4];
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) {
        const int j = j4 + c;
        float z[S], a[S];
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) z[s] = 0.f;
        if (j < dm.H) {
            float acc = __ldg(b1 + j);
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) acc = fmaf(__ldg(W1 + j * D + d), xv[d], acc);
            z[0] = dm.omega * acc;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < C::ND; ++d) z[1 + d] = dm.omega * __ldg(W1 + j * D + d);
        }
        insr_sine_fwd<D, ORDER>(z, a);
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) { zo[s][c] = z[s]; ao[s][c] = a[s]; }
    }
    INSR_PRAGMA_UNROLL
    for (int s = 0; s < S; ++s) {
        const int64_t off = ((int64_t)s * NCp + p) * HP + j4;
        if (Zpre) st4(Zpre + off, zo[s][0], zo[s][1], zo[s][2], zo[s][3]);
        st4(Act + off, ao[s][0], ao[s][1], ao[s][2], ao[s][3]);
    }
}

// ---------------------------------------------------------------------------------------------
// T2: hidden sine layer forward: GEMM + bias + activation.  grid = (NCp / BP, panels)
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER>
__global__ void __launch_bounds__(NT, (StreamCfg<D, ORDER>::S <= 3) ? 2 : 1) k_tiled_fwd(SirenDims dm, Geo g, const float *__restrict__ W,
                                                  const float *__restrict__ bias, const float *__restrict__ Ain,
                                                  int64_t NCp, float *__restrict__ Zpre, float *__restrict__ Aout) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int TP = Tile<S>::TP;
    INSR_DYN_SMEM(float, sm);
    float *smA = sm, *smB = sm + BK * g.AST;
    const int tid = threadIdx.x, ct = tid % g.CT, rt = tid / g.CT;
    const bool active = rt < g.RT;
    const int64_t p_tile0 = (int64_t)blockIdx.x * g.BP;
    const int j0 = blockIdx.y * g.BN;
    const int H = dm.H;
    float acc[S][TP][TN];
    auto bfetch = [&](int k, int col) -> float {     // B[k][col] = W[j0+col][k]
        const int j = j0 + col;
        return (j < H && k < H) ? __ldg(W + (size_t)j * H + k) : 0.f;
    };
    gemm_mainloop<S, TP>(g, Ain, NCp, p_tile0, g.HP, bfetch, smA, smB, rt, ct, active, acc);
    if (!active) return;
    const int jcA = j0 + 4 * ct, jcB = j0 + 4 * g.CT + 4 * ct;      // the thread's two groups of 4 columns
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TP; ++t) {
        const int64_t p = p_tile0 + TP * rt + t;
        float zo[S][TN], ao[S][TN];
        INSR_PRAGMA_UNROLL
        for (int c = 0; c < TN; ++c) {
            const int j = (c < 4) ? (jcA + c) : (jcB + c - 4);
            float z[S], a[S];
            const float bj = (j < H) ? __ldg(bias + j) : 0.f;
            z[0] = dm.omega * (acc[0][t][c] + bj);
            INSR_PRAGMA_UNROLL
            for (int s = 1; s < S; ++s) z[s] = dm.omega * acc[s][t][c];
            insr_sine_fwd<D, ORDER>(z, a);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) { zo[s][c] = z[s]; ao[s][c] = a[s]; }
        }
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) {
            const int64_t row = ((int64_t)s * NCp + p) * g.HP;
            if (jcA < g.HP) {
                if (Zpre) st4(Zpre + row + jcA, zo[s][0], zo[s][1], zo[s][2], zo[s][3]);
                st4(Aout + row + jcA, ao[s][0], ao[s][1], ao[s][2], ao[s][3]);
            }
            if (jcB < g.HP) {
                if (Zpre) st4(Zpre + row + jcB, zo[s][4], zo[s][5], zo[s][6], zo[s][7]);
                st4(Aout + row + jcB, ao[s][4], ao[s][5], ao[s][6], ao[s][7]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// T3: output layer forward.  One warp per point: lanes stride the neurons, shuffle-reduce O*S sums.
// ---------------------------------------------------------------------------------------------
template <int D, int O, int ORDER>
__global__ void __launch_bounds__(NT) k_tiled_out_fwd(SirenDims dm, int HP, const float *__restrict__ theta,
                                                      const float *__restrict__ Act, int64_t NCp, int64_t n0, int nv,
                                                      float *__restrict__ y, float *__restrict__ jac,
                                                      float *__restrict__ h2) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= nv) return;                                    // warp-uniform
    const float *Wo = theta + insr_w_offset(dm, dm.L + 1);
    const float *bo = theta + insr_b_offset(dm, dm.L + 1);
    float out[O][S];
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o)
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) out[o][s] = 0.f;
    for (int j = lane; j < dm.H; j += 32) {
        float a[S];
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) a[s] = Act[((int64_t)s * NCp + p) * HP + j];
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) {
            const float w = __ldg(Wo + o * dm.H + j);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) out[o][s] = fmaf(w, a[s], out[o][s]);
        }
    }
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o) {
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) {
            float v = out[o][s];
            INSR_PRAGMA_UNROLL
            for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
            out[o][s] = v;
        }
        out[o][0] += __ldg(bo + o);
        if (lane == o) insr_store_outputs<D, O, ORDER>(n0 + p, o, out[o], y, jac, h2);
    }
}

// ---------------------------------------------------------------------------------------------
// T4: output layer backward.  One thread per (point, 4 neurons): cotangent of the last sine layer
// Wo^T g, activation adjoint with the tape -> zbar_L; also stores g into G[o][s][NCp] for the
// output-layer weight gradient.
// ---------------------------------------------------------------------------------------------
template <int D, int O, int ORDER>
__global__ void __launch_bounds__(NT) k_tiled_out_bwd(SirenDims dm, int HP, const float *__restrict__ theta,
                                                      const float *__restrict__ gy, const float *__restrict__ gjac,
                                                      const float *__restrict__ gh2, int64_t n0, int nv, int64_t NCp,
                                                      const float *__restrict__ ZpreL, float *__restrict__ ZB,
                                                      float *__restrict__ G) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    const int quads = HP / 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= NCp * quads) return;
    const int64_t p = gid / quads;
    const int q = (int)(gid - p * quads);
    const int j4 = q * 4;
    float g[O][S];
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o) {
        if (p < nv) {
            insr_load_cotangents<D, O, ORDER>(n0 + p, o, gy, gjac, gh2, g[o]);
        } else {
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) g[o][s] = 0.f;
        }
        if (q == 0) {
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) G[((int64_t)o * S + s) * NCp + p] = g[o][s];
        }
    }
    const float *Wo = theta + insr_w_offset(dm, dm.L + 1);
    float zb[S][4];
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) {
        const int j = j4 + c;
        float ab[S], z[S], zo[S];
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) ab[s] = 0.f;
        if (j < dm.H) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                const float w = __ldg(Wo + o * dm.H + j);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) ab[s] = fmaf(w, g[o][s], ab[s]);
            }
        }
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) z[s] = ZpreL[((int64_t)s * NCp + p) * HP + j];
        insr_sine_bwd<D, ORDER>(z, ab, zo);
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) zb[s][c] = zo[s];
    }
    INSR_PRAGMA_UNROLL
    for (int s = 0; s < S; ++s) st4(ZB + ((int64_t)s * NCp + p) * HP + j4, zb[s][0], zb[s][1], zb[s][2], zb[s][3]);
}

// ---------------------------------------------------------------------------------------------
// T5: data gradient of hidden layer l + activation adjoint of layer l-1:
//     abar[m][k] = omega * sum_j zbar_l[m][j] W_l[j][k]   ->   zbar_{l-1} = adj(Zpre_{l-1}, abar)
// ---------------------------------------------------------------------------------------------
template <int D, int ORDER>
__global__ void __launch_bounds__(NT, (StreamCfg<D, ORDER>::S <= 3) ? 2 : 1) k_tiled_dgrad(SirenDims dm, Geo g, const float *__restrict__ W,
                                                    const float *__restrict__ ZBin, int64_t NCp,
                                                    const float *__restrict__ ZprePrev, float *__restrict__ ZBout) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int TP = Tile<S>::TP;
    INSR_DYN_SMEM(float, sm);
    float *smA = sm, *smB = sm + BK * g.AST;
    const int tid = threadIdx.x, ct = tid % g.CT, rt = tid / g.CT;
    const bool active = rt < g.RT;
    const int64_t p_tile0 = (int64_t)blockIdx.x * g.BP;
    const int k0c = blockIdx.y * g.BN;
    const int H = dm.H;
    float acc[S][TP][TN];
    auto bfetch = [&](int j, int col) -> float {     // B[j][col] = W[j][k0c+col]  (reduction over j)
        const int k = k0c + col;
        return (j < H && k < H) ? __ldg(W + (size_t)j * H + k) : 0.f;
    };
    gemm_mainloop<S, TP>(g, ZBin, NCp, p_tile0, g.HP, bfetch, smA, smB, rt, ct, active, acc);
    if (!active) return;
    const int kcA = k0c + 4 * ct, kcB = k0c + 4 * g.CT + 4 * ct;
    const bool inA = kcA < g.HP, inB = kcB < g.HP;
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TP; ++t) {
        const int64_t p = p_tile0 + TP * rt + t;
        float zin[S][TN], zo[S][TN];
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) {
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v0 = inA ? ld4(ZprePrev + ((int64_t)s * NCp + p) * g.HP + kcA) : zero4;
            const float4 v1 = inB ? ld4(ZprePrev + ((int64_t)s * NCp + p) * g.HP + kcB) : zero4;
            zin[s][0] = v0.x; zin[s][1] = v0.y; zin[s][2] = v0.z; zin[s][3] = v0.w;
            zin[s][4] = v1.x; zin[s][5] = v1.y; zin[s][6] = v1.z; zin[s][7] = v1.w;
        }
        INSR_PRAGMA_UNROLL
        for (int c = 0; c < TN; ++c) {
            float ab[S], z[S], zb[S];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) { ab[s] = dm.omega * acc[s][t][c]; z[s] = zin[s][c]; }
            insr_sine_bwd<D, ORDER>(z, ab, zb);
            const int k = (c < 4) ? (kcA + c) : (kcB + c - 4);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) zo[s][c] = (k < H) ? zb[s] : 0.f;
        }
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) {
            const int64_t row = ((int64_t)s * NCp + p) * g.HP;
            if (inA) st4(ZBout + row + kcA, zo[s][0], zo[s][1], zo[s][2], zo[s][3]);
            if (inB) st4(ZBout + row + kcB, zo[s][4], zo[s][5], zo[s][6], zo[s][7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// T6: weight gradient of hidden layer l:  gW[j][k] += omega * sum_{s,p} zbar[s][p][j] * act[s][p][k]
// 64 x 64 output tile per CTA; the 256 threads form 4 groups that split the 32-row slabs of the
// reduction and are combined through shared memory.  grid = (ceil(HP/64), ceil(HP/64), row splits)
// ---------------------------------------------------------------------------------------------
constexpr int WG_T = 64;     // output tile
constexpr int WG_R = 32;     // reduction rows per slab

__global__ void __launch_bounds__(NT) k_tiled_wgrad(SirenDims dm, int HP, int S, const float *__restrict__ ZB,
                                                    const float *__restrict__ Act, int64_t NCp, int nv, int rsplit,
                                                    float *__restrict__ gW, float *__restrict__ gb) {
    __shared__ float4 smbuf4[2 * WG_R * (WG_T + 4) / 4];            // float4-typed for 16-byte alignment
    float *smbuf = reinterpret_cast<float *>(smbuf4);     // 4352 floats >= the 64x64 combine tile
    float (*smZ)[WG_T + 4] = reinterpret_cast<float (*)[WG_T + 4]>(smbuf);
    float (*smA)[WG_T + 4] = reinterpret_cast<float (*)[WG_T + 4]>(smbuf + WG_R * (WG_T + 4));
    const int tid = threadIdx.x;
    const int grp = tid >> 6, lt = tid & 63;
    const int jt = lt >> 3, kt = lt & 7;
    const int j0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int H = dm.H;
    float acc[8][8];
    INSR_PRAGMA_UNROLL
    for (int a = 0; a < 8; ++a)
        INSR_PRAGMA_UNROLL
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    float bsum = 0.f;
    // rows of the reduction are (s, p) with p < nv; slabs of WG_R points of one stream
    const int slabs_per_stream = (nv + WG_R - 1) / WG_R;
    const int total = S * slabs_per_stream;
    // software pipeline: the next slab's 2+2 float4 are fetched into registers while the current
    // slab is being contracted out of shared memory
    float4 rz[2], ra[2];
    auto gload = [&](int sl) {
        const int s = sl / slabs_per_stream;
        const int p0 = (sl - s * slabs_per_stream) * WG_R;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * NT;
            const int r = idx >> 4, c4 = (idx & 15) * 4;
            const int64_t row = (int64_t)s * NCp + p0 + r;
            const bool rv = p0 + r < nv;
            rz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            ra[i] = rz[i];
            if (rv && j0 + c4 < HP) rz[i] = ld4(ZB + row * HP + j0 + c4);
            if (rv && k0 + c4 < HP) ra[i] = ld4(Act + row * HP + k0 + c4);
        }
    };
    auto sstore = [&]() {
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * NT;
            const int r = idx >> 4, c4 = (idx & 15) * 4;
            st4(&smZ[r][c4], rz[i].x, rz[i].y, rz[i].z, rz[i].w);
            st4(&smA[r][c4], ra[i].x, ra[i].y, ra[i].z, ra[i].w);
        }
    };
    int sl = blockIdx.z;
    if (sl < total) { gload(sl); sstore(); }
    __syncthreads();
    for (; sl < total; sl += rsplit) {
        const int s = sl / slabs_per_stream;
        const bool more = sl + rsplit < total;
        if (more) gload(sl + rsplit);
        INSR_PRAGMA_UNROLL
        for (int rr = 0; rr < WG_R / 4; ++rr) {
            const int r = grp * (WG_R / 4) + rr;
            const float4 z0 = ld4(&smZ[r][4 * jt]), z1 = ld4(&smZ[r][32 + 4 * jt]);     // interleaved 4+4 columns:
            const float4 a0 = ld4(&smA[r][4 * kt]), a1 = ld4(&smA[r][32 + 4 * kt]);     // conflict-free LDS.128
            const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            INSR_PRAGMA_UNROLL
            for (int a = 0; a < 8; ++a)
                INSR_PRAGMA_UNROLL
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(zz[a], aa[b], acc[a][b]);
        }
        if (s == 0 && blockIdx.y == 0 && tid < WG_T) {       // bias gradient: column sums of the value stream
            INSR_PRAGMA_UNROLL
            for (int r = 0; r < WG_R; ++r) bsum += smZ[r][tid];
        }
        __syncthreads();
        if (more) { sstore(); __syncthreads(); }
    }
    // combine the 4 groups through shared memory (reuse smZ/smA as a 64x64 tile), then red.global
    float *tile = smbuf;
    for (int gsel = 0; gsel < 4; ++gsel) {
        if (grp == gsel) {
            INSR_PRAGMA_UNROLL
            for (int a = 0; a < 8; ++a)
                INSR_PRAGMA_UNROLL
                for (int b = 0; b < 8; ++b) {
                    const int jj = (a < 4) ? (4 * jt + a) : (32 + 4 * jt + a - 4);
                    const int kk = (b < 4) ? (4 * kt + b) : (32 + 4 * kt + b - 4);
                    float *dst = tile + jj * WG_T + kk;
                    *dst = (gsel == 0 ? 0.f : *dst) + acc[a][b];
                }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < WG_T * WG_T; idx += NT) {
        const int j = j0 + idx / WG_T, k = k0 + idx % WG_T;
        const float v = tile[idx];
        if (j < H && k < H && v != 0.f) atomicAdd(gW + (size_t)j * H + k, dm.omega * v);
    }
    if (blockIdx.y == 0 && tid < WG_T && j0 + tid < H && bsum != 0.f) atomicAdd(gb + j0 + tid, dm.omega * bsum);
}

// ---------------------------------------------------------------------------------------------
// T7: thin reductions over points (first layer, output layer) and d loss / d x.
//   gW1[j][d] += omega * sum_p ( zb0[0][p][j] x[p][d] + zb0[1+d][p][j] ),  gb1[j] += omega * sum_p zb0[0][p][j]
//   gWo[o][j] += sum_p sum_s G[o][s][p] actL[s][p][j],                      gbo[o] += sum_p G[o][0][p]
// One thread per neuron j (coalesced over j), each CTA reduces a slice of points, then red.global.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int edge_lanes(int H) { const int w = (H + 31) & ~31; return w < 128 ? w : 128; }

template <int D, int O, int ORDER>
__global__ void __launch_bounds__(NT) k_tiled_edge(SirenDims dm, int HP, const float *__restrict__ x, int64_t n0,
                                                   int nv, int64_t NCp, const float *__restrict__ ZB0,
                                                   const float *__restrict__ ActL, const float *__restrict__ G,
                                                   int pslice, float *__restrict__ gtheta) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    // JW lanes over the neurons, NT / JW groups of lanes striding through the CTA's slice of points
    const int JW = edge_lanes(dm.H), nsub = NT / JW;
    const int j = blockIdx.x * JW + (int)threadIdx.x % JW, sub = (int)threadIdx.x / JW;
    const int p_begin = blockIdx.y * pslice;
    const int p_end = (p_begin + pslice < nv) ? (p_begin + pslice) : nv;
    if (j < dm.H && sub < nsub) {
        float gw1[D], gb1 = 0.f, gwo[O];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) gw1[d] = 0.f;
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) gwo[o] = 0.f;
        INSR_PRAGMA_UNROLL_N(unroll 4)
        for (int p = p_begin + sub; p < p_end; p += nsub) {
            const float z0 = ZB0[(int64_t)p * HP + j];
            gb1 += z0;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) {
                float v = z0 * __ldg(x + (n0 + p) * D + d);
                if (C::ND > 0) v += ZB0[((int64_t)(1 + d) * NCp + p) * HP + j];
                gw1[d] += v;
            }
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                const float a = ActL[((int64_t)s * NCp + p) * HP + j];
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o) gwo[o] = fmaf(__ldg(G + ((int64_t)o * S + s) * NCp + p), a, gwo[o]);
            }
        }
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) atomicAdd(gtheta + insr_w_offset(dm, 0) + j * D + d, dm.omega * gw1[d]);
        atomicAdd(gtheta + insr_b_offset(dm, 0) + j, dm.omega * gb1);
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) atomicAdd(gtheta + insr_w_offset(dm, dm.L + 1) + o * dm.H + j, gwo[o]);
    }
    if (blockIdx.x == 0 && threadIdx.x < O) {                 // output bias
        float sgb = 0.f;
        for (int p = p_begin; p < p_end; ++p) sgb += G[((int64_t)threadIdx.x * S) * NCp + p];
        atomicAdd(gtheta + insr_b_offset(dm, dm.L + 1) + threadIdx.x, sgb);
    }
}

// d loss / d x[p][d] = omega * sum_j W1[j][d] zbar_0[0][p][j]   (one warp per point)
template <int D>
__global__ void __launch_bounds__(NT) k_tiled_gx(SirenDims dm, int HP, const float *__restrict__ theta,
                                                 const float *__restrict__ ZB0, int64_t n0, int nv,
                                                 float *__restrict__ gx) {
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= nv) return;
    float acc[D];
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    for (int j = lane; j < dm.H; j += 32) {
        const float z0 = ZB0[p * HP + j];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) acc[d] = fmaf(__ldg(theta + j * D + d), z0, acc[d]);
    }
    INSR_PRAGMA_UNROLL
    for (int d = 0; d < D; ++d) {
        float v = acc[d];
        INSR_PRAGMA_UNROLL
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) gx[(n0 + p) * D + d] = dm.omega * v;
    }
}

// (host-side sizing helpers live in siren_tiled_api.h)

// the forward sweep that leaves the tape (pre-activations and activations of every sine layer) in the backward
// workspace layout: Zpre[l], Act[l] = [stream][NCp][HP] for l = 0..L
template <int D, int ORDER>
void taped_forward(const SirenDims &dm, const Geo &g, const float *theta, const float *x, int64_t n0, int nv, int64_t NCp,
                   int64_t rows, size_t buf, float *Zpre, float *Act, void *stream, int64_t *launches, bool tensor,
                   float *y = nullptr, float *jac = nullptr, float *h2 = nullptr, bool *fused_out = nullptr) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    (void)S;
#ifndef INSR_CPU_EMU
    // widths of the elasticity scripts with <= 3 streams: the whole network in ONE kernel that also leaves the tape
    if (tensor && insr_mid_supported(dm, ORDER)) {     // (the caller has laid Zpre / Act out for the fused kernels' tape)
        const int rc = insr_mid_forward(dm, ORDER, theta, x + n0 * D, nv, y ? y + n0 * dm.O : nullptr,
                                        jac ? jac + n0 * dm.O * D : nullptr,
                                        h2 ? h2 + n0 * dm.O * (ORDER == 3 ? D * D : 1) : nullptr, Zpre, Act, buf, stream, launches);
        if (rc == 0) {
            if (fused_out) *fused_out = true;
            return;
        }
    }
#endif
    const int64_t items = rows * (g.HP / 4);
    const unsigned eg = (unsigned)((items + NT - 1) / NT);
    const size_t smem = gemm_smem_floats(g) * sizeof(float);
    auto kfwd = k_tiled_fwd<D, ORDER>;
    auto k0 = k_tiled_layer0<D, ORDER>;
    INSR_LAUNCH(k0, dim3(eg), dim3(NT), 0, stream, dm, g.HP, theta, x, n0, nv, NCp, Zpre, Act);
    ++*launches;
    for (int l = 1; l <= dm.L; ++l) {
#ifndef INSR_CPU_EMU
        if constexpr (S <= 4) {
            if (tensor) {
                if (l == dm.L && y && insr_wide::wide_out_ok(dm.H)) {      // output layer rides on the last hidden layer
                    insr_wide::WideOut wo{theta + insr_w_offset(dm, dm.L + 1), theta + insr_b_offset(dm, dm.L + 1), dm.O, nv, n0, y, jac, h2};
                    insr_wide::launch_wide<D, ORDER, 0, true>(dm, theta + insr_w_offset(dm, l), theta + insr_b_offset(dm, l),
                                                              Act + (size_t)(l - 1) * buf, NCp, rows, nullptr, Zpre + (size_t)l * buf,
                                                              Act + (size_t)l * buf, stream, launches, wo);
                    if (fused_out) *fused_out = true;
                    continue;
                }
                insr_wide::launch_wide<D, ORDER, 0>(dm, theta + insr_w_offset(dm, l), theta + insr_b_offset(dm, l),
                                                    Act + (size_t)(l - 1) * buf, NCp, rows, nullptr, Zpre + (size_t)l * buf,
                                                    Act + (size_t)l * buf, stream, launches);
                continue;
            }
        }
#endif
        INSR_LAUNCH(kfwd, dim3((unsigned)(rows / g.BP), g.panels), dim3(NT), smem, stream, dm, g,
                    theta + insr_w_offset(dm, l), theta + insr_b_offset(dm, l), Act + (size_t)(l - 1) * buf, NCp,
                    Zpre + (size_t)l * buf, Act + (size_t)l * buf);
        ++*launches;
    }
}

// forward that keeps its tape for a following run_backward(..., have_tape = true) on the SAME workspace (the backward
// layout and size); only for batches of one workspace chunk (tape_fits)
template <int D, int O, int ORDER>
int run_forward_tape(const SirenDims &dm, const float *theta, const float *x, int64_t N, float *y, float *jac,
                     float *h2, float *ws, void *stream, int64_t *launches, bool tensor) {
    constexpr int S = StreamCfg<D, ORDER>::S;
#ifdef INSR_CPU_EMU
    tensor = false;
#else
    tensor = tensor && S <= 4;
#endif
    constexpr int TP = Tile<S>::TP;
    const Geo g = make_geo(dm.H, S, TP);
    const int64_t chunk = chunk_points(dm, S, true, N);
    if (N > chunk) return -6;
    const int64_t NCp = capacity(chunk);
    size_t buf = (size_t)S * NCp * g.HP;
#ifndef INSR_CPU_EMU
    if (tensor && insr_mid_supported(dm, ORDER)) buf = (size_t)S * NCp * insr_mid_width(dm);      // the fused kernels' tape
#endif
    float *Zpre = ws, *Act = ws + (size_t)(dm.L + 1) * buf;
    const size_t smem = gemm_smem_floats(g) * sizeof(float);
    auto kfwd = k_tiled_fwd<D, ORDER>;
    cudaFuncSetAttribute(kfwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int nv = (int)N;
    const int64_t rows = round_up(nv, tensor ? 128 : g.BP);
    bool fused_out = false;
    taped_forward<D, ORDER>(dm, g, theta, x, 0, nv, NCp, rows, buf, Zpre, Act, stream, launches, tensor, y, jac, h2, &fused_out);
    if (!fused_out) {
        auto ko = k_tiled_out_fwd<D, O, ORDER>;
        INSR_LAUNCH(ko, dim3((unsigned)(((int64_t)nv * 32 + NT - 1) / NT)), dim3(NT), 0, stream, dm, g.HP, theta,
                    Act + (size_t)dm.L * buf, NCp, (int64_t)0, nv, y, jac, h2);
        ++*launches;
    }
    return 0;
}

template <int D, int O, int ORDER>
int run_forward(const SirenDims &dm, const float *theta, const float *x, int64_t N, float *y, float *jac,
                float *h2, float *ws, void *stream, int64_t *launches, bool tensor, bool tape) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    if (tape) return run_forward_tape<D, O, ORDER>(dm, theta, x, N, y, jac, h2, ws, stream, launches, tensor);
#ifdef INSR_CPU_EMU
    tensor = false;
#else
    tensor = tensor && S <= 4;                              // S * NCOL TMEM columns
    if (tensor && insr_mid_supported(dm, ORDER) &&         // one kernel, no workspace, any batch
        insr_mid_forward(dm, ORDER, theta, x, N, y, jac, h2, nullptr, nullptr, 0, stream, launches) == 0)
        return 0;
#endif
    constexpr int TP = Tile<S>::TP;
    const Geo g = make_geo(dm.H, S, TP);
    const int64_t chunk = chunk_points(dm, S, false, N);
    const int64_t NCp = capacity(chunk);
    const size_t buf = (size_t)S * NCp * g.HP;
    float *A0 = ws, *A1 = ws + buf;
    const size_t smem = gemm_smem_floats(g) * sizeof(float);
    auto kfwd = k_tiled_fwd<D, ORDER>;
    cudaFuncSetAttribute(kfwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
        const int nv = (int)((N - n0 < chunk) ? (N - n0) : chunk);
        const int64_t rows = round_up(nv, tensor ? 128 : g.BP);  // tiles actually computed (<= NCp)
        auto k0 = k_tiled_layer0<D, ORDER>;
        const int64_t items = rows * (g.HP / 4);
        INSR_LAUNCH(k0, dim3((unsigned)((items + NT - 1) / NT)), dim3(NT), 0, stream, dm, g.HP, theta, x, n0, nv,
                    NCp, (float *)nullptr, A0);
        ++*launches;
        float *in = A0, *out = A1;
        bool fused_out = false;
        for (int l = 1; l <= dm.L; ++l) {
#ifndef INSR_CPU_EMU
            if constexpr (S <= 4) {
                if (tensor) {
                    if (l == dm.L && insr_wide::wide_out_ok(dm.H)) {   // output layer in the epilogue; the last activations never reach HBM
                        insr_wide::WideOut wo{theta + insr_w_offset(dm, dm.L + 1), theta + insr_b_offset(dm, dm.L + 1), dm.O, nv, n0, y, jac, h2};
                        insr_wide::launch_wide<D, ORDER, 0, true>(dm, theta + insr_w_offset(dm, l), theta + insr_b_offset(dm, l), in, NCp, rows,
                                                                  nullptr, nullptr, nullptr, stream, launches, wo);
                        fused_out = true;
                        continue;
                    }
                    insr_wide::launch_wide<D, ORDER, 0>(dm, theta + insr_w_offset(dm, l), theta + insr_b_offset(dm, l), in, NCp, rows,
                                                        nullptr, nullptr, out, stream, launches);
                    float *t = in; in = out; out = t;
                    continue;
                }
            }
#endif
            INSR_LAUNCH(kfwd, dim3((unsigned)(rows / g.BP), g.panels), dim3(NT), smem, stream, dm, g,
                        theta + insr_w_offset(dm, l), theta + insr_b_offset(dm, l), in, NCp, (float *)nullptr, out);
            ++*launches;
            float *t = in; in = out; out = t;
        }
        if (!fused_out) {
            auto ko = k_tiled_out_fwd<D, O, ORDER>;
            INSR_LAUNCH(ko, dim3((unsigned)(((int64_t)nv * 32 + NT - 1) / NT)), dim3(NT), 0, stream, dm, g.HP, theta, in,
                        NCp, n0, nv, y, jac, h2);
            ++*launches;
        }
    }
    return 0;
}

template <int D, int O, int ORDER>
int run_backward(const SirenDims &dm, const float *theta, const float *x, int64_t N, const float *gy,
                 const float *gjac, const float *gh2, float *gtheta, float *gx, float *ws, void *stream,
                 int64_t *launches, bool tensor, bool have_tape) {
    constexpr int S = StreamCfg<D, ORDER>::S;
#ifdef INSR_CPU_EMU
    tensor = false;
#else
    tensor = tensor && S <= 4;
#endif
    constexpr int TP = Tile<S>::TP;
    const Geo g = make_geo(dm.H, S, TP);
    const int L = dm.L;
    const int64_t chunk = chunk_points(dm, S, true, N);
    const int64_t NCp = capacity(chunk);
#ifndef INSR_CPU_EMU
    if (tensor && insr_mid_supported(dm, ORDER)) {
        // widths of the elasticity scripts, <= 3 streams: (taped forward +) data-gradient chain + weight gradients of all
        // layers + thin layers = 3-4 kernels per chunk on the fused kernels' private tape layout
        if (have_tape && N > chunk) return -6;
        const size_t mbuf = (size_t)S * NCp * insr_mid_width(dm);
        float *Zp = ws, *Ac = ws + (size_t)(L + 1) * mbuf;
        for (int64_t n0 = 0; n0 < N; n0 += chunk) {
            const int64_t nv = (N - n0 < chunk) ? (N - n0) : chunk;
            int rc = 0;
            if (!have_tape)
                rc = insr_mid_forward(dm, ORDER, theta, x + n0 * D, nv, nullptr, nullptr, nullptr, Zp, Ac, mbuf, stream, launches);
            if (rc == 0)
                rc = insr_mid_backward(dm, ORDER, theta, x + n0 * D, nv, gy ? gy + n0 * dm.O : nullptr,
                                       gjac ? gjac + n0 * dm.O * D : nullptr,
                                       gh2 ? gh2 + n0 * dm.O * (ORDER == 3 ? D * D : 1) : nullptr, gtheta,
                                       gx ? gx + n0 * D : nullptr, Zp, Ac, mbuf, stream, launches);
            if (rc) return rc;
        }
        return 0;
    }
#endif
    const size_t buf = (size_t)S * NCp * g.HP;
    float *Zpre = ws;                               // [L+1] buffers
    float *Act = ws + (size_t)(L + 1) * buf;        // [L+1] buffers
    float *ZB0 = ws + (size_t)2 * (L + 1) * buf, *ZB1 = ZB0 + buf;
    float *G = ZB1 + buf;
    const size_t smem = gemm_smem_floats(g) * sizeof(float);
    auto kfwd = k_tiled_fwd<D, ORDER>;
    auto kdg = k_tiled_dgrad<D, ORDER>;
    cudaFuncSetAttribute(kfwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kdg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (have_tape && N > chunk) return -6;
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
        const int nv = (int)((N - n0 < chunk) ? (N - n0) : chunk);
        const int64_t rows = round_up(nv, tensor ? 128 : g.BP);
        const int64_t items = rows * (g.HP / 4);
        const unsigned eg = (unsigned)((items + NT - 1) / NT);
        // ---- forward with tape (skipped when the caller kept the tape of its forward call in this workspace)
        if (!have_tape) taped_forward<D, ORDER>(dm, g, theta, x, n0, nv, NCp, rows, buf, Zpre, Act, stream, launches, tensor);
        // ---- output layer backward -> zbar_L
        auto kob = k_tiled_out_bwd<D, O, ORDER>;
        INSR_LAUNCH(kob, dim3(eg), dim3(NT), 0, stream, dm, g.HP, theta, gy, gjac, gh2, n0, nv, NCp,
                    Zpre + (size_t)L * buf, ZB0, G);
        ++*launches;
        float *zin = ZB0, *zout = ZB1;
        const int wt = (g.HP + WG_T - 1) / WG_T;
        for (int l = L; l >= 1; --l) {
            // weight gradient of layer l needs zbar_l (zin) and act_{l-1}
            const int slabs = S * ((nv + WG_R - 1) / WG_R);
            int rsplit = (2 * sms + wt * wt - 1) / (wt * wt);
            if (rsplit > slabs) rsplit = slabs;
            if (rsplit < 1) rsplit = 1;
#ifndef INSR_CPU_EMU
            if (tensor && insr_wide::wide_wgrad_ok(dm.H)) {
                insr_wide::launch_wide_wgrad(dm, S, zin, Act + (size_t)(l - 1) * buf, NCp, nv, gtheta + insr_w_offset(dm, l),
                                             gtheta + insr_b_offset(dm, l), stream, launches);
            } else
#endif
            {
                auto kwg = k_tiled_wgrad;
                INSR_LAUNCH(kwg, dim3(wt, wt, rsplit), dim3(NT), 0, stream, dm, g.HP, S, zin,
                            Act + (size_t)(l - 1) * buf, NCp, nv, rsplit, gtheta + insr_w_offset(dm, l),
                            gtheta + insr_b_offset(dm, l));
                ++*launches;
            }
#ifndef INSR_CPU_EMU
            bool done = false;
            if constexpr (S <= 4) {
                if (tensor) {
                    insr_wide::launch_wide<D, ORDER, 1>(dm, theta + insr_w_offset(dm, l), nullptr, zin, NCp, rows,
                                                        Zpre + (size_t)(l - 1) * buf, nullptr, zout, stream, launches);
                    done = true;
                }
            }
            if (!done)
#endif
            {
                INSR_LAUNCH(kdg, dim3((unsigned)(rows / g.BP), g.panels), dim3(NT), smem, stream, dm, g,
                            theta + insr_w_offset(dm, l), zin, NCp, Zpre + (size_t)(l - 1) * buf, zout);
                ++*launches;
            }
            float *t = zin; zin = zout; zout = t;
        }
        // ---- thin layers + gx  (zin now holds zbar_0)
        const int jw = edge_lanes(dm.H);
        int pslice = (nv + 4 * sms - 1) / (4 * sms);
        if (pslice < 8 * (NT / jw)) pslice = 8 * (NT / jw);
        auto ke = k_tiled_edge<D, O, ORDER>;
        INSR_LAUNCH(ke, dim3((dm.H + jw - 1) / jw, (nv + pslice - 1) / pslice), dim3(NT), 0, stream, dm, g.HP, x, n0,
                    nv, NCp, zin, Act + (size_t)L * buf, G, pslice, gtheta);
        ++*launches;
        if (gx) {
            auto kgx = k_tiled_gx<D>;
            INSR_LAUNCH(kgx, dim3((unsigned)(((int64_t)nv * 32 + NT - 1) / NT)), dim3(NT), 0, stream, dm, g.HP, theta,
                        zin, n0, nv, gx);
            ++*launches;
        }
    }
    return 0;
}

}  // namespace insr_tiled

