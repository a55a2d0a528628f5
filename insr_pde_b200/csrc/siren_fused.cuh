// siren_fused.cuh -- fused, resident-weights SIREN kernels for hidden width H <= 32
// (advect1D H=20, fluid2Dtlgn H=32: the headline shapes).
//
// Design (B200, FP32 FFMA pipe):
//   * persistent CTAs, one per SM; all layer weights (omega folded in, zero padded to 32x32,
//     XOR-swizzled 16-byte chunks) stay resident in shared memory for the whole launch;
//   * each WARP autonomously pushes tiles of 8 collocation points through the whole network:
//     lane (pg, jg) = (lane>>2, lane&3) owns point pg and the 8 contiguous neurons 8*jg..8*jg+7,
//     for ALL forward-mode streams (value, D tangents, Laplacian trace), so the sine-layer
//     stream algebra and its adjoint run in registers with sin/cos evaluated once per
//     activation;
//   * the three contractions of a hidden layer -- forward  Z = A W^T, data gradient
//     Abar = Zbar W, weight gradient gW += Zbar^T A -- are register-tiled FFMA GEMMs whose
//     operands are read from shared memory with 128-bit loads (10.7 FFMA per LDS.128, every
//     warp-wide load touches <= 128 distinct bytes -> one wavefront);
//   * the backward kernel recomputes the forward pass into a per-lane tape (sin, cos and the
//     pre-activation derivative streams), keeps the weight-gradient partial sums of all layers
//     in REGISTERS across the whole persistent loop, and reduces them once per CTA at the end
//     (shared-memory atomics, then one red.global per parameter per CTA);
//   * warps only ever __syncwarp(); there is no __syncthreads() in the tile loop.
//
// Reference semantics: base/networks.py:21-71 (MLP/Sine), base/diff_ops.py:33-82 (derivative
// streams), base/baseModel.py:77 (backward); lsq mode additionally folds the mean-square
// residual losses of advection/model.py:43-52,68-91 and fluid/model.py:43-52,72-151.
#pragma once
#include "siren_common.cuh"

namespace insr_fused {

constexpr int HP = 32;          // padded hidden width
constexpr int PW = 8;           // points per warp tile
constexpr int LMAX_BWD = 3;     // hidden layers whose gW partials live in registers
constexpr int LMAX_FWD = 8;
constexpr int MAX_WARPS = 8;
constexpr int MAX_COEF = 4 * 3 * 5;
constexpr size_t SMEM_LIMIT = 227 * 1024;

struct Params {
    SirenDims dm;
    const float *theta;
    const float *x;
    int64_t N;
    float *y, *jac, *h2;                 // forward outputs
    const float *gy, *gjac, *gh2;        // backward cotangents (nullable)
    float *gtheta, *gx;                  // backward outputs (gx nullable)
    const float *target;                 // lsq: (N, n_res) or NULL
    float *loss_out;                     // lsq: device scalar, accumulated
    float scale;                         // lsq: loss = scale * sum r^2
    int n_res;
    float coef[MAX_COEF];                // lsq: coef[c][o][s]
    int nwarps;
    void *ws;                            // tensor-core backward: tape scratch
};

// ---- shared-memory map (float offsets) ------------------------------------------------
__host__ __device__ inline int w_floats(int L) { return L * HP * HP + L * HP + HP * 4 + 3 * HP + 4; }
__host__ __device__ inline int off_W(int l /*0-based hidden*/) { return l * HP * HP; }
__host__ __device__ inline int off_B(int L, int l) { return L * HP * HP + l * HP; }
__host__ __device__ inline int off_W1(int L) { return L * HP * HP + L * HP; }
__host__ __device__ inline int off_WO(int L) { return off_W1(L) + HP * 4; }
__host__ __device__ inline int off_BO(int L) { return off_WO(L) + 3 * HP; }
// per-warp region: A[S*TP*256] | T[(L+1)*(S+1)*256] (bwd only, TP == 1) | XS[32] | G[128]
// TP = points per lane (a warp tile covers 8*TP points)
__host__ __device__ inline int warp_floats(int S, int TP, int L, bool bwd) {
    return S * TP * 256 + (bwd ? (L + 1) * (S + 1) * 256 : 0) + 32 + 128;
}

// swizzled float offset of element k of a 32-float row whose swizzle key is `key`
__device__ __forceinline__ int swz(int key, int k) { return ((((k >> 2) ^ key) & 7) << 2) | (k & 3); }

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void sts4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ float f4get(const float4 &v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

// explicit shared-memory byte addresses: lets the swizzle be ONE xor on the address and the row
// offset an immediate of the load (operand buffers are 128-byte aligned, rows are 128 bytes)
#ifdef INSR_CPU_EMU
typedef uintptr_t saddr_t;
__device__ inline saddr_t saddr(const void *p) { return reinterpret_cast<uintptr_t>(p); }
template <int OFF>
__device__ inline float4 lds4a(saddr_t a) { return *reinterpret_cast<const float4 *>(a + OFF); }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t saddr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int OFF>
__device__ __forceinline__ float4 lds4a(saddr_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
    return v;
}
#endif

// ---- stage all weights into shared memory (omega folded into the sine layers) -----------
__device__ inline void stage_weights(const Params &p, float *sm) {
    const SirenDims &dm = p.dm;
    const int H = dm.H, L = dm.L, D = dm.D, O = dm.O;
    const float w = dm.omega;
    const int nt = blockDim.x, tid = threadIdx.x;
    for (int idx = tid; idx < L * HP * HP; idx += nt) {
        const int l = idx / (HP * HP), j = (idx / HP) % HP, k = idx % HP;
        float v = 0.f;
        if (j < H && k < H) v = w * p.theta[insr_w_offset(dm, l + 1) + (int64_t)j * H + k];
        sm[off_W(l) + j * HP + swz(j >> 3, k)] = v;
    }
    for (int idx = tid; idx < L * HP; idx += nt) {
        const int l = idx / HP, j = idx % HP;
        sm[off_B(L, l) + j] = (j < H) ? w * p.theta[insr_b_offset(dm, l + 1) + j] : 0.f;
    }
    for (int idx = tid; idx < HP * 4; idx += nt) {
        const int j = idx >> 2, d = idx & 3;
        float v = 0.f;
        if (j < H) {
            if (d < D) v = w * p.theta[insr_w_offset(dm, 0) + (int64_t)j * D + d];
            else if (d == 3) v = w * p.theta[insr_b_offset(dm, 0) + j];
        }
        sm[off_W1(L) + idx] = v;
    }
    for (int idx = tid; idx < 3 * HP; idx += nt) {
        const int o = idx / HP, j = idx % HP;
        sm[off_WO(L) + idx] = (o < O && j < H) ? p.theta[insr_w_offset(dm, L + 1) + (int64_t)o * H + j] : 0.f;
    }
    if (tid < 4) sm[off_BO(L) + tid] = (tid < O) ? p.theta[insr_b_offset(dm, L + 1) + tid] : 0.f;
}

// per-warp view of shared memory
struct WarpSm {
    const float *W, *Bv, *W1, *WO, *BO;
    float *A, *T, *XS, *G;
};

template <int S, int TP>
__device__ inline WarpSm warp_view(float *sm, int L, int warp, bool bwd) {
    WarpSm v;
    v.W = sm; v.Bv = sm + off_B(L, 0); v.W1 = sm + off_W1(L); v.WO = sm + off_WO(L); v.BO = sm + off_BO(L);
    float *base = sm + ((w_floats(L) + 31) & ~31) + (size_t)warp * warp_floats(S, TP, L, bwd);
    v.A = base;
    v.T = base + S * TP * 256;
    v.XS = v.T + (bwd ? (L + 1) * (S + 1) * 256 : 0);
    v.G = v.XS + 32;
    return v;
}

// ---- sine layer on 4 neurons at once ----------------------------------------------------
// z[s][c]: pre-activations (omega folded) of neurons c=0..3 -> a[s][c]; tape tv[t][c] = (sin, cos,
// zdot_1..D, zddot)
template <int D, int ORDER>
__device__ __forceinline__ void act4(const float (&z)[StreamCfg<D, ORDER>::S][4],
                                     float (&a)[StreamCfg<D, ORDER>::S][4],
                                     float (&tv)[StreamCfg<D, ORDER>::S + 1][4]) {
    typedef StreamCfg<D, ORDER> C;
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) {
        float s, co;
        insr_sincos(z[0][c], s, co);
        tv[0][c] = s; tv[1][c] = co;
        a[0][c] = s;
        float quad = 0.f;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < C::ND; ++d) {
            tv[2 + d][c] = z[1 + d][c];
            a[1 + d][c] = co * z[1 + d][c];
            quad = fmaf(z[1 + d][c], z[1 + d][c], quad);
        }
        if constexpr (ORDER == 2) {
            tv[2 + C::ND][c] = z[1 + C::ND][c];
            a[1 + C::ND][c] = co * z[1 + C::ND][c] - s * quad;
        }
    }
}

// post-activations from the tape
template <int D, int ORDER>
__device__ __forceinline__ void a_from_tape4(const float (&tv)[StreamCfg<D, ORDER>::S + 1][4],
                                             float (&a)[StreamCfg<D, ORDER>::S][4]) {
    typedef StreamCfg<D, ORDER> C;
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) {
        const float s = tv[0][c], co = tv[1][c];
        a[0][c] = s;
        float quad = 0.f;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < C::ND; ++d) {
            a[1 + d][c] = co * tv[2 + d][c];
            quad = fmaf(tv[2 + d][c], tv[2 + d][c], quad);
        }
        if constexpr (ORDER == 2) a[1 + C::ND][c] = co * tv[2 + C::ND][c] - s * quad;
    }
}

// adjoint of the sine layer for 4 neurons: ab[s][c] (cotangent of a) -> zb[s][c] (in place)
template <int D, int ORDER>
__device__ __forceinline__ void adj4(const float (&tv)[StreamCfg<D, ORDER>::S + 1][4],
                                     float (&ab)[StreamCfg<D, ORDER>::S][4]) {
    typedef StreamCfg<D, ORDER> C;
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) {
        const float s = tv[0][c], co = tv[1][c];
        float zb0 = co * ab[0][c];
        float quad = 0.f;
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < C::ND; ++d) {
            const float zd = tv[2 + d][c];
            zb0 = fmaf(-s * zd, ab[1 + d][c], zb0);
            quad = fmaf(zd, zd, quad);
        }
        if constexpr (ORDER == 2) {
            const float aq = ab[1 + C::ND][c];
            const float zq = tv[2 + C::ND][c];
            zb0 = fmaf(aq, -(s * zq + co * quad), zb0);
            const float m = -2.f * s * aq;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < C::ND; ++d) ab[1 + d][c] = fmaf(m, tv[2 + d][c], co * ab[1 + d][c]);
            ab[1 + C::ND][c] = co * aq;
        } else {
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < C::ND; ++d) ab[1 + d][c] = co * ab[1 + d][c];
        }
        ab[0][c] = zb0;
    }
}

// tape <-> shared memory: lane-private float4 slots, index (h*TV + t)*32 + lane
template <int TV>
__device__ __forceinline__ void tape_store(float *Tl, int lane, int h, const float (&tv)[TV][4]) {
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TV; ++t) sts4(Tl + ((h * TV + t) * 32 + lane) * 4, tv[t][0], tv[t][1], tv[t][2], tv[t][3]);
}
template <int TV>
__device__ __forceinline__ void tape_load(const float *Tl, int lane, int h, float (&tv)[TV][4]) {
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TV; ++t) {
        const float4 v = lds4(Tl + ((h * TV + t) * 32 + lane) * 4);
        tv[t][0] = v.x; tv[t][1] = v.y; tv[t][2] = v.z; tv[t][3] = v.w;
    }
}
// operand rows [(s*TP+t)*8+pg][32] (swizzle key pg): write the lane's neurons 8*jg+4h .. +3 of every
// stream of its point t
template <int S, int TP = 1>
__device__ __forceinline__ void operand_store(float *buf, int pg, int jg, int h, const float (&a)[S][4], int t = 0) {
    INSR_PRAGMA_UNROLL
    for (int s = 0; s < S; ++s)
        sts4(buf + ((s * TP + t) * 8 + pg) * HP + ((((2 * jg + h) ^ pg) & 7) << 2), a[s][0], a[s][1], a[s][2], a[s][3]);
}

// ---- first sine layer (D -> H): tangents are the columns of W1, second order is zero ------
template <int D, int ORDER, bool STASH, int TP = 1>
__device__ __forceinline__ void layer0(const WarpSm &ws, int lane, const float (&xv)[TP][D]) {
    typedef StreamCfg<D, ORDER> C;
    const int pg = lane >> 2, jg = lane & 3;
    INSR_PRAGMA_UNROLL
    for (int h = 0; h < 2; ++h) {
        float4 w[4];
        INSR_PRAGMA_UNROLL
        for (int c = 0; c < 4; ++c) w[c] = lds4(ws.W1 + (8 * jg + 4 * h + c) * 4);
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < TP; ++t) {
            float z[C::S][4], a[C::S][4], tv[C::S + 1][4];
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < 4; ++c) {
                float acc = w[c].w;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) acc = fmaf(f4get(w[c], d), xv[t][d], acc);
                z[0][c] = acc;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < C::ND; ++d) z[1 + d][c] = f4get(w[c], d);
                if constexpr (ORDER == 2) z[1 + C::ND][c] = 0.f;
            }
            act4<D, ORDER>(z, a, tv);
            operand_store<C::S, TP>(ws.A, pg, jg, h, a, t);
            if (STASH) tape_store<C::S + 1>(ws.T, lane, h, tv);
        }
    }
}

// ---- forward contraction: acc[t][i][s] = sum_k W[8jg+i][k] * A[s][t][pg][k] -----------------
// rolled over the 8 chunks of 4 reduction indices: the body (TP*S + 8 LDS.128, 32*TP*S FFMA)
// stays resident in the instruction cache
template <int S, int TP>
__device__ __forceinline__ void gemm_fwd(const float *Wl, const float *A, int pg, int jg, float (&acc)[TP][8][S]) {
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TP; ++t)
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < 8; ++i)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) acc[t][i][s] = 0.f;
    const float *arow = A + pg * HP;
    const float *wrow = Wl + (8 * jg) * HP;
    INSR_PRAGMA_UNROLL_N(unroll 1)
    for (int c = 0; c < 8; ++c) {
        const float *ap = arow + (((c ^ pg) & 7) << 2);
        const float *wp = wrow + (((c ^ jg) & 7) << 2);
        float4 a[TP][S];
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < TP; ++t)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) a[t][s] = lds4(ap + (s * TP + t) * 8 * HP);
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < 8; ++i) {
            const float4 w = lds4(wp + i * HP);
            INSR_PRAGMA_UNROLL
            for (int t = 0; t < TP; ++t)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    float v = acc[t][i][s];
                    v = fmaf(w.x, a[t][s].x, v); v = fmaf(w.y, a[t][s].y, v);
                    v = fmaf(w.z, a[t][s].z, v); v = fmaf(w.w, a[t][s].w, v);
                    acc[t][i][s] = v;
                }
        }
    }
}

// ---- data-gradient contraction: ab[k=8jg+i][s] = sum_j W[j][8jg+i] * Zb[s][pg][j] ---------
template <int S>
__device__ __forceinline__ void gemm_dgrad(const float *Wl, const float *Zb, int pg, int jg, float (&ab)[2][S][4]) {
    INSR_PRAGMA_UNROLL
    for (int h = 0; h < 2; ++h)
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s)
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < 4; ++c) ab[h][s][c] = 0.f;
    const float *zrow = Zb + pg * HP;
    INSR_PRAGMA_UNROLL_N(unroll 1)
    for (int c = 0; c < 8; ++c) {            // chunk of 4 reduction indices j = 4c .. 4c+3 (row key c>>1)
        const float *zp = zrow + (((c ^ pg) & 7) << 2);
        float4 z[S];
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) z[s] = lds4(zp + s * 8 * HP);
        const float *w0p = Wl + (4 * c) * HP + ((((2 * jg) ^ (c >> 1)) & 7) << 2);
        const float *w1p = Wl + (4 * c) * HP + ((((2 * jg + 1) ^ (c >> 1)) & 7) << 2);
        INSR_PRAGMA_UNROLL
        for (int jj = 0; jj < 4; ++jj) {
            const float4 w0 = lds4(w0p + jj * HP);
            const float4 w1 = lds4(w1p + jj * HP);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                const float zj = f4get(z[s], jj);
                ab[0][s][0] = fmaf(w0.x, zj, ab[0][s][0]); ab[0][s][1] = fmaf(w0.y, zj, ab[0][s][1]);
                ab[0][s][2] = fmaf(w0.z, zj, ab[0][s][2]); ab[0][s][3] = fmaf(w0.w, zj, ab[0][s][3]);
                ab[1][s][0] = fmaf(w1.x, zj, ab[1][s][0]); ab[1][s][1] = fmaf(w1.y, zj, ab[1][s][1]);
                ab[1][s][2] = fmaf(w1.z, zj, ab[1][s][2]); ab[1][s][3] = fmaf(w1.w, zj, ab[1][s][3]);
            }
        }
    }
}

// ---- weight-gradient contraction over the tile's S*8 rows.  lane (jt, kt) = (lane>>2, lane&3)
// accumulates gw[a][b] += sum_rows Zb[row][4jt+a] * A[row][8kt+b]  and (kt==0) gb[a] += value-stream rows.
// Row r has swizzle key r&7, so with the row index static inside a group of 8 the swizzle is one xor
// of an immediate on a lane-constant address and the row offset is an immediate of the load.
template <int RR, bool BIAS>
__device__ __forceinline__ void wgrad_row(saddr_t zb, saddr_t ab, bool bias_lane, float (&gw)[4][8], float (&gb)[4]) {
    const saddr_t za = zb ^ (saddr_t)(RR << 4);
    const saddr_t aa = ab ^ (saddr_t)(RR << 4);
    const float4 z = lds4a<RR * 128>(za);
    const float4 a0 = lds4a<RR * 128>(aa);
    const float4 a1 = lds4a<RR * 128 + ((RR & 1) ? -16 : 16)>(aa);     // chunk 2kt+1 == (2kt) ^ 1
    const float zz[4] = {z.x, z.y, z.z, z.w};
    INSR_PRAGMA_UNROLL
    for (int a = 0; a < 4; ++a) {
        gw[a][0] = fmaf(zz[a], a0.x, gw[a][0]); gw[a][1] = fmaf(zz[a], a0.y, gw[a][1]);
        gw[a][2] = fmaf(zz[a], a0.z, gw[a][2]); gw[a][3] = fmaf(zz[a], a0.w, gw[a][3]);
        gw[a][4] = fmaf(zz[a], a1.x, gw[a][4]); gw[a][5] = fmaf(zz[a], a1.y, gw[a][5]);
        gw[a][6] = fmaf(zz[a], a1.z, gw[a][6]); gw[a][7] = fmaf(zz[a], a1.w, gw[a][7]);
    }
    if (BIAS) {
        if (bias_lane) {
            INSR_PRAGMA_UNROLL
            for (int a = 0; a < 4; ++a) gb[a] += zz[a];
        }
    }
}
template <bool BIAS>
__device__ __forceinline__ void wgrad_rows8(saddr_t zb, saddr_t ab, bool bias_lane, float (&gw)[4][8], float (&gb)[4]) {
    wgrad_row<0, BIAS>(zb, ab, bias_lane, gw, gb); wgrad_row<1, BIAS>(zb, ab, bias_lane, gw, gb);
    wgrad_row<2, BIAS>(zb, ab, bias_lane, gw, gb); wgrad_row<3, BIAS>(zb, ab, bias_lane, gw, gb);
    wgrad_row<4, BIAS>(zb, ab, bias_lane, gw, gb); wgrad_row<5, BIAS>(zb, ab, bias_lane, gw, gb);
    wgrad_row<6, BIAS>(zb, ab, bias_lane, gw, gb); wgrad_row<7, BIAS>(zb, ab, bias_lane, gw, gb);
}
template <int S>
__device__ __forceinline__ void gemm_wgrad(const float *Zb, const float *A, int lane, float (&gw)[4][8], float (&gb)[4]) {
    const int jt = lane >> 2, kt = lane & 3;
    saddr_t zb = saddr(Zb) + (saddr_t)(jt << 4);            // chunk jt of row 0 (key 0)
    saddr_t ab = saddr(A) + (saddr_t)(kt << 5);             // chunk 2kt of row 0
    wgrad_rows8<true>(zb, ab, kt == 0, gw, gb);             // rows 0..7: the value stream (bias gradient)
    INSR_PRAGMA_UNROLL_N(unroll 1)
    for (int g = 1; g < S; ++g) {
        zb += 8 * 128; ab += 8 * 128;
        wgrad_rows8<false>(zb, ab, false, gw, gb);
    }
}

// ---- hidden sine layers, forward (with optional tape) ---------------------------------------
template <int D, int ORDER, bool STASH, int TP = 1>
__device__ __forceinline__ void hidden_forward(const WarpSm &ws, int L, int lane) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    const int pg = lane >> 2, jg = lane & 3;
    INSR_PRAGMA_UNROLL_N(unroll 1)
    for (int l = 0; l < L; ++l) {
        float acc[TP][8][S];
        gemm_fwd<S, TP>(ws.W + off_W(l), ws.A, pg, jg, acc);
        __syncwarp();                                   // every lane has finished reading A
        INSR_PRAGMA_UNROLL
        for (int h = 0; h < 2; ++h) {
            const float4 b = lds4(ws.Bv + l * HP + 8 * jg + 4 * h);
            INSR_PRAGMA_UNROLL
            for (int t = 0; t < TP; ++t) {
                float z[S][4], a[S][4], tv[S + 1][4];
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    z[0][c] = acc[t][4 * h + c][0] + f4get(b, c);
                    INSR_PRAGMA_UNROLL
                    for (int s = 1; s < S; ++s) z[s][c] = acc[t][4 * h + c][s];
                }
                act4<D, ORDER>(z, a, tv);
                operand_store<S, TP>(ws.A, pg, jg, h, a, t);
                if (STASH) tape_store<S + 1>(ws.T + (l + 1) * (S + 1) * 256, lane, h, tv);
            }
        }
        __syncwarp();
    }
}

// ---- output layer: out[t][o][s] = sum_j WO[o][j] * a_L[s][t][pg][j] (+ bias on the value stream)
template <int O, int S, int TP = 1>
__device__ __forceinline__ void output_forward(const WarpSm &ws, int lane, float (&out)[TP][O][S]) {
    const int pg = lane >> 2, jg = lane & 3;
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TP; ++t)
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) out[t][o][s] = 0.f;
    INSR_PRAGMA_UNROLL
    for (int h = 0; h < 2; ++h) {
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < TP; ++t) {
            float4 a[S];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                a[s] = lds4(ws.A + ((s * TP + t) * 8 + pg) * HP + ((((2 * jg + h) ^ pg) & 7) << 2));
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                const float4 w = lds4(ws.WO + o * HP + 8 * jg + 4 * h);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    float v = out[t][o][s];
                    v = fmaf(w.x, a[s].x, v); v = fmaf(w.y, a[s].y, v);
                    v = fmaf(w.z, a[s].z, v); v = fmaf(w.w, a[s].w, v);
                    out[t][o][s] = v;
                }
            }
        }
    }
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < TP; ++t)
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) {
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                float v = out[t][o][s];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                out[t][o][s] = v;
            }
            out[t][o][0] += ws.BO[o];
        }
}

// points per lane of the forward kernel: 8 x 8 register tiles (TP*S ~ 8 rows x 8 neurons) give
// >= 4 FFMA per float delivered by shared memory, the balance point of the 128 B/clk LSU pipe
template <int S> struct FwdTile { static constexpr int TP = (S >= 3) ? 2 : (S == 2 ? 4 : 8); };
constexpr int FWD_WARPS = 12;

// =============================================================================================
// forward kernel
// =============================================================================================
template <int D, int O, int ORDER>
__global__ void __launch_bounds__(FWD_WARPS * 32, 1) k_fused_fwd(Params p) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int TP = FwdTile<S>::TP;
    INSR_DYN_SMEM(float, sm);
    stage_weights(p, sm);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pg = lane >> 2, jg = lane & 3;
    const int L = p.dm.L;
    const WarpSm ws = warp_view<S, TP>(sm, L, warp, false);
    const int64_t ntiles = (p.N + PW * TP - 1) / (PW * TP);
    const int64_t stride = (int64_t)gridDim.x * p.nwarps;
    for (int64_t tile = (int64_t)blockIdx.x * p.nwarps + warp; tile < ntiles; tile += stride) {
        float xv[TP][D];
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < TP; ++t) {
            const int64_t n = tile * (PW * TP) + t * 8 + pg;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) xv[t][d] = (n < p.N) ? __ldg(p.x + n * D + d) : 0.f;
        }
        layer0<D, ORDER, false, TP>(ws, lane, xv);
        __syncwarp();
        hidden_forward<D, ORDER, false, TP>(ws, L, lane);
        float out[TP][O][S];
        output_forward<O, S, TP>(ws, lane, out);
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < TP; ++t) {
            const int64_t n = tile * (PW * TP) + t * 8 + pg;
            if (n < p.N) {
                // the 4 lanes of a point share the stores: lane jg writes output o == jg (O <= 3)
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o)
                    if (jg == o) insr_store_outputs<D, O, ORDER>(n, o, out[t][o], p.y, p.jac, p.h2);
            }
        }
        __syncwarp();                                   // A is rewritten by the next tile
    }
}

// =============================================================================================
// backward kernel (LSQ = false: cotangents given; LSQ = true: fused residual/loss epilogue)
// =============================================================================================
template <int D, int O, int ORDER, bool LSQ>
__global__ void __launch_bounds__(MAX_WARPS * 32, 1) k_fused_bwd(Params p) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int TV = S + 1;
    INSR_DYN_SMEM(float, sm);
    stage_weights(p, sm);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pg = lane >> 2, jg = lane & 3;     // forward / dgrad ownership
    const int jt = lane >> 2, kt = lane & 3;     // weight-gradient ownership
    const int L = p.dm.L;
    const WarpSm ws = warp_view<S, 1>(sm, L, warp, true);

    // persistent partial sums (registers, whole launch)
    float gw[LMAX_BWD][4][8], gb[LMAX_BWD][4];
    float g1[4] = {0.f, 0.f, 0.f, 0.f};      // first layer: kt==0 -> bias, kt==1+d -> column d of W1
    float gwo[4] = {0.f, 0.f, 0.f, 0.f};     // output layer row o == kt
    float gbo = 0.f, loss_acc = 0.f;
    INSR_PRAGMA_UNROLL
    for (int l = 0; l < LMAX_BWD; ++l)
        INSR_PRAGMA_UNROLL
        for (int a = 0; a < 4; ++a) {
            gb[l][a] = 0.f;
            INSR_PRAGMA_UNROLL
            for (int b = 0; b < 8; ++b) gw[l][a][b] = 0.f;
        }

    const int64_t ntiles = (p.N + PW - 1) / PW;
    const int64_t stride = (int64_t)gridDim.x * p.nwarps;
    for (int64_t tile = (int64_t)blockIdx.x * p.nwarps + warp; tile < ntiles; tile += stride) {
        const int64_t n = tile * PW + pg;
        const bool valid = n < p.N;
        float xv[1][D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) xv[0][d] = valid ? __ldg(p.x + n * D + d) : 0.f;
        if (jg == 0) {
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) ws.XS[pg * 4 + d] = xv[0][d];
        }
        // ---------------- forward with tape
        layer0<D, ORDER, true, 1>(ws, lane, xv);
        __syncwarp();
        hidden_forward<D, ORDER, true, 1>(ws, L, lane);

        // ---------------- output-layer cotangents g[o][s]
        float g[O][S];
        if constexpr (LSQ) {
            float out1[1][O][S];
            output_forward<O, S, 1>(ws, lane, out1);
            float (&out)[O][S] = out1[0];
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) g[o][s] = 0.f;
            for (int c = 0; c < p.n_res; ++c) {
                float r = (valid && p.target) ? -__ldg(p.target + n * p.n_res + c) : 0.f;
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o)
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) r = fmaf(p.coef[(c * O + o) * S + s], out[o][s], r);
                if (!valid) r = 0.f;
                if (jg == 0) loss_acc = fmaf(r, r, loss_acc);
                const float r2 = 2.f * p.scale * r;
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o)
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) g[o][s] = fmaf(p.coef[(c * O + o) * S + s], r2, g[o][s]);
            }
        } else {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                if (valid) {
                    insr_load_cotangents<D, O, ORDER>(n, o, p.gy, p.gjac, p.gh2, g[o]);
                } else {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) g[o][s] = 0.f;
                }
            }
        }
        if (jg == 0) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) ws.G[pg * 16 + o * S + s] = g[o][s];
        }
        // cotangent of the last sine layer's outputs: ab[h][s][c] for neuron 8jg+4h+c
        float ab[2][S][4];
        INSR_PRAGMA_UNROLL
        for (int h = 0; h < 2; ++h) {
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) ab[h][s][c] = 0.f;
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                const float4 w = lds4(ws.WO + o * HP + 8 * jg + 4 * h);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    ab[h][s][0] = fmaf(w.x, g[o][s], ab[h][s][0]); ab[h][s][1] = fmaf(w.y, g[o][s], ab[h][s][1]);
                    ab[h][s][2] = fmaf(w.z, g[o][s], ab[h][s][2]); ab[h][s][3] = fmaf(w.w, g[o][s], ab[h][s][3]);
                }
            }
        }
        __syncwarp();                                   // G visible (A == a_L already is)
        // ---------------- output-layer gradients: row o == kt, neurons 4jt .. 4jt+3
        if (kt < O) {
            INSR_PRAGMA_UNROLL
            for (int r = 0; r < S * 8; ++r) {
                const int s = r >> 3, q = r & 7;
                const float gv = ws.G[q * 16 + kt * S + s];
                const float4 a = lds4(ws.A + r * HP + (((jt ^ q) & 7) << 2));
                gwo[0] = fmaf(gv, a.x, gwo[0]); gwo[1] = fmaf(gv, a.y, gwo[1]);
                gwo[2] = fmaf(gv, a.z, gwo[2]); gwo[3] = fmaf(gv, a.w, gwo[3]);
                if (s == 0 && jt == 0) gbo += gv;
            }
        }

        // ---------------- reverse sweep through the hidden layers.  Runtime loop (small code, resident in
        // the instruction cache); only the weight-gradient call is switched statically because its
        // accumulators gw[l] are registers
        INSR_PRAGMA_UNROLL_N(unroll 1)
        for (int l = L; l >= 1; --l) {
            float *Tl = ws.T + l * TV * 256;
            INSR_PRAGMA_UNROLL
            for (int h = 0; h < 2; ++h) {
                float tv[TV][4];
                tape_load<TV>(Tl, lane, h, tv);
                adj4<D, ORDER>(tv, ab[h]);
            }
            __syncwarp();                               // tape of layer l consumed; A / previous Zb no longer read
            float *Zb = Tl;                             // zbar operand reuses the tape slot of layer l
            INSR_PRAGMA_UNROLL
            for (int h = 0; h < 2; ++h) {
                operand_store<S>(Zb, pg, jg, h, ab[h]);
                float tv[TV][4], a[S][4];
                tape_load<TV>(Tl - TV * 256, lane, h, tv);
                a_from_tape4<D, ORDER>(tv, a);
                operand_store<S>(ws.A, pg, jg, h, a);   // a_{l-1}: the input of hidden layer l
            }
            __syncwarp();
            gemm_dgrad<S>(ws.W + off_W(l - 1), Zb, pg, jg, ab);
            switch (l) {
                case 1: gemm_wgrad<S>(Zb, ws.A, lane, gw[0], gb[0]); break;
                case 2: gemm_wgrad<S>(Zb, ws.A, lane, gw[1], gb[1]); break;
                default: gemm_wgrad<S>(Zb, ws.A, lane, gw[2], gb[2]); break;
            }
        }
        // ---------------- first sine layer
        INSR_PRAGMA_UNROLL
        for (int h = 0; h < 2; ++h) {
            float tv[TV][4];
            tape_load<TV>(ws.T, lane, h, tv);
            adj4<D, ORDER>(tv, ab[h]);
        }
        __syncwarp();
        INSR_PRAGMA_UNROLL
        for (int h = 0; h < 2; ++h) operand_store<S>(ws.T, pg, jg, h, ab[h]);
        if (p.gx) {                                     // exact d loss / d x (all streams funnel into zbar_0)
            float px[D];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) px[d] = 0.f;
            INSR_PRAGMA_UNROLL
            for (int h = 0; h < 2; ++h)
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const float4 w = lds4(ws.W1 + (8 * jg + 4 * h + c) * 4);
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) px[d] = fmaf(f4get(w, d), ab[h][0][c], px[d]);
                }
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) {
                px[d] += __shfl_xor_sync(0xffffffffu, px[d], 1);
                px[d] += __shfl_xor_sync(0xffffffffu, px[d], 2);
            }
            if (valid && jg == 0) {
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) p.gx[n * D + d] = px[d];
            }
        }
        __syncwarp();
        // first-layer gradients: lane column kt: 0 -> bias, 1+d -> W1[:, d]; neurons 4jt .. 4jt+3
        if (kt <= D) {
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < 8; ++q) {
                const float4 z0 = lds4(ws.T + q * HP + (((jt ^ q) & 7) << 2));
                float m = 1.f;
                float4 zd = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kt > 0) {
                    m = ws.XS[q * 4 + kt - 1];
                    if (C::ND > 0) zd = lds4(ws.T + (kt * 8 + q) * HP + (((jt ^ q) & 7) << 2));
                }
                g1[0] += fmaf(z0.x, m, zd.x); g1[1] += fmaf(z0.y, m, zd.y);
                g1[2] += fmaf(z0.z, m, zd.z); g1[3] += fmaf(z0.w, m, zd.w);
            }
        }
        __syncwarp();                                   // XS / A / tape are rewritten by the next tile
    }

    // ---------------- CTA reduction of the register partials, then one red.global per parameter.
    // Deterministic inside the CTA: warps take turns adding their registers into one shared copy
    // (inside a warp every lane owns distinct parameters, so plain read-modify-write suffices).
    __syncthreads();
    const SirenDims dm = p.dm;
    const int H = dm.H;
    const int P = (int)insr_theta_size(dm);
    float *red = sm + ((w_floats(L) + 31) & ~31);       // reuse the (now idle) per-warp regions
    for (int i = threadIdx.x; i < P + 1; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const float w = dm.omega;
    for (int turn = 0; turn < p.nwarps; ++turn) {
        if (warp == turn) {
            INSR_PRAGMA_UNROLL
            for (int l = 1; l <= LMAX_BWD; ++l) {
                if (l <= L) {
                    const int wo = (int)insr_w_offset(dm, l), bo = (int)insr_b_offset(dm, l);
                    INSR_PRAGMA_UNROLL
                    for (int a = 0; a < 4; ++a) {
                        const int j = 4 * jt + a;
                        if (j < H) {
                            INSR_PRAGMA_UNROLL
                            for (int b = 0; b < 8; ++b) {
                                const int k = 8 * kt + b;
                                if (k < H) red[wo + j * H + k] += w * gw[l - 1][a][b];
                            }
                            if (kt == 0) red[bo + j] += w * gb[l - 1][a];
                        }
                    }
                }
            }
            INSR_PRAGMA_UNROLL
            for (int a = 0; a < 4; ++a) {
                const int j = 4 * jt + a;
                if (j < H) {
                    if (kt == 0) red[(int)insr_b_offset(dm, 0) + j] += w * g1[a];
                    else if (kt <= D) red[(int)insr_w_offset(dm, 0) + j * D + (kt - 1)] += w * g1[a];
                    if (kt < O) red[(int)insr_w_offset(dm, L + 1) + kt * H + j] += gwo[a];
                }
            }
            if (jt == 0 && kt < O) red[(int)insr_b_offset(dm, L + 1) + kt] += gbo;
            if (LSQ) {
                float t = loss_acc;
                INSR_PRAGMA_UNROLL
                for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
                if (lane == 0) red[P] += t;
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const float v = red[i];
        if (v != 0.f) atomicAdd(p.gtheta + i, v);
    }
    if (LSQ && threadIdx.x == 0) atomicAdd(p.loss_out, p.scale * red[P]);
}

// =============================================================================================
// host side
// =============================================================================================
inline bool shape_instantiated(int D, int O) { return (D == 1 && O == 1) || (D == 2 && O == 1) || (D == 2 && O == 2); }

inline size_t smem_bytes(int S, int TP, int L, bool bwd, int nwarps) {
    return ((size_t)((w_floats(L) + 31) & ~31) + (size_t)nwarps * warp_floats(S, TP, L, bwd)) * sizeof(float);
}
inline int pick_warps(int S, int TP, int L, bool bwd, int max_warps) {
    int nw = max_warps;
    while (nw > 1 && smem_bytes(S, TP, L, bwd, nw) > SMEM_LIMIT) --nw;
    return nw;
}

inline int sm_count() {
    static thread_local int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int D, int O, int ORDER>
int launch_fwd(Params &p, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    constexpr int TP = FwdTile<S>::TP;
    p.nwarps = pick_warps(S, TP, p.dm.L, false, FWD_WARPS);
    const size_t smem = smem_bytes(S, TP, p.dm.L, false, p.nwarps);
    auto kfn = k_fused_fwd<D, O, ORDER>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t tiles = (p.N + PW * TP - 1) / (PW * TP);
    int64_t ctas = (tiles + p.nwarps - 1) / p.nwarps;
    const int64_t cap = (int64_t)sm_count();
    if (ctas > cap) ctas = cap;
    INSR_LAUNCH(kfn, dim3((unsigned)ctas), dim3(p.nwarps * 32), smem, stream, p);
    ++*launches;
    return 0;
}

template <int D, int O, int ORDER, bool LSQ>
int launch_bwd(Params &p, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    p.nwarps = pick_warps(S, 1, p.dm.L, true, MAX_WARPS);
    const size_t smem = smem_bytes(S, 1, p.dm.L, true, p.nwarps);
    auto kfn = k_fused_bwd<D, O, ORDER, LSQ>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t tiles = (p.N + PW - 1) / PW;
    int64_t ctas = (tiles + p.nwarps - 1) / p.nwarps;
    if (ctas > sm_count()) ctas = sm_count();
    INSR_LAUNCH(kfn, dim3((unsigned)ctas), dim3(p.nwarps * 32), smem, stream, p);
    ++*launches;
    return 0;
}

}  // namespace insr_fused

inline bool insr_fused_supported(const SirenDims &dm, int order, int backward) {
    using namespace insr_fused;
    if (dm.H > HP || order > 2 || dm.L < 1) return false;
    if (!shape_instantiated(dm.D, dm.O)) return false;
    return dm.L <= (backward ? LMAX_BWD : LMAX_FWD);
}

// the tensor-core backward keeps its tape in a per-CTA global scratch: (L+1) * 4 (S+1) float4 per thread
inline size_t insr_fused_ws_bytes(const SirenDims &dm, int64_t, int order, int backward) {
    if (!backward) return 0;
    const int S = insr_nstreams(dm.D, order);
    return (size_t)insr_fused::sm_count() * (size_t)(dm.L + 1) * 4 * (S + 1) * 256 * 16 + 256;
}

// one translation unit per (D, O) pair instantiates the kernels (siren_fused_inst.cuh) so that
// nvcc can build them in parallel; kind: 0 = forward, 1 = backward, 2 = lsq step
int insr_fused_run_11(int kind, insr_fused::Params &p, int order, void *stream, int64_t *launches);
int insr_fused_run_21(int kind, insr_fused::Params &p, int order, void *stream, int64_t *launches);
int insr_fused_run_22(int kind, insr_fused::Params &p, int order, void *stream, int64_t *launches);

inline int insr_fused_run(int kind, insr_fused::Params &p, int order, void *stream, int64_t *launches) {
    if (p.dm.D == 1 && p.dm.O == 1) return insr_fused_run_11(kind, p, order, stream, launches);
    if (p.dm.D == 2 && p.dm.O == 1) return insr_fused_run_21(kind, p, order, stream, launches);
    if (p.dm.D == 2 && p.dm.O == 2) return insr_fused_run_22(kind, p, order, stream, launches);
    return -6;
}

inline int insr_fused_forward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N,
                              float *y, float *jac, float *h2, float *, void *stream, int64_t *launches,
                              bool tensor = false) {
    insr_fused::Params p{};
    p.dm = dm; p.theta = theta; p.x = x; p.N = N; p.y = y; p.jac = jac; p.h2 = h2;
    return insr_fused_run(tensor ? 3 : 0, p, order, stream, launches);
}

inline int insr_fused_backward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N,
                               const float *gy, const float *gjac, const float *gh2, float *gtheta, float *gx,
                               float *ws, void *stream, int64_t *launches, bool tensor = false) {
    insr_fused::Params p{};
    p.dm = dm; p.theta = theta; p.x = x; p.N = N; p.gy = gy; p.gjac = gjac; p.gh2 = gh2;
    p.gtheta = gtheta; p.gx = gx; p.ws = ws;
    return insr_fused_run(tensor ? 4 : 1, p, order, stream, launches);
}

// coef_host: cy (n_res x O) | cj (n_res x O x D) | cl (n_res x O)  ->  coef[c][o][s]
inline int insr_fused_lsq_step(const SirenDims &dm, int order, int n_res, const float *coef_host,
                               const float *theta, const float *x, int64_t N, const float *target, float scale,
                               float *loss_out, float *gtheta, float *ws, size_t, void *stream, int64_t *launches,
                               bool tensor = false) {
    if (!insr_fused_supported(dm, order, 1)) return -6;
    insr_fused::Params p{};
    p.dm = dm; p.theta = theta; p.x = x; p.N = N; p.gtheta = gtheta; p.gx = nullptr; p.ws = ws;
    p.target = target; p.scale = scale; p.loss_out = loss_out; p.n_res = n_res;
    const int D = dm.D, O = dm.O, S = insr_nstreams(D, order);
    const float *cy = coef_host, *cj = coef_host + n_res * O, *cl = cj + n_res * O * D;
    for (int c = 0; c < n_res; ++c)
        for (int o = 0; o < O; ++o) {
            float *dst = p.coef + (c * O + o) * S;
            dst[0] = cy[c * O + o];
            if (order >= 1)
                for (int d = 0; d < D; ++d) dst[1 + d] = cj[(c * O + o) * D + d];
            if (order == 2) dst[1 + D] = cl[c * O + o];
        }
    return insr_fused_run(tensor ? 5 : 2, p, order, stream, launches);
}
