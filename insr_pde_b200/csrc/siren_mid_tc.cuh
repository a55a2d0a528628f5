// siren_mid_tc.cuh -- fully fused tcgen05 / TMEM kernels for the widths of the elasticity scripts (32 < H <= 80:
// hidden_features = 64 / 66 / 68, scripts/elasticity2Dstretch.sh, elasticity3D*.sh; S <= 3 forward-mode streams).
//
// Why: the layer-by-layer kernels of this family (siren_wide_tc.cuh) move 2.6 KB per point per hidden layer through
// HBM and need L + 2 launches per pass; at these widths a 128-point tile of ALL streams fits on chip, so the whole
// network -- first sine layer, L hidden layers, output layer -- runs in ONE kernel with the activations never leaving
// the SM (k_mid_fwd), and the reverse sweep's data-gradient chain likewise (k_mid_dgrad).
//
// Mapping (the H <= 32 family's, siren_tc.cuh, at width HP16 = H rounded up to 16):
//   * CTA tile = 128 collocation points = M of tcgen05.mma, one TMEM lane per point; 512 threads = 4 per point:
//     thread (row, quarter) owns the neuron columns {16 i + 4 quarter + c : i < NQ, c < 4}, so every 16-column round of
//     the epilogue covers a contiguous 64-byte run of each row of the [stream][point][HP] tape planes.
//   * every stream s is its own MMA chain  D_s[128 x HP16] = A_s[128 x HP16] . B^T  into TMEM columns
//     [s HP16, (s+1) HP16); 3 products per k-step keep FP32-level accuracy:
//         a_hi . w_hi + a_hi . w_lo   kind::tf32, a_hi (13 low mantissa bits cleared) read from TMEM columns
//                                     [S HP16 + s HP16, ..) -- an MMA with A in TMEM runs at the math rate --,
//         a_lo . w_hi                 kind::f16 (bf16 x bf16, FP32 accumulate): a_lo = a - a_hi carries 2^-11 of the
//                                     value, so its bf16 rounding and that of w_hi cost 2^-20 relative -- and the
//                                     operand takes 2 bytes per element of shared memory instead of 4.
//   * weights are NOT resident (3 layers x 64 KB of split operands do not fit beside the activations): one layer at a
//     time is loaded from L2 into registers while the tensor core works on the previous layer, then split (omega folded
//     in) into the single operand buffer  W_hi | W_lo (tf32, K-major canonical, no swizzle) | bf16(W).
//   * shared memory: 64 KB weights + S x 20 KB bf16 lo operands + partials; TMEM: 2 S HP16 <= 480 columns.
//   * tape (pre-activations / activations of every sine layer, kept for the reverse sweep): a layout private to this
//     file -- float4 of 4 neurons at [layer][tile][stream][round][quarter][row] -- so that every tape access of every
//     kernel is a fully coalesced 512-byte warp transaction straight from / to registers (no staging, no barriers).
//
// Only compiled by nvcc (inline PTX); the host-side SIMT emulation keeps using the FFMA kernels.
#pragma once
#include <cstdlib>
#include "siren_tc.cuh"

#ifndef INSR_CPU_EMU
namespace insr_mid {

using insr_tc::s32;
constexpr int MT = 512;                 // threads: 4 per point
constexpr int TILE = 128;               // points per CTA tile
constexpr int MAX_HP16 = 80;
constexpr int MAX_S = 4;

// tape element: the 4 neurons 16 i + 4 q .. + 3 of (tile T, stream s, row r); index in float4 units inside one layer buffer
__host__ __device__ inline size_t tape_f4(int64_t T, int s, int i, int q, int r, int S, int NQ) {
    return ((((size_t)T * S + s) * NQ + i) * 4 + q) * TILE + r;
}

__host__ __device__ inline int hp16_of(int H) { return (H + 15) & ~15; }

// K-major canonical layouts without swizzle: core matrix = 8 rows x 16 bytes; LBO (next core matrix along K) = 128 B,
// SBO (next 8-row group) = 8 rows x row bytes
__device__ __forceinline__ int off32(int row, int k, int sbo) { return (row >> 3) * sbo + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ int off16(int row, int k, int sbo) { return (row >> 3) * sbo + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
// tcgen05.ld is asynchronous: values may only be consumed after tcgen05.wait::ld.  This empty volatile asm, placed after
// the wait, makes every later use of the four registers depend on it (volatile asms keep their order).
__device__ __forceinline__ void tmem_ld_ready4(float (&v)[4]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                    "r"(__float_as_uint(v[3])) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    return w;
}

// streams whose hi operand lives in TMEM (beside the S HP16 accumulator columns); the others keep it in shared memory
// (fp32, K-major canonical) and pay the operand fetch: 32 + N/4 instead of N/2 cycles per MMA
__host__ __device__ constexpr int nht_of(int HP16, int S) { return (512 - S * HP16) / HP16 < S ? (512 - S * HP16) / HP16 : S; }

struct MidSmem {
    int w_hi, w_lo, w_hb, a_lo, a_hi, bias, w1, wo, bo, part, mbar, tmem, total;
};
__host__ __device__ inline MidSmem mid_smem(int HP16, int S, int L, int PS) {
    MidSmem m;
    int o = 0;
    m.w_hi = o; o += HP16 * HP16 * 4;
    m.w_lo = o; o += HP16 * HP16 * 4;
    m.w_hb = o; o += HP16 * HP16 * 2;
    m.a_lo = o; o += S * TILE * HP16 * 2;
    m.a_hi = o; o += (S - nht_of(HP16, S)) * TILE * HP16 * 4;
    m.bias = o; o += L * HP16 * 4;
    m.w1 = o; o += HP16 * 16;
    m.wo = o; o += 3 * HP16 * 4;
    m.bo = o; o += 16;
    m.part = m.w_hi; (void)PS;                 // [4][128][PS] partials alias the weight buffer (idle between the tiles' contractions)
    m.mbar = o; o += 16;
    m.tmem = o; o += 16;
    m.total = o;
    return m;
}

struct MidParams {
    SirenDims dm;
    const float *theta;
    const float *x;            // (N, D), already offset to the chunk's first point for taped runs
    int64_t N;                 // points of this launch
    float *y, *jac, *h2;       // outputs (offset like x)
    // tape (see tape_f4): pre-activations (turned into zbar in place by k_mid_dgrad) and activations of every sine layer
    float *Zpre, *Act;
    int64_t buf;               // floats per layer buffer: S * rows capacity * HP16
    // reverse sweep
    const float *gy, *gjac, *gh2;   // output cotangents, already offset to the chunk's first point (NULL = zero)
    float *gx;                 // (N, D) or NULL
    float *gtheta;
};

// one layer's weights, raw from global memory (L2).  A warp handles one 8 x 8 block (n8, k8) of the operand B per
// iteration: the lane's element pair (n, k), (n, k + 1) lands in the K-major canonical layout such that the 32 lanes
// cover two full 128-byte core-matrix rows (tf32 operands) or one (bf16): conflict-free shared-memory stores, and every
// global load touches 8 sectors of 32 bytes.
template <int HP16, bool TRANSPOSED>
struct WRegs {
    static constexpr int NB8 = HP16 / 8;
    static constexpr int NU = NB8 * NB8;                 // 8 x 8 blocks
    static constexpr int NI = (NU + MT / 32 - 1) / (MT / 32);
    float2 v[NI];
    __device__ __forceinline__ static void coords(int unit, int lane, int &n, int &k) {
        const int n8 = unit / NB8, k8 = unit % NB8;
        const int nl = TRANSPOSED ? (lane & 7) : (lane >> 2), kp = TRANSPOSED ? (lane >> 3) : (lane & 3);
        n = 8 * n8 + nl;
        k = 8 * k8 + 2 * kp;
    }
    // B[n][k] = W[n][k] (forward: n = output neuron) or W[k][n] (data gradient: n = input neuron)
    __device__ __forceinline__ void load(const float *__restrict__ W, int H, int tid) {
        const int warp = tid >> 5, lane = tid & 31;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NI; ++i) {
            const int unit = warp + (MT / 32) * i;
            float a = 0.f, b = 0.f;
            if (NU % (MT / 32) == 0 || unit < NU) {
                int n, k;
                coords(unit, lane, n, k);
                if (n < H) {
                    if (!TRANSPOSED) {
                        if (k < H) a = __ldg(W + (size_t)n * H + k);
                        if (k + 1 < H) b = __ldg(W + (size_t)n * H + k + 1);
                    } else {
                        if (k < H) a = __ldg(W + (size_t)k * H + n);
                        if (k + 1 < H) b = __ldg(W + (size_t)(k + 1) * H + n);
                    }
                }
            }
            v[i] = make_float2(a, b);
        }
    }
    __device__ __forceinline__ void store(unsigned char *whi, unsigned char *wlo, unsigned char *whb, float omega, int tid) const {
        constexpr int SBO32 = 32 * HP16, SBO16 = 16 * HP16;
        const int warp = tid >> 5, lane = tid & 31;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NI; ++i) {
            const int unit = warp + (MT / 32) * i;
            if (NU % (MT / 32) == 0 || unit < NU) {
                int n, k;
                coords(unit, lane, n, k);
                const float a = omega * v[i].x, b = omega * v[i].y;
                const float ah = insr_tc::tf32_hi(a), bh = insr_tc::tf32_hi(b);
                *reinterpret_cast<float2 *>(whi + off32(n, k, SBO32)) = make_float2(ah, bh);
                *reinterpret_cast<float2 *>(wlo + off32(n, k, SBO32)) = make_float2(a - ah, b - bh);
                *reinterpret_cast<uint32_t *>(whb + off16(n, k, SBO16)) = pack_bf16x2(a, b);
            }
        }
    }
};

// the MMAs of one hidden-layer contraction (all streams), issued by one elected lane
template <int HP16, int S>
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, uint32_t whi, uint32_t wlo, uint32_t whb, uint32_t alo, uint32_t ahi_sm, int H) {
    const int nk8 = (H + 7) >> 3, nk16 = (H + 15) >> 4;      // k-steps that hold data (the padding beyond H is zero on both sides)
    constexpr int NHT = nht_of(HP16, S);
    constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HP16 >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HP16 >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t SBO32 = 32 * HP16, SBO16 = 16 * HP16;
    INSR_PRAGMA_UNROLL
    for (int s = 0; s < S; ++s) {
        const uint32_t d = tmem_base + (uint32_t)(s * HP16);
        if (s < NHT) {
            const uint32_t ahi = tmem_base + (uint32_t)((S + s) * HP16);
            INSR_PRAGMA_UNROLL
            for (int ks = 0; ks < HP16 / 8; ++ks) {    // K = 8 per tf32 instruction: 8 TMEM columns / two core matrices
                if (ks < nk8) {
                    mma_tf32_ts(d, ahi + 8 * ks, desc_kmajor(whi + 256 * ks, SBO32), IDESC_TF32, ks > 0 ? 1u : 0u);
                    mma_tf32_ts(d, ahi + 8 * ks, desc_kmajor(wlo + 256 * ks, SBO32), IDESC_TF32, 1u);
                }
            }
        } else {
            const uint32_t ahi = ahi_sm + (uint32_t)((s - NHT) * TILE * HP16 * 4);
            INSR_PRAGMA_UNROLL
            for (int ks = 0; ks < HP16 / 8; ++ks) {
                if (ks < nk8) {
                    mma_tf32_ss(d, desc_kmajor(ahi + 256 * ks, SBO32), desc_kmajor(whi + 256 * ks, SBO32), IDESC_TF32, ks > 0 ? 1u : 0u);
                    mma_tf32_ss(d, desc_kmajor(ahi + 256 * ks, SBO32), desc_kmajor(wlo + 256 * ks, SBO32), IDESC_TF32, 1u);
                }
            }
        }
        INSR_PRAGMA_UNROLL
        for (int kb = 0; kb < HP16 / 16; ++kb)         // K = 16 per bf16 instruction: two core matrices
            if (kb < nk16) mma_bf16_ss(d, desc_kmajor(alo + (uint32_t)(s * TILE * HP16 * 2) + 256 * kb, SBO16), desc_kmajor(whb + 256 * kb, SBO16),
                        IDESC_BF16, 1u);
    }
}

// 4 consecutive neurons j0 .. j0 + 3 of one stream of row `row` become the operand of the next contraction:
// hi part -> TMEM, lo part -> bf16 shared-memory operand
template <int HP16>
__device__ __forceinline__ void store_operand4(uint32_t hi_tmem, unsigned char *ahi_s, unsigned char *alo_s, int row, int j0, const float (&a)[4]) {
    float h[4];
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) h[c] = insr_tc::tf32_hi(a[c]);
    if (ahi_s == nullptr) tmem_st4(hi_tmem + j0, h);
    else *reinterpret_cast<float4 *>(ahi_s + off32(row, j0, 32 * HP16)) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint2 *>(alo_s + off16(row, j0, 16 * HP16)) =
        make_uint2(pack_bf16x2(a[0] - h[0], a[1] - h[1]), pack_bf16x2(a[2] - h[2], a[3] - h[3]));
}

// =============================================================================================
// forward: x -> first sine layer (FFMA) -> L hidden layers (tcgen05) -> output layer (FFMA) -> y / J / h2
// TAPE: additionally leaves the pre-activations and activations of every sine layer in the caller's planes
// =============================================================================================
template <int D, int O, int ORDER, int HP16, bool TAPE>
__global__ void __launch_bounds__(MT, 1) k_mid_fwd(MidParams p, int tmem_cols) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int NQ = HP16 / 16;                        // 16-column rounds = quads per thread
    constexpr int PS = O * S;
    constexpr int NHT = nht_of(HP16, S);
    static_assert(S <= MAX_S && HP16 <= MAX_HP16 && NHT >= 1 && (S + NHT) * HP16 <= 512, "TMEM budget");
    extern __shared__ __align__(1024) unsigned char smraw_[];
    unsigned char *sm = smraw_ + ((128u - (s32(smraw_) & 127u)) & 127u);
    const SirenDims dm = p.dm;
    const int L = dm.L, H = dm.H;
    const MidSmem M = mid_smem(HP16, S, L, PS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;              // TMEM lane == point inside the tile
    const int quarter = warp >> 2;
    float *biasS = reinterpret_cast<float *>(sm + M.bias);
    float *w1S = reinterpret_cast<float *>(sm + M.w1);
    float *woS = reinterpret_cast<float *>(sm + M.wo);
    float *boS = reinterpret_cast<float *>(sm + M.bo);
    float *partS = reinterpret_cast<float *>(sm + M.part);
    const uint32_t mbar = s32(sm + M.mbar);
    const float w = dm.omega;

    // ---- small operands
    for (int idx = tid; idx < L * HP16; idx += MT) {
        const int l = idx / HP16, j = idx % HP16;
        biasS[idx] = (j < H) ? w * p.theta[insr_b_offset(dm, l + 1) + j] : 0.f;
    }
    for (int idx = tid; idx < HP16 * 4; idx += MT) {
        const int j = idx >> 2, d = idx & 3;
        float v = 0.f;
        if (j < H) {
            if (d < D) v = w * p.theta[insr_w_offset(dm, 0) + (int64_t)j * D + d];
            else if (d == 3) v = w * p.theta[insr_b_offset(dm, 0) + j];
        }
        w1S[idx] = v;
    }
    for (int idx = tid; idx < 3 * HP16; idx += MT) {
        const int o = idx / HP16, j = idx % HP16;
        woS[idx] = (o < O && j < H) ? p.theta[insr_w_offset(dm, L + 1) + (int64_t)o * H + j] : 0.f;
    }
    if (tid < 4) boS[tid] = (tid < O) ? p.theta[insr_b_offset(dm, L + 1) + tid] : 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s32(sm + M.tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        insr_tc::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    insr_tc::fence_async_smem();
    insr_tc::tc_fence_before();
    __syncthreads();
    insr_tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + M.tmem);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t phase = 0;

    WRegs<HP16, false> wr;
    wr.load(p.theta + insr_w_offset(dm, 1), H, tid);     // layer 1 of the first tile

    const int64_t ntiles = (p.N + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * TILE;
        const int64_t n = p0 + row;
        const bool valid = n < p.N;
        float xv[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) xv[d] = valid ? __ldg(p.x + n * D + d) : 0.f;

        // ---- first sine layer (FFMA)
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NQ; ++i) {
            const int j0 = 16 * i + 4 * quarter;
            float zq[S][4], aq[S][4];
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < 4; ++c) {
                const float4 wv = *reinterpret_cast<const float4 *>(w1S + (j0 + c) * 4);
                float z[S], a[S];
                float acc = wv.w;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) acc = fmaf(insr_fused::f4get(wv, d), xv[d], acc);
                z[0] = acc;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < C::ND; ++d) z[1 + d] = insr_fused::f4get(wv, d);
                INSR_PRAGMA_UNROLL
                for (int q = 1 + C::ND; q < S; ++q) z[q] = 0.f;
                insr_sine_fwd<D, ORDER>(z, a);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) { zq[s][c] = z[s]; aq[s][c] = a[s]; }
            }
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                store_operand4<HP16>(tmem_row + (uint32_t)((S + s) * HP16), s < NHT ? nullptr : sm + M.a_hi + (s - NHT) * TILE * HP16 * 4, sm + M.a_lo + s * TILE * HP16 * 2, row, j0, aq[s]);
                if (TAPE) {
                    const size_t t4 = tape_f4(tile, s, i, quarter, row, S, NQ);
                    reinterpret_cast<float4 *>(p.Zpre)[t4] = make_float4(zq[s][0], zq[s][1], zq[s][2], zq[s][3]);
                    reinterpret_cast<float4 *>(p.Act)[t4] = make_float4(aq[s][0], aq[s][1], aq[s][2], aq[s][3]);
                }
            }
        }
        float out[O][S];
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) out[o][s] = 0.f;

        // ---- hidden layers on the tensor cores
        for (int l = 1; l <= L; ++l) {
            wr.store(sm + M.w_hi, sm + M.w_lo, sm + M.w_hb, w, tid);     // the previous layer's MMAs have completed
            insr_tc::tmem_st_wait();
            insr_tc::fence_async_smem();                  // generic-proxy operand writes -> visible to the async proxy
            insr_tc::tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                insr_tc::tc_fence_after();
                if (insr_tc::elect_one()) {
                    issue_layer<HP16, S>(tmem_base, s32(sm + M.w_hi), s32(sm + M.w_lo), s32(sm + M.w_hb), s32(sm + M.a_lo), s32(sm + M.a_hi), H);
                    insr_tc::mma_commit(mbar);
                }
                __syncwarp();
            }
            // the next contraction's weights travel from L2 to registers while the tensor core works
            {
                const int ln = (l < L) ? l + 1 : 1;
                wr.load(p.theta + insr_w_offset(dm, ln), H, tid);
            }
            insr_tc::mbar_wait(mbar, phase);
            phase ^= 1;
            insr_tc::tc_fence_after();
            const bool last = (l == L);
            // accumulators of round i + 1 are requested before round i is processed (tcgen05.ld latency under the arithmetic)
            float znext[S][4];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) tmem_ld4(tmem_row + (uint32_t)(s * HP16 + 4 * quarter), znext[s]);
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < NQ; ++i) {
                const int j0 = 16 * i + 4 * quarter;
                float zz[S][4], aq[S][4];
                insr_tc::tmem_ld_wait();
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    tmem_ld_ready4(znext[s]);
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) zz[s][c] = znext[s][c];
                }
                if (i + 1 < NQ) {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) tmem_ld4(tmem_row + (uint32_t)(s * HP16 + j0 + 16), znext[s]);
                }
                const float4 bv = *reinterpret_cast<const float4 *>(biasS + (l - 1) * HP16 + j0);
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    float z[S], a[S];
                    z[0] = zz[0][c] + insr_fused::f4get(bv, c);
                    INSR_PRAGMA_UNROLL
                    for (int s = 1; s < S; ++s) z[s] = zz[s][c];
                    insr_sine_fwd<D, ORDER>(z, a);
                    zz[0][c] = z[0];
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) aq[s][c] = a[s];
                }
                if (last) {                               // output layer (FFMA) folded into the last epilogue
                    INSR_PRAGMA_UNROLL
                    for (int o = 0; o < O; ++o) {
                        const float4 wv = *reinterpret_cast<const float4 *>(woS + o * HP16 + j0);
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s)
                            INSR_PRAGMA_UNROLL
                            for (int c = 0; c < 4; ++c) out[o][s] = fmaf(insr_fused::f4get(wv, c), aq[s][c], out[o][s]);
                    }
                } else {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        store_operand4<HP16>(tmem_row + (uint32_t)((S + s) * HP16), s < NHT ? nullptr : sm + M.a_hi + (s - NHT) * TILE * HP16 * 4, sm + M.a_lo + s * TILE * HP16 * 2, row, j0, aq[s]);
                }
                if (TAPE) {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) {
                        const size_t t4 = tape_f4(tile, s, i, quarter, row, S, NQ);
                        reinterpret_cast<float4 *>(p.Zpre + (size_t)l * p.buf)[t4] = make_float4(zz[s][0], zz[s][1], zz[s][2], zz[s][3]);
                        reinterpret_cast<float4 *>(p.Act + (size_t)l * p.buf)[t4] = make_float4(aq[s][0], aq[s][1], aq[s][2], aq[s][3]);
                    }
                }
            }
        }
        // ---- combine the four column quarters of every point
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) partS[((size_t)quarter * TILE + row) * PS + o * S + s] = out[o][s];
        insr_tc::tc_fence_before();
        __syncthreads();
        if (quarter == 0 && valid && p.y) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                float r[S];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    float acc = 0.f;
                    INSR_PRAGMA_UNROLL
                    for (int q = 0; q < 4; ++q) acc += partS[((size_t)q * TILE + row) * PS + o * S + s];
                    r[s] = acc;
                }
                r[0] += boS[o];
                insr_store_outputs<D, O, ORDER>(n, o, r, p.y, p.jac, p.h2);
            }
        }
        __syncthreads();                                  // partS is rewritten by the next tile
        insr_tc::tc_fence_after();
    }
    // ---- TMEM release
    insr_tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// =============================================================================================
// reverse sweep, part 1: the data-gradient chain.  Output cotangents -> adjoint of the last sine layer -> for l = L .. 1:
// abar_{l-1} = zbar_l . (omega W_l) on the tensor cores (same operand scheme as the forward pass, against the transposed
// weights) -> adjoint of sine layer l-1 against its taped pre-activations.  zbar_l REPLACES the pre-activations of layer
// l in the tape (each thread overwrites exactly the float4 it has just read), which is what the weight-gradient kernels
// below consume; d loss / d x rides on the last adjoint.  Replaces k_tiled_out_bwd + L x k_wide_tc<MODE 1> + k_tiled_gx.
// =============================================================================================
template <int D, int O, int ORDER, int HP16>
__global__ void __launch_bounds__(MT, 1) k_mid_dgrad(MidParams p, int tmem_cols) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int NQ = HP16 / 16;
    constexpr int PS = D;
    constexpr int NHT = nht_of(HP16, S);
    extern __shared__ __align__(1024) unsigned char smraw_[];
    unsigned char *sm = smraw_ + ((128u - (s32(smraw_) & 127u)) & 127u);
    const SirenDims dm = p.dm;
    const int L = dm.L, H = dm.H;
    const MidSmem M = mid_smem(HP16, S, L, PS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;
    const int quarter = warp >> 2;
    float *w1S = reinterpret_cast<float *>(sm + M.w1);
    float *woS = reinterpret_cast<float *>(sm + M.wo);
    float *partS = reinterpret_cast<float *>(sm + M.part);
    const uint32_t mbar = s32(sm + M.mbar);
    const float w = dm.omega;

    for (int idx = tid; idx < HP16 * 4; idx += MT) {       // omega W1[j][d] (for d loss / d x)
        const int j = idx >> 2, d = idx & 3;
        w1S[idx] = (j < H && d < D) ? w * p.theta[insr_w_offset(dm, 0) + (int64_t)j * D + d] : 0.f;
    }
    for (int idx = tid; idx < 3 * HP16; idx += MT) {
        const int o = idx / HP16, j = idx % HP16;
        woS[idx] = (o < O && j < H) ? p.theta[insr_w_offset(dm, L + 1) + (int64_t)o * H + j] : 0.f;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s32(sm + M.tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        insr_tc::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    insr_tc::fence_async_smem();
    insr_tc::tc_fence_before();
    __syncthreads();
    insr_tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + M.tmem);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t phase = 0;

    WRegs<HP16, true> wr;
    wr.load(p.theta + insr_w_offset(dm, L), H, tid);

    const int64_t ntiles = (p.N + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n = tile * TILE + row;
        const bool valid = n < p.N;
        float g[O][S];
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) {
            if (valid) {
                insr_load_cotangents<D, O, ORDER>(n, o, p.gy, p.gjac, p.gh2, g[o]);
            } else {
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) g[o][s] = 0.f;
            }
        }
        // DRAM latency of the tape is taken off the critical path: the layers that are needed later are pulled into L2 now
        for (int l = L - 1; l >= 0; --l) {
            const float4 *zt = reinterpret_cast<const float4 *>(p.Zpre + (size_t)l * p.buf);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < NQ; ++i)
                    if ((lane & 1) == 0)        // one prefetch per 32-byte sector
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(zt + tape_f4(tile, s, i, quarter, row, S, NQ)));
        }
        // ---- adjoint of the last sine layer: abar_L = Wo^T g  ->  zbar_L
        {
            float4 *zt = reinterpret_cast<float4 *>(p.Zpre + (size_t)L * p.buf);
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < NQ; ++i) {
                const int j0 = 16 * i + 4 * quarter;
                float4 z4[S];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) z4[s] = zt[tape_f4(tile, s, i, quarter, row, S, NQ)];
                float zbq[S][4];
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    float z[S], ab[S], zb[S];
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) { z[s] = insr_fused::f4get(z4[s], c); ab[s] = 0.f; }
                    INSR_PRAGMA_UNROLL
                    for (int o = 0; o < O; ++o) {
                        const float wv = woS[o * HP16 + j0 + c];
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s) ab[s] = fmaf(wv, g[o][s], ab[s]);
                    }
                    insr_sine_bwd<D, ORDER>(z, ab, zb);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) zbq[s][c] = zb[s];
                }
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    zt[tape_f4(tile, s, i, quarter, row, S, NQ)] = make_float4(zbq[s][0], zbq[s][1], zbq[s][2], zbq[s][3]);
                    store_operand4<HP16>(tmem_row + (uint32_t)((S + s) * HP16), s < NHT ? nullptr : sm + M.a_hi + (s - NHT) * TILE * HP16 * 4, sm + M.a_lo + s * TILE * HP16 * 2, row, j0, zbq[s]);
                }
            }
        }
        float gxp[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) gxp[d] = 0.f;
        // ---- hidden layers, last to first
        for (int l = L; l >= 1; --l) {
            wr.store(sm + M.w_hi, sm + M.w_lo, sm + M.w_hb, w, tid);
            insr_tc::tmem_st_wait();
            insr_tc::fence_async_smem();
            insr_tc::tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                insr_tc::tc_fence_after();
                if (insr_tc::elect_one()) {
                    issue_layer<HP16, S>(tmem_base, s32(sm + M.w_hi), s32(sm + M.w_lo), s32(sm + M.w_hb), s32(sm + M.a_lo), s32(sm + M.a_hi), H);
                    insr_tc::mma_commit(mbar);
                }
                __syncwarp();
            }
            {
                const int ln = (l > 1) ? l - 1 : L;
                wr.load(p.theta + insr_w_offset(dm, ln), H, tid);
            }
            float4 *zt = reinterpret_cast<float4 *>(p.Zpre + (size_t)(l - 1) * p.buf);
            // the tape of the previous layer travels while the tensor core works: the first round's pre-activations are
            // requested now, every later round one round ahead of its use
            float4 znext[S];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) znext[s] = zt[tape_f4(tile, s, 0, quarter, row, S, NQ)];
            insr_tc::mbar_wait(mbar, phase);
            phase ^= 1;
            insr_tc::tc_fence_after();
            float accn[S][4];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) tmem_ld4(tmem_row + (uint32_t)(s * HP16 + 4 * quarter), accn[s]);
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < NQ; ++i) {
                const int j0 = 16 * i + 4 * quarter;
                float acc[S][4];
                float4 z4[S];
                insr_tc::tmem_ld_wait();
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    tmem_ld_ready4(accn[s]);
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) acc[s][c] = accn[s][c];
                    z4[s] = znext[s];
                }
                if (i + 1 < NQ) {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) {
                        tmem_ld4(tmem_row + (uint32_t)(s * HP16 + j0 + 16), accn[s]);
                        znext[s] = zt[tape_f4(tile, s, i + 1, quarter, row, S, NQ)];
                    }
                }
                float zbq[S][4];
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    float z[S], ab[S], zb[S];
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) { z[s] = insr_fused::f4get(z4[s], c); ab[s] = acc[s][c]; }
                    insr_sine_bwd<D, ORDER>(z, ab, zb);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) zbq[s][c] = zb[s];
                }
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    zt[tape_f4(tile, s, i, quarter, row, S, NQ)] = make_float4(zbq[s][0], zbq[s][1], zbq[s][2], zbq[s][3]);
                    if (l > 1) store_operand4<HP16>(tmem_row + (uint32_t)((S + s) * HP16), s < NHT ? nullptr : sm + M.a_hi + (s - NHT) * TILE * HP16 * 4, sm + M.a_lo + s * TILE * HP16 * 2, row, j0, zbq[s]);
                }
                if (l == 1 && p.gx) {                      // d loss / d x = (omega W1)^T zbar_0 (value stream)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) {
                        const float4 wv = *reinterpret_cast<const float4 *>(w1S + (j0 + c) * 4);
                        INSR_PRAGMA_UNROLL
                        for (int d = 0; d < D; ++d) gxp[d] = fmaf(insr_fused::f4get(wv, d), zbq[0][c], gxp[d]);
                    }
                }
            }
        }
        if (p.gx) {
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) partS[((size_t)quarter * TILE + row) * PS + d] = gxp[d];
            __syncthreads();
            if (quarter == 0 && valid) {
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) {
                    float acc = 0.f;
                    INSR_PRAGMA_UNROLL
                    for (int q = 0; q < 4; ++q) acc += partS[((size_t)q * TILE + row) * PS + d];
                    p.gx[n * D + d] = acc;
                }
            }
            __syncthreads();
        }
    }
    insr_tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// =============================================================================================
// reverse sweep, part 2: hidden-layer weight gradients of ALL layers in one launch (grid.y = layer).
//     gW_l[j][k] += omega sum_{s,p} zbar_l[s][p][j] act_{l-1}[s][p][k],      gb_l[j] += omega sum_p zbar_l[0][p][j]
// The arithmetic of k_wide_wgrad (siren_wide_tc.cuh: both operands MN-major, two bf16 levels each, all four cross
// products, accumulators resident in TMEM, one red.global per weight per CTA) reading the private tape layout: a
// stage = 64 points of one stream, every warp load a contiguous 512-byte run.
// =============================================================================================
constexpr int WGT = 256;
constexpr int WG_PTS = 64;
constexpr int WG_ATOM = WG_PTS * 128;
// kind::f16, D = F32, A / B = BF16, both MN-major, N = 128, M = 128; MN-major operand in the 128-byte swizzle (siren_wide_tc.cuh)
constexpr uint32_t IDESC_WG128 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ uint64_t desc_mn128(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// thin layers (first / output layer gradients): see k_mid_edge_role below
template <int D, int O, int ORDER, int HP16>
__device__ __forceinline__ void mid_edge_role(const MidParams &p, int cta, int nctas);

template <int D, int O, int ORDER, int HP16>
__global__ void __launch_bounds__(WGT, 2) k_mid_wgrad(MidParams p) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int NQ = HP16 / 16;
    constexpr int NCH = HP16 / 4;                            // float4 chunks per row
    constexpr int NI = NCH / 4;                              // chunks per thread
    constexpr int NB = (HP16 + 63) / 64;                     // blocks of 64 neurons on either side
    const SirenDims dm = p.dm;
    if ((int)blockIdx.y == dm.L) {                           // the extra grid row: thin layers, beside the tensor-core CTAs
        mid_edge_role<D, O, ORDER, HP16>(p, (int)blockIdx.x, (int)gridDim.x);
        return;
    }
    extern __shared__ __align__(1024) unsigned char smraw_[];
    unsigned char *sm = smraw_ + ((1024u - (s32(smraw_) & 1023u)) & 1023u);
    const int layer = 1 + (int)blockIdx.y;
    const int64_t buf = p.buf;
    const int nv = (int)p.N;
    const float4 *ZB = reinterpret_cast<const float4 *>(p.Zpre + (size_t)layer * buf);
    const float4 *AC = reinterpret_cast<const float4 *>(p.Act + (size_t)(layer - 1) * buf);
    float *gW = p.gtheta + insr_w_offset(dm, layer), *gb = p.gtheta + insr_b_offset(dm, layer);
    unsigned char *zt = sm, *at = sm + 4 * WG_ATOM;          // [block][level][64 x 128 B] each
    float *bsumS = reinterpret_cast<float *>(sm + 8 * WG_ATOM);
    const uint32_t mbar = s32(sm + 8 * WG_ATOM + 512), tslot = mbar + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = dm.H;
    constexpr int TCOLS = NB * NB * 128;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tslot), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        insr_tc::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 128) bsumS[tid] = 0.f;
    insr_tc::tc_fence_before();
    __syncthreads();
    insr_tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + 8 * WG_ATOM + 512 + 16);

    // thread -> (row rr of the 64-point group, chunks ch(t)): a warp covers 16 rows x 2 adjacent chunks, so that its
    // 8-byte operand stores touch every bank pair exactly twice (two wavefronts: the minimum) and every global load is
    // two contiguous 256-byte runs
    const int rr = 16 * (warp & 3) + (lane & 7) + 8 * (lane >> 4);
    const int ch0 = 2 * (warp >> 2) + ((lane >> 3) & 1);      // chunks ch0 + 4 t
    float4 rz[2][NI], ra[2][NI], bsum[NI];
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < NI; ++t) bsum[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ngroups = (nv + WG_PTS - 1) / WG_PTS;
    const int64_t nstages = (int64_t)((ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * S;
    auto stage_coords = [&](int64_t st, int &g, int &s) { g = (int)blockIdx.x + (int)(st / S) * (int)gridDim.x; s = (int)(st % S); };
    auto gload = [&](int64_t st, float4 (&z)[NI], float4 (&a)[NI]) {
        int g, s;
        stage_coords(st, g, s);
        const int64_t T = g >> 1;
        const int r = 64 * (g & 1) + rr;
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < NI; ++t) {
            const int ch = ch0 + 4 * t;                       // neurons 4 ch .. = round ch / 4, quarter ch % 4
            const size_t idx = tape_f4(T, s, ch >> 2, ch & 3, r, S, NQ);
            z[t] = __ldg(ZB + idx);
            a[t] = __ldg(AC + idx);
        }
    };
    auto split4 = [&](const float4 &v, uint2 &l1, uint2 &l2) {
        const uint32_t a = pack_bf16x2(v.x, v.y), b = pack_bf16x2(v.z, v.w);
        const float r0_ = v.x - __uint_as_float(a << 16), r1_ = v.y - __uint_as_float(a & 0xFFFF0000u);
        const float r2_ = v.z - __uint_as_float(b << 16), r3_ = v.w - __uint_as_float(b & 0xFFFF0000u);
        l1 = make_uint2(a, b);
        l2 = make_uint2(pack_bf16x2(r0_, r1_), pack_bf16x2(r2_, r3_));
    };
    auto sstore = [&](int64_t st, const float4 (&z)[NI], const float4 (&a)[NI]) {
        int g, s;
        stage_coords(st, g, s);
        INSR_PRAGMA_UNROLL
        for (int t = 0; t < NI; ++t) {
            const int j = 4 * (ch0 + 4 * t);
            const int jb = j >> 6, jj = j & 63;
            const int off = rr * 128 + ((((jj >> 3) ^ rr) & 7) << 4) + (jj & 7) * 2;
            uint2 l1, l2;
            split4(z[t], l1, l2);
            *reinterpret_cast<uint2 *>(zt + (jb * 2 + 0) * WG_ATOM + off) = l1;
            *reinterpret_cast<uint2 *>(zt + (jb * 2 + 1) * WG_ATOM + off) = l2;
            if (s == 0) { bsum[t].x += z[t].x; bsum[t].y += z[t].y; bsum[t].z += z[t].z; bsum[t].w += z[t].w; }
            split4(a[t], l1, l2);
            *reinterpret_cast<uint2 *>(at + (jb * 2 + 0) * WG_ATOM + off) = l1;
            *reinterpret_cast<uint2 *>(at + (jb * 2 + 1) * WG_ATOM + off) = l2;
        }
    };
    if (HP16 % 64 != 0) {                                    // the unused neuron slots of the last block stay zero
        for (int idx = tid; idx < 8 * WG_ATOM / 16; idx += WGT) reinterpret_cast<uint4 *>(sm)[idx] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
    }

    // two stages in flight in registers (the loads of stage st + 2 are issued while stage st is in the tensor core)
    uint32_t phase = 0;
    if (nstages > 0) gload(0, rz[0], ra[0]);
    if (nstages > 1) gload(1, rz[1], ra[1]);
    for (int64_t st = 0; st < nstages; ++st) {
        if (st & 1) sstore(st, rz[1], ra[1]); else sstore(st, rz[0], ra[0]);
        insr_tc::fence_async_smem();
        insr_tc::tc_fence_before();
        __syncthreads();
        if (st + 2 < nstages) { if (st & 1) gload(st + 2, rz[1], ra[1]); else gload(st + 2, rz[0], ra[0]); }
        if (warp == 0) {
            insr_tc::tc_fence_after();
            if (insr_tc::elect_one()) {
                INSR_PRAGMA_UNROLL
                for (int jb = 0; jb < NB; ++jb)
                    INSR_PRAGMA_UNROLL
                    for (int kb = 0; kb < NB; ++kb) {
                        const uint32_t d = tmem_base + (uint32_t)((jb * NB + kb) * 128);
                        const uint32_t za = s32(zt + jb * 2 * WG_ATOM), aa = s32(at + kb * 2 * WG_ATOM);
                        INSR_PRAGMA_UNROLL
                        for (int q = 0; q < 4; ++q) {          // 16 points per instruction = two 1024-byte row groups
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                :: "r"(d), "l"(desc_mn128(za + 2048 * q, WG_ATOM)), "l"(desc_mn128(aa + 2048 * q, WG_ATOM)),
                                   "r"(IDESC_WG128), "r"((st > 0 || q > 0) ? 1u : 0u) : "memory");
                        }
                    }
                insr_tc::mma_commit(mbar);
            }
            __syncwarp();
        }
        insr_tc::mbar_wait(mbar, phase);
        phase ^= 1;
        insr_tc::tc_fence_after();
    }

    // ---- flush: bias gradient, then the weight blocks (the four level-products of a weight are held by four warps)
    INSR_PRAGMA_UNROLL
    for (int t = 0; t < NI; ++t) {
        const int j = 4 * (ch0 + 4 * t);
        float v[4] = {bsum[t].x, bsum[t].y, bsum[t].z, bsum[t].w};
        INSR_PRAGMA_UNROLL
        for (int c = 0; c < 4; ++c) {                           // lanes with the same chunk: bit 3 of the lane index is the chunk
            float a = v[c];
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            if ((lane & 23) == 0 && j + c < 128) atomicAdd(bsumS + j + c, a);
        }
    }
    __syncthreads();
    if (tid < 128 && tid < H && bsumS[tid] != 0.f) atomicAdd(gb + tid, dm.omega * bsumS[tid]);
    if (nstages > 0) {
        float *piece = reinterpret_cast<float *>(sm);            // [8 warps][32 rows][33] floats (operand region is free)
        const int half = warp >> 2;                              // columns 64 half .. + 63: level of `a`
        const uint32_t trow = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);   // rows: level of z (warp & 2), neurons 32 (warp & 1) ..
        for (int jb = 0; jb < NB; ++jb)
            for (int kb = 0; kb < NB; ++kb)
                for (int cc = 0; cc < 2; ++cc) {                 // 32 input columns at a time
                    __syncthreads();
                    INSR_PRAGMA_UNROLL
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float v[8];
                        insr_tc::tmem_ld8(trow + (uint32_t)((jb * NB + kb) * 128 + 64 * half + 32 * cc + 8 * c8), v);
                        insr_tc::tmem_ld_wait();
                        INSR_PRAGMA_UNROLL
                        for (int i = 0; i < 8; ++i) piece[(warp * 32 + lane) * 33 + 8 * c8 + i] = v[i];
                    }
                    __syncthreads();
                    for (int idx = tid; idx < 64 * 32; idx += WGT) {
                        const int jl = idx >> 5, kl = idx & 31;
                        const int j = jb * 64 + jl, k = kb * 64 + 32 * cc + kl;
                        if (j < H && k < H) {
                            const int wq = jl >> 5, rrow = (jl & 31) * 33 + kl;
                            const float v = (piece[(wq + 0) * 32 * 33 + rrow] + piece[(wq + 2) * 32 * 33 + rrow]) +
                                            (piece[(wq + 4) * 32 * 33 + rrow] + piece[(wq + 6) * 32 * 33 + rrow]);
                            if (v != 0.f) atomicAdd(gW + (size_t)j * H + k, dm.omega * v);
                        }
                    }
                }
    }
    insr_tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

// =============================================================================================
// reverse sweep, part 3: the thin layers.  First sine layer (gW1, gb1 from zbar_0 and the points) and output layer
// (gWo from the cotangents and the last activations, gbo): reductions over points, lanes along the rows of a tile (every
// load a contiguous 512-byte run of the tape), one warp per group of 4 neurons, a butterfly over the 32 lanes at the end.
// Replaces k_tiled_edge.
// =============================================================================================
template <int D, int O, int ORDER, int HP16>
__device__ __forceinline__ void mid_edge_role(const MidParams &p, int cta, int nctas) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int NQ = HP16 / 16;
    constexpr int NCH = HP16 / 4;
    const SirenDims dm = p.dm;
    const int L = dm.L, H = dm.H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float4 *ZB0 = reinterpret_cast<const float4 *>(p.Zpre);
    const float4 *ACL = reinterpret_cast<const float4 *>(p.Act + (size_t)L * p.buf);
    const int64_t ntiles = (p.N + TILE - 1) / TILE;
    auto wsum = [](float v) {
        INSR_PRAGMA_UNROLL
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        return v;
    };
    for (int ch = warp; ch < NCH; ch += nwarps) {           // neurons 4 ch .. 4 ch + 3
        const int i = ch >> 2, q = ch & 3;
        float gb1[4], gw1[D][4], gwo[O][4], gbo[O];
        INSR_PRAGMA_UNROLL
        for (int c = 0; c < 4; ++c) {
            gb1[c] = 0.f;
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) gw1[d][c] = 0.f;
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) gwo[o][c] = 0.f;
        }
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) gbo[o] = 0.f;
        for (int64_t T = cta; T < ntiles; T += nctas) {
            INSR_PRAGMA_UNROLL
            for (int rq = 0; rq < 4; ++rq) {
                const int r = 32 * rq + lane;
                const int64_t n = T * TILE + r;
                const bool valid = n < p.N;                    // (zbar of the rows beyond the batch is zero anyway)
                const int64_t nc = valid ? n : 0;
                float xv[D];
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) xv[d] = __ldg(p.x + nc * D + d);
                const float4 z0 = __ldg(ZB0 + tape_f4(T, 0, i, q, r, S, NQ));
                float4 zd[C::ND > 0 ? C::ND : 1], a4[S];
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < C::ND; ++d) zd[d] = __ldg(ZB0 + tape_f4(T, 1 + d, i, q, r, S, NQ));
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) a4[s] = __ldg(ACL + tape_f4(T, s, i, q, r, S, NQ));
                float g[O][S];
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o) {
                    insr_load_cotangents<D, O, ORDER>(nc, o, p.gy, p.gjac, p.gh2, g[o]);
                    if (!valid) {
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s) g[o][s] = 0.f;
                    }
                }
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const float zc = valid ? insr_fused::f4get(z0, c) : 0.f;
                    gb1[c] += zc;
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) gw1[d][c] = fmaf(zc, xv[d], gw1[d][c]);
                }
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < C::ND; ++d)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) gw1[d][c] += valid ? insr_fused::f4get(zd[d], c) : 0.f;
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int o = 0; o < O; ++o)
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) gwo[o][c] = fmaf(g[o][s], insr_fused::f4get(a4[s], c), gwo[o][c]);
                if (ch == 0) {
                    INSR_PRAGMA_UNROLL
                    for (int o = 0; o < O; ++o) gbo[o] += g[o][0];
                }
            }
        }
        INSR_PRAGMA_UNROLL
        for (int c = 0; c < 4; ++c) {
            const int j = 4 * ch + c;
            const float b1 = wsum(gb1[c]);
            float w1[D], wo[O];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) w1[d] = wsum(gw1[d][c]);
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) wo[o] = wsum(gwo[o][c]);
            if (lane == 0 && j < H) {
                atomicAdd(p.gtheta + insr_b_offset(dm, 0) + j, dm.omega * b1);
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) atomicAdd(p.gtheta + insr_w_offset(dm, 0) + j * D + d, dm.omega * w1[d]);
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o) atomicAdd(p.gtheta + insr_w_offset(dm, L + 1) + o * H + j, wo[o]);
            }
        }
        if (ch == 0) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                const float v = wsum(gbo[o]);
                if (lane == 0) atomicAdd(p.gtheta + insr_b_offset(dm, L + 1) + o, v);
            }
        }
    }
}

inline bool mid_enabled() {
    static const bool on = [] { const char *e = getenv("INSR_MID"); return !(e && e[0] == '0'); }();
    return on;
}
// shapes the fused kernels serve: widths up to 80 (the elasticity scripts; H <= 32 nets that the resident-weights family does not
// instantiate -- D = 3, deeper than 3 hidden layers -- run at the padded width 64) with at most 4 streams
inline bool mid_supported(const SirenDims &dm, int order) {
    const int S = insr_nstreams(dm.D, order);
    if (!(mid_enabled() && dm.H >= 1 && hp16_of(dm.H) <= MAX_HP16 && S <= MAX_S && dm.L >= 1 && dm.L <= 8)) return false;
    const int hp = hp16_of(dm.H) <= 64 ? 64 : 80;
    return mid_smem(hp, S, dm.L, 0).total + 128 <= 232448;      // 227 KB of shared memory per CTA
}

template <int D, int O, int ORDER, int HP16, bool TAPE>
int launch_mid_fwd_hp(MidParams &p, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    if constexpr (S > MAX_S) {
        return -6;
    } else {
        const MidSmem M = mid_smem(HP16, S, p.dm.L, O * S);
        auto kfn = k_mid_fwd<D, O, ORDER, HP16, TAPE>;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total + 128);
        const int64_t tiles = (p.N + TILE - 1) / TILE;
        const int64_t sms = insr_fused::sm_count();
        const int64_t ctas = tiles < sms ? tiles : sms;
        kfn<<<dim3((unsigned)ctas), dim3(MT), M.total + 128, reinterpret_cast<cudaStream_t>(stream)>>>(p, insr_tc::pow2_cols((S + nht_of(HP16, S)) * HP16));
        ++*launches;
        return 0;
    }
}
template <int D, int O, int ORDER, bool TAPE>
int launch_mid_fwd(MidParams &p, void *stream, int64_t *launches) {
    const int hp = hp16_of(p.dm.H);                     // widths instantiated: 64 (H <= 64, padded) and 80
    if (hp <= 64) return launch_mid_fwd_hp<D, O, ORDER, 64, TAPE>(p, stream, launches);
    if (hp <= 80) return launch_mid_fwd_hp<D, O, ORDER, 80, TAPE>(p, stream, launches);
    return -6;
}

template <int D, int O, int ORDER, int HP16>
int launch_mid_bwd_hp(MidParams &p, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    if constexpr (S > MAX_S) {
        return -6;
    } else {
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        const int64_t tiles = (p.N + TILE - 1) / TILE;
        const int64_t sms = insr_fused::sm_count();
        {   // data-gradient chain (+ d loss / d x)
            const MidSmem M = mid_smem(HP16, S, p.dm.L, D);
            auto kfn = k_mid_dgrad<D, O, ORDER, HP16>;
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total + 128);
            const int64_t ctas = tiles < sms ? tiles : sms;
            kfn<<<dim3((unsigned)ctas), dim3(MT), M.total + 128, st>>>(p, insr_tc::pow2_cols((S + nht_of(HP16, S)) * HP16));
            ++*launches;
        }
        {   // hidden-layer weight gradients of all layers (grid rows 0 .. L-1) and, beside them, the thin layers (row L)
            const size_t smem = (size_t)8 * WG_ATOM + 512 + 64 + 1024;
            auto kfn = k_mid_wgrad<D, O, ORDER, HP16>;
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const int groups = (int)((p.N + WG_PTS - 1) / WG_PTS);
            int ctas = (int)(sms / p.dm.L);
            if (ctas < 1) ctas = 1;
            if (ctas > groups) ctas = groups;
            kfn<<<dim3((unsigned)ctas, (unsigned)p.dm.L + 1), dim3(WGT), smem, st>>>(p);
            ++*launches;
        }
        return 0;
    }
}
template <int D, int O, int ORDER>
int launch_mid_bwd(MidParams &p, void *stream, int64_t *launches) {
    const int hp = hp16_of(p.dm.H);
    if (hp <= 64) return launch_mid_bwd_hp<D, O, ORDER, 64>(p, stream, launches);
    if (hp <= 80) return launch_mid_bwd_hp<D, O, ORDER, 80>(p, stream, launches);
    return -6;
}

}  // namespace insr_mid
#endif  // !INSR_CPU_EMU
