// siren_wide_tc.cuh -- tcgen05 / TMEM hidden-layer GEMMs of the tiled family (32 < H <= 512, S <= 4 streams).
//
// Replaces the FFMA GEMM + epilogue kernels k_tiled_fwd / k_tiled_dgrad (siren_tiled.cuh) -- 75 % of that family's
// time -- with one kernel per hidden layer on the tensor cores, same buffers ([stream][point][HP] fp32 in the
// caller's workspace), same epilogues, 3xTF32 precision (a = a_hi + a_lo, w = w_hi + w_lo, three products, FP32
// accumulation in TMEM; see siren_tc.cuh for the measured hardware facts this relies on).
//
//   MODE 0  forward layer   D_s[128 x NCOL] = A_s[128 x K] . (omega W)^T      epilogue: + omega b, sine-stream
//                           activation (insr_sine_fwd), writes the pre-activations (tape) and the post-activations
//   MODE 1  data gradient   D_s[128 x NCOL] = Zbar_s[128 x K] . (omega W)     epilogue: activation adjoint of the
//                           previous layer against its tape (insr_sine_bwd), writes zbar_{l-1}
//
// CTA = 128 points (TMEM lanes) x one pass of NCOL <= 128 output columns, all S streams (S NCOL <= 512 TMEM columns).
// The reduction runs in slabs of 32 (one 128-byte swizzle row of fp32): every thread loads its share of the next slab
// (S A-tiles of 128 x 32 + the weight tile) from HBM / L2 into REGISTERS while the tensor core works on the current
// one, then splits hi / lo and stores both halves in the K-major 128-byte-swizzle layout (conflict-free 16-byte
// stores: 8 lanes cover one 128-byte row).  36 MMAs (S = 3) or 48 (S = 4) of N = NCOL per slab, issued by one elected
// lane of warp 0.  Per layer the kernel moves 4 S HP bytes/point in and 4-8 S HP out against 2 S HP^2 flops: it is
// HBM-bound for H <= 256.
//
// Only compiled by nvcc (inline PTX); the host-side SIMT emulation keeps using the FFMA kernels.
#pragma once
#include <cstdlib>
#include "siren_tc.cuh"

#ifndef INSR_CPU_EMU
namespace insr_wide {

using insr_tc::s32;
constexpr int WT = 512;                 // threads of k_wide_tc: 16 warps (4 per TMEM lane quadrant) hide the epilogue's latency chains
constexpr int WGT = 256;                // threads of k_wide_wgrad
constexpr int TILE = 128;               // points per CTA
constexpr int KS = 32;                  // reduction slab
constexpr int A_TILE = TILE * KS * 4;   // 16 KB: one 128 x 32 fp32 operand tile

struct WGeo {
    int HP, NK, NCOL, passes;
};
inline WGeo make_wgeo(int H) {
    WGeo g;
    g.HP = (H + 7) & ~7;
    g.NK = (g.HP + KS - 1) / KS;
    const int hp16 = (g.HP + 15) & ~15;
    if (hp16 <= 128) { g.NCOL = hp16; g.passes = 1; }
    else { g.NCOL = 128; g.passes = (g.HP + 127) / 128; }
    return g;
}
inline size_t smem_bytes(int S, int NCOL, bool k16 = false) {
    const int ks = k16 ? 16 : KS, est = k16 ? 20 : 36;
    const size_t operands = (size_t)2 * S * TILE * ks * 4 + (size_t)2 * NCOL * ks * 4 + 64;  // + mbarrier / TMEM slot
    const size_t staging = (size_t)2 * S * TILE * est * 4;                                   // epilogue tiles (reuse the operand region)
    return (operands > staging ? operands : staging) + 1024;
}

// byte offset of element (row, kk) of a K-major operand tile in the 128-byte swizzle: rows of 128 B (32 fp32 along K),
// 8-row atoms of 1024 B, 16-byte chunk index XORed with (row & 7)   (probe2: verified for A and B, k-steps at +32 B)
__device__ __forceinline__ int sw_off(int row, int kk) {
    return (row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ row) & 7) << 4) + (kk & 3) * 4;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// byte offset of the 16-byte chunk `ch` of operand row `row`: 128-byte swizzle (rows of 128 B, chunk ^ (row & 7)) or, K16, the
// 64-byte swizzle (rows of 64 B, 8-row atoms of 512 B, chunk ^ ((row >> 1) & 3); profiles/r1_tcgen05_probes.txt)
template <bool K16>
__device__ __forceinline__ int a_chunk_off(int row, int ch) {
    if (K16) return (row >> 3) * 512 + (row & 7) * 64 + (((ch ^ (row >> 1)) & 3) << 4);
    return (row >> 3) * 1024 + (row & 7) * 128 + (((ch ^ row) & 7) << 4);
}
template <bool K16>
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
    if (K16) return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
                    ((uint64_t)4 << 61);
    return desc_sw128(saddr);
}

// S <= 2 (value-only evaluations of the frozen nets, 1-D fields): 2 CTAs per SM (<= 256 TMEM columns, <= 97 KB each), so one
// CTA's MMAs and loads overlap the other's epilogue
#ifdef INSR_WIDE_PROFILE
__device__ unsigned long long g_wide_prof[16];
#endif

// OUT = true (last hidden layer of a forward, single column pass): the output layer  y = Wo act + bo  rides on the
// epilogue's cooperative copy -- each thread dots its 4 staged activations of a row with Wo, 8 lanes reduce by shuffle --
// so the (N, S, HP) activations are not read again by a separate kernel, and not even written when no tape is kept.
struct WideOut {
    const float *Wo, *bo;          // (O, H), (O)
    int O, nv;                     // outputs, valid rows of this chunk
    int64_t n0;                    // first point of the chunk in y / jac / h2
    float *y, *jac, *h2;
};

template <int D, int ORDER, int MODE, bool OUT = false, bool K16 = false>
__global__ void __launch_bounds__(K16 ? 256 : WT, (K16 || StreamCfg<D, ORDER>::S <= 2) ? 2 : 1)
k_wide_tc(SirenDims dm, int HP, int NK, int NCOL, int tmem_cols, const float *__restrict__ W,
                                                   const float *__restrict__ bias, const float *__restrict__ Ain,
                                                   int64_t NCp, int64_t p_base, const float *__restrict__ Ztape,
                                                   float *__restrict__ Zout, float *__restrict__ Aout, WideOut wo) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    static_assert(S <= 4, "S * NCOL must fit the 512 TMEM columns");
    // K16 (opt-in, S >= 3): 16-wide K slabs in the 64-byte swizzle, 256 threads, 16-column epilogue stage -- half the
    // shared memory, so that two CTAs share an SM (one CTA's loads and MMAs under the other's epilogue)
    constexpr int KSL = K16 ? 16 : KS;                  // reduction slab
    constexpr int NT_ = K16 ? 256 : WT;                 // threads
    constexpr int ATL = TILE * KSL * 4;                 // one 128 x KSL fp32 operand tile
    constexpr int CPR = KSL / 4;                        // 16-byte chunks per operand row
    constexpr int LCPR = K16 ? 2 : 3;                   // log2(CPR)
    extern __shared__ __align__(1024) unsigned char smraw_[];
    unsigned char *sm = smraw_ + ((1024u - (s32(smraw_) & 1023u)) & 1023u);
    unsigned char *a_hi = sm, *a_lo = sm + S * ATL;
    unsigned char *b_hi = sm + 2 * S * ATL, *b_lo = b_hi + NCOL * KSL * 4;
    const uint32_t mbar = s32(b_lo + NCOL * KSL * 4), tslot = mbar + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = dm.H;
    const float w = dm.omega;
    const int64_t p0 = p_base + (int64_t)blockIdx.x * TILE;
    const int j0 = blockIdx.y * NCOL;                   // first output column of this pass

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tslot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        insr_tc::mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    insr_tc::tc_fence_before();
    __syncthreads();
    insr_tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(b_lo + NCOL * KSL * 4 + 16);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);

    // ---- staging assignment: A = S tiles of 128 rows x 8 chunks (16 B) -> 4 S chunks per thread; B = NCOL x 32 scalars
    constexpr int NA = 128 * CPR * S / NT_;
    constexpr int NB = 128 * KSL / NT_;
    float4 ra[NA];
    float rb[NB];
    auto gload = [&](int k0) {
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NA; ++i) {
            const int c = tid + NT_ * i;                // chunk index over (stream, row, chunk)
            const int s = c >> (7 + LCPR), row = (c >> LCPR) & 127, ch = c & (CPR - 1);
            const int k = k0 + 4 * ch;
            ra[i] = (k < HP) ? __ldg(reinterpret_cast<const float4 *>(Ain + ((int64_t)s * NCp + p0 + row) * HP + k))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NB; ++i) {                  // raw weights: omega is applied in sstore (no use of the value here, so the
            const int e = tid + NT_ * i;                // loads stay in flight while the tensor core works)
            float v = 0.f;
            if (MODE == 0) {                            // B[n][kk] = omega W[j0 + n][k0 + kk]
                const int n = e >> (LCPR + 2), kk = e & (KSL - 1);
                const int j = j0 + n, k = k0 + kk;
                if (n < NCOL && j < H && k < H) v = __ldg(W + (size_t)j * H + k);
            } else {                                    // B[n][kk] = omega W[k0 + kk][j0 + n]   (n = input neuron of the layer)
                const int n = e & 127, kk = e >> 7;
                const int kin = j0 + n, j = k0 + kk;
                if (n < NCOL && j < H && kin < H) v = __ldg(W + (size_t)j * H + kin);
            }
            rb[i] = v;
        }
    };
    auto sstore = [&]() {
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NA; ++i) {
            const int c = tid + NT_ * i;
            const int s = c >> (7 + LCPR), row = (c >> LCPR) & 127, ch = c & (CPR - 1);
            const float4 v = ra[i];
            const float4 h = make_float4(insr_tc::tf32_hi(v.x), insr_tc::tf32_hi(v.y), insr_tc::tf32_hi(v.z), insr_tc::tf32_hi(v.w));
            const int off = s * ATL + a_chunk_off<K16>(row, ch);
            *reinterpret_cast<float4 *>(a_hi + off) = h;
            *reinterpret_cast<float4 *>(a_lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < NB; ++i) {
            const int e = tid + NT_ * i;
            int n, kk;
            if (MODE == 0) { n = e >> (LCPR + 2); kk = e & (KSL - 1); } else { n = e & 127; kk = e >> 7; }
            if (n < NCOL) {
                const float v = w * rb[i];
                const float h = insr_tc::tf32_hi(v);
                *reinterpret_cast<float *>(b_hi + a_chunk_off<K16>(n, kk >> 2) + (kk & 3) * 4) = h;
                *reinterpret_cast<float *>(b_lo + a_chunk_off<K16>(n, kk >> 2) + (kk & 3) * 4) = v - h;
            }
        }
    };

    uint32_t phase = 0;
#ifdef INSR_WIDE_PROFILE
    unsigned long long *gprof = g_wide_prof;            // phase timing of warp 1 (debug builds; insr_debug_wide_prof)
    long long tprev = clock64();
#define WTP(k) do { if (tid == 32 && gprof) { const long long t_ = clock64(); atomicAdd(gprof + (k), (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define WTP(k) do { } while (0)
#endif
    gload(0);
    WTP(0);
    for (int ks = 0; ks < NK; ++ks) {
        sstore();
        WTP(1);
        insr_tc::fence_async_smem();
        insr_tc::tc_fence_before();
        __syncthreads();
        WTP(2);
        if (ks + 1 < NK) gload((ks + 1) * KSL);          // in flight while the tensor core works on slab ks
        if (warp == 0) {
            insr_tc::tc_fence_after();
            if (insr_tc::elect_one()) {
                const uint32_t bh = s32(b_hi), bl = s32(b_lo);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    const uint32_t ah = s32(a_hi + s * ATL), al = s32(a_lo + s * ATL);
                    const uint32_t d = tmem_base + (uint32_t)(s * NCOL);
                    INSR_PRAGMA_UNROLL
                    for (int q = 0; q < KSL / 8; ++q) { // K = 8 per instruction: 32 bytes inside the swizzled row
                        mma_tf32(d, desc_k<K16>(ah + 32 * q), desc_k<K16>(bh + 32 * q), idesc, (ks > 0 || q > 0) ? 1u : 0u);
                        mma_tf32(d, desc_k<K16>(al + 32 * q), desc_k<K16>(bh + 32 * q), idesc, 1u);
                        mma_tf32(d, desc_k<K16>(ah + 32 * q), desc_k<K16>(bl + 32 * q), idesc, 1u);
                    }
                }
                insr_tc::mma_commit(mbar);
            }
            __syncwarp();
        }
        WTP(3);
        insr_tc::mbar_wait(mbar, phase);
        phase ^= 1;
        insr_tc::tc_fence_after();
        WTP(4);
    }

    // ---- epilogue, 32 output columns at a time, staged through shared memory (the operand tiles are free now) so that
    // every global access is a coalesced 128-byte row segment: a thread owns a point ROW of the accumulators, and a
    // row-per-lane store to [stream][point][HP] would touch 32 different lines per instruction (measured: 3x slower).
    //   staging tiles: [tile][128 rows][36 floats] (row stride 144 B = 16 mod 128: conflict-free 16-byte accesses),
    //   tiles 0..S-1 = tape (MODE 0: Zout, written; MODE 1: Ztape, read), tiles S..2S-1 = Aout
    constexpr int EC = NT_ / 16, EST = EC + 4;         // 32 columns / row stride 36 floats (K16: 16 / 20)
    float *stage = reinterpret_cast<float *>(sm);
    auto tile_at = [&](int t, int r, int c) -> float * { return stage + ((size_t)t * TILE + r) * EST + c; };
    const int row = 32 * (warp & 3) + lane;
    const int quarter = warp >> 2;                                // 8 of the 32 columns of a chunk
    const uint32_t trow = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    constexpr int LPR = EC / 4;                                   // lanes per staged row segment (8; K16: 4)
    const int crow = tid >> (LCPR), cch = tid & (LPR - 1);                  // cooperative copies: LPR lanes per row segment, 64 rows per sweep
    constexpr int CR = TILE * LPR / NT_;                          // row sweeps of a cooperative copy (2)
    float4 zpre[MODE == 1 ? S : 1][CR];                           // MODE 1: tape chunk in registers
    auto load_tape_chunk = [&](int jc_) {
        if (MODE == 1) {
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < CR; ++i) {
                    const int r = crow + (NT_ / LPR) * i, j = jc_ + 4 * cch;
                    zpre[MODE == 1 ? s : 0][i] = (j < HP) ? __ldg(reinterpret_cast<const float4 *>(Ztape + ((int64_t)s * NCp + p0 + r) * HP + j))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                }
        }
    };
    load_tape_chunk(j0);
    float oacc[OUT ? CR : 1][OUT ? S : 1][3];
    if (OUT) {
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < CR; ++i)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) { oacc[OUT ? i : 0][OUT ? s : 0][0] = 0.f; oacc[OUT ? i : 0][OUT ? s : 0][1] = 0.f; oacc[OUT ? i : 0][OUT ? s : 0][2] = 0.f; }
    }
    for (int cc = 0; cc < NCOL; cc += EC) {
        const int jc = j0 + cc;                                   // first output column of this chunk
        if (jc >= HP) break;
        if (MODE == 1) {                                          // stage the tape of the previous layer (loaded one chunk ahead)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < CR; ++i)
                    *reinterpret_cast<float4 *>(tile_at(s, crow + (NT_ / LPR) * i, 4 * cch)) = zpre[MODE == 1 ? s : 0][i];
            __syncthreads();
            if (cc + EC < NCOL) load_tape_chunk(jc + EC);         // in flight during this chunk's arithmetic
        }
        {
            const int c0 = 8 * quarter;                           // column inside the chunk
            float acc[S][8], o1[S][8], o2[S][8];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) insr_tc::tmem_ld8(trow + (uint32_t)(s * NCOL + cc + c0), acc[s]);
            insr_tc::tmem_ld_wait();
            if (MODE == 0) {
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < 8; ++i) {
                    float z[S], a[S];
                    const int j = jc + c0 + i;
                    const float bj = (j < H) ? w * __ldg(bias + j) : 0.f;
                    z[0] = acc[0][i] + bj;
                    INSR_PRAGMA_UNROLL
                    for (int s = 1; s < S; ++s) z[s] = acc[s][i];
                    insr_sine_fwd<D, ORDER>(z, a);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) { o1[s][i] = z[s]; o2[s][i] = a[s]; }
                }
                if (Zout) {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) {
                        *reinterpret_cast<float4 *>(tile_at(s, row, c0)) = make_float4(o1[s][0], o1[s][1], o1[s][2], o1[s][3]);
                        *reinterpret_cast<float4 *>(tile_at(s, row, c0 + 4)) = make_float4(o1[s][4], o1[s][5], o1[s][6], o1[s][7]);
                    }
                }
            } else {
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    const float4 z0 = *reinterpret_cast<const float4 *>(tile_at(s, row, c0));
                    const float4 z1 = *reinterpret_cast<const float4 *>(tile_at(s, row, c0 + 4));
                    o1[s][0] = z0.x; o1[s][1] = z0.y; o1[s][2] = z0.z; o1[s][3] = z0.w;
                    o1[s][4] = z1.x; o1[s][5] = z1.y; o1[s][6] = z1.z; o1[s][7] = z1.w;
                }
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < 8; ++i) {
                    float z[S], ab[S], zb[S];
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) { z[s] = o1[s][i]; ab[s] = acc[s][i]; }
                    insr_sine_bwd<D, ORDER>(z, ab, zb);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) o2[s][i] = zb[s];
                }
            }
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                *reinterpret_cast<float4 *>(tile_at(S + s, row, c0)) = make_float4(o2[s][0], o2[s][1], o2[s][2], o2[s][3]);
                *reinterpret_cast<float4 *>(tile_at(S + s, row, c0 + 4)) = make_float4(o2[s][4], o2[s][5], o2[s][6], o2[s][7]);
            }
        }
        __syncthreads();
        float4 wv[OUT ? 3 : 1];
        if (OUT) {
            const int j = jc + 4 * cch;
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < 3; ++o) {
                const float *wr = wo.Wo + (size_t)o * H + j;
                const bool oo = o < wo.O;
                wv[OUT ? o : 0] = make_float4(oo && j + 0 < H ? __ldg(wr + 0) : 0.f, oo && j + 1 < H ? __ldg(wr + 1) : 0.f,
                                              oo && j + 2 < H ? __ldg(wr + 2) : 0.f, oo && j + 3 < H ? __ldg(wr + 3) : 0.f);
            }
        }
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s)
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < CR; ++i) {
                const int r = crow + (NT_ / LPR) * i, j = jc + 4 * cch;
                if (j < HP) {
                    const int64_t g = ((int64_t)s * NCp + p0 + r) * HP + j;
                    if (MODE == 0 && Zout) *reinterpret_cast<float4 *>(Zout + g) = *reinterpret_cast<const float4 *>(tile_at(s, r, 4 * cch));
                    const float4 av = *reinterpret_cast<const float4 *>(tile_at(S + s, r, 4 * cch));
                    if (!OUT || Aout) *reinterpret_cast<float4 *>(Aout + g) = av;
                    if (OUT) {
                        INSR_PRAGMA_UNROLL
                        for (int o = 0; o < 3; ++o) {
                            const float4 wq = wv[OUT ? o : 0];
                            float &t = oacc[OUT ? i : 0][OUT ? s : 0][o];
                            t = fmaf(av.x, wq.x, fmaf(av.y, wq.y, fmaf(av.z, wq.z, fmaf(av.w, wq.w, t))));
                        }
                    }
                }
            }
        __syncthreads();                                          // tiles are rewritten by the next chunk
    }
    if (OUT) {
        typedef StreamCfg<D, ORDER> CO;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < CR; ++i) {
            const int64_t pr = p0 - p_base + crow + (NT_ / LPR) * i;          // row inside the chunk
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < 3; ++o) {
                float out[S];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    float v = oacc[OUT ? i : 0][OUT ? s : 0][o];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (LPR == 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
                    out[s] = v;
                }
                if (cch == 0 && o < wo.O && pr < wo.nv) {
                    const int64_t n = wo.n0 + pr;
                    const int O = wo.O;
                    wo.y[n * O + o] = out[0] + __ldg(wo.bo + o);
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < CO::ND; ++d) wo.jac[(n * O + o) * D + d] = out[1 + d];
                    if (ORDER == 2) wo.h2[n * O + o] = out[S - 1];
                    if (ORDER == 3) {                                       // S <= 4 here means D = 1: one second derivative
                        wo.h2[n * O + o] = out[S - 1];
                    }
                }
            }
        }
    }
    WTP(5);
#ifdef INSR_WIDE_PROFILE
    if (tid == 32 && gprof) atomicAdd(gprof + 7, 1ull);
#endif
    insr_tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// one hidden layer on the tensor cores; rows = number of point rows to process (multiple of 128, within the buffers)
inline bool wide_out_ok(int H) { return make_wgeo(H).passes == 1; }      // the output layer can ride on the last hidden layer

template <int D, int ORDER, int MODE, bool OUT = false>
int launch_wide(const SirenDims &dm, const float *W, const float *bias, const float *Ain, int64_t NCp, int64_t rows,
                const float *Ztape, float *Zout, float *Aout, void *stream, int64_t *launches, WideOut wo = WideOut{}) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    const WGeo g = make_wgeo(dm.H);
    if constexpr (S >= 3) {
        // opt-in (INSR_WIDE_K16=1) until measured: 16-wide K slabs, two CTAs per SM where S * NCOL fits 256 TMEM columns
        static const bool want_k16 = [] { const char *e = getenv("INSR_WIDE_K16"); return e && e[0] == '1'; }();
        if (want_k16 && insr_tc::pow2_cols(S * g.NCOL) <= 256) {
            auto k16 = k_wide_tc<D, ORDER, MODE, OUT, true>;
            const size_t smem16 = smem_bytes(S, g.NCOL, true);
            cudaFuncSetAttribute(k16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16);
            cudaFuncSetAttribute(k16, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            k16<<<dim3((unsigned)(rows / TILE), (unsigned)g.passes), dim3(256), smem16, reinterpret_cast<cudaStream_t>(stream)>>>(
                dm, g.HP, (g.HP + 15) / 16, g.NCOL, insr_tc::pow2_cols(S * g.NCOL), W, bias, Ain, NCp, (int64_t)0, Ztape, Zout, Aout, wo);
            ++*launches;
            return 0;
        }
    }
    auto kfn = k_wide_tc<D, ORDER, MODE, OUT>;
    const size_t smem = smem_bytes(S, g.NCOL);
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (S <= 2)      // room for two CTAs; otherwise leave the split to the driver: the weight loads like a large L1 (measured: 10 %)
        cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kfn<<<dim3((unsigned)(rows / TILE), (unsigned)g.passes), dim3(WT), smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        dm, g.HP, g.NK, g.NCOL, insr_tc::pow2_cols(S * g.NCOL), W, bias, Ain, NCp, (int64_t)0, Ztape, Zout, Aout, wo);
    ++*launches;
    return 0;
}

// =============================================================================================
// weight gradient of a hidden layer of the tiled family (HP <= 128: one 128 x 128 group; wider: one grid row per group):
//     gW[j][k] += omega * sum_{s,p} zbar[s][p][j] * act[s][p][k],      gb[j] += omega * sum_p zbar[0][p][j]
// The reduction runs over POINTS, i.e. over the rows of the [stream][point][HP] buffers: both operands are MN-major,
// which kind::tf32 does not support (siren_tc.cuh), so -- as in the H <= 32 family -- the operands are split in two
// bf16 levels (z = z1 + z2, a = a1 + a2, residual <= 2^-17) and all four cross products are kept:
//     D[jb][kb][128 x 128] += [z1 | z2]^T . [a1 | a2]      (M = 2 levels x 64 neurons, N likewise, K = 16 points)
// for the <= 2 x 2 blocks of 64 x 64 weights; all accumulators (<= 512 TMEM columns) stay resident while the persistent
// CTA streams its share of the points (64 points of one stream per stage: loaded to registers while the previous stage
// is in the tensor core, split, stored as [point][64 bf16] rows in the 128-byte swizzle), and are flushed once:
// quadrants combined in shared memory, one red.global per weight per CTA.
// =============================================================================================
constexpr int WG_PTS = 64;                                   // points per stage
constexpr int WG_ATOM = WG_PTS * 128;                        // one (block, level) operand: 64 rows x 128 B = 8 KB
constexpr uint32_t IDESC_WG128 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint64_t desc_mn128(uint32_t saddr, uint32_t lbo) {          // MN-major, 128-byte swizzle
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(WGT, 1) k_wide_wgrad(SirenDims dm, int HP, int S, const float *__restrict__ ZB,
                                                      const float *__restrict__ Act, int64_t NCp, int nv,
                                                      float *__restrict__ gW, float *__restrict__ gb) {
    extern __shared__ __align__(1024) unsigned char smraw_[];
    unsigned char *sm = smraw_ + ((1024u - (s32(smraw_) & 1023u)) & 1023u);
    // blockIdx.y = 128 x 128 group of the weight matrix (H > 128): output neurons jz0.., input neurons ka0..
    const int G1 = (HP + 127) >> 7;
    const int jz0 = ((int)blockIdx.y / G1) * 128, ka0 = ((int)blockIdx.y % G1) * 128;
    const int HPz = HP - jz0 < 128 ? HP - jz0 : 128, HPa = HP - ka0 < 128 ? HP - ka0 : 128;
    const int NBZ = (HPz + 63) >> 6, NBA = (HPa + 63) >> 6;  // 1 or 2 blocks of 64 neurons on either side
    unsigned char *zt = sm, *at = sm + 4 * WG_ATOM;          // [block][level][64 x 128 B] each
    float *bsumS = reinterpret_cast<float *>(sm + 8 * WG_ATOM);                 // 128 floats
    const uint32_t mbar = s32(sm + 8 * WG_ATOM + 512), tslot = mbar + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = dm.H;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tslot), "r"(512) : "memory");
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

    // staging: per operand 64 rows x 32 chunk slots (16 B of fp32 = 4 neurons); thread -> chunk slot (tid & 31), rows (tid >> 5) + 8 i
    const int ch = tid & 31, r0 = tid >> 5;
    const bool z_ok = 4 * ch < HPz, a_ok = 4 * ch < HPa;
    float4 rz[8], ra[8];
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ngroups = (nv + WG_PTS - 1) / WG_PTS;
    const int64_t nstages = (int64_t)((ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * S;   // this CTA's stages
    auto stage_coords = [&](int64_t st, int &g, int &s) { g = (int)blockIdx.x + (int)(st / S) * (int)gridDim.x; s = (int)(st % S); };
    auto gload = [&](int64_t st) {
        int g, s;
        stage_coords(st, g, s);
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + 8 * i;
            const int64_t pnt = (int64_t)g * WG_PTS + r;
            rz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            ra[i] = rz[i];
            if (pnt < nv) {
                const int64_t off = ((int64_t)s * NCp + pnt) * HP + 4 * ch;
                if (z_ok) rz[i] = __ldg(reinterpret_cast<const float4 *>(ZB + off + jz0));
                if (a_ok) ra[i] = __ldg(reinterpret_cast<const float4 *>(Act + off + ka0));
            }
        }
    };
    auto split4 = [&](const float4 &v, uint2 &l1, uint2 &l2) {
        uint32_t a, b, c, d;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(v.w), "f"(v.z));
        const float r0_ = v.x - __uint_as_float(a << 16), r1_ = v.y - __uint_as_float(a & 0xFFFF0000u);
        const float r2_ = v.z - __uint_as_float(b << 16), r3_ = v.w - __uint_as_float(b & 0xFFFF0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(c) : "f"(r1_), "f"(r0_));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(r3_), "f"(r2_));
        l1 = make_uint2(a, b); l2 = make_uint2(c, d);
    };
    auto sstore = [&](int64_t st) {
        int g, s;
        stage_coords(st, g, s);
        const int jb = ch >> 4, jj = (4 * ch) & 63;              // block, neuron inside the block (multiple of 4)
        const bool zs = 4 * ch < 64 * NBZ, as = 4 * ch < 64 * NBA;
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + 8 * i;
            const int off = r * 128 + ((((jj >> 3) ^ r) & 7) << 4) + (jj & 7) * 2;
            uint2 l1, l2;
            if (zs) {
                split4(rz[i], l1, l2);
                *reinterpret_cast<uint2 *>(zt + (jb * 2 + 0) * WG_ATOM + off) = l1;
                *reinterpret_cast<uint2 *>(zt + (jb * 2 + 1) * WG_ATOM + off) = l2;
                if (s == 0 && ka0 == 0) { bsum.x += rz[i].x; bsum.y += rz[i].y; bsum.z += rz[i].z; bsum.w += rz[i].w; }
            }
            if (as) {
                split4(ra[i], l1, l2);
                *reinterpret_cast<uint2 *>(at + (jb * 2 + 0) * WG_ATOM + off) = l1;
                *reinterpret_cast<uint2 *>(at + (jb * 2 + 1) * WG_ATOM + off) = l2;
            }
        }
    };

    uint32_t phase = 0;
    if (nstages > 0) gload(0);
    for (int64_t st = 0; st < nstages; ++st) {
        sstore(st);
        insr_tc::fence_async_smem();
        insr_tc::tc_fence_before();
        __syncthreads();
        if (st + 1 < nstages) gload(st + 1);
        if (warp == 0) {
            insr_tc::tc_fence_after();
            if (insr_tc::elect_one()) {
                for (int jb = 0; jb < NBZ; ++jb)
                    for (int kb = 0; kb < NBA; ++kb) {
                        const uint32_t d = tmem_base + (uint32_t)((jb * 2 + kb) * 128);
                        const uint32_t za = s32(zt + jb * 2 * WG_ATOM), aa = s32(at + kb * 2 * WG_ATOM);
                        INSR_PRAGMA_UNROLL
                        for (int q = 0; q < 4; ++q) {              // 16 points per instruction = two 1024-byte row groups
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

    // ---- flush: bias gradient, then the weight blocks (quadrants combined in shared memory)
    if (z_ok && ka0 == 0) {
        atomicAdd(bsumS + 4 * ch + 0, bsum.x); atomicAdd(bsumS + 4 * ch + 1, bsum.y);
        atomicAdd(bsumS + 4 * ch + 2, bsum.z); atomicAdd(bsumS + 4 * ch + 3, bsum.w);
    }
    __syncthreads();
    if (tid < 128 && jz0 + tid < H && bsumS[tid] != 0.f) atomicAdd(gb + jz0 + tid, dm.omega * bsumS[tid]);
    if (nstages > 0) {
        // every (neuron j, input k) of a block has four partial sums -- (level of z) x (level of a) -- held by four
        // different warps: each warp parks its 32 x 32 piece in its own padded slab (no atomics, no bank conflicts),
        // then all threads add the four pieces and issue ONE global reduction per element.
        float *piece = reinterpret_cast<float *>(sm);            // [8 warps][32 rows][33] floats (operand region is free)
        const int half = warp >> 2;                              // columns 64 half .. + 63: level of `a`
        const uint32_t trow = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);   // rows: level of z (warp & 2), neurons 32 (warp & 1) ..
        for (int jb = 0; jb < NBZ; ++jb)
            for (int kb = 0; kb < NBA; ++kb)
                for (int cc = 0; cc < 2; ++cc) {                 // 32 input columns at a time
                    INSR_PRAGMA_UNROLL
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float v[8];
                        insr_tc::tmem_ld8(trow + (uint32_t)((jb * 2 + kb) * 128 + 64 * half + 32 * cc + 8 * c8), v);
                        insr_tc::tmem_ld_wait();
                        INSR_PRAGMA_UNROLL
                        for (int i = 0; i < 8; ++i) piece[(warp * 32 + lane) * 33 + 8 * c8 + i] = v[i];
                    }
                    __syncthreads();
                    for (int idx = tid; idx < 64 * 32; idx += WGT) {
                        const int jl = idx >> 5, kl = idx & 31;
                        const int j = jz0 + jb * 64 + jl, k = ka0 + kb * 64 + 32 * cc + kl;
                        if (j < H && k < H) {
                            const int wq = jl >> 5, rr = (jl & 31) * 33 + kl;      // warp & 1, row inside the piece
                            const float v = (piece[(wq + 0) * 32 * 33 + rr] + piece[(wq + 2) * 32 * 33 + rr]) +
                                            (piece[(wq + 4) * 32 * 33 + rr] + piece[(wq + 6) * 32 * 33 + rr]);
                            if (v != 0.f) atomicAdd(gW + (size_t)j * H + k, dm.omega * v);
                        }
                    }
                    __syncthreads();
                }
    }
    insr_tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
    }
}

inline bool wide_wgrad_ok(int H) { return ((H + 7) & ~7) <= 512; }

inline int launch_wide_wgrad(const SirenDims &dm, int S, const float *ZB, const float *Act, int64_t NCp, int nv, float *gW,
                             float *gb, void *stream, int64_t *launches) {
    const int HP = (dm.H + 7) & ~7;
    const size_t smem = (size_t)8 * WG_ATOM + 512 + 64 + 1024;
    cudaFuncSetAttribute(k_wide_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int groups = (nv + WG_PTS - 1) / WG_PTS;
    const int sms = insr_fused::sm_count();
    const int G1 = (HP + 127) >> 7, G = G1 * G1;             // 128 x 128 groups of the weight matrix, one grid row each
    int ctas = sms / G;                                      // one CTA per SM (512 TMEM columns each)
    if (ctas < 1) ctas = 1;
    if (ctas > groups) ctas = groups;
    k_wide_wgrad<<<dim3((unsigned)ctas, (unsigned)G), dim3(WGT), smem, reinterpret_cast<cudaStream_t>(stream)>>>(dm, HP, S, ZB, Act, NCp, nv, gW, gb);
    ++*launches;
    return 0;
}

}  // namespace insr_wide
#endif  // !INSR_CPU_EMU
