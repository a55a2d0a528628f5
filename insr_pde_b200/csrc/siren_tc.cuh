// siren_tc.cuh -- tcgen05 / TMEM kernels for the resident-weights family (H <= 32).
//
// Why: ncu shows the FFMA kernels limited by shared-memory delivery (LDS wavefronts), not by the
// FP32 pipe; the 5th-gen tensor cores take their operands from shared memory descriptors or from TMEM.
// Single-pass TF32 misses the 1e-4 parity target, so every hidden-layer contraction of
// the forward pass and of the data gradient is a 3xTF32 split:  a = a_hi + a_lo (a_hi = fp32 with the 13
// low mantissa bits cleared, a_lo = a - a_hi exact),  a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi
// accumulated in FP32 in TMEM (dropped term ~2^-22 relative).
//
// Mapping.  CTA tile = 128 collocation points = the M dimension of tcgen05.mma (one TMEM lane per
// point).  Each forward-mode stream s is its own MMA chain  D_s[128 x 32] = A_s[128 x 32] . W^T
// into TMEM columns [32 s, 32 s + 32), so a thread (= one point) reads ALL streams of its neurons
// with tcgen05.ld and runs the sine-stream algebra in registers.  256 threads: warps w and w+4 share
// TMEM lanes 32 (w%4) .. +31 and split the 32 neurons in two halves.
//
// Measured on the B200 (tools/probe/*.cu, profiles/r1_tcgen05_probes.txt) and designed around:
//   * an MMA whose A operand comes from shared memory costs 32 + N/4 cycles (operand fetch at 128 B/clk:
//     40 cycles at N = 32), the same MMA with A in TMEM costs N/2 = 16 cycles.  So the hi part of every
//     activation operand is written back to TMEM with tcgen05.st (columns [32 S + 32 s, ..)) and used by
//     two MMAs (x W_hi, x W_lo); only the lo part goes through shared memory (UMMA K-major canonical
//     layout: 8 x 16-byte core matrices, LBO = 128 B along K, SBO = 1024 B along M/N).
//   * kind::tf32 ignores MN-major ("transposed") operands: the MMA contributes exactly zero.  The data
//     gradient therefore uses a second, explicitly transposed copy of the weights.  kind::f16 (bf16) does
//     honour MN-major operands, so the weight gradient -- a reduction over POINTS, i.e. over the rows of
//     the natural [point][neuron] layout -- runs in bf16 with both operands split in two bf16 levels
//     (z = z1 + z2, a = a1 + a2, |error| <= 2^-17 per operand, all four cross products kept):
//         D[64 x 64] += [z1 | z2]^T . [a1 | a2]      (M = 64 = 2 levels x 32 neurons, N likewise, K = points)
//     read MN-major from rows of 128 bytes in the 128-byte swizzle; the accumulators of all layers stay in
//     TMEM for the whole launch.
//   * tcgen05.mma must be issued from a warp-uniform branch by ONE elected lane (elect.sync): a
//     `tid == 0` branch makes nvcc wrap every MMA in a divergence "waterfall" loop (~100 cycles each).
//
// Only compiled by nvcc (inline PTX); the host-side SIMT emulation does not cover this file.
#pragma once
#include <cstdlib>
#include "siren_fused.cuh"

#ifndef INSR_CPU_EMU
namespace insr_tc {

using insr_fused::Params;
constexpr int HP = 32;
constexpr int TILE_M = 128;
constexpr int THREADS = 256;
constexpr int OP_BYTES = TILE_M * HP * 4;          // one 128 x 32 fp32 operand: 16 KB
constexpr int W_BYTES = HP * HP * 4;               // one 32 x 32 weight operand: 4 KB
// kind::tf32, D = F32, A/B = TF32, both K-major, N = 32, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
// kind::f16, D = F32, A/B = BF16, both MN-major, N = 64, M = 64
constexpr uint32_t IDESC_WG = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);

__device__ __forceinline__ uint32_t s32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// UMMA shared-memory descriptor: K-major, no swizzle, LBO = 128 B, SBO = 1024 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// MN-major, 128-byte swizzle: rows of 128 B (64 bf16 along M/N), 8 rows (K) per 1024-byte atom, SBO = 1024 B
// between 8-row groups along K (probe2: element (n, k) at (k>>3)*1024 + (k&7)*128 + (((n>>3) ^ (k&7)) << 4) + (n&7)*2)
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// byte offset of element (row, k) inside a K-major canonical operand
__device__ __forceinline__ int op_off(int row, int k) { return (row >> 3) * 1024 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_wg(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC_WG), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_LOOP;\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                    "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                    "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

// 8 consecutive neurons (k = 8 g .. 8 g + 7; g = 0..3) of one stream of row `row`: the hi part goes to TMEM
// (8 columns at a_tmem), the lo part to the K-major shared-memory operand
__device__ __forceinline__ void store_split8(uint32_t a_tmem, unsigned char *op_lo, int row, int g, const float (&a)[8]) {
    float h[8], l[8];
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 8; ++c) { h[c] = tf32_hi(a[c]); l[c] = a[c] - h[c]; }
    tmem_st8(a_tmem, h);
    const int off = (row >> 3) * 1024 + (2 * g) * 128 + (row & 7) * 16;
    *reinterpret_cast<float4 *>(op_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<float4 *>(op_lo + off + 128) = make_float4(l[4], l[5], l[6], l[7]);
}

// the 12 MMAs of one stream of one hidden-layer contraction: D (32 columns) = A . B^T with A_hi in TMEM
// (32 columns at a_tmem), A_lo and B_hi / B_lo K-major in shared memory
__device__ __forceinline__ void issue_stream(uint32_t d_tmem, uint32_t a_tmem, uint32_t alo, uint32_t bhi, uint32_t blo) {
    INSR_PRAGMA_UNROLL
    for (int ks = 0; ks < 4; ++ks) {                  // K = 8 per instruction: 8 TMEM columns / two 16-byte core matrices
        mma_tf32_ts(d_tmem, a_tmem + 8 * ks, umma_desc(bhi + 256 * ks), ks > 0);
        mma_tf32_ts(d_tmem, a_tmem + 8 * ks, umma_desc(blo + 256 * ks), 1);
        mma_tf32_ss(d_tmem, umma_desc(alo + 256 * ks), umma_desc(bhi + 256 * ks), 1);
    }
}

__host__ __device__ inline int pow2_cols(int c) {
    int r = 32;
    while (r < c) r <<= 1;
    return r;
}

// Staging.  At the scripts' batch sizes (one wave of tiles) the prologue is a visible part of a launch: a loop of
// "load theta -> dependent shared-memory stores" pays one L2 (first touch: HBM) round trip per trip.  So every staging loop
// issues its loads together, through the read-only path, before the first store (theta is never written by these kernels).
constexpr int STAGE_BATCH = 8;

// stage the small operands shared by the forward and the backward kernel (NTH threads; single trip for L <= NTH / 32)
template <int NTH>
__device__ __forceinline__ void stage_small_t(const Params &p, float *biasS, float *w1S, float *woS, float *boS) {
    const SirenDims &dm = p.dm;
    const int L = dm.L, H = dm.H, D = dm.D, O = dm.O, tid = threadIdx.x;
    const float w = dm.omega;
    // first trip of every vector: loads, then stores
    float vb = 0.f, v1 = 0.f, vo = 0.f, vbo = 0.f;
    {
        if (tid < L * HP) { const int l = tid / HP, j = tid % HP; if (j < H) vb = __ldg(p.theta + insr_b_offset(dm, l + 1) + j); }
        if (tid < HP * 4) {
            const int j = tid >> 2, d = tid & 3;
            if (j < H) {
                if (d < D) v1 = __ldg(p.theta + insr_w_offset(dm, 0) + (int64_t)j * D + d);
                else if (d == 3) v1 = __ldg(p.theta + insr_b_offset(dm, 0) + j);
            }
        }
        if (tid < 3 * HP) { const int o = tid / HP, j = tid % HP; if (o < O && j < H) vo = __ldg(p.theta + insr_w_offset(dm, L + 1) + (int64_t)o * H + j); }
        if (tid < 4 && tid < O) vbo = __ldg(p.theta + insr_b_offset(dm, L + 1) + tid);
    }
    if (tid < L * HP) biasS[tid] = w * vb;
    if (tid < HP * 4) w1S[tid] = w * v1;
    if (tid < 3 * HP) woS[tid] = vo;
    if (tid < 4) boS[tid] = vbo;
    for (int idx = tid + NTH; idx < L * HP; idx += NTH) {          // deep nets with few threads only
        const int l = idx / HP, j = idx % HP;
        biasS[idx] = (j < H) ? w * __ldg(p.theta + insr_b_offset(dm, l + 1) + j) : 0.f;
    }
    static_assert(NTH >= HP * 4 && NTH >= 3 * HP, "w1 / wo are staged in one trip");
}
__device__ __forceinline__ void stage_small(const Params &p, float *biasS, float *w1S, float *woS, float *boS) {
    stage_small_t<THREADS>(p, biasS, w1S, woS, boS);
}
// hidden-layer weights split hi / lo in the UMMA K-major layout (omega folded in); the transposed copy (wthi / wtlo, may be
// NULL) is the B operand of the data gradient (row = input neuron k, reduction index = output neuron j)
template <int NTH>
__device__ __forceinline__ void stage_hidden_t(const Params &p, unsigned char *whi, unsigned char *wlo, unsigned char *wthi,
                                               unsigned char *wtlo) {
    const SirenDims &dm = p.dm;
    const int L = dm.L, H = dm.H, total = L * HP * HP, tid = threadIdx.x;
    const float w = dm.omega;
    for (int base = 0; base < total; base += STAGE_BATCH * NTH) {
        float vv[STAGE_BATCH];
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < STAGE_BATCH; ++i) {
            const int idx = base + i * NTH + tid;
            vv[i] = 0.f;
            if (idx < total) {
                const int l = idx / (HP * HP), j = (idx / HP) % HP, k = idx % HP;
                if (j < H && k < H) vv[i] = __ldg(p.theta + insr_w_offset(dm, l + 1) + (int64_t)j * H + k);
            }
        }
        INSR_PRAGMA_UNROLL
        for (int i = 0; i < STAGE_BATCH; ++i) {
            const int idx = base + i * NTH + tid;
            if (idx < total) {
                const int l = idx / (HP * HP), j = (idx / HP) % HP, k = idx % HP;
                const float v = w * vv[i];
                const float hi = tf32_hi(v);
                if (whi) {
                    *reinterpret_cast<float *>(whi + l * W_BYTES + op_off(j, k)) = hi;
                    *reinterpret_cast<float *>(wlo + l * W_BYTES + op_off(j, k)) = v - hi;
                }
                if (wthi) {
                    *reinterpret_cast<float *>(wthi + l * W_BYTES + op_off(k, j)) = hi;
                    *reinterpret_cast<float *>(wtlo + l * W_BYTES + op_off(k, j)) = v - hi;
                }
            }
        }
    }
}
__device__ __forceinline__ void stage_hidden(const Params &p, unsigned char *whi, unsigned char *wlo, bool transposed) {
    if (transposed) stage_hidden_t<THREADS>(p, nullptr, nullptr, whi, wlo);
    else stage_hidden_t<THREADS>(p, whi, wlo, nullptr, nullptr);
}

// =============================================================================================
// forward kernel: 2 CTAs per SM (one's MMAs overlap the other's epilogue), 256 TMEM columns each
// =============================================================================================
struct Smem {
    int w_hi, w_lo, a_lo, bias, w1, wo, bo, part, mbar, tmem, total;
};
__host__ __device__ inline Smem smem_map(int L, int S) {
    Smem m;
    int o = 0;
    m.w_hi = o; o += L * W_BYTES;
    m.w_lo = o; o += L * W_BYTES;
    m.a_lo = o; o += S * OP_BYTES;
    m.bias = o; o += L * HP * 4;
    m.w1 = o; o += HP * 16;
    m.wo = o; o += 3 * HP * 4;
    m.bo = o; o += 16;
    m.part = o; o += TILE_M * 16 * 4;          // output-layer partials of the upper neuron half: [128][<=16]
    m.mbar = o; o += 16;
    m.tmem = o; o += 16;
    m.total = o;
    return m;
}

template <int D, int O, int ORDER>
__global__ void __launch_bounds__(THREADS, 2) k_tc_fwd(Params p, int tmem_cols) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int ABASE = 32 * S;                      // TMEM columns of the hi operand
    extern __shared__ __align__(1024) unsigned char smraw[];
    const SirenDims dm = p.dm;
    const int L = dm.L;
    const Smem M = smem_map(L, S);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;            // TMEM lane == point inside the tile
    const int half = warp >> 2;                        // neurons 16 half .. 16 half + 15
    float *biasS = reinterpret_cast<float *>(smraw + M.bias);
    float *w1S = reinterpret_cast<float *>(smraw + M.w1);
    float *woS = reinterpret_cast<float *>(smraw + M.wo);
    float *boS = reinterpret_cast<float *>(smraw + M.bo);
    float *partS = reinterpret_cast<float *>(smraw + M.part);
    const uint32_t mbar = s32(smraw + M.mbar);

    stage_hidden(p, smraw + M.w_hi, smraw + M.w_lo, false);
    stage_small(p, biasS, w1S, woS, boS);

    // ---- TMEM allocation (warp 0) + mbarrier init
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s32(smraw + M.tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smraw + M.tmem);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t phase = 0;

    const int64_t ntiles = (p.N + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n = tile * TILE_M + row;
        const bool valid = n < p.N;
        float xv[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) xv[d] = valid ? __ldg(p.x + n * D + d) : 0.f;

        // ---- first sine layer (FFMA): 2 groups of 8 neurons
        INSR_PRAGMA_UNROLL
        for (int g8 = 0; g8 < 2; ++g8) {
            float a8[S][8];
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < 2; ++q) {
                float z[S][4], a[S][4], tv[S + 1][4];
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const int j = 16 * half + 8 * g8 + 4 * q + c;
                    const float4 wv = *reinterpret_cast<const float4 *>(w1S + j * 4);
                    float acc = wv.w;
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) acc = fmaf(insr_fused::f4get(wv, d), xv[d], acc);
                    z[0][c] = acc;
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < C::ND; ++d) z[1 + d][c] = insr_fused::f4get(wv, d);
                    if constexpr (ORDER == 2) z[1 + C::ND][c] = 0.f;
                }
                insr_fused::act4<D, ORDER>(z, a, tv);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) a8[s][4 * q + c] = a[s][c];
            }
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                store_split8(tmem_row + ABASE + 32 * s + 16 * half + 8 * g8, smraw + M.a_lo + s * OP_BYTES, row, 2 * half + g8, a8[s]);
        }
        float out[O][S];
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) out[o][s] = 0.f;

        // ---- hidden layers on the tensor cores
        for (int l = 0; l < L; ++l) {
            tmem_st_wait();
            fence_async_smem();                          // generic-proxy operand writes -> visible to the async proxy
            tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t whi = s32(smraw + M.w_hi + l * W_BYTES), wlo = s32(smraw + M.w_lo + l * W_BYTES);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        issue_stream(tmem_base + 32 * s, tmem_base + ABASE + 32 * s, s32(smraw + M.a_lo + s * OP_BYTES), whi, wlo);
                    mma_commit(mbar);
                }
                __syncwarp();
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            tc_fence_after();
            const bool last = (l == L - 1);
            // ---- epilogue: 2 groups of 8 neurons of this thread's half
            INSR_PRAGMA_UNROLL
            for (int g8 = 0; g8 < 2; ++g8) {
                float zz[S][8], a8[S][8];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) tmem_ld8(tmem_row + 32 * s + 16 * half + 8 * g8, zz[s]);
                tmem_ld_wait();
                INSR_PRAGMA_UNROLL
                for (int q = 0; q < 2; ++q) {
                    float z[S][4], a[S][4], tv[S + 1][4];
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) {
                        const int j = 16 * half + 8 * g8 + 4 * q + c;
                        z[0][c] = zz[0][4 * q + c] + biasS[l * HP + j];
                        INSR_PRAGMA_UNROLL
                        for (int s = 1; s < S; ++s) z[s][c] = zz[s][4 * q + c];
                    }
                    insr_fused::act4<D, ORDER>(z, a, tv);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) a8[s][4 * q + c] = a[s][c];
                }
                if (last) {                              // output layer (FFMA) folded into the last epilogue
                    INSR_PRAGMA_UNROLL
                    for (int o = 0; o < O; ++o)
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s)
                            INSR_PRAGMA_UNROLL
                            for (int i = 0; i < 8; ++i) out[o][s] = fmaf(woS[o * HP + 16 * half + 8 * g8 + i], a8[s][i], out[o][s]);
                } else {
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        store_split8(tmem_row + ABASE + 32 * s + 16 * half + 8 * g8, smraw + M.a_lo + s * OP_BYTES, row, 2 * half + g8, a8[s]);
                }
            }
        }
        // ---- combine the two neuron halves
        if (half == 1) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) partS[row * 16 + o * S + s] = out[o][s];
        }
        tc_fence_before();
        __syncthreads();
        if (half == 0 && valid) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) out[o][s] += partS[row * 16 + o * S + s];
                out[o][0] += boS[o];
                insr_store_outputs<D, O, ORDER>(n, o, out[o], p.y, p.jac, p.h2);
            }
        }
        __syncthreads();                                 // partS is rewritten by the next tile
        tc_fence_after();
    }
    // ---- TMEM release
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

template <int D, int O, int ORDER>
int launch_tc_fwd(Params &p, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    static_assert(O * S <= 16, "output partial buffer holds 16 values per point");
    const Smem M = smem_map(p.dm.L, S);
    auto kfn = k_tc_fwd<D, O, ORDER>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total);
    cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int64_t tiles = (p.N + TILE_M - 1) / TILE_M;
    const int64_t slots = 2 * (int64_t)insr_fused::sm_count();
    const int64_t ctas = tiles < slots ? tiles : slots;
    kfn<<<dim3((unsigned)ctas), dim3(THREADS), M.total, reinterpret_cast<cudaStream_t>(stream)>>>(p, pow2_cols(64 * S));
    ++*launches;
    return 0;
}


// =============================================================================================
// backward / fused-closure kernel on the tensor cores (one CTA of 512 threads per SM, 512 TMEM columns)
//   forward recompute   D_s = A_s W^T        3xTF32, A_hi in TMEM, W K-major in shared memory
//   data gradient       D_s = Zbar_s W       the same, against the transposed weight copy
//   weight gradient     Wacc_l[64 x 64] += [z1|z2]^T [a1|a2]   bf16 x 2 levels, both operands MN-major;
//                       accumulators stay in TMEM for the whole launch and are read out once at the end
// Four threads per point (8 neurons each): 16 warps hide the latency of the tape (L2), TMEM and MUFU round trips.
// TMEM columns: [0, 32 S) accumulators | [32 S, 64 S) hi operand | 32 NLO_T lo operand of the first NLO_T streams
// (the other streams keep their lo operand in shared memory) | 64 L weight-gradient accumulators.
// Shared memory: one weight-gradient operand slot per stream (ZT | AT, 32 KB), so that a reverse layer needs a
// single CTA barrier: operands -> barrier -> data-gradient MMAs, then weight-gradient MMAs (which overlap the next
// layer's adjoint).  The tape (sin, cos, t_d, t_q per activation) lives in a per-CTA global scratch that stays
// L2-resident.  Thin layers (first / output layer, biases, d loss/d x) are reduced over the 32 points of a warp
// with a halving butterfly into a few persistent registers.
// =============================================================================================
constexpr int BT = 512;                             // threads of the backward kernel
constexpr int NT = BT / TILE_M;                     // threads per point
constexpr int NPT = HP / NT;                        // neurons per thread (8)
constexpr int LMAX_TC = 3;                          // hidden layers (TMEM budget of the weight-gradient accumulators)
constexpr int SLOT_BYTES = 2 * TILE_M * 128;        // ZT (128 points x 128 B) | AT (128 points x 128 B)
__host__ __device__ constexpr int nlo_tmem(int S) { return (512 - 64 * S - 64 * LMAX_TC) / 32 < S ? (512 - 64 * S - 64 * LMAX_TC) / 32 : S; }

__device__ __forceinline__ float reduce8(float (&v)[8], int lane) {
    // after the call the lane holds the warp-wide sum of value index (lane >> 2) & 7
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 16;
        const float send = up ? v[i] : v[i + 4];
        const float keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 8;
        const float send = up ? v[i] : v[i + 2];
        const float keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const bool up = lane & 4;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct SmemB {
    int slots, x_lo, w_hi, w_lo, wt_hi, wt_lo, bias, w1, wo, bo, part, mbar, tmem, total;
};
__host__ __device__ inline SmemB smem_map_bwd(int L, int S, int PS) {
    SmemB m;
    int o = 0;
    m.slots = o; o += S * SLOT_BYTES;              // first: 1024-byte aligned (128-byte swizzle atoms)
    m.x_lo = o; o += (S - nlo_tmem(S)) * OP_BYTES;
    m.w_hi = o; o += L * W_BYTES;
    m.w_lo = o; o += L * W_BYTES;
    m.wt_hi = o; o += L * W_BYTES;
    m.wt_lo = o; o += L * W_BYTES;
    m.bias = o; o += L * HP * 4;
    m.w1 = o; o += HP * 16;
    m.wo = o; o += 3 * HP * 4;
    m.bo = o; o += 16;
    m.part = o; o += NT * TILE_M * PS * 4;         // per-thread partial outputs / gx: [NT][128][PS]
    m.mbar = o; o += 32;
    m.tmem = o; o += 16;
    m.total = o;
    return m;
}
// tape: per CTA  [(L+1)][2 (S+1)][512] float4
__host__ __device__ inline size_t tape_float4_per_cta(int L, int S) { return (size_t)(L + 1) * 2 * (S + 1) * BT; }

// two bf16 levels of 8 values -> 2 x 4 packed words
__device__ __forceinline__ void split_bf16x2(const float (&v)[8], uint32_t (&l1)[4], uint32_t (&l2)[4]) {
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 4; ++i) {
        uint32_t w1, w2;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        const float r0 = v[2 * i] - __uint_as_float(w1 << 16);
        const float r1 = v[2 * i + 1] - __uint_as_float(w1 & 0xFFFF0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w2) : "f"(r1), "f"(r0));
        l1[i] = w1; l2[i] = w2;
    }
}
// write this thread's 8 neurons (two levels) of one stream into a [point][64 bf16] operand with 128-byte swizzle:
// chunk `part` holds level 1, chunk 4 + part level 2
__device__ __forceinline__ void store_wg_operand(unsigned char *buf, int row, int part, const float (&v)[8]) {
    uint32_t l1[4], l2[4];
    split_bf16x2(v, l1, l2);
    unsigned char *r = buf + row * 128;
    const int key = row & 7;
    *reinterpret_cast<uint4 *>(r + ((part ^ key) << 4)) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
    *reinterpret_cast<uint4 *>(r + (((4 + part) ^ key) << 4)) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
}
// hi part -> TMEM; lo part -> TMEM (lo_tmem != 0) or the K-major shared-memory operand
template <bool LO_TMEM>
__device__ __forceinline__ void store_split8b(uint32_t hi_tmem, uint32_t lo_tmem, unsigned char *op_lo, int row, int g, const float (&a)[8]) {
    float h[8], l[8];
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 8; ++c) { h[c] = tf32_hi(a[c]); l[c] = a[c] - h[c]; }
    tmem_st8(hi_tmem, h);
    if constexpr (LO_TMEM) {
        tmem_st8(lo_tmem, l);
    } else {
        const int off = (row >> 3) * 1024 + (2 * g) * 128 + (row & 7) * 16;
        *reinterpret_cast<float4 *>(op_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<float4 *>(op_lo + off + 128) = make_float4(l[4], l[5], l[6], l[7]);
    }
}
template <bool LO_TMEM>
__device__ __forceinline__ void issue_stream_b(uint32_t d_tmem, uint32_t hi_tmem, uint32_t lo_tmem, uint32_t alo, uint32_t bhi, uint32_t blo) {
    INSR_PRAGMA_UNROLL
    for (int ks = 0; ks < 4; ++ks) {
        mma_tf32_ts(d_tmem, hi_tmem + 8 * ks, umma_desc(bhi + 256 * ks), ks > 0);
        mma_tf32_ts(d_tmem, hi_tmem + 8 * ks, umma_desc(blo + 256 * ks), 1);
        if constexpr (LO_TMEM) mma_tf32_ts(d_tmem, lo_tmem + 8 * ks, umma_desc(bhi + 256 * ks), 1);
        else mma_tf32_ss(d_tmem, umma_desc(alo + 256 * ks), umma_desc(bhi + 256 * ks), 1);
    }
}

// phase timing of warp 0 (debug builds with -DINSR_TC_PROFILE): cycles per phase accumulated into 16 counters that
// sit right behind the tape scratch (tools/tc_phase_profile.py)
#ifdef INSR_TC_PROFILE
#define TCP(k) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(gprof + (k), (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define TCP(k) do { } while (0)
#endif

template <int D, int O, int ORDER, bool LSQ>
__global__ void __launch_bounds__(BT, 1) k_tc_bwd(Params p, float4 *__restrict__ tape_all) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int TV = S + 1;
    constexpr int NQ = 2 * TV;                       // float4 tape slots per thread per layer
    constexpr int NLO = nlo_tmem(S);
    constexpr int ABASE = 32 * S, LBASE = 64 * S, WBASE = 64 * S + 32 * NLO;
    constexpr int PS = (O * S > D ? O * S : D);      // partial-buffer stride (floats)
    extern __shared__ __align__(1024) unsigned char smraw_[];
    unsigned char *smraw = smraw_ + ((1024u - (s32(smraw_) & 1023u)) & 1023u);     // 128-byte swizzle atoms: 1024-byte aligned
    const SirenDims dm = p.dm;
    const int L = dm.L, H = dm.H;
    const SmemB M = smem_map_bwd(L, S, PS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;          // TMEM lane == point inside the tile
    const int part = warp >> 2;                      // neurons 8 part .. 8 part + 7
    float *biasS = reinterpret_cast<float *>(smraw + M.bias);
    float *w1S = reinterpret_cast<float *>(smraw + M.w1);
    float *woS = reinterpret_cast<float *>(smraw + M.wo);
    float *boS = reinterpret_cast<float *>(smraw + M.bo);
    float *partS = reinterpret_cast<float *>(smraw + M.part);
    const uint32_t mbarD = s32(smraw + M.mbar), mbarW = mbarD + 8;
    float4 *tape = tape_all + (size_t)blockIdx.x * tape_float4_per_cta(L, S);
#ifdef INSR_TC_PROFILE
    unsigned long long *gprof = reinterpret_cast<unsigned long long *>(tape_all + (size_t)gridDim.x * tape_float4_per_cta(L, S));
    long long tprev = 0;
#endif

    // ---- stage the weights (BT threads): loads first, stores after (see stage_hidden_t)
    stage_hidden_t<BT>(p, smraw + M.w_hi, smraw + M.w_lo, smraw + M.wt_hi, smraw + M.wt_lo);
    stage_small_t<BT>(p, biasS, w1S, woS, boS);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s32(smraw + M.tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(mbarD, 1); mbar_init(mbarW, S);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smraw + M.tmem);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t phD = 0, phW = 0;
    bool pendW = false;
    uint32_t wacc_mask = 0;                                       // bit l-1: Wacc_l holds data
    {   // zero the weight-gradient accumulators: every weight-gradient MMA accumulates, whichever warp issues it first
        float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = 8 * part; c < 64 * L; c += 8 * NT) tmem_st8(tmem_row + WBASE + c, zero8);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    // persistent thin-layer partial sums: the lane holds value (lane >> 2) & 7 of this thread's 8 neurons
    float acc_gwo[O], acc_g1[1 + D], acc_gb[LMAX_TC] = {0.f, 0.f, 0.f};
    float acc_gbo = 0.f, loss_acc = 0.f;
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o) acc_gwo[o] = 0.f;
    INSR_PRAGMA_UNROLL
    for (int d = 0; d <= D; ++d) acc_g1[d] = 0.f;

    auto publish_and_sync = [&]() {          // operand writes (generic proxy / tcgen05.st) -> MMA, then CTA barrier
        tmem_st_wait();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    // store the operand (all streams) of the next contraction: this thread's 8 neurons
    auto store_operand = [&](const float (&a)[S][8]) {
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) {
            if (s < NLO)
                store_split8b<true>(tmem_row + ABASE + 32 * s + 8 * part, tmem_row + LBASE + 32 * (s < NLO ? s : 0) + 8 * part, nullptr, row, part, a[s]);
            else
                store_split8b<false>(tmem_row + ABASE + 32 * s + 8 * part, 0, smraw + M.x_lo + (s >= NLO ? s - NLO : 0) * OP_BYTES, row, part, a[s]);
        }
    };
    auto issue_contraction = [&](uint32_t bhi, uint32_t blo) {
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) {
            if (s < NLO)
                issue_stream_b<true>(tmem_base + 32 * s, tmem_base + ABASE + 32 * s, tmem_base + LBASE + 32 * (s < NLO ? s : 0), 0, bhi, blo);
            else
                issue_stream_b<false>(tmem_base + 32 * s, tmem_base + ABASE + 32 * s, 0, s32(smraw + M.x_lo + (s >= NLO ? s - NLO : 0) * OP_BYTES), bhi, blo);
        }
    };

    const int64_t ntiles = (p.N + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n = tile * TILE_M + row;
        const bool valid = n < p.N;
        float xv[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) xv[d] = valid ? __ldg(p.x + n * D + d) : 0.f;
#ifdef INSR_TC_PROFILE
        if (tid == 0) { tprev = clock64(); atomicAdd(gprof + 15, 1ull); }
#endif
        // HBM latency of the per-point streams is taken off the critical path: this tile's cotangents / targets (needed
        // after the forward recompute) and the next tile's points are pulled into L2 now (no registers held)
        if (part == 0) {
            auto pf = [](const void *q) { asm volatile("prefetch.global.L2 [%0];" :: "l"(q)); };
            if (valid) {
                if constexpr (LSQ) {
                    if (p.target) pf(p.target + n * p.n_res);
                } else {
                    if (p.gy) pf(p.gy + n * O);
                    if (ORDER >= 1 && p.gjac) pf(p.gjac + n * O * D);
                    if (ORDER >= 2 && p.gh2) pf(p.gh2 + n * O);
                }
            }
            const int64_t nn = n + (int64_t)gridDim.x * TILE_M;
            if (nn < p.N) pf(p.x + nn * D);
        }

        // ================= forward with tape =================
        // Tape traffic (L1 <-> L2) is the largest single cost of this kernel (ablation: 1.8 of 5.0 ms), so the first sine
        // layer is not taped: it is recomputed from x where it is needed (8 sincos per thread per use).
        auto layer0 = [&](int q, float (&a)[S][4], float (&tv)[TV][4]) {      // neurons 8 part + 4 q .. + 3
            float z[S][4];
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < 4; ++c) {
                const int j = NPT * part + 4 * q + c;
                const float4 wv = *reinterpret_cast<const float4 *>(w1S + j * 4);
                float acc = wv.w;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) acc = fmaf(insr_fused::f4get(wv, d), xv[d], acc);
                z[0][c] = acc;
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < C::ND; ++d) z[1 + d][c] = insr_fused::f4get(wv, d);
                if constexpr (ORDER == 2) z[1 + C::ND][c] = 0.f;
            }
            insr_fused::act4<D, ORDER>(z, a, tv);
        };
        {
            float a8[S][8];
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < 2; ++q) {
                float a[S][4], tv[TV][4];
                layer0(q, a, tv);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) a8[s][4 * q + c] = a[s][c];
            }
            store_operand(a8);
        }
        TCP(0);
        float out[O][S];                                   // LSQ: this thread's share of the outputs
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) out[o][s] = 0.f;
        for (int l = 0; l < L; ++l) {
            publish_and_sync();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    issue_contraction(s32(smraw + M.w_hi + l * W_BYTES), s32(smraw + M.w_lo + l * W_BYTES));
                    mma_commit(mbarD);
                }
                __syncwarp();
            }
            TCP(1);
            mbar_wait(mbarD, phD);
            phD ^= 1;
            tc_fence_after();
            TCP(2);
            const bool last = (l == L - 1);
            float4 *tl = tape + (size_t)(l + 1) * NQ * BT;
            float zz[S][8], a8[S][8];
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) tmem_ld8(tmem_row + 32 * s + 8 * part, zz[s]);
            tmem_ld_wait();
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < 2; ++q) {
                float z[S][4], a[S][4], tv[TV][4];
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const int j = NPT * part + 4 * q + c;
                    z[0][c] = zz[0][4 * q + c] + biasS[l * HP + j];
                    INSR_PRAGMA_UNROLL
                    for (int s = 1; s < S; ++s) z[s][c] = zz[s][4 * q + c];
                }
                insr_fused::act4<D, ORDER>(z, a, tv);
                INSR_PRAGMA_UNROLL
                for (int t = 0; t < TV; ++t) tl[(size_t)(q * TV + t) * BT + tid] = make_float4(tv[t][0], tv[t][1], tv[t][2], tv[t][3]);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) a8[s][4 * q + c] = a[s][c];
            }
            if (last) {
                if constexpr (LSQ) {
                    INSR_PRAGMA_UNROLL
                    for (int o = 0; o < O; ++o)
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s)
                            INSR_PRAGMA_UNROLL
                            for (int i = 0; i < 8; ++i) out[o][s] = fmaf(woS[o * HP + NPT * part + i], a8[s][i], out[o][s]);
                }
            } else {
                store_operand(a8);
            }
            TCP(3);
        }
        // ================= output layer: cotangents g[o][s] =================
        float g[O][S];
        if constexpr (LSQ) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) partS[(part * TILE_M + row) * PS + o * S + s] = out[o][s];
            __syncthreads();
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    float acc = 0.f;
                    INSR_PRAGMA_UNROLL
                    for (int q = 0; q < NT; ++q) acc += partS[(q * TILE_M + row) * PS + o * S + s];
                    out[o][s] = acc;
                }
                out[o][0] += boS[o];
            }
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
                if (part == 0) loss_acc = fmaf(r, r, loss_acc);
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
        // output-layer gradients (this thread's 8 neurons; a_L recomputed from the tape), the cotangent of the last sine
        // layer and -- from the SAME tape values, one L2 round trip instead of two -- its activation adjoint zbar_L
        float ab[S][8];
        {
            const float4 *tl = tape + (size_t)L * NQ * BT;
            float v[O][8];
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < 2; ++q) {
                float tv[TV][4], a[S][4], abq[S][4];
                INSR_PRAGMA_UNROLL
                for (int t = 0; t < TV; ++t) {
                    const float4 u = tl[(size_t)(q * TV + t) * BT + tid];
                    tv[t][0] = u.x; tv[t][1] = u.y; tv[t][2] = u.z; tv[t][3] = u.w;
                }
                insr_fused::a_from_tape4<D, ORDER>(tv, a);
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) {
                        float acc = 0.f;
                        INSR_PRAGMA_UNROLL
                        for (int s = 0; s < S; ++s) acc = fmaf(g[o][s], a[s][c], acc);
                        v[o][4 * q + c] = acc;
                    }
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) {
                        float acc = 0.f;
                        INSR_PRAGMA_UNROLL
                        for (int o = 0; o < O; ++o) acc = fmaf(woS[o * HP + NPT * part + 4 * q + c], g[o][s], acc);
                        abq[s][c] = acc;
                    }
                insr_fused::adj4<D, ORDER>(tv, abq);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) ab[s][4 * q + c] = abq[s][c];
            }
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                acc_gwo[o] += reduce8(v[o], lane);
                if (part == 0) {
                    float t = g[o][0];
                    INSR_PRAGMA_UNROLL
                    for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
                    if (lane == o) acc_gbo += t;
                }
            }
        }
        TCP(4);
        // ================= reverse sweep =================
        for (int l = L; l >= 1; --l) {
            // ---- activation adjoint of layer l: ab (cotangent of a_l) -> zbar_l (in place); layer L's was taken above
            const float4 *tl = tape + (size_t)l * NQ * BT;
            if (l < L) {
                INSR_PRAGMA_UNROLL
                for (int q = 0; q < 2; ++q) {
                    float tv[TV][4], abq[S][4];
                    INSR_PRAGMA_UNROLL
                    for (int t = 0; t < TV; ++t) {
                        const float4 v = tl[(size_t)(q * TV + t) * BT + tid];
                        tv[t][0] = v.x; tv[t][1] = v.y; tv[t][2] = v.z; tv[t][3] = v.w;
                    }
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) abq[s][c] = ab[s][4 * q + c];
                    insr_fused::adj4<D, ORDER>(tv, abq);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) ab[s][4 * q + c] = abq[s][c];
                }
            }
            TCP(5);
            // bias gradient of layer l: sum over points of the value-stream zbar
            {
                float v[8];
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < 8; ++i) v[i] = ab[0][i];
                const float r = reduce8(v, lane);
                if (l == 1) acc_gb[0] += r; else if (l == 2) acc_gb[1] += r; else acc_gb[2] += r;
            }
            // ---- data-gradient operands: zbar_l hi / lo (all streams)
            store_operand(ab);
            TCP(6);
            // ---- weight-gradient operands, one slot per stream: ZT <- zbar_{l,s}, AT <- a_{l-1,s} (from the tape)
            {
                const float4 *tp = tape + (size_t)(l - 1) * NQ * BT;
                float av[S][8];
                INSR_PRAGMA_UNROLL
                for (int q = 0; q < 2; ++q) {
                    float tv[TV][4], a[S][4];
                    if (l == 1) {
                        layer0(q, a, tv);
                    } else {
                        INSR_PRAGMA_UNROLL
                        for (int t = 0; t < TV; ++t) {
                            const float4 v = tp[(size_t)(q * TV + t) * BT + tid];
                            tv[t][0] = v.x; tv[t][1] = v.y; tv[t][2] = v.z; tv[t][3] = v.w;
                        }
                        insr_fused::a_from_tape4<D, ORDER>(tv, a);
                    }
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) av[s][4 * q + c] = a[s][c];
                }
                TCP(7);
                if (pendW) { mbar_wait(mbarW, phW); phW ^= 1; pendW = false; }     // previous weight-gradient MMAs read the slots
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    unsigned char *slot = smraw + M.slots + s * SLOT_BYTES;
                    store_wg_operand(slot, row, part, ab[s]);
                    store_wg_operand(slot + TILE_M * 128, row, part, av[s]);
                }
            }
            TCP(8);
            publish_and_sync();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    // data gradient: D_s = Zbar_s . W_l  (B = transposed copy)
                    issue_contraction(s32(smraw + M.wt_hi + (l - 1) * W_BYTES), s32(smraw + M.wt_lo + (l - 1) * W_BYTES));
                    mma_commit(mbarD);
                }
                __syncwarp();
            }
            TCP(9);
            // ---- cotangent of a_{l-1} from the data-gradient accumulators
            mbar_wait(mbarD, phD);
            phD ^= 1;
            tc_fence_after();
            TCP(10);
            // ---- weight gradient of layer l: warp s issues stream s.  A warp that issues blocks on the shallow MMA
            // queue until most of its MMAs have executed, so the 8 S instructions are spread over S warps (one per
            // scheduler) and issued when the tensor pipe is idle; they execute under the next layer's adjoint.
            if (warp < S) {
                if (elect_one()) {
                    const uint32_t dw = tmem_base + WBASE + 64 * (l - 1);
                    const uint32_t zt = s32(smraw + M.slots + warp * SLOT_BYTES), at = zt + TILE_M * 128;
                    INSR_PRAGMA_UNROLL
                    for (int kp = 0; kp < 8; ++kp)              // 16 points per instruction = two 1024-byte atoms
                        mma_bf16_wg(dw, umma_desc_mn128(zt + 2048 * kp), umma_desc_mn128(at + 2048 * kp), 1);
                    mma_commit(mbarW);
                }
                __syncwarp();
            }
            pendW = true;
            wacc_mask |= 1u << (l - 1);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) tmem_ld8(tmem_row + 32 * s + 8 * part, ab[s]);
            tmem_ld_wait();
            TCP(11);
        }
        // ================= first sine layer =================
        {
            float gx_part[D];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) gx_part[d] = 0.f;
            float v0[8], vd[D > 0 ? D : 1][8];
            INSR_PRAGMA_UNROLL
            for (int q = 0; q < 2; ++q) {
                float tv[TV][4], abq[S][4], a0[S][4];
                layer0(q, a0, tv);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) abq[s][c] = ab[s][4 * q + c];
                insr_fused::adj4<D, ORDER>(tv, abq);
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const int i = 4 * q + c;
                    const float4 wv = *reinterpret_cast<const float4 *>(w1S + (NPT * part + i) * 4);
                    v0[i] = abq[0][c];
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) {
                        vd[d][i] = abq[0][c] * xv[d] + (C::ND > 0 ? abq[(C::ND > 0) ? 1 + d : 0][c] : 0.f);
                        gx_part[d] = fmaf(insr_fused::f4get(wv, d), abq[0][c], gx_part[d]);
                    }
                }
            }
            acc_g1[0] += reduce8(v0, lane);
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) acc_g1[1 + d] += reduce8(vd[d], lane);
            if (p.gx) {
                __syncthreads();                           // partS may still be read by the LSQ combine of slow warps
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) partS[(part * TILE_M + row) * PS + d] = gx_part[d];
                __syncthreads();
                if (part == 0 && valid) {
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) {
                        float acc = 0.f;
                        INSR_PRAGMA_UNROLL
                        for (int q = 0; q < NT; ++q) acc += partS[(q * TILE_M + row) * PS + d];
                        p.gx[n * D + d] = acc;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                   // partS reuse by the next tile
        tc_fence_after();
        TCP(12);
    }

    // ================= flush =================
    if (pendW) { mbar_wait(mbarW, phW); phW ^= 1; pendW = false; }
    tc_fence_after();
    const float wsc = dm.omega;
#ifdef INSR_TC_DEBUG_FLUSH
    if (p.nwarps & 1) wacc_mask = 0;                    // timing experiment: no hidden-layer weight-gradient atomics
#endif
    // hidden-layer weight gradients.  M = 64 accumulator layout: row r of D sits in TMEM lane 32 (r >> 4) + (r & 15),
    // so lanes 0..15 of warp quadrant q hold rows 16 q .. 16 q + 15 = level (q >> 1), neurons 16 (q & 1) + lane;
    // the four warps of a quadrant split the 32 input neurons; columns k (x a1) and 32 + k (x a2) are added.
    // Every CTA of a one-wave launch reaches this point at the same time and would walk the SAME addresses in the same
    // order: same-address reductions serialise in L2 (measured: 39 of 58 us of a 128-tile launch).  So the CTAs start at
    // different layers / halves (blockIdx-rotated order) and each thread reduces 4 consecutive weights per instruction
    // (red.global.add.v4.f32) where the row is 16-byte aligned.
    {
        const int j = 16 * (warp & 1) + lane;
        const bool vec = (H & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gtheta) & 15u) == 0;
        const int rot = (int)blockIdx.x;
        for (int ll = 0; ll < L; ++ll) {
            const int l = 1 + (ll + rot) % L;
            if (!((wacc_mask >> (l - 1)) & 1u)) continue;
            float *gW = p.gtheta + insr_w_offset(dm, l);
            float v1[8], v2[8];
            tmem_ld8(tmem_row + WBASE + 64 * (l - 1) + 8 * part, v1);
            tmem_ld8(tmem_row + WBASE + 64 * (l - 1) + 32 + 8 * part, v2);
            tmem_ld_wait();
            if (lane < 16 && j < H) {
                INSR_PRAGMA_UNROLL
                for (int hh = 0; hh < 2; ++hh) {
                    const int h4 = (hh + (rot / L)) & 1;          // which half of the 8 weights first
                    const int k0 = 8 * part + 4 * h4;
                    float q[4];
                    INSR_PRAGMA_UNROLL
                    for (int i = 0; i < 4; ++i) q[i] = h4 ? wsc * (v1[4 + i] + v2[4 + i]) : wsc * (v1[i] + v2[i]);
                    float *dst = gW + (size_t)j * H + k0;
                    if (vec && k0 + 3 < H) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                     :: "l"(dst), "f"(q[0]), "f"(q[1]), "f"(q[2]), "f"(q[3]) : "memory");
                    } else {
                        INSR_PRAGMA_UNROLL
                        for (int i = 0; i < 4; ++i)
                            if (k0 + i < H) atomicAdd(dst + i, q[i]);
                    }
                }
            }
        }
    }
    // thin layers: lanes with (lane & 3) == 0 hold the totals of neuron 8 part + (lane >> 2)
    if ((lane & 3) == 0) {
        const int j = NPT * part + (lane >> 2);
        if (j < H) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) atomicAdd(p.gtheta + insr_w_offset(dm, L + 1) + o * H + j, acc_gwo[o]);
            atomicAdd(p.gtheta + insr_b_offset(dm, 0) + j, wsc * acc_g1[0]);
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) atomicAdd(p.gtheta + insr_w_offset(dm, 0) + j * D + d, wsc * acc_g1[1 + d]);
            for (int l = 1; l <= L; ++l) atomicAdd(p.gtheta + insr_b_offset(dm, l) + j, wsc * acc_gb[l - 1]);
        }
    }
    if (part == 0 && lane < O) atomicAdd(p.gtheta + insr_b_offset(dm, L + 1) + lane, acc_gbo);
    if (LSQ) {
        float t = loss_acc;
        INSR_PRAGMA_UNROLL
        for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (lane == 0 && part == 0) atomicAdd(p.loss_out, p.scale * t);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
    }
}

inline size_t tc_bwd_ws_bytes(int L, int S) {
    return (size_t)insr_fused::sm_count() * tape_float4_per_cta(L, S) * 16 + 256;
}

template <int D, int O, int ORDER, bool LSQ>
int launch_tc_bwd(Params &p, float *ws, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    constexpr int PS = (O * S > D ? O * S : D);
    if (p.dm.L > LMAX_TC) return -6;
    const SmemB M = smem_map_bwd(p.dm.L, S, PS);
    auto kfn = k_tc_bwd<D, O, ORDER, LSQ>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total + 1024);
    const int64_t tiles = (p.N + TILE_M - 1) / TILE_M;
    int64_t ctas = tiles < insr_fused::sm_count() ? tiles : insr_fused::sm_count();
    float4 *tape = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(ws) + 15) & ~uintptr_t(15));
#ifdef INSR_TC_DEBUG_FLUSH
    { const char *e = getenv("INSR_TC_DEBUG"); p.nwarps = e ? atoi(e) : 0; }
#endif
    kfn<<<dim3((unsigned)ctas), dim3(BT), M.total + 1024, reinterpret_cast<cudaStream_t>(stream)>>>(p, tape);
    ++*launches;
    return 0;
}

}  // namespace insr_tc
#endif  // !INSR_CPU_EMU
