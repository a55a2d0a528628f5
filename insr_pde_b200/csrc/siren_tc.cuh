// siren_tc.cuh -- tcgen05 / TMEM forward kernel for the resident-weights family (H <= 32).
//
// Why: ncu shows the FFMA kernels limited by shared-memory delivery (LDS wavefronts), not by the
// FP32 pipe; the 5th-gen tensor cores read their operands straight from shared memory through
// descriptors.  Single-pass TF32 misses the 1e-4 parity target, so every hidden-layer
// contraction is done as a 3xTF32 split:  a = a_hi + a_lo (a_hi = fp32 with the 13 low mantissa bits
// cleared, a_lo = a - a_hi exact),  a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo  accumulated in FP32 in
// TMEM (dropped term ~2^-22 relative).
//
// Mapping.  CTA tile = 128 collocation points = the M dimension of tcgen05.mma (one TMEM lane per
// point).  Each forward-mode stream s is its own MMA chain  D_s[128 x 32] = A_s[128 x 32] . W^T
// into TMEM columns [32 s, 32 s + 32), so a thread (= one point) reads ALL streams of its neurons
// with tcgen05.ld and runs the sine-stream algebra in registers, then writes the next layer's
// operand -- already split into hi / lo -- back to shared memory in the UMMA K-major canonical
// layout (8 x 16-byte core matrices, LBO = 128 B along K, SBO = 1024 B along M/N).  256 threads:
// warps w and w+4 share TMEM lanes 32 (w%4) .. +31 and split the 32 neurons in two halves.
// One elected thread issues the 3 x 4 x S MMAs of a layer and commits them to an mbarrier.
//
// Only compiled by nvcc (inline PTX); the host-side SIMT emulation does not cover this file.
#pragma once
#include "siren_fused.cuh"

#ifndef INSR_CPU_EMU
namespace insr_tc {

using insr_fused::Params;
constexpr int HP = 32;
constexpr int TILE_M = 128;
constexpr int THREADS = 256;
constexpr int OP_BYTES = TILE_M * HP * 4;          // one 128 x 32 fp32 operand: 16 KB
constexpr int W_BYTES = HP * HP * 4;               // one 32 x 32 weight operand: 4 KB
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
// kind::tf32, D = F32, A/B = TF32, both K-major, N = 32, M = 128
constexpr uint32_t IDESC_DGRAD = IDESC | (1u << 16);                 // B operand MN-major (W read "transposed")
constexpr uint32_t IDESC_WGRAD = IDESC | (1u << 15) | (1u << 16);    // A and B MN-major (reduction over points)

__device__ __forceinline__ uint32_t s32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// UMMA shared-memory descriptor: K-major, no swizzle, LBO = 128 B, SBO = 1024 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// byte offset of element (row, k) inside a K-major canonical operand
__device__ __forceinline__ int op_off(int row, int k) { return (row >> 3) * 1024 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }

// the same 8 x 16-byte core matrices read MN-major (M/N index contiguous inside the 16 bytes, the 8 rows
// are the K index): leading (K-group) offset 1024 B, stride (M/N-group) offset 128 B
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate,
                                         uint32_t idesc = IDESC) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

// write 4 consecutive neurons (k = 4 kg .. 4 kg + 3) of one stream of row `row`, split hi / lo
__device__ __forceinline__ void store_split4(unsigned char *op_hi, unsigned char *op_lo, int row, int kg,
                                             const float (&a)[4]) {
    float h[4], l[4];
    INSR_PRAGMA_UNROLL
    for (int c = 0; c < 4; ++c) { h[c] = tf32_hi(a[c]); l[c] = a[c] - h[c]; }
    const int off = (row >> 3) * 1024 + kg * 128 + (row & 7) * 16;
    *reinterpret_cast<float4 *>(op_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(op_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// shared-memory map (bytes)
struct Smem {
    int w_hi, w_lo, bias, w1, wo, bo, a_hi, a_lo, part, mbar, tmem, total;
};
__host__ __device__ inline Smem smem_map(int L, int S, int O) {
    Smem m;
    int o = 0;
    m.w_hi = o; o += L * W_BYTES;
    m.w_lo = o; o += L * W_BYTES;
    m.a_hi = o; o += S * OP_BYTES;
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
__global__ void __launch_bounds__(THREADS, 1) k_tc_fwd(Params p, int tmem_cols) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const SirenDims dm = p.dm;
    const int L = dm.L, H = dm.H;
    const Smem M = smem_map(L, S, O);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;            // TMEM lane == point inside the tile
    const int half = warp >> 2;                        // neurons 16 half .. 16 half + 15
    float *biasS = reinterpret_cast<float *>(smraw + M.bias);
    float *w1S = reinterpret_cast<float *>(smraw + M.w1);
    float *woS = reinterpret_cast<float *>(smraw + M.wo);
    float *boS = reinterpret_cast<float *>(smraw + M.bo);
    float *partS = reinterpret_cast<float *>(smraw + M.part);
    const uint32_t mbar = s32(smraw + M.mbar);

    // ---- stage weights: hidden layers split hi / lo in UMMA K-major layout (omega folded in)
    const float w = dm.omega;
    for (int idx = tid; idx < L * HP * HP; idx += THREADS) {
        const int l = idx / (HP * HP), j = (idx / HP) % HP, k = idx % HP;
        float v = 0.f;
        if (j < H && k < H) v = w * p.theta[insr_w_offset(dm, l + 1) + (int64_t)j * H + k];
        const float hi = tf32_hi(v);
        *reinterpret_cast<float *>(smraw + M.w_hi + l * W_BYTES + op_off(j, k)) = hi;
        *reinterpret_cast<float *>(smraw + M.w_lo + l * W_BYTES + op_off(j, k)) = v - hi;
    }
    for (int idx = tid; idx < L * HP; idx += THREADS) {
        const int l = idx / HP, j = idx % HP;
        biasS[idx] = (j < H) ? w * p.theta[insr_b_offset(dm, l + 1) + j] : 0.f;
    }
    for (int idx = tid; idx < HP * 4; idx += THREADS) {
        const int j = idx >> 2, d = idx & 3;
        float v = 0.f;
        if (j < H) {
            if (d < D) v = w * p.theta[insr_w_offset(dm, 0) + (int64_t)j * D + d];
            else if (d == 3) v = w * p.theta[insr_b_offset(dm, 0) + j];
        }
        w1S[idx] = v;
    }
    for (int idx = tid; idx < 3 * HP; idx += THREADS) {
        const int o = idx / HP, j = idx % HP;
        woS[idx] = (o < O && j < H) ? p.theta[insr_w_offset(dm, L + 1) + (int64_t)o * H + j] : 0.f;
    }
    if (tid < 4) boS[tid] = (tid < O) ? p.theta[insr_b_offset(dm, L + 1) + tid] : 0.f;

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

        float alast[S][16];                              // post-activations of the current layer (this thread's 16 neurons)
        // ---- first sine layer (FFMA): 4 groups of 4 neurons
        INSR_PRAGMA_UNROLL
        for (int g4 = 0; g4 < 4; ++g4) {
            float z[S][4], a[S][4], tv[S + 1][4];
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < 4; ++c) {
                const int j = 16 * half + 4 * g4 + c;
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
            for (int s = 0; s < S; ++s) {
                store_split4(smraw + M.a_hi + s * OP_BYTES, smraw + M.a_lo + s * OP_BYTES, row, 4 * half + g4, a[s]);
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) alast[s][4 * g4 + c] = a[s][c];
            }
        }
        // ---- hidden layers on the tensor cores
        for (int l = 0; l < L; ++l) {
            fence_async_smem();                          // generic-proxy operand writes -> visible to the async proxy
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t whi = s32(smraw + M.w_hi + l * W_BYTES), wlo = s32(smraw + M.w_lo + l * W_BYTES);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    const uint32_t ahi = s32(smraw + M.a_hi + s * OP_BYTES), alo = s32(smraw + M.a_lo + s * OP_BYTES);
                    const uint32_t d = tmem_base + 32 * s;
                    INSR_PRAGMA_UNROLL
                    for (int ks = 0; ks < 4; ++ks) {      // K = 8 per instruction: two 16-byte core matrices = 256 B
                        mma_tf32(d, umma_desc(ahi + 256 * ks), umma_desc(whi + 256 * ks), ks > 0);
                        mma_tf32(d, umma_desc(alo + 256 * ks), umma_desc(whi + 256 * ks), 1);
                        mma_tf32(d, umma_desc(ahi + 256 * ks), umma_desc(wlo + 256 * ks), 1);
                    }
                }
                mma_commit(mbar);
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            tc_fence_after();
            // ---- epilogue: 2 groups of 8 neurons of this thread's half
            INSR_PRAGMA_UNROLL
            for (int g8 = 0; g8 < 2; ++g8) {
                float zz[S][8];
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
                    for (int s = 0; s < S; ++s) {
                        store_split4(smraw + M.a_hi + s * OP_BYTES, smraw + M.a_lo + s * OP_BYTES, row,
                                     4 * half + 2 * g8 + q, a[s]);
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) alast[s][8 * g8 + 4 * q + c] = a[s][c];
                    }
                }
            }
        }
        // ---- output layer (FFMA): this thread's 16 neurons, then combine the two halves
        float out[O][S];
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o)
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                float acc = 0.f;
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < 16; ++i) acc = fmaf(woS[o * HP + 16 * half + i], alast[s][i], acc);
                out[o][s] = acc;
            }
        if (half == 1) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) partS[row * 16 + o * S + s] = out[o][s];
        }
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
        __syncthreads();                                 // partS / operands are rewritten by the next tile
    }
    // ---- TMEM release
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

inline int tmem_columns(int S) {
    int c = 32;
    while (c < 32 * S) c <<= 1;
    return c;
}

template <int D, int O, int ORDER>
int launch_tc_fwd(Params &p, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    static_assert(O * S <= 16, "output partial buffer holds 16 values per point");
    const Smem M = smem_map(p.dm.L, S, O);
    auto kfn = k_tc_fwd<D, O, ORDER>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total + 1024);
    const int64_t tiles = (p.N + TILE_M - 1) / TILE_M;
    int64_t ctas = tiles < insr_fused::sm_count() ? tiles : insr_fused::sm_count();
    kfn<<<dim3((unsigned)ctas), dim3(THREADS), M.total + 1024, reinterpret_cast<cudaStream_t>(stream)>>>(p, tmem_columns(S));
    ++*launches;
    return 0;
}


// =============================================================================================
// backward / fused-closure kernel on the tensor cores
//   forward recompute   D_s = A_s W^T                 (A K-major, W K-major)
//   data gradient       D_s = Zbar_s W                (Zbar K-major, the SAME W buffer read MN-major)
//   weight gradient     Wacc_l[j][k] += sum_p Zbar_s[p][j] A_s[p][k]   (both operands read MN-major; the
//                       accumulators stay in TMEM for the whole launch and are read out once at the end)
// The tape (sin, cos, t_d, t_q per activation) lives in a per-CTA global scratch that stays L2-resident.
// Thin layers (first / output layer, biases, d loss/d x) are reduced over the 32 points of a warp with a
// halving butterfly (16 shuffles per 16 values) into 7 persistent registers.
// =============================================================================================
__device__ __forceinline__ float reduce16(float (&v)[16], int lane) {
    // after the call the lane holds the warp-wide sum of value index (lane >> 1) & 15
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 16;
        const float send = up ? v[i] : v[i + 8];
        const float keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 8;
        const float send = up ? v[i] : v[i + 4];
        const float keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    INSR_PRAGMA_UNROLL
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 4;
        const float send = up ? v[i] : v[i + 2];
        const float keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool up = lane & 2;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct SmemB {
    int w_hi, w_lo, x_hi, x_lo, y_hi[2], y_lo[2], bias, w1, wo, bo, part, mbar, tmem, total;
};
__host__ __device__ inline SmemB smem_map_bwd(int L, int S) {
    SmemB m;
    int o = 0;
    m.w_hi = o; o += L * W_BYTES;
    m.w_lo = o; o += L * W_BYTES;
    m.x_hi = o; o += S * OP_BYTES;
    m.x_lo = o; o += S * OP_BYTES;
    for (int i = 0; i < 2; ++i) { m.y_hi[i] = o; o += OP_BYTES; m.y_lo[i] = o; o += OP_BYTES; }
    m.bias = o; o += L * HP * 4;
    m.w1 = o; o += HP * 16;
    m.wo = o; o += 3 * HP * 4;
    m.bo = o; o += 16;
    m.part = o; o += 2 * TILE_M * 8 * 4;       // per-half partial outputs / gx: [2][128][8]
    m.mbar = o; o += 32;
    m.tmem = o; o += 16;
    m.total = o;
    return m;
}
// tape: per CTA  [(L+1)][4 (S+1)][256] float4
__host__ __device__ inline size_t tape_float4_per_cta(int L, int S) { return (size_t)(L + 1) * 4 * (S + 1) * THREADS; }

template <int D, int O, int ORDER, bool LSQ>
__global__ void __launch_bounds__(THREADS, 1) k_tc_bwd(Params p, float4 *__restrict__ tape_all, int tmem_cols) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int TV = S + 1;
    constexpr int NQ = 4 * TV;                       // float4 tape slots per thread per layer
    static_assert(O * S <= 8, "partial buffers hold 8 values per point and half");
    extern __shared__ __align__(1024) unsigned char smraw[];
    const SirenDims dm = p.dm;
    const int L = dm.L, H = dm.H;
    const SmemB M = smem_map_bwd(L, S);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;
    const int half = warp >> 2;
    float *biasS = reinterpret_cast<float *>(smraw + M.bias);
    float *w1S = reinterpret_cast<float *>(smraw + M.w1);
    float *woS = reinterpret_cast<float *>(smraw + M.wo);
    float *boS = reinterpret_cast<float *>(smraw + M.bo);
    float *partS = reinterpret_cast<float *>(smraw + M.part);
    const uint32_t mbarD = s32(smraw + M.mbar), mbarY0 = mbarD + 8, mbarY1 = mbarD + 16;
    float4 *tape = tape_all + (size_t)blockIdx.x * tape_float4_per_cta(L, S);

    // ---- stage weights (as in the forward kernel)
    const float w = dm.omega;
    for (int idx = tid; idx < L * HP * HP; idx += THREADS) {
        const int l = idx / (HP * HP), j = (idx / HP) % HP, k = idx % HP;
        float v = 0.f;
        if (j < H && k < H) v = w * p.theta[insr_w_offset(dm, l + 1) + (int64_t)j * H + k];
        const float hi = tf32_hi(v);
        *reinterpret_cast<float *>(smraw + M.w_hi + l * W_BYTES + op_off(j, k)) = hi;
        *reinterpret_cast<float *>(smraw + M.w_lo + l * W_BYTES + op_off(j, k)) = v - hi;
    }
    for (int idx = tid; idx < L * HP; idx += THREADS) {
        const int l = idx / HP, j = idx % HP;
        biasS[idx] = (j < H) ? w * p.theta[insr_b_offset(dm, l + 1) + j] : 0.f;
    }
    for (int idx = tid; idx < HP * 4; idx += THREADS) {
        const int j = idx >> 2, d = idx & 3;
        float v = 0.f;
        if (j < H) {
            if (d < D) v = w * p.theta[insr_w_offset(dm, 0) + (int64_t)j * D + d];
            else if (d == 3) v = w * p.theta[insr_b_offset(dm, 0) + j];
        }
        w1S[idx] = v;
    }
    for (int idx = tid; idx < 3 * HP; idx += THREADS) {
        const int o = idx / HP, j = idx % HP;
        woS[idx] = (o < O && j < H) ? p.theta[insr_w_offset(dm, L + 1) + (int64_t)o * H + j] : 0.f;
    }
    if (tid < 4) boS[tid] = (tid < O) ? p.theta[insr_b_offset(dm, L + 1) + tid] : 0.f;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s32(smraw + M.tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(mbarD, 1); mbar_init(mbarY0, 1); mbar_init(mbarY1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smraw + M.tmem);
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t tmem_wacc = tmem_base + 32 * S;                // L blocks of 32 columns
    uint32_t phD = 0, phY[2] = {0, 0};
    bool pendY[2] = {false, false};
    uint32_t wacc_mask = 0;                                       // bit l-1: Wacc_l holds data

    // persistent thin-layer partial sums: the lane holds value (lane >> 1) & 15 of this thread's neuron half
    float acc_gwo[O], acc_g1[1 + D], acc_gb[3] = {0.f, 0.f, 0.f};
    float acc_gbo = 0.f, loss_acc = 0.f;
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o) acc_gwo[o] = 0.f;
    INSR_PRAGMA_UNROLL
    for (int d = 0; d <= D; ++d) acc_g1[d] = 0.f;

    auto wait_y = [&](int slot) {
        if (pendY[slot]) { mbar_wait(slot ? mbarY1 : mbarY0, phY[slot]); phY[slot] ^= 1; pendY[slot] = false; }
    };
    auto publish_and_sync = [&]() {          // operand writes (generic proxy) -> async proxy, then CTA barrier
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };

    const int64_t ntiles = (p.N + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n = tile * TILE_M + row;
        const bool valid = n < p.N;
        float xv[D];
        INSR_PRAGMA_UNROLL
        for (int d = 0; d < D; ++d) xv[d] = valid ? __ldg(p.x + n * D + d) : 0.f;
        wait_y(0); wait_y(1);                              // previous tile's weight-gradient MMAs still read X

        float alast[S][16];
        // ================= forward with tape =================
        INSR_PRAGMA_UNROLL
        for (int g4 = 0; g4 < 4; ++g4) {
            float z[S][4], a[S][4], tv[TV][4];
            INSR_PRAGMA_UNROLL
            for (int c = 0; c < 4; ++c) {
                const int j = 16 * half + 4 * g4 + c;
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
            for (int t = 0; t < TV; ++t) tape[(size_t)(g4 * TV + t) * THREADS + tid] = make_float4(tv[t][0], tv[t][1], tv[t][2], tv[t][3]);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                store_split4(smraw + M.x_hi + s * OP_BYTES, smraw + M.x_lo + s * OP_BYTES, row, 4 * half + g4, a[s]);
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) alast[s][4 * g4 + c] = a[s][c];
            }
        }
        for (int l = 0; l < L; ++l) {
            publish_and_sync();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t whi = s32(smraw + M.w_hi + l * W_BYTES), wlo = s32(smraw + M.w_lo + l * W_BYTES);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    const uint32_t ahi = s32(smraw + M.x_hi + s * OP_BYTES), alo = s32(smraw + M.x_lo + s * OP_BYTES);
                    const uint32_t d = tmem_base + 32 * s;
                    INSR_PRAGMA_UNROLL
                    for (int ks = 0; ks < 4; ++ks) {
                        mma_tf32(d, umma_desc(ahi + 256 * ks), umma_desc(whi + 256 * ks), ks > 0);
                        mma_tf32(d, umma_desc(alo + 256 * ks), umma_desc(whi + 256 * ks), 1);
                        mma_tf32(d, umma_desc(ahi + 256 * ks), umma_desc(wlo + 256 * ks), 1);
                    }
                }
                mma_commit(mbarD);
            }
            mbar_wait(mbarD, phD);
            phD ^= 1;
            tc_fence_after();
            float4 *tl = tape + (size_t)(l + 1) * NQ * THREADS;
            INSR_PRAGMA_UNROLL
            for (int g8 = 0; g8 < 2; ++g8) {
                float zz[S][8];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) tmem_ld8(tmem_row + 32 * s + 16 * half + 8 * g8, zz[s]);
                tmem_ld_wait();
                INSR_PRAGMA_UNROLL
                for (int q = 0; q < 2; ++q) {
                    float z[S][4], a[S][4], tv[TV][4];
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) {
                        const int j = 16 * half + 8 * g8 + 4 * q + c;
                        z[0][c] = zz[0][4 * q + c] + biasS[l * HP + j];
                        INSR_PRAGMA_UNROLL
                        for (int s = 1; s < S; ++s) z[s][c] = zz[s][4 * q + c];
                    }
                    insr_fused::act4<D, ORDER>(z, a, tv);
                    const int g4 = 2 * g8 + q;
                    INSR_PRAGMA_UNROLL
                    for (int t = 0; t < TV; ++t) tl[(size_t)(g4 * TV + t) * THREADS + tid] = make_float4(tv[t][0], tv[t][1], tv[t][2], tv[t][3]);
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s) {
                        store_split4(smraw + M.x_hi + s * OP_BYTES, smraw + M.x_lo + s * OP_BYTES, row, 4 * half + g4, a[s]);
                        INSR_PRAGMA_UNROLL
                        for (int c = 0; c < 4; ++c) alast[s][4 * g4 + c] = a[s][c];
                    }
                }
            }
        }
        // ================= output layer: cotangents g[o][s] =================
        float g[O][S];
        if constexpr (LSQ) {
            float out[O][S];
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    float acc = 0.f;
                    INSR_PRAGMA_UNROLL
                    for (int i = 0; i < 16; ++i) acc = fmaf(woS[o * HP + 16 * half + i], alast[s][i], acc);
                    out[o][s] = acc;
                    partS[(half * TILE_M + row) * 8 + o * S + s] = acc;
                }
            __syncthreads();
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) {
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) out[o][s] += partS[((half ^ 1) * TILE_M + row) * 8 + o * S + s];
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
                if (half == 0) loss_acc = fmaf(r, r, loss_acc);
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
        // output-layer gradients (this thread's 16 neurons) and the cotangent of the last sine layer
        float ab[S][16];
        INSR_PRAGMA_UNROLL
        for (int o = 0; o < O; ++o) {
            float v[16];
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < 16; ++i) {
                float acc = 0.f;
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) acc = fmaf(g[o][s], alast[s][i], acc);
                v[i] = acc;
            }
            acc_gwo[o] += reduce16(v, lane);
            if (half == 0) {
                float t = g[o][0];
                INSR_PRAGMA_UNROLL
                for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
                if (lane == o) acc_gbo += t;
            }
        }
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s)
            INSR_PRAGMA_UNROLL
            for (int i = 0; i < 16; ++i) {
                float acc = 0.f;
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o) acc = fmaf(woS[o * HP + 16 * half + i], g[o][s], acc);
                ab[s][i] = acc;
            }

        // ================= reverse sweep =================
        for (int l = L; l >= 1; --l) {
            // ---- activation adjoint of layer l: ab (cotangent of a_l) -> zbar_l (in place)
            const float4 *tl = tape + (size_t)l * NQ * THREADS;
            INSR_PRAGMA_UNROLL
            for (int g4 = 0; g4 < 4; ++g4) {
                float tv[TV][4], abq[S][4];
                INSR_PRAGMA_UNROLL
                for (int t = 0; t < TV; ++t) {
                    const float4 v = tl[(size_t)(g4 * TV + t) * THREADS + tid];
                    tv[t][0] = v.x; tv[t][1] = v.y; tv[t][2] = v.z; tv[t][3] = v.w;
                }
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) abq[s][c] = ab[s][4 * g4 + c];
                insr_fused::adj4<D, ORDER>(tv, abq);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) ab[s][4 * g4 + c] = abq[s][c];
            }
            // bias gradient of layer l: sum over points of the value-stream zbar
            {
                float v[16];
                INSR_PRAGMA_UNROLL
                for (int i = 0; i < 16; ++i) v[i] = ab[0][i];
                const float r = reduce16(v, lane);
                if (l == 1) acc_gb[0] += r; else if (l == 2) acc_gb[1] += r; else acc_gb[2] += r;
            }
            // ---- X <- zbar_l (all streams).  The previous layer's weight-gradient MMAs read X: wait for them.
            wait_y(0); wait_y(1);
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s)
                INSR_PRAGMA_UNROLL
                for (int g4 = 0; g4 < 4; ++g4) {
                    const float a4[4] = {ab[s][4 * g4], ab[s][4 * g4 + 1], ab[s][4 * g4 + 2], ab[s][4 * g4 + 3]};
                    store_split4(smraw + M.x_hi + s * OP_BYTES, smraw + M.x_lo + s * OP_BYTES, row, 4 * half + g4, a4);
                }
            publish_and_sync();
            const uint32_t whi = s32(smraw + M.w_hi + (l - 1) * W_BYTES), wlo = s32(smraw + M.w_lo + (l - 1) * W_BYTES);
            if (tid == 0) {                            // data gradient: D_s = Zbar_s . W_l   (B = W read MN-major)
                tc_fence_after();
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) {
                    const uint32_t zhi = s32(smraw + M.x_hi + s * OP_BYTES), zlo = s32(smraw + M.x_lo + s * OP_BYTES);
                    const uint32_t d = tmem_base + 32 * s;
                    INSR_PRAGMA_UNROLL
                    for (int ks = 0; ks < 4; ++ks) {   // 8 reduction indices j per instruction = one 1024-byte row group of W
                        mma_tf32(d, umma_desc(zhi + 256 * ks), umma_desc_mn(whi + 1024 * ks), ks > 0, IDESC_DGRAD);
                        mma_tf32(d, umma_desc(zlo + 256 * ks), umma_desc_mn(whi + 1024 * ks), 1, IDESC_DGRAD);
                        mma_tf32(d, umma_desc(zhi + 256 * ks), umma_desc_mn(wlo + 1024 * ks), 1, IDESC_DGRAD);
                    }
                }
                mma_commit(mbarD);
            }
            // ---- weight gradient, one stream at a time through the two Y slots: Y <- a_{l-1,s}
            const float4 *tp = tape + (size_t)(l - 1) * NQ * THREADS;
            float aprev[S][16];
            INSR_PRAGMA_UNROLL
            for (int g4 = 0; g4 < 4; ++g4) {
                float tv[TV][4], a[S][4];
                INSR_PRAGMA_UNROLL
                for (int t = 0; t < TV; ++t) {
                    const float4 v = tp[(size_t)(g4 * TV + t) * THREADS + tid];
                    tv[t][0] = v.x; tv[t][1] = v.y; tv[t][2] = v.z; tv[t][3] = v.w;
                }
                insr_fused::a_from_tape4<D, ORDER>(tv, a);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) aprev[s][4 * g4 + c] = a[s][c];
            }
            INSR_PRAGMA_UNROLL
            for (int s = 0; s < S; ++s) {
                const int slot = s & 1;
                wait_y(slot);
                INSR_PRAGMA_UNROLL
                for (int g4 = 0; g4 < 4; ++g4) {
                    const float a4[4] = {aprev[s][4 * g4], aprev[s][4 * g4 + 1], aprev[s][4 * g4 + 2], aprev[s][4 * g4 + 3]};
                    store_split4(smraw + M.y_hi[slot], smraw + M.y_lo[slot], row, 4 * half + g4, a4);
                }
                publish_and_sync();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t zhi = s32(smraw + M.x_hi + s * OP_BYTES), zlo = s32(smraw + M.x_lo + s * OP_BYTES);
                    const uint32_t yhi = s32(smraw + M.y_hi[slot]), ylo = s32(smraw + M.y_lo[slot]);
                    const uint32_t d = tmem_wacc + 32 * (l - 1);
                    const bool fresh = !((wacc_mask >> (l - 1)) & 1u) && s == 0;
                    INSR_PRAGMA_UNROLL
                    for (int pg = 0; pg < 16; ++pg) {   // 8 points per instruction = one 1024-byte row group of both operands
                        mma_tf32(d, umma_desc_mn(zhi + 1024 * pg), umma_desc_mn(yhi + 1024 * pg), !(fresh && pg == 0), IDESC_WGRAD);
                        mma_tf32(d, umma_desc_mn(zlo + 1024 * pg), umma_desc_mn(yhi + 1024 * pg), 1, IDESC_WGRAD);
                        mma_tf32(d, umma_desc_mn(zhi + 1024 * pg), umma_desc_mn(ylo + 1024 * pg), 1, IDESC_WGRAD);
                    }
                    mma_commit(slot ? mbarY1 : mbarY0);
                }
                pendY[slot] = true;
            }
            wacc_mask |= 1u << (l - 1);
            // ---- cotangent of a_{l-1} from the data-gradient accumulators
            mbar_wait(mbarD, phD);
            phD ^= 1;
            tc_fence_after();
            INSR_PRAGMA_UNROLL
            for (int g8 = 0; g8 < 2; ++g8) {
                float zz[S][8];
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s) tmem_ld8(tmem_row + 32 * s + 16 * half + 8 * g8, zz[s]);
                tmem_ld_wait();
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int i = 0; i < 8; ++i) ab[s][8 * g8 + i] = zz[s][i];
            }
        }
        // ================= first sine layer =================
        {
            float gx_part[D];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) gx_part[d] = 0.f;
            float v0[16], vd[D > 0 ? D : 1][16];
            INSR_PRAGMA_UNROLL
            for (int g4 = 0; g4 < 4; ++g4) {
                float tv[TV][4], abq[S][4];
                INSR_PRAGMA_UNROLL
                for (int t = 0; t < TV; ++t) {
                    const float4 v = tape[(size_t)(g4 * TV + t) * THREADS + tid];
                    tv[t][0] = v.x; tv[t][1] = v.y; tv[t][2] = v.z; tv[t][3] = v.w;
                }
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    INSR_PRAGMA_UNROLL
                    for (int c = 0; c < 4; ++c) abq[s][c] = ab[s][4 * g4 + c];
                insr_fused::adj4<D, ORDER>(tv, abq);
                INSR_PRAGMA_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const int i = 4 * g4 + c;
                    const float4 wv = *reinterpret_cast<const float4 *>(w1S + (16 * half + i) * 4);
                    v0[i] = abq[0][c];
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) {
                        vd[d][i] = abq[0][c] * xv[d] + (C::ND > 0 ? abq[(C::ND > 0) ? 1 + d : 0][c] : 0.f);
                        gx_part[d] = fmaf(insr_fused::f4get(wv, d), abq[0][c], gx_part[d]);
                    }
                }
            }
            acc_g1[0] += reduce16(v0, lane);
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) acc_g1[1 + d] += reduce16(vd[d], lane);
            if (p.gx) {
                __syncthreads();                           // partS may still be read by the LSQ combine of slow warps
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) partS[(half * TILE_M + row) * 8 + d] = gx_part[d];
                __syncthreads();
                if (half == 0 && valid) {
                    INSR_PRAGMA_UNROLL
                    for (int d = 0; d < D; ++d) p.gx[n * D + d] = gx_part[d] + partS[(TILE_M + row) * 8 + d];
                }
            }
        }
        __syncthreads();                                   // partS reuse by the next tile
    }

    // ================= flush =================
    wait_y(0); wait_y(1);
    tc_fence_after();
    const float wsc = dm.omega;
    // hidden-layer weight gradients: TMEM lanes 0..31 = output neuron j, 32 columns = input k (warp 0 reads them)
    if (warp == 0) {
        for (int l = 1; l <= L; ++l) {
            if (!((wacc_mask >> (l - 1)) & 1u)) continue;
            float *gW = p.gtheta + insr_w_offset(dm, l);
            INSR_PRAGMA_UNROLL
            for (int c8 = 0; c8 < 4; ++c8) {
                float vals[8];
                tmem_ld8(tmem_wacc + 32 * (l - 1) + 8 * c8, vals);
                tmem_ld_wait();
                if (lane < H) {
                    INSR_PRAGMA_UNROLL
                    for (int i = 0; i < 8; ++i)
                        if (8 * c8 + i < H) atomicAdd(gW + (size_t)lane * H + 8 * c8 + i, wsc * vals[i]);
                }
            }
        }
    }
    // thin layers: the even lanes hold the totals of neuron 16 half + (lane >> 1); one warp per (half, row block)
    if ((lane & 1) == 0) {
        const int j = 16 * half + (lane >> 1);
        if (j < H) {
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < O; ++o) atomicAdd(p.gtheta + insr_w_offset(dm, L + 1) + o * H + j, acc_gwo[o]);
            atomicAdd(p.gtheta + insr_b_offset(dm, 0) + j, wsc * acc_g1[0]);
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) atomicAdd(p.gtheta + insr_w_offset(dm, 0) + j * D + d, wsc * acc_g1[1 + d]);
            for (int l = 1; l <= L; ++l) atomicAdd(p.gtheta + insr_b_offset(dm, l) + j, wsc * acc_gb[l - 1]);
        }
    }
    if (half == 0 && lane < O) atomicAdd(p.gtheta + insr_b_offset(dm, L + 1) + lane, acc_gbo);
    if (LSQ) {
        float t = loss_acc;
        INSR_PRAGMA_UNROLL
        for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (lane == 0 && half == 0) atomicAdd(p.loss_out, p.scale * t);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

inline size_t tc_bwd_ws_bytes(int L, int S) {
    return (size_t)insr_fused::sm_count() * tape_float4_per_cta(L, S) * 16 + 256;
}

template <int D, int O, int ORDER, bool LSQ>
int launch_tc_bwd(Params &p, float *ws, void *stream, int64_t *launches) {
    constexpr int S = StreamCfg<D, ORDER>::S;
    const SmemB M = smem_map_bwd(p.dm.L, S);
    auto kfn = k_tc_bwd<D, O, ORDER, LSQ>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total);
    const int64_t tiles = (p.N + TILE_M - 1) / TILE_M;
    int64_t ctas = tiles < insr_fused::sm_count() ? tiles : insr_fused::sm_count();
    int cols = 32;
    while (cols < 32 * S + 32 * p.dm.L) cols <<= 1;
    float4 *tape = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(ws) + 15) & ~uintptr_t(15));
    kfn<<<dim3((unsigned)ctas), dim3(THREADS), M.total, reinterpret_cast<cudaStream_t>(stream)>>>(p, tape, cols);
    ++*launches;
    return 0;
}

}  // namespace insr_tc
#endif  // !INSR_CPU_EMU
