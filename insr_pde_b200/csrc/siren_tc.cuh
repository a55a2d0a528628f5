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

__device__ __forceinline__ uint32_t s32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// UMMA shared-memory descriptor: K-major, no swizzle, LBO = 128 B, SBO = 1024 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// byte offset of element (row, k) inside a K-major canonical operand
__device__ __forceinline__ int op_off(int row, int k) { return (row >> 3) * 1024 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
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

}  // namespace insr_tc
#endif  // !INSR_CPU_EMU
