// siren_tc_target.cuh -- the frozen-net side of the fluid / advection closures in ONE kernel.
//
// Every least-squares closure of the reference compares the trainable field with a TARGET computed from frozen
// networks at the same collocation points:
//     fluid/model.py:78-87     u_adv  = u_prev(clamp(x - u_prev(x) dt, -1, 1))        semi-Lagrangian backtrace
//     fluid/model.py:108-109   div u  = d u_x / dx + d u_y / dy   of the (detached) velocity
//     fluid/model.py:131-137   u_prev(x) - grad p(x)
//     advection/model.py:78-84 u_prev / dt - vel / 2 * d u_prev / dx
// All of them are "evaluate one or two frozen SIREN fields (H <= 32), optionally feed the first one's value back as the
// second one's position, take a fixed linear combination of the outputs".  k_tc_target does exactly that per 128-point
// tile with the tcgen05 forward machinery of siren_tc.cuh (weights of both nets resident in shared memory, hi operand in
// TMEM, 3xTF32): the intermediate values, the clamp and the combination never leave the registers; only target (N, R) is
// written.  Replaces 2 evaluate kernels + 2-4 elementwise torch kernels per closure.
#pragma once
#include "siren_tc.cuh"

#ifndef INSR_CPU_EMU
namespace insr_tc {

struct NetSmem {
    int w_hi, w_lo, bias, w1, wo, bo;
};
struct TargetSmem {
    NetSmem a, b;
    int a_lo, part, mbar, tmem, total;
};
__host__ __device__ inline int net_smem(NetSmem &m, int o, int L) {
    m.w_hi = o; o += L * W_BYTES;
    m.w_lo = o; o += L * W_BYTES;
    m.bias = o; o += L * HP * 4;
    m.w1 = o; o += HP * 16;
    m.wo = o; o += 3 * HP * 4;
    m.bo = o; o += 16;
    return o;
}
__host__ __device__ inline TargetSmem target_smem(int LA, int LB, int Smax) {
    TargetSmem m;
    int o = 0;
    m.a_lo = o; o += Smax * OP_BYTES;
    o = net_smem(m.a, o, LA);
    o = net_smem(m.b, o, LB);
    m.part = o; o += 2 * TILE_M * 16 * 4;          // both neuron halves park their partial outputs: [2][128][16]
    m.mbar = o; o += 16;
    m.tmem = o; o += 16;
    m.total = o;
    return m;
}

struct TargetParams {
    SirenDims dmA, dmB;
    const float *thetaA, *thetaB;
    const float *x;
    int64_t N;
    float *target;              // (N, n_res)
    int n_res;
    float coefA[2 * 3 * 4];     // [c][o][s], stream stride = S of the evaluation
    float coefB[2 * 3 * 4];
    float dt, lo, hi;           // backtrace: x' = clamp(x - dt * y_A(x), lo, hi)
};

__device__ __forceinline__ void stage_net(const SirenDims &dm, const float *theta, unsigned char *smraw, const NetSmem &M) {
    Params p{};
    p.dm = dm;
    p.theta = theta;
    stage_hidden(p, smraw + M.w_hi, smraw + M.w_lo, false);
    stage_small(p, reinterpret_cast<float *>(smraw + M.bias), reinterpret_cast<float *>(smraw + M.w1),
                reinterpret_cast<float *>(smraw + M.wo), reinterpret_cast<float *>(smraw + M.bo));
}

// one network on one 128-point tile (the tile body of k_tc_fwd); every thread of the CTA calls it and ends up with the
// COMPLETE outputs of its point (both neuron halves combined)
template <int D, int O, int ORDER>
__device__ __forceinline__ void eval_tile(unsigned char *smraw, const NetSmem &M, int a_lo_off, float *partS, uint32_t tmem_base,
                                          uint32_t tmem_row, uint32_t mbar, uint32_t &phase, int L, int row, int half, int warp,
                                          const float (&xv)[D], float (&out)[O][StreamCfg<D, ORDER>::S]) {
    typedef StreamCfg<D, ORDER> C;
    constexpr int S = C::S;
    constexpr int ABASE = 32 * S;
    const float *biasS = reinterpret_cast<const float *>(smraw + M.bias);
    const float *w1S = reinterpret_cast<const float *>(smraw + M.w1);
    const float *woS = reinterpret_cast<const float *>(smraw + M.wo);
    const float *boS = reinterpret_cast<const float *>(smraw + M.bo);
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
            store_split8(tmem_row + ABASE + 32 * s + 16 * half + 8 * g8, smraw + a_lo_off + s * OP_BYTES, row, 2 * half + g8, a8[s]);
    }
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o)
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) out[o][s] = 0.f;
    for (int l = 0; l < L; ++l) {
        tmem_st_wait();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t whi = s32(smraw + M.w_hi + l * W_BYTES), wlo = s32(smraw + M.w_lo + l * W_BYTES);
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    issue_stream(tmem_base + 32 * s, tmem_base + ABASE + 32 * s, s32(smraw + a_lo_off + s * OP_BYTES), whi, wlo);
                mma_commit(mbar);
            }
            __syncwarp();
        }
        mbar_wait(mbar, phase);
        phase ^= 1;
        tc_fence_after();
        const bool last = (l == L - 1);
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
            if (last) {
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < O; ++o)
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < S; ++s)
                        INSR_PRAGMA_UNROLL
                        for (int i = 0; i < 8; ++i) out[o][s] = fmaf(woS[o * HP + 16 * half + 8 * g8 + i], a8[s][i], out[o][s]);
            } else {
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < S; ++s)
                    store_split8(tmem_row + ABASE + 32 * s + 16 * half + 8 * g8, smraw + a_lo_off + s * OP_BYTES, row, 2 * half + g8, a8[s]);
            }
        }
    }
    // ---- both halves get the complete outputs
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o)
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) partS[(half * TILE_M + row) * 16 + o * S + s] = out[o][s];
    tc_fence_before();
    __syncthreads();
    INSR_PRAGMA_UNROLL
    for (int o = 0; o < O; ++o) {
        INSR_PRAGMA_UNROLL
        for (int s = 0; s < S; ++s) out[o][s] = partS[row * 16 + o * S + s] + partS[(TILE_M + row) * 16 + o * S + s];
        out[o][0] += boS[o];
    }
    __syncthreads();                                     // partS is rewritten by the next evaluation
    tc_fence_after();
}

// MODE 0: one evaluation (net A at x).  MODE 1: backtrace -- net A (value only) at x, then net A again at
// clamp(x - dt y_A, lo, hi); coefB applies to the second evaluation.  MODE 2: net A and net B, both at x.
template <int D, int OA, int ORDA, int OB, int ORDB, int MODE>
__global__ void __launch_bounds__(THREADS, 1) k_tc_target(TargetParams p, int tmem_cols) {
    constexpr int SA = StreamCfg<D, ORDA>::S, SB = StreamCfg<D, ORDB>::S;
    constexpr int SMAX = (MODE == 0) ? SA : (SA > SB ? SA : SB);
    static_assert(OA * SA <= 8 && OB * SB <= 8, "partial buffer holds 2 x 8 values per point");
    static_assert(MODE != 1 || (ORDA == 0 && OA == D && OB == OA), "backtrace: a D -> D field evaluated for its value");
    extern __shared__ __align__(1024) unsigned char smraw[];
    const TargetSmem M = target_smem(p.dmA.L, MODE == 2 ? p.dmB.L : 0, SMAX);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;
    const int half = warp >> 2;
    float *partS = reinterpret_cast<float *>(smraw + M.part);
    const uint32_t mbar = s32(smraw + M.mbar);

    stage_net(p.dmA, p.thetaA, smraw, M.a);
    if (MODE == 2) stage_net(p.dmB, p.thetaB, smraw, M.b);
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
        float t[2] = {0.f, 0.f};
        float outA[OA][SA];
        eval_tile<D, OA, ORDA>(smraw, M.a, M.a_lo, partS, tmem_base, tmem_row, mbar, phase, p.dmA.L, row, half, warp, xv, outA);
        for (int c = 0; c < p.n_res; ++c)
            INSR_PRAGMA_UNROLL
            for (int o = 0; o < OA; ++o)
                INSR_PRAGMA_UNROLL
                for (int s = 0; s < SA; ++s) t[c] = fmaf(p.coefA[(c * OA + o) * SA + s], outA[o][s], t[c]);
        if constexpr (MODE != 0) {
            float xb[D];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < D; ++d) xb[d] = xv[d];
            if constexpr (MODE == 1) {
                INSR_PRAGMA_UNROLL
                for (int d = 0; d < D; ++d) xb[d] = fminf(fmaxf(xv[d] - p.dt * outA[d < OA ? d : 0][0], p.lo), p.hi);
            }
            float outB[OB][SB];
            eval_tile<D, OB, ORDB>(smraw, MODE == 1 ? M.a : M.b, M.a_lo, partS, tmem_base, tmem_row, mbar, phase,
                                   MODE == 1 ? p.dmA.L : p.dmB.L, row, half, warp, xb, outB);
            for (int c = 0; c < p.n_res; ++c)
                INSR_PRAGMA_UNROLL
                for (int o = 0; o < OB; ++o)
                    INSR_PRAGMA_UNROLL
                    for (int s = 0; s < SB; ++s) t[c] = fmaf(p.coefB[(c * OB + o) * SB + s], outB[o][s], t[c]);
        }
        if (half == 0 && valid)
            for (int c = 0; c < p.n_res; ++c) p.target[n * p.n_res + c] = t[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

template <int D, int OA, int ORDA, int OB, int ORDB, int MODE>
int launch_tc_target(TargetParams &p, void *stream, int64_t *launches) {
    constexpr int SA = StreamCfg<D, ORDA>::S, SB = StreamCfg<D, ORDB>::S;
    constexpr int SMAX = (MODE == 0) ? SA : (SA > SB ? SA : SB);
    const TargetSmem M = target_smem(p.dmA.L, MODE == 2 ? p.dmB.L : 0, SMAX);
    auto kfn = k_tc_target<D, OA, ORDA, OB, ORDB, MODE>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, M.total);
    const int64_t tiles = (p.N + TILE_M - 1) / TILE_M;
    const int64_t slots = 2 * (int64_t)insr_fused::sm_count();
    const int64_t ctas = tiles < slots ? tiles : slots;
    kfn<<<dim3((unsigned)ctas), dim3(THREADS), M.total, reinterpret_cast<cudaStream_t>(stream)>>>(p, pow2_cols(64 * SMAX));
    ++*launches;
    return 0;
}

}  // namespace insr_tc
#endif  // !INSR_CPU_EMU
