// peer_kernels.cuh -- the data-parallel exchange of one training iteration over NVLink / NVSwitch PEER MEMORY
// (points sharded, weights replicated, the flat parameter gradient + the loss scalars reduced once per
// iteration, base/baseModel.py:73-81 is where the reduced gradient meets the optimiser).
//
// The payload is 3.6-56 KB for the script configurations: latency, not bandwidth.  NCCL's all-reduce of such a buffer
// costs 20-40 us inside an iteration graph whose single-GPU length is 50 us.  Here every rank owns one IPC-exported
// allocation  [header | flat gradients of all nets | loss slots]  that all ranks of the box map (cudaIpc*), and the
// reduction is ONE-SHOT: after a flag barrier every rank reads the W buffers directly (P2P loads through the switch)
// and adds them in rank order -- so all replicas compute bit-identical sums -- and a second barrier releases the
// buffers for the next iteration.  Two kernels share the machinery:
//   k_peer_allreduce          out[i] = scale * sum_r buf_r[i]                           (bench.py's step, eager loops)
//   k_iteration_update_peer   the reduction FUSED into the tail of the iteration (insr_iteration_update): Adam of every net
//                             on the reduced gradient, zero_grad, ReduceLROnPlateau on the reduced main loss, loss log --
//                             the reduced gradient is never written anywhere.
// Barrier: flags[cta][peer] words in every rank's header, written remotely with st.release.sys and polled locally with
// ld.acquire.sys; values are a monotonically increasing epoch kept in the owner's header (no reset, so a rank that runs
// ahead cannot be confused with the previous barrier).  CTA k of rank r synchronises with CTA k of every other rank; all
// ranks launch the same grid.  A poll that sees nothing for INSR_PEER_TIMEOUT_NS sets the header's status word and gives
// up (the result is then garbage and the host must look at the status word: insr_peer_status) instead of hanging the GPU.
#pragma once
#include "insr_platform.h"
#include "optim_kernels.cuh"
#ifdef INSR_CPU_EMU
#include <atomic>
#include <chrono>
#endif

#define INSR_PEER_MAX_WORLD 16
#define INSR_PEER_MAX_CTAS 32
#define INSR_PEER_HEADER_BYTES 4096            /* flags 32 x 16 x 4 B | epoch | status | ticket | pad */
#define INSR_PEER_EPOCH_WORD (INSR_PEER_MAX_CTAS * INSR_PEER_MAX_WORLD)
#define INSR_PEER_STATUS_WORD (INSR_PEER_EPOCH_WORD + 1)
#define INSR_PEER_TICKET_WORD (INSR_PEER_EPOCH_WORD + 2)
#define INSR_PEER_TIMEOUT_NS 60000000000ull    /* 60 s */

struct insr_peer_set {
    int world, rank;
    unsigned char *base[INSR_PEER_MAX_WORLD];  // this process's mappings of every rank's allocation (base[rank] = own)
};

namespace insr_peer {

#ifdef INSR_CPU_EMU
// host emulation (tests/emu): the ranks are host threads of one process that launch concurrently, "peer memory" is ordinary memory
inline void st_release_sys(uint32_t *p, uint32_t v) { reinterpret_cast<std::atomic<uint32_t> *>(p)->store(v, std::memory_order_release); }
inline uint32_t ld_acquire_sys(const uint32_t *p) {
    return reinterpret_cast<const std::atomic<uint32_t> *>(p)->load(std::memory_order_acquire);
}
inline float ld_peer(const float *p) { return *reinterpret_cast<const volatile float *>(p); }
inline float4 ld_peer4(const float *p) { return make_float4(ld_peer(p), ld_peer(p + 1), ld_peer(p + 2), ld_peer(p + 3)); }
inline unsigned long long now_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#else
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif
__device__ __forceinline__ uint32_t *header(const insr_peer_set &ps, int r) { return reinterpret_cast<uint32_t *>(ps.base[r]); }

// CTA blockIdx.x of this rank meets CTA blockIdx.x of every other rank at epoch value e
__device__ __forceinline__ void barrier(const insr_peer_set &ps, uint32_t e) {
    __syncthreads();                                   // every thread's earlier peer loads / local stores are performed
    const int t = threadIdx.x;
    if (t < ps.world && t != ps.rank) {
        __threadfence_system();
        st_release_sys(header(ps, t) + blockIdx.x * INSR_PEER_MAX_WORLD + ps.rank, e);
        const uint32_t *mine = header(ps, ps.rank) + blockIdx.x * INSR_PEER_MAX_WORLD + t;
        const unsigned long long t0 = now_ns();
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            if (now_ns() - t0 > INSR_PEER_TIMEOUT_NS) {
                atomicExch(header(ps, ps.rank) + INSR_PEER_STATUS_WORD, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

// the last CTA of the grid to get here publishes the next epoch (every CTA has read the current one long before)
__device__ __forceinline__ bool finish(const insr_peer_set &ps, uint32_t e) {
    bool last = false;
    if (threadIdx.x == 0) {
        uint32_t *h = header(ps, ps.rank);
        __threadfence();
        const uint32_t tk = atomicAdd(h + INSR_PEER_TICKET_WORD, 1u);
        if (tk == gridDim.x - 1) {
            h[INSR_PEER_TICKET_WORD] = 0u;
            h[INSR_PEER_EPOCH_WORD] = e + 2u;
            last = true;
        }
    }
    return last;                                      // meaningful in thread 0 only
}

// out[i] = scale * sum_r buf_r[off + i], i < n; `off` (floats from the allocation's base) is 16-byte aligned
__global__ void __launch_bounds__(512) k_peer_allreduce(insr_peer_set ps, int64_t off, int64_t n, float scale, float *__restrict__ out) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t *>(header(ps, ps.rank) + INSR_PEER_EPOCH_WORD);
    barrier(ps, e);                                   // every rank's buffer is final (its producers precede this kernel in stream order)
    const int64_t n4 = n >> 2;
    const bool out16 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < ps.world; ++r) {
            const float4 v = ld_peer4(reinterpret_cast<const float *>(ps.base[r]) + off + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        if (out16) *reinterpret_cast<float4 *>(out + 4 * i) = acc;
        else { out[4 * i] = acc.x; out[4 * i + 1] = acc.y; out[4 * i + 2] = acc.z; out[4 * i + 3] = acc.w; }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = 4 * n4 + threadIdx.x;
        float acc = 0.f;
        for (int r = 0; r < ps.world; ++r) acc += ld_peer(reinterpret_cast<const float *>(ps.base[r]) + off + i);
        out[i] = acc * scale;
    }
    barrier(ps, e + 1u);                              // every rank has finished reading: the buffers may be rewritten
    finish(ps, e);
}

// insr_iteration_update with the cross-GPU reduction folded in.  The gradient slots and the loss slots live in this
// rank's peer allocation at the same offsets as on every other rank.
__global__ void __launch_bounds__(512) k_iteration_update_peer(insr_peer_set ps, insr_opt_slots sl, float scale, float *__restrict__ sched,
                                                               float *__restrict__ losses, int n_losses, int main_index,
                                                               float *__restrict__ losses_red, float *__restrict__ hist,
                                                               int64_t hist_capacity, int64_t *hist_idx, float beta1, float beta2,
                                                               float eps, float factor, int patience, float threshold, float min_lr,
                                                               float eps_lr, int zero_grad, int clear_losses) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t *>(header(ps, ps.rank) + INSR_PEER_EPOCH_WORD);
    const float lr = sched[0];
    const float t = sched[3] + 1.f;
    const insr_adam_consts c = insr_adam_prepare(lr, t, beta1, beta2);
    int64_t total = 0;
    for (int k = 0; k < sl.n_slots; ++k) total += sl.n[k];
    barrier(ps, e);
    const unsigned char *own = ps.base[ps.rank];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        int64_t j = i;
        while (k < sl.n_slots - 1 && j >= sl.n[k]) { j -= sl.n[k]; ++k; }
        const int64_t byte_off = reinterpret_cast<const unsigned char *>(sl.grad[k] + j) - own;
        float g = 0.f;
        for (int r = 0; r < ps.world; ++r) g += ld_peer(reinterpret_cast<const float *>(ps.base[r] + byte_off));
        g *= scale;
        float mi = sl.m[k][j], vi = sl.v[k][j];
        sl.theta[k][j] = insr_adam_element(sl.theta[k][j], g, mi, vi, beta1, beta2, eps, c);
        sl.m[k][j] = mi;
        sl.v[k][j] = vi;
    }
    if (blockIdx.x == 0 && threadIdx.x < n_losses) {
        const int64_t byte_off = reinterpret_cast<const unsigned char *>(losses + threadIdx.x) - own;
        float v = 0.f;
        for (int r = 0; r < ps.world; ++r) v += ld_peer(reinterpret_cast<const float *>(ps.base[r] + byte_off));
        losses_red[threadIdx.x] = v * scale;
    }
    barrier(ps, e + 1u);                              // nobody reads this rank's gradients / losses any more
    if (zero_grad) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            int k = 0;
            int64_t j = i;
            while (k < sl.n_slots - 1 && j >= sl.n[k]) { j -= sl.n[k]; ++k; }
            sl.grad[k][j] = 0.f;
        }
    }
    if (clear_losses && blockIdx.x == 0 && threadIdx.x < n_losses) losses[threadIdx.x] = 0.f;
    __syncthreads();
    if (finish(ps, e)) {                              // thread 0 of the last CTA: schedule + log on the REDUCED losses
        __threadfence();
        const float cur = *reinterpret_cast<volatile float *>(losses_red + main_index);
        float nlr = lr, best = sched[1], bad = sched[2];
        if (cur < __fmul_rn(best, __fsub_rn(1.f, threshold))) { best = cur; bad = 0.f; }
        else bad += 1.f;
        if (bad > (float)patience) {
            const float nl = fmaxf(__fmul_rn(nlr, factor), min_lr);
            if (__fsub_rn(nlr, nl) > eps_lr) nlr = nl;
            bad = 0.f;
        }
        sched[0] = nlr; sched[1] = best; sched[2] = bad; sched[3] = t;
        if (hist && hist_idx) {
            const int64_t idx = *hist_idx;
            if (idx < hist_capacity)
                for (int q = 0; q < n_losses; ++q) hist[idx * n_losses + q] = *reinterpret_cast<volatile float *>(losses_red + q);
            *hist_idx = idx + 1;
        }
    }
}

}  // namespace insr_peer
