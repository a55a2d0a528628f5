// optim_kernels.cuh -- the optimiser side of one training iteration, kept on the device so that
// a whole iteration (sampling, fused loss closures, update, LR schedule) can be replayed as ONE
// CUDA graph with no host synchronisation.
//
// Reference semantics restated: torch.optim.Adam (amsgrad=False, weight_decay=0) and
// torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown=0) as the
// reference builds them in base/baseModel.py:55-62 and steps them in :73-81.
//
// The library is compiled with --use_fast_math (the field kernels want MUFU sin / cos and FTZ); the optimiser must not
// inherit that: its bias corrections are evaluated in double precision (1 - 0.999^t loses 1e-4 relative under __powf
// at t = 1) and every division / square root is the IEEE round-to-nearest intrinsic, so that one step equals
// torch.optim.Adam's to the last bit or two.
#pragma once
#include "insr_platform.h"

struct insr_adam_consts { float step_size, bc2_sqrt; };
// torch/optim/adam.py (_single_tensor_adam): bias_correction{1,2} = 1 - beta^step; step_size = lr / bias_correction1;
// denom = sqrt(v) / sqrt(bias_correction2) + eps -- python floats, i.e. double precision
__device__ __forceinline__ insr_adam_consts insr_adam_prepare(float lr, float t, float beta1, float beta2) {
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    insr_adam_consts c;
    c.step_size = (float)((double)lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    return c;
}
// one element: m.lerp_(g, 1 - beta1); v.mul_(beta2).addcmul_(g, g, 1 - beta2); theta.addcdiv_(m, denom, -step_size)
__device__ __forceinline__ float insr_adam_element(float theta, float g, float &m, float &v, float beta1, float beta2, float eps,
                                                   const insr_adam_consts &c) {
    m = __fmaf_rn(__fsub_rn(g, m), __fsub_rn(1.f, beta1), m);
    v = __fmaf_rn(__fmul_rn(g, g), __fsub_rn(1.f, beta2), __fmul_rn(v, beta2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), eps);
    return __fsub_rn(theta, __fmul_rn(c.step_size, __fdiv_rn(m, denom)));
}

// sched[0] = lr, sched[1] = best, sched[2] = num_bad_epochs, sched[3] = step count (Adam's t)
__global__ void k_adam_step(float *__restrict__ theta, const float *__restrict__ grad, float *__restrict__ m,
                            float *__restrict__ v, int64_t n, const float *__restrict__ sched, float beta1,
                            float beta2, float eps) {
    const float lr = sched[0];
    const float t = sched[3] + 1.f;                       // this step's index (the counter is bumped by k_plateau_step)
    const insr_adam_consts c = insr_adam_prepare(lr, t, beta1, beta2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float mi = m[i], vi = v[i];
        theta[i] = insr_adam_element(theta[i], grad[i], mi, vi, beta1, beta2, eps, c);
        m[i] = mi;
        v[i] = vi;
    }
}

// one thread: scheduler.step(loss) + bump the shared step counter
__global__ void k_plateau_step(const float *__restrict__ loss, float *__restrict__ sched, float factor, int patience,
                               float threshold, float min_lr, float eps) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const float cur = loss[0];
    float lr = sched[0], best = sched[1], bad = sched[2];
    if (cur < __fmul_rn(best, __fsub_rn(1.f, threshold))) { best = cur; bad = 0.f; }
    else bad += 1.f;
    if (bad > (float)patience) {
        const float nl = fmaxf(__fmul_rn(lr, factor), min_lr);
        if (__fsub_rn(lr, nl) > eps) lr = nl;
        bad = 0.f;
    }
    sched[0] = lr; sched[1] = best; sched[2] = bad; sched[3] += 1.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole tail of a training iteration in ONE kernel (base/baseModel.py:73-81 + the loss logging of :116-118):
// Adam for every trainable net, gradient buffers left zeroed for the next iteration (the reference's zero_grad),
// ReduceLROnPlateau on the main loss, the step counter, and the iteration's loss values appended to a device ring.
// Every thread reads lr / t at entry; the LAST CTA to finish (ticket) applies the scheduler and writes the log, so no
// thread can see the updated schedule.  Replaces 2 memsets + 2 Adam launches + plateau + stack / index_copy / counter
// nodes of the iteration graph.  clear_losses: the loss slots (accumulated into by the closures' kernels) are zeroed after
// they have been consumed, so that the iteration graph needs no fill node for them either.
#define INSR_MAX_OPT_SLOTS 8
struct insr_opt_slots {
    int n_slots;
    float *theta[INSR_MAX_OPT_SLOTS], *grad[INSR_MAX_OPT_SLOTS], *m[INSR_MAX_OPT_SLOTS], *v[INSR_MAX_OPT_SLOTS];
    int64_t n[INSR_MAX_OPT_SLOTS];
};

__global__ void k_iteration_update(insr_opt_slots sl, float *__restrict__ sched, float *__restrict__ losses, int n_losses,
                                   int main_index, float *__restrict__ hist, int64_t hist_capacity, int64_t *hist_idx,
                                   unsigned int *ticket, float beta1, float beta2, float eps, float factor, int patience,
                                   float threshold, float min_lr, float eps_lr, int zero_grad, int clear_losses) {
    const float lr = sched[0];
    const float t = sched[3] + 1.f;
    const insr_adam_consts c = insr_adam_prepare(lr, t, beta1, beta2);
    int64_t total = 0;
    for (int k = 0; k < sl.n_slots; ++k) total += sl.n[k];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        int64_t j = i;
        while (k < sl.n_slots - 1 && j >= sl.n[k]) { j -= sl.n[k]; ++k; }
        float mi = sl.m[k][j], vi = sl.v[k][j];
        sl.theta[k][j] = insr_adam_element(sl.theta[k][j], sl.grad[k][j], mi, vi, beta1, beta2, eps, c);
        sl.m[k][j] = mi;
        sl.v[k][j] = vi;
        if (zero_grad) sl.grad[k][j] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        if (tk == gridDim.x - 1) {
            const float cur = losses[main_index];
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
                    for (int q = 0; q < n_losses; ++q) hist[idx * n_losses + q] = losses[q];
                *hist_idx = idx + 1;
            }
            // the loss slots are accumulators of the closures' kernels: left zeroed for the next iteration, like the gradients
            if (clear_losses)
                for (int q = 0; q < n_losses; ++q) losses[q] = 0.f;
            *ticket = 0u;
        }
    }
}
