// optim_kernels.cuh -- the optimiser side of one training iteration, kept on the device so that
// a whole iteration (sampling, fused loss closures, update, LR schedule) can be replayed as ONE
// CUDA graph with no host synchronisation.
//
// Reference semantics restated: torch.optim.Adam (amsgrad=False, weight_decay=0) and
// torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown=0) as the
// reference builds them in base/baseModel.py:55-62 and steps them in :73-81.
#pragma once
#include "insr_platform.h"

// sched[0] = lr, sched[1] = best, sched[2] = num_bad_epochs, sched[3] = step count (Adam's t)
__global__ void k_adam_step(float *__restrict__ theta, const float *__restrict__ grad, float *__restrict__ m,
                            float *__restrict__ v, int64_t n, const float *__restrict__ sched, float beta1,
                            float beta2, float eps) {
    const float lr = sched[0];
    const float t = sched[3] + 1.f;                       // this step's index (the counter is bumped by k_plateau_step)
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i];
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        theta[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// one thread: scheduler.step(loss) + bump the shared step counter
__global__ void k_plateau_step(const float *__restrict__ loss, float *__restrict__ sched, float factor, int patience,
                               float threshold, float min_lr, float eps) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const float cur = loss[0];
    float lr = sched[0], best = sched[1], bad = sched[2];
    if (cur < best * (1.f - threshold)) { best = cur; bad = 0.f; }
    else bad += 1.f;
    if (bad > (float)patience) {
        const float nl = fmaxf(lr * factor, min_lr);
        if (lr - nl > eps) lr = nl;
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
// nodes of the iteration graph.
#define INSR_MAX_OPT_SLOTS 8
struct insr_opt_slots {
    int n_slots;
    float *theta[INSR_MAX_OPT_SLOTS], *grad[INSR_MAX_OPT_SLOTS], *m[INSR_MAX_OPT_SLOTS], *v[INSR_MAX_OPT_SLOTS];
    int64_t n[INSR_MAX_OPT_SLOTS];
};

__global__ void k_iteration_update(insr_opt_slots sl, float *__restrict__ sched, const float *__restrict__ losses, int n_losses,
                                   int main_index, float *__restrict__ hist, int64_t hist_capacity, int64_t *hist_idx,
                                   unsigned int *ticket, float beta1, float beta2, float eps, float factor, int patience,
                                   float threshold, float min_lr, float eps_lr, int zero_grad) {
    const float lr = sched[0];
    const float t = sched[3] + 1.f;
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    const float step_size = lr / bc1;
    int64_t total = 0;
    for (int k = 0; k < sl.n_slots; ++k) total += sl.n[k];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        int64_t j = i;
        while (k < sl.n_slots - 1 && j >= sl.n[k]) { j -= sl.n[k]; ++k; }
        const float g = sl.grad[k][j];
        const float mi = beta1 * sl.m[k][j] + (1.f - beta1) * g;
        const float vi = beta2 * sl.v[k][j] + (1.f - beta2) * g * g;
        sl.m[k][j] = mi;
        sl.v[k][j] = vi;
        sl.theta[k][j] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
        if (zero_grad) sl.grad[k][j] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        if (tk == gridDim.x - 1) {
            const float cur = losses[main_index];
            float nlr = lr, best = sched[1], bad = sched[2];
            if (cur < best * (1.f - threshold)) { best = cur; bad = 0.f; }
            else bad += 1.f;
            if (bad > (float)patience) {
                const float nl = fmaxf(nlr * factor, min_lr);
                if (nlr - nl > eps_lr) nlr = nl;
                bad = 0.f;
            }
            sched[0] = nlr; sched[1] = best; sched[2] = bad; sched[3] = t;
            if (hist && hist_idx) {
                const int64_t idx = *hist_idx;
                if (idx < hist_capacity)
                    for (int q = 0; q < n_losses; ++q) hist[idx * n_losses + q] = losses[q];
                *hist_idx = idx + 1;
            }
            *ticket = 0u;
        }
    }
}
