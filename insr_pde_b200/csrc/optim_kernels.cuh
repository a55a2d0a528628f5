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
