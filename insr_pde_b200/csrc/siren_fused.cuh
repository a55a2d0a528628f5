// siren_fused.cuh -- placeholder; replaced by the resident-weights fused kernels.
#pragma once
#include "siren_common.cuh"
inline bool insr_fused_supported(const SirenDims &, int, int) { return false; }
inline size_t insr_fused_ws_bytes(const SirenDims &, int64_t, int, int) { return 0; }
inline int insr_fused_forward(const SirenDims &, int, const float *, const float *, int64_t, float *, float *,
                              float *, float *, void *, int64_t *) { return -6; }
inline int insr_fused_backward(const SirenDims &, int, const float *, const float *, int64_t, const float *,
                               const float *, const float *, float *, float *, float *, void *, int64_t *) { return -6; }
inline int insr_fused_lsq_step(const SirenDims &, int, int, const float *, const float *, const float *, int64_t,
                               const float *, float, float *, float *, float *, size_t, void *, int64_t *) { return -6; }
