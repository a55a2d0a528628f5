// siren_mid_api.h -- non-template entry points of the fused mid-width kernels (siren_mid_tc.cuh), compiled in their own
// translation unit (siren_mid_inst.cu); called by the 32 < H <= 512 family's drivers (siren_tiled.cuh)
#pragma once
#include "siren_common.cuh"

// 1 if the fused kernels serve this shape (32 < H <= 80, at most 3 streams, 1 <= L <= 8; INSR_MID=0 switches them off)
bool insr_mid_supported(const SirenDims &dm, int order);
// kernel width (64 or 80): a layer buffer of the tape holds S * rows * width floats, rows a multiple of 128
int insr_mid_width(const SirenDims &dm);

// whole network in one kernel.  Zpre == nullptr: plain evaluation (no workspace).  Otherwise the pre-activations and
// activations of every sine layer are left in Zpre / Act = [layer][buf floats] (private layout), x / y / jac / h2
// pointing at the chunk's first point; y == nullptr: tape only.
int insr_mid_forward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N, float *y, float *jac,
                     float *h2, float *Zpre, float *Act, size_t buf, void *stream, int64_t *launches);
// reverse sweep from the tape of insr_mid_forward (which it overwrites): gtheta += d loss / d theta; gx (nullable) written.
// Three kernels: data-gradient chain, hidden-layer weight gradients, first / output layer gradients.
int insr_mid_backward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N, const float *gy,
                      const float *gjac, const float *gh2, float *gtheta, float *gx, float *Zpre, float *Act, size_t buf,
                      void *stream, int64_t *launches);
