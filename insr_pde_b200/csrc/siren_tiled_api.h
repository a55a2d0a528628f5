// siren_tiled_api.h -- what the ABI unit needs from the tiled family: workspace sizing and the
// non-template entry points (the kernels themselves are in siren_tiled.cuh, compiled in their own unit)
#pragma once
#include "siren_common.cuh"

namespace insr_tiled {
inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// row width the workspace is sized for: H rounded up to 8, or -- for the widths the fused mid-width kernels serve
// (siren_mid_tc.cuh: 32 < H <= 80) -- the kernel width 64 / 80 of their private tape layout
inline int hp_alloc(const SirenDims &dm) {
    const int hp = (dm.H + 7) & ~7, h16 = (dm.H + 15) & ~15;
    if (h16 <= 80) return h16 <= 64 ? 64 : 80;
    return hp;
}

// points per chunk: bound the workspace to ~1.5 GB
inline int64_t chunk_points(const SirenDims &dm, int S, bool bwd, int64_t N) {
    const int HP = hp_alloc(dm);
    const size_t per_point = (size_t)S * HP * sizeof(float) * (bwd ? (2 * (dm.L + 1) + 2) : 2) + (bwd ? 16 * 4 * 4 : 0);
    int64_t nc = (int64_t)(((size_t)1536 << 20) / per_point);
    nc = nc / 1024 * 1024;
    if (nc < 1024) nc = 1024;
    const int64_t need = round_up(N, 1024);
    return nc < need ? nc : need;
}

// row capacity of the workspace buffers: the chunk plus slack for the last (partial) CTA tile
inline int64_t capacity(int64_t chunk) { return chunk + 256; }

inline size_t ws_bytes(const SirenDims &dm, int order, bool bwd, int64_t N) {
    const int S = insr_nstreams(dm.D, order);
    const int HP = hp_alloc(dm);
    const int64_t NCp = capacity(chunk_points(dm, S, bwd, N));
    const size_t buf = (size_t)S * NCp * HP;
    size_t floats = bwd ? buf * (2 * (dm.L + 1) + 2) + (size_t)dm.O * S * NCp : buf * 2;
    return floats * sizeof(float) + 256;
}

}  // namespace insr_tiled

// non-template entry points, defined in siren_tiled_inst.cuh (own translation unit under nvcc)
// tensor = true: hidden-layer GEMMs (forward + data gradient) on tcgen05 where the shape allows (siren_wide_tc.cuh)
int insr_tiled_forward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N, float *y,
                       float *jac, float *h2, float *ws, void *stream, int64_t *launches, bool tensor = false,
                       bool keep_tape = false);
int insr_tiled_backward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N,
                        const float *gy, const float *gjac, const float *gh2, float *gtheta, float *gx, float *ws,
                        void *stream, int64_t *launches, bool tensor = false, bool have_tape = false);
// keep_tape / have_tape: the forward leaves its tape in the (backward-sized) workspace and the backward call on the same
// workspace skips the recomputation; possible when the batch is a single workspace chunk
inline bool insr_tiled_tape_fits(const SirenDims &dm, int64_t N, int order) {
    return N <= insr_tiled::chunk_points(dm, insr_nstreams(dm.D, order), true, N);
}
#ifndef INSR_CPU_EMU
bool insr_mid_supported(const SirenDims &dm, int order);      // siren_mid_api.h
#endif
// 32 < H <= 512; and those H <= 32 shapes outside the resident-weights family (D = 3, more than 3 hidden layers in the
// backward) that the fused mid-width kernels serve at the padded width 64 instead of the generic thread-per-point kernels
inline bool insr_tiled_supported(const SirenDims &dm, int order) {
    if (order > 3) return false;
    if (dm.H > 32) return dm.H <= 512;
#ifndef INSR_CPU_EMU
    return insr_mid_supported(dm, order);
#else
    return false;
#endif
}
inline size_t insr_tiled_ws_bytes(const SirenDims &dm, int64_t N, int order, int backward) {
    return insr_tiled::ws_bytes(dm, order, backward != 0, N);
}
