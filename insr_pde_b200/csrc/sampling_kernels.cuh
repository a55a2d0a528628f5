// sampling_kernels.cuh -- all collocation-point sets of one training iteration in ONE kernel.
//
// Reference semantics restated (base/sampling.py:14-18, 45-64): every set is i.i.d. uniform in an axis-aligned box --
// the interior U[-1,1]^D and the epsilon bands [-1-eps,-1+eps] x [-1,1] etc. -- which the reference draws with
// torch.rand + scale + shift + cat (about 40 tiny kernels for the three sets of a fluid iteration).  Here: one thread
// per point, counter-based Philox4x32-10 keyed by (seed; point index, iteration), so that a CUDA-graph replay draws
// fresh points every iteration (the iteration counter lives in device memory and is bumped by the last CTA to finish)
// and a data-parallel rank can draw exactly its shard of a global set (point_offset).
// The distributions are the reference's; the random STREAM is not torch's (parity tests that replay the reference's
// recorded samples keep using insr_pde_b200.sampling's torch restatement).
#pragma once
#include "insr_platform.h"

#define INSR_MAX_BOXES 8

struct insr_box_set {
    int n_boxes;
    int dim;                              // 1..3, same for every box of a call
    int count[INSR_MAX_BOXES];            // points per box
    float lo[INSR_MAX_BOXES][3], hi[INSR_MAX_BOXES][3];
};

__device__ __forceinline__ void insr_philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    INSR_PRAGMA_UNROLL
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// out: boxes concatenated, box b = count[b] points of `dim` floats.  counter (nullable): device iteration counter,
// read by every thread at entry and incremented by the last CTA to finish (ticket must be a zeroed device word).
__global__ void k_sample_boxes(insr_box_set bs, uint64_t seed, int64_t *counter, unsigned int *ticket,
                               int64_t point_offset, float *__restrict__ out) {
    const uint64_t iter = counter ? (uint64_t)*counter : 0ull;
    int64_t total = 0;
    for (int b = 0; b < bs.n_boxes; ++b) total += bs.count[b];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int b = 0;
        int64_t j = i;
        while (b < bs.n_boxes - 1 && j >= bs.count[b]) { j -= bs.count[b]; ++b; }
        const uint64_t pid = (uint64_t)(i + point_offset);
        uint32_t c[4] = {(uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)iter, (uint32_t)(iter >> 32)};
        insr_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int d = 0; d < bs.dim; ++d) {
            const float u = (float)(c[d] >> 8) * (1.0f / 16777216.0f);            // [0, 1), 24 bits like torch.rand
            out[i * bs.dim + d] = fmaf(u, bs.hi[b][d] - bs.lo[b][d], bs.lo[b][d]);
        }
    }
    if (counter) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            if (t == gridDim.x - 1) { *counter = (int64_t)(iter + 1); *ticket = 0u; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Points on a triangle / tetrahedron mesh (elasticity/sampling.py:4-9 -> torchgp/sample_surface.py:28-52,
// torchgp/sample_volume.py:9-43).  The reference picks an element from a Categorical over area / volume
// (random_face.py / random_tet.py), then draws barycentric weights: triangles (1 - sqrt(u), sqrt(u)(1 - v), sqrt(u) v),
// tetrahedra Dirichlet(1,1,1,1) via numpy on the HOST followed by a copy to the device.  Here: one thread per point,
// element by binary search in the inclusive cumulative distribution `cdf` (n_elem floats, last entry ~ 1), Dirichlet
// weights as normalised exponentials -log(u_i); same Philox keying as k_sample_boxes.  Only the first `dim_out`
// coordinates are written (the reference slices [:, 0:dim]).
template <int K>
__global__ void k_sample_mesh(const float *__restrict__ V, const int32_t *__restrict__ elem, const float *__restrict__ cdf,
                              int n_elem, int64_t n, int dim_out, uint64_t seed, int64_t *counter, unsigned int *ticket,
                              int64_t point_offset, float *__restrict__ out) {
    const uint64_t iter = counter ? (uint64_t)*counter : 0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t pid = (uint64_t)(i + point_offset);
        uint32_t c[4] = {(uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)iter, (uint32_t)(iter >> 32)};
        insr_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        // element: first e with cdf[e] > u  (u in [0, 1) scaled to the table's total, so a tail of rounding is harmless)
        const float ue = (float)(c[0] >> 8) * (1.0f / 16777216.0f) * cdf[n_elem - 1];
        int lo = 0, hi = n_elem - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] > ue) hi = mid; else lo = mid + 1;
        }
        float w[K];
        if (K == 3) {
            const float su = sqrtf((float)(c[1] >> 8) * (1.0f / 16777216.0f)), v = (float)(c[2] >> 8) * (1.0f / 16777216.0f);
            w[0] = 1.0f - su; w[1] = su * (1.0f - v); w[2] = su * v;
        } else {
            // a second Philox block for the fourth uniform (counter word 3 flipped keeps the streams disjoint)
            uint32_t c2[4] = {(uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)iter, ~(uint32_t)(iter >> 32)};
            insr_philox4x32_10(c2, (uint32_t)seed, (uint32_t)(seed >> 32));
            float tot = 0.0f;
            INSR_PRAGMA_UNROLL
            for (int k = 0; k < K; ++k) {
                const uint32_t r = k < 3 ? c[k + 1] : c2[0];
                w[k] = -__logf(((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f));      // (0, 1): no log(0)
                tot += w[k];
            }
            const float inv = 1.0f / tot;
            INSR_PRAGMA_UNROLL
            for (int k = 0; k < K; ++k) w[k] *= inv;
        }
        float p[3] = {0.0f, 0.0f, 0.0f};
        INSR_PRAGMA_UNROLL
        for (int k = 0; k < K; ++k) {
            const int64_t vi = elem[(int64_t)lo * K + k];
            INSR_PRAGMA_UNROLL
            for (int d = 0; d < 3; ++d) p[d] = fmaf(w[k], V[vi * 3 + d], p[d]);
        }
        for (int d = 0; d < dim_out; ++d) out[i * dim_out + d] = p[d];
    }
    if (counter) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            if (t == gridDim.x - 1) { *counter = (int64_t)(iter + 1); *ticket = 0u; }
        }
    }
}
