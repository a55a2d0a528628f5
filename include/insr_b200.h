/*
 * insr_b200.h -- C ABI of the B200-native INSR-PDE hot path (libinsr_b200.so).
 *
 * Plain C: pointers and sizes only, no torch types.  All tensors are FP32, row-major,
 * contiguous DEVICE memory owned by the caller (PyTorch's caching allocator in the
 * shipped host side); the library allocates nothing persistent, keeps no pointers between
 * calls, enqueues all work on the cudaStream_t passed as `stream` and never synchronises.
 *
 * What each entry point replaces in the reference (paths relative to the reference root):
 *
 *   insr_siren_forward   base/networks.py:67-71  MLP.forward  (nn.Sequential of nn.Linear +
 *                        Sine, base/networks.py:21-27,50-60), and -- for order >= 1 -- the
 *                        torch.autograd.grad(create_graph=True) sweeps of
 *                        base/diff_ops.py:53-58 (gradient), :44-50 (divergence),
 *                        :33-41 (laplace), :61-82 (jacobian), :6-30 (hessian).
 *   insr_siren_backward  the autograd backward the reference triggers with
 *                        loss.backward() at base/baseModel.py:77 (first-, second- and
 *                        third-order chains through the graphs built by diff_ops).
 *   insr_siren_lsq_step  the fused loss closures: residual + mean-square loss + backward in
 *                        one pass, covering advection/model.py:43-52,68-91 and
 *                        fluid/model.py:43-52,72-151 (see INTEGRATION.md for the mapping).
 *
 * Parameter vector `theta` (flat, the order of nn.Module.parameters()):
 *   W1 (H x D), b1 (H), W2 (H x H), b2 (H), ..., W_{L+1} (H x H), b_{L+1} (H),
 *   Wout (O x H), bout (O);  each W is (out, in) row-major exactly as nn.Linear stores it.
 *
 * Error behaviour: 0 on success; negative insr_status for argument errors; positive values
 * are cudaError_t codes.  insr_last_error() returns a thread-local message.  There is no
 * CPU fallback and no silent degradation: an unsupported shape is an error.
 */
#ifndef INSR_B200_H
#define INSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INSR_ABI_VERSION 1
#define INSR_MAX_IN 3       /* D: spatial dimension of the collocation points   */
#define INSR_MAX_OUT 3      /* O: field components                              */
#define INSR_MAX_HIDDEN 512 /* H: hidden_features (config.py:99)                */
#define INSR_MAX_LAYERS 16  /* L: num_hidden_layers (config.py:98)              */

typedef enum insr_status {
    INSR_OK = 0,
    INSR_ERR_NULL = -1,        /* a required pointer is NULL                        */
    INSR_ERR_SHAPE = -2,       /* D/O/H/L/N outside the supported range             */
    INSR_ERR_ORDER = -3,       /* unknown derivative order                          */
    INSR_ERR_ALIGN = -4,       /* a buffer is not 16-byte aligned                   */
    INSR_ERR_WORKSPACE = -5,   /* workspace missing or too small                    */
    INSR_ERR_UNSUPPORTED = -6, /* valid request that this build cannot serve        */
    INSR_ERR_NO_DEVICE = -7    /* no CUDA device / wrong architecture               */
} insr_status;

/* which derivative streams are propagated forward-mode through the layers */
typedef enum insr_order {
    INSR_ORDER_VALUE = 0, /* y                                            S = 1           */
    INSR_ORDER_JAC = 1,   /* y, J[n,o,d] = dy_o/dx_d                      S = 1 + D       */
    INSR_ORDER_LAP = 2,   /* y, J, lap[n,o] = sum_d d2y_o/dx_d^2          S = 2 + D       */
    INSR_ORDER_HESS = 3   /* y, J, hess[n,o,d,e]                          S = 1+D+D(D+1)/2 */
} insr_order;

/* one SIREN field: MLP(in, out, num_hidden_layers, hidden, nonlinearity='sine') */
typedef struct insr_siren_desc {
    int32_t in_features;       /* D */
    int32_t out_features;      /* O */
    int32_t hidden_features;   /* H */
    int32_t num_hidden_layers; /* L  (=> L+1 sine layers, L+2 Linear layers) */
    float omega;               /* 30.0f in the reference (base/networks.py:27) */
    int32_t flags;             /* INSR_FLAG_* */
} insr_siren_desc;

#define INSR_FLAG_NONE 0
#define INSR_FLAG_FORCE_GENERIC 1 /* use the generic (any-H) kernels even where a fused one exists */
#define INSR_FLAG_NO_TENSOR 2     /* keep the H <= 32 family on the FP32 FFMA kernels.  Default: forward, backward and
                                     lsq_step run on tcgen05/TMEM (3xTF32 split for the forward pass and the data
                                     gradient, 2-level bf16 split for the weight gradient, FP32 accumulation; same
                                     1e-4 parity; the backward keeps its tape in the workspace) */
#define INSR_FLAG_FFMA_BWD 4      /* tcgen05 forward, FP32 FFMA backward / lsq_step (A/B measurements) */
#define INSR_FLAG_KEEP_TAPE 8     /* 32 < H <= 512 family: insr_siren_forward leaves the activations of every layer in the
                                   * workspace (which must then have the BACKWARD size) and insr_siren_backward on the same,
                                   * untouched workspace skips recomputing them -- what autograd's saved activations are to
                                   * the reference (base/baseModel.py:77).  Only where insr_siren_tape_supported() says 1;
                                   * ignored by the H <= 32 family (its backward kernel recomputes inside registers). */

int insr_version(void);
const char *insr_last_error(void);

/* number of floats in theta for this descriptor (0 if the descriptor is invalid) */
int64_t insr_siren_theta_size(const insr_siren_desc *desc);

/* bytes of scratch the caller must pass to forward (backward=0) / backward (backward=1) */
size_t insr_siren_workspace_bytes(const insr_siren_desc *desc, int64_t n_points, int order,
                                  int backward);

/*
 * Evaluate the field and its spatial derivatives at n_points collocation points.
 *   x     (N, D)            y    (N, O)
 *   jac   (N, O, D)         required for order >= 1, ignored otherwise
 *   h2    (N, O)            per-output Laplacian for order == INSR_ORDER_LAP
 *         (N, O, D, D)      full symmetric Hessian for order == INSR_ORDER_HESS
 */
int insr_siren_forward(const insr_siren_desc *desc, const float *theta, const float *x,
                       int64_t n_points, int order, float *y, float *jac, float *h2,
                       void *workspace, size_t workspace_bytes, void *stream);

/*
 * Reverse sweep: given the cotangents of the outputs of insr_siren_forward (any of them may
 * be NULL = zero), ACCUMULATE (+=) the parameter gradient into gtheta (flat, same layout as
 * theta) and, if gx != NULL, WRITE the gradient w.r.t. the points into gx (N, D).
 * The forward activations are recomputed inside the kernel; nothing is saved between calls.
 */
int insr_siren_backward(const insr_siren_desc *desc, const float *theta, const float *x,
                        int64_t n_points, int order, const float *gy, const float *gjac,
                        const float *gh2, float *gtheta, float *gx, void *workspace,
                        size_t workspace_bytes, void *stream);

/*
 * Fused least-squares step (one pass: forward streams -> residual -> loss partial -> reverse
 * sweep -> parameter gradient; no activation recompute, no output round trip through HBM).
 *
 *   r[n,c] = sum_o ( cy[c,o] * y[n,o] + sum_d cj[c,o,d] * J[n,o,d] + cl[c,o] * lap[n,o] )
 *            - target[n,c]                                   c = 0 .. n_res-1  (<= 4)
 *   loss   = scale * sum_{n,c} r[n,c]^2          (scale = 1/(N*n_res) gives torch.mean)
 *
 * loss_out[0] += loss (device scalar, accumulated), gtheta += d loss / d theta.
 * coef = { cy (n_res x O), cj (n_res x O x D), cl (n_res x O) } packed host floats.
 */
int insr_siren_lsq_step(const insr_siren_desc *desc, const float *theta, const float *x,
                        int64_t n_points, int order, int n_res, const float *coef_host,
                        const float *target, float scale, float *loss_out, float *gtheta,
                        void *workspace, size_t workspace_bytes, void *stream);

/*
 * The frozen-network side of a least-squares closure in one kernel: the target that insr_siren_lsq_step compares the
 * trainable field with.  One or two frozen fields of the H <= 32 family are evaluated at the points x and a fixed
 * linear combination of their outputs is written to target (n, n_res), n_res <= 2:
 *     target[n,c] = sum_{o} ( cy_A[c,o] y_A + cj_A[c,o,:] . J_A + cl_A[c,o] lap_A )  +  the same for B
 *   mode 0  field A at x                                fluid/model.py:108-109  div u  (cj = trace pattern)
 *                                                       advection/model.py:78-84  u_prev / dt - vel/2 du_prev/dx
 *   mode 1  backtrace: A (value, D -> D) at x, then A again at clamp(x - dt y_A(x), lo, hi); b's coefficients apply to
 *           the second evaluation (b->desc / theta are ignored)   fluid/model.py:78-87  semi-Lagrangian advection
 *   mode 2  A and B, both at x                          fluid/model.py:131-137  u_prev - grad p
 * coef_* are HOST arrays (nullable = zeros): cy (n_res x O), cj (n_res x O x D), cl (n_res x O).
 * Returns INSR_ERR_UNSUPPORTED for shapes outside the resident-weights family or combinations not instantiated.
 */
typedef struct insr_target_eval {
    insr_siren_desc desc;
    const float *theta;      /* device */
    int32_t order;           /* INSR_ORDER_VALUE .. INSR_ORDER_JAC */
    const float *coef_y, *coef_jac, *coef_lap;
} insr_target_eval;
int insr_siren_target(const insr_target_eval *a, const insr_target_eval *b, int mode, float dt, float lo, float hi,
                      const float *x, int64_t n_points, int n_res, float *target, void *stream);

/*
 * Optimiser side of one iteration, on the device (so a whole iteration can be one CUDA graph):
 * `sched` is 4 device floats {lr, best, num_bad_epochs, step}.
 *   insr_adam_step     torch.optim.Adam(amsgrad=False, weight_decay=0) on a flat vector, reading lr and the
 *                      step index from `sched`                           (base/baseModel.py:60, :79)
 *   insr_plateau_step  ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown=0).step(loss) updating
 *                      sched[0..2], and sched[3] += 1                    (base/baseModel.py:61-62, :81)
 * Call order per iteration: insr_adam_step for every net, then insr_plateau_step once.
 */
int insr_adam_step(float *theta, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                   const float *sched, float beta1, float beta2, float eps, void *stream);
int insr_plateau_step(const float *loss, float *sched, float factor, int patience, float threshold,
                      float min_lr, float eps, void *stream);

/*
 * The tail of one training iteration in one kernel: Adam for every trainable net (n_slots <= 8 flat vectors; host arrays of
 * device pointers), the gradient buffers zeroed for the next iteration (optimizer.zero_grad, base/baseModel.py:76), the
 * ReduceLROnPlateau step on losses[main_index] and the shared step counter (:79-81), and the iteration's n_losses loss
 * values appended to the device log hist[hist_idx++] (the .item() logging of :116-118 without the host sync; hist may be
 * NULL).  Same arithmetic as insr_adam_step + insr_plateau_step.  ticket: a zeroed device word (last-CTA detection).
 * clear_losses != 0: the loss slots -- accumulators of the closures' kernels -- are zeroed once they have been consumed.
 */
int insr_iteration_update(int n_slots, float *const *theta, float *const *grad, float *const *exp_avg,
                          float *const *exp_avg_sq, const int64_t *n, float *sched, float *losses, int n_losses,
                          int main_index, float *hist, int64_t hist_capacity, int64_t *hist_idx, uint32_t *ticket, float beta1,
                          float beta2, float eps, float factor, int patience, float threshold, float min_lr, float eps_lr,
                          int zero_grad, int clear_losses, void *stream);

/*
 * Data parallelism over the GPUs of one box without a communication library on the path (the reference
 * itself is single-GPU: base/baseModel.py:25 hard-codes cuda:0, and :73-81 is where a reduced gradient must meet the
 * optimiser).  PEER MEMORY is the one exception to "the library allocates nothing": a buffer that other processes map
 * must come from a plain cudaMalloc (cudaIpcGetMemHandle), which a caching allocator does not guarantee.
 *
 *   insr_peer_alloc   one allocation  [4096-byte header | data_bytes]  on the current device, zeroed; *base = its address,
 *                     handle64 = the 64-byte cudaIpcMemHandle_t to hand to the other ranks (any host-side channel)
 *   insr_peer_open    map another rank's allocation into this process (cudaIpcOpenMemHandle, lazy peer access)
 *   insr_peer_close / insr_peer_free   unmap a peer's / release the own allocation (after all ranks have stopped using it)
 *   insr_peer_status  1 if a barrier of a kernel below gave up after 60 s without its peers (results are then invalid),
 *                     0 otherwise; synchronises the device; reset != 0 clears the flag
 *
 * Layout contract: every rank allocates the same size and keeps the same quantities at the same offsets.  bases[r] is THIS
 * process's mapping of rank r's allocation (bases[rank] = the own one).  Every rank must issue the same sequence of the two
 * calls below on the same allocation set, with the same sizes (they contain a flag barrier across the ranks).
 *
 *   insr_peer_allreduce         out[i] = scale * sum_r data_r[offset_floats + i], i < n, summed in rank order on every rank
 *                               (bit-identical results); out is ordinary device memory, NOT part of a peer allocation;
 *                               offset_floats counts from the allocation's base (>= 1024: behind the header), multiple of 4
 *   insr_iteration_update_peer  insr_iteration_update with that reduction folded in: grad[k] and losses point INTO the own
 *                               peer allocation (peer_bytes = its total size); the kernel reads every rank's gradients and
 *                               loss slots, applies `scale`, runs Adam / the plateau schedule / the loss log on the reduced
 *                               values (losses_red: n_losses floats of ordinary device memory) and leaves the own
 *                               gradient (zero_grad) and loss slots (clear_losses) zeroed once no peer reads them any more
 */
int insr_peer_alloc(int64_t data_bytes, void **base, unsigned char *handle64);
int insr_peer_open(const unsigned char *handle64, void **base);
int insr_peer_close(void *base);
int insr_peer_free(void *base);
int insr_peer_status(void *base, int reset);
int insr_peer_allreduce(int world, int rank, void *const *bases, int64_t offset_floats, int64_t n, float scale, float *out,
                        void *stream);
int insr_iteration_update_peer(int world, int rank, void *const *bases, int64_t peer_bytes, float scale, int n_slots,
                               float *const *theta, float *const *grad, float *const *exp_avg, float *const *exp_avg_sq,
                               const int64_t *n, float *sched, float *losses, int n_losses, int main_index, float *losses_red,
                               float *hist, int64_t hist_capacity, int64_t *hist_idx, float beta1, float beta2, float eps,
                               float factor, int patience, float threshold, float min_lr, float eps_lr, int zero_grad,
                               int clear_losses, void *stream);

/*
 * Batched 2x2 / 3x3 singular value decomposition and the fused elasticity energy.
 * Replaces, in the elasticity closure (elasticity/model.py:143-149):
 *     U_x, S_x, V_x = torch.svd(jac_x)                                  -> insr_svd_small
 *     E_arap = r_a * sum((S_x - 1)^2); E_volume = r_v * sum((prod(S_x, 1) - 1)^2)  and their adjoint d E / d jac_x
 *                                                                       -> insr_elastic_energy
 *   F (n, d, d) row-major, d in {2, 3};  S (n, d) descending, >= 0;  U, V (n, d, d) with F = U diag(S) V^T (nullable)
 *   energy: device scalar, ACCUMULATED (+=);  gF (n, d, d) = d energy / d F  (nullable: energy only)
 */
int insr_svd_small(const float *F, int64_t n, int d, float *U, float *S, float *V, void *stream);
int insr_elastic_energy(const float *F, int64_t n, int d, float ratio_arap, float ratio_volume, float *energy,
                        float *gF, void *stream);

/*
 * The whole loss of ElasticityModel._solve_deformation (elasticity/model.py:127-189 with elasticity/losses.py:6-39) and
 * its cotangents w.r.t. the trainable field's outputs, in one kernel.  Rows [0, n) are the interior samples (value y,
 * Jacobian J of the field, points x, the two previous-frame fields' values); rows [n, n + n_left) the clamped left face
 * (ratio_constraint |y|^2); rows [n + n_left, n + n_left + n_right) the right face (ratio_constraint |y - offset_right|^2,
 * offset carrying the sign of 'constraint_right' / 'constraint_right_compress').  A ratio of zero switches a term off;
 * external_force must be zeroed by the caller once timestep > external_force_timesteps (model.py:155).
 * The sphere term is the 2-D form of losses.py:22-39 (force = ratio * (q - center)); the reference's 3-D form broadcasts
 * (M,1,1) * (M,3) into an (M,M,3) product and is not offered here.
 * loss: device scalar, ACCUMULATED.  gy (n_all, d), gJ (n_all, d, d; may be NULL when r_arap = r_volume = 0): overwritten.
 */
typedef struct insr_elastic_terms_desc {
    int64_t n, n_left, n_right;
    float dt;
    float r_arap, r_volume, r_kinematics;
    float r_left, r_right;
    float r_plane, plane_height;
    float r_sphere, radius;
    float external_force[3], offset_right[3], center[3];
} insr_elastic_terms_desc;
int insr_elastic_terms(const insr_elastic_terms_desc *t, int d, const float *y, const float *J, const float *x,
                       const float *y_prev, const float *y_pp, float *loss, float *gy, float *gJ, void *stream);

/*
 * All collocation-point sets of one iteration in one kernel.  Replaces the torch.rand / scale / shift / cat sequences
 * of base/sampling.py:14-18 (sample_random) and :21-64 (sample_boundary, sample_boundary2D_separate): every set is
 * i.i.d. uniform in an axis-aligned box.  Box b holds count[b] points in [lo[b], hi[b])^dim (lo, hi: n_boxes x dim,
 * host arrays); `out` receives the boxes back to back.  Counter-based Philox4x32-10 keyed by (seed; point index +
 * point_offset, iteration): `counter` (nullable) is a DEVICE iteration counter read at entry and incremented by the
 * last CTA to finish (`ticket`: a zeroed device word), so a CUDA-graph replay draws fresh points every time;
 * point_offset lets a data-parallel rank draw exactly its shard of a global set.
 * Same distributions as the reference; the random stream is Philox's, not torch's generator.
 */
int insr_sample_boxes(int n_boxes, int dim, const int32_t *count, const float *lo, const float *hi, uint64_t seed,
                      int64_t *counter, uint32_t *ticket, int64_t point_offset, float *out, void *stream);

/*
 * Points on a triangle (verts_per_elem 3) or tetrahedron (4) mesh.  Replaces elasticity/sampling.py:4-9 ->
 * torchgp/sample_surface.py:28-52 (Categorical over face areas, weights (1 - sqrt(u), sqrt(u)(1 - v), sqrt(u) v)) and
 * torchgp/sample_volume.py:9-43 (Categorical over tet volumes, numpy Dirichlet(1,1,1,1) on the host + copy).
 * V (n_vert, 3) float, elem (n_elem, verts_per_elem) int32, cdf (n_elem) inclusive cumulative area / volume (any
 * positive normalisation; the kernel scales by its last entry) -- all DEVICE arrays; out (n, dim_out) receives the
 * first dim_out coordinates (the reference slices [:, 0:dim]).  seed / counter / ticket / point_offset as above.
 */
int insr_sample_mesh(const float *V, const int32_t *elem, const float *cdf, int n_elem, int verts_per_elem, int64_t n,
                     int dim_out, uint64_t seed, int64_t *counter, uint32_t *ticket, int64_t point_offset, float *out,
                     void *stream);

/* 1 if a forward / backward pair with INSR_FLAG_KEEP_TAPE shares its tape for this shape and batch (the batch must be a
 * single workspace chunk of the tiled family), 0 if the flag would be refused, negative = error */
int insr_siren_tape_supported(const insr_siren_desc *desc, int64_t n_points, int order);

/* introspection used by bench.py / tests: which kernel family a call would dispatch to.
 * returns 0 = generic, 1 = fused resident-weights kernels (H <= 32), 2 = tiled shared-memory GEMM
 * kernels (32 < H <= 512); negative = error. */
int insr_siren_kernel_family(const insr_siren_desc *desc, int order, int backward);

/* number of kernel launches issued by this library on the calling thread since the last
 * call with reset != 0 (bench.py's gpu_launches counter) */
int64_t insr_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* INSR_B200_H */
