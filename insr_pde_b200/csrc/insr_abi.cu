// insr_abi.cu -- the extern "C" boundary of libinsr_b200.so (see include/insr_b200.h).
//
// Argument validation, kernel-family dispatch and launch bookkeeping.  No allocation, no
// synchronisation, no state kept between calls except thread-local error text / counters.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "insr_b200.h"
#include "siren_generic.cuh"
#include "siren_fused.cuh"
#include "siren_tiled_api.h"
#include "optim_kernels.cuh"
#include "peer_kernels.cuh"
#include "linalg_kernels.cuh"
#include "sampling_kernels.cuh"
#include "siren_tc_target.cuh"
#ifdef INSR_SINGLE_TU
#include "siren_tiled.cuh"   // emulation build: everything in one translation unit
#define INSR_INST_D 1
#define INSR_INST_O 1
#include "siren_fused_inst.cuh"
#define INSR_INST_D 2
#define INSR_INST_O 1
#include "siren_fused_inst.cuh"
#define INSR_INST_D 2
#define INSR_INST_O 2
#include "siren_fused_inst.cuh"
#include "siren_tiled_inst.cuh"
#endif

namespace {

thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

int validate(const insr_siren_desc *d, int64_t N, int order, SirenDims *dm) {
    if (!d) return fail(INSR_ERR_NULL, "desc is NULL");
    if (d->in_features < 1 || d->in_features > INSR_MAX_IN)
        return fail(INSR_ERR_SHAPE, "in_features=%d outside [1,%d]", d->in_features, INSR_MAX_IN);
    if (d->out_features < 1 || d->out_features > INSR_MAX_OUT)
        return fail(INSR_ERR_SHAPE, "out_features=%d outside [1,%d]", d->out_features, INSR_MAX_OUT);
    if (d->hidden_features < 1 || d->hidden_features > INSR_MAX_HIDDEN)
        return fail(INSR_ERR_SHAPE, "hidden_features=%d outside [1,%d]", d->hidden_features, INSR_MAX_HIDDEN);
    if (d->num_hidden_layers < 0 || d->num_hidden_layers > INSR_MAX_LAYERS)
        return fail(INSR_ERR_SHAPE, "num_hidden_layers=%d outside [0,%d]", d->num_hidden_layers, INSR_MAX_LAYERS);
    if (N < 0 || N > ((int64_t)1 << 31) - 1024)
        return fail(INSR_ERR_SHAPE, "n_points=%lld outside [0, 2^31)", (long long)N);
    if (order < INSR_ORDER_VALUE || order > INSR_ORDER_HESS)
        return fail(INSR_ERR_ORDER, "order=%d is not one of 0,1,2,3", order);
    dm->D = d->in_features; dm->O = d->out_features; dm->H = d->hidden_features;
    dm->L = d->num_hidden_layers; dm->omega = d->omega;
    return 0;
}

int check_device() {
    static thread_local int cached = -1;
    if (cached == 1) return 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(INSR_ERR_NO_DEVICE, "no CUDA device");
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10)
        return fail(INSR_ERR_NO_DEVICE, "libinsr_b200 is built for sm_100a only (device is sm_%d x)", major);
    cached = 1;
    return 0;
}

// streaming multiprocessors of the current device (cudaDevAttrMultiProcessorCount; 148 on the B200)
int n_sm() { return insr_fused::sm_count(); }

// number of workspace slots (= resident threads) of the generic kernels
int generic_slots(int64_t N, size_t bytes_per_slot) {
    const size_t budget = (size_t)768 << 20;
    int64_t tmax = (int64_t)(budget / (bytes_per_slot ? bytes_per_slot : 1));
    if (tmax > n_sm() * 512) tmax = n_sm() * 512;
    if (tmax < 1024) tmax = 1024;
    int64_t t = ((N + INSR_GEN_THREADS - 1) / INSR_GEN_THREADS) * INSR_GEN_THREADS;
    if (t > tmax) t = (tmax / INSR_GEN_THREADS) * INSR_GEN_THREADS;
    if (t < INSR_GEN_THREADS) t = INSR_GEN_THREADS;
    return (int)t;
}

size_t generic_ws_bytes(const SirenDims &dm, int64_t N, int order, int backward) {
    const int S = insr_nstreams(dm.D, order);
    if (!backward) {
        const size_t per = (size_t)2 * S * dm.H * sizeof(float);
        return per * generic_slots(N, per);
    }
    const size_t per = insr_gen_bwd_ws_floats(S, dm.H, dm.L, dm.O, 1) * sizeof(float);
    return per * generic_slots(N, per);
}

// ---------------------------------------------------------------- generic launches
template <int D, int O, int ORDER>
int launch_generic_fwd(const SirenDims &dm, const float *theta, const float *x, int64_t N, float *y,
                       float *jac, float *h2, float *ws, void *stream) {
    const int S = StreamCfg<D, ORDER>::S;
    const int T = generic_slots(N, (size_t)2 * S * dm.H * sizeof(float));
    auto kfn = k_generic_fwd<D, O, ORDER>;
    INSR_LAUNCH(kfn, dim3(T / INSR_GEN_THREADS), dim3(INSR_GEN_THREADS), 0, stream, dm, theta, x, N, y,
                jac, h2, ws, T);
    ++g_launches;
    return check_cuda("k_generic_fwd");
}

template <int D, int O, int ORDER>
int launch_generic_bwd(const SirenDims &dm, const float *theta, const float *x, int64_t N,
                       const float *gy, const float *gjac, const float *gh2, float *gtheta, float *gx,
                       float *ws, void *stream) {
    typedef StreamCfg<D, ORDER> C;
    const int S = C::S;
    const int T = generic_slots(N, insr_gen_bwd_ws_floats(S, dm.H, dm.L, dm.O, 1) * sizeof(float));
    for (int64_t n0 = 0; n0 < N; n0 += T) {
        const int nv = (int)((N - n0 < T) ? (N - n0) : T);
        auto ksweep = k_generic_bwd_sweep<D, O, ORDER>;
        INSR_LAUNCH(ksweep, dim3((nv + INSR_GEN_THREADS - 1) / INSR_GEN_THREADS), dim3(INSR_GEN_THREADS), 0,
                    stream, dm, theta, x, N, n0, gy, gjac, gh2, gx, ws, T);
        ++g_launches;
        if (dm.L > 0) {
            const int tiles = (dm.H + 31) / 32;
            int zsplit = (2 * n_sm() + tiles * tiles * dm.L - 1) / (tiles * tiles * dm.L);
            const int zmax = (nv + 255) / 256;
            if (zsplit > zmax) zsplit = zmax;
            if (zsplit < 1) zsplit = 1;
            auto kw = k_generic_wgrad_hidden;
            INSR_LAUNCH(kw, dim3(tiles, tiles, dm.L * zsplit), dim3(256), 0, stream, dm, S, ws, T, nv,
                        zsplit, gtheta);
            ++g_launches;
        }
        const int n_elem = dm.H * dm.D + dm.H + dm.O * dm.H + dm.O;
        auto ke = k_generic_wgrad_edge;
        INSR_LAUNCH(ke, dim3((n_elem + 7) / 8), dim3(256), 0, stream, dm, S, C::ND, ws, T, nv, x, n0,
                    gtheta);
        ++g_launches;
        int rc = check_cuda("generic backward");
        if (rc) return rc;
    }
    return 0;
}

#define INSR_DISPATCH_DO(D_, O_, ORD_, CALL)                                                     \
    if (dm.D == D_ && dm.O == O_ && order == ORD_) { constexpr int D = D_, O = O_, ORDER = ORD_;  \
        (void)D; (void)O; (void)ORDER; return CALL; }
#define INSR_DISPATCH_O(D_, ORD_, CALL) \
    INSR_DISPATCH_DO(D_, 1, ORD_, CALL) INSR_DISPATCH_DO(D_, 2, ORD_, CALL) INSR_DISPATCH_DO(D_, 3, ORD_, CALL)
#define INSR_DISPATCH_ORD(D_, CALL) \
    INSR_DISPATCH_O(D_, 0, CALL) INSR_DISPATCH_O(D_, 1, CALL) INSR_DISPATCH_O(D_, 2, CALL) INSR_DISPATCH_O(D_, 3, CALL)
#define INSR_DISPATCH_ALL(CALL) INSR_DISPATCH_ORD(1, CALL) INSR_DISPATCH_ORD(2, CALL) INSR_DISPATCH_ORD(3, CALL)

int generic_forward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N,
                    float *y, float *jac, float *h2, float *ws, void *stream) {
    INSR_DISPATCH_ALL((launch_generic_fwd<D, O, ORDER>(dm, theta, x, N, y, jac, h2, ws, stream)))
    return fail(INSR_ERR_UNSUPPORTED, "no generic forward kernel for D=%d O=%d order=%d", dm.D, dm.O, order);
}

int generic_backward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N,
                     const float *gy, const float *gjac, const float *gh2, float *gtheta, float *gx,
                     float *ws, void *stream) {
    INSR_DISPATCH_ALL((launch_generic_bwd<D, O, ORDER>(dm, theta, x, N, gy, gjac, gh2, gtheta, gx, ws, stream)))
    return fail(INSR_ERR_UNSUPPORTED, "no generic backward kernel for D=%d O=%d order=%d", dm.D, dm.O, order);
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// the 32 < H <= 512 family serves this call (H <= 32 only through its tensor-core kernels: no FFMA variant at those widths)
bool tiled_ok(const insr_siren_desc *d, const SirenDims &dm, int order) {
    if (d->flags & INSR_FLAG_FORCE_GENERIC) return false;
    if (!insr_tiled_supported(dm, order)) return false;
    return dm.H > 32 || !(d->flags & INSR_FLAG_NO_TENSOR);
}
// the resident-weights family (H <= 32) serves this forward call -- unless the call keeps its tape for a backward that only
// the other family can run (e.g. L = 5: forward instantiated, backward not)
bool fused_fwd_ok(const insr_siren_desc *d, const SirenDims &dm, int order) {
    if ((d->flags & INSR_FLAG_FORCE_GENERIC) || !insr_fused_supported(dm, order, 0)) return false;
    if ((d->flags & INSR_FLAG_KEEP_TAPE) && !insr_fused_supported(dm, order, 1) && tiled_ok(d, dm, order)) return false;
    return true;
}

}  // namespace

extern "C" {

int insr_version(void) { return INSR_ABI_VERSION; }

const char *insr_last_error(void) { return g_err; }

int64_t insr_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int64_t insr_siren_theta_size(const insr_siren_desc *desc) {
    SirenDims dm;
    if (validate(desc, 0, 0, &dm)) return 0;
    return insr_theta_size(dm);
}

int insr_siren_kernel_family(const insr_siren_desc *desc, int order, int backward) {
    SirenDims dm;
    int rc = validate(desc, 0, order, &dm);
    if (rc) return rc;
    if (desc->flags & INSR_FLAG_FORCE_GENERIC) return 0;
    if (backward ? insr_fused_supported(dm, order, 1) : fused_fwd_ok(desc, dm, order)) return 1;
    return tiled_ok(desc, dm, order) ? 2 : 0;
}

int insr_siren_tape_supported(const insr_siren_desc *desc, int64_t n_points, int order) {
    SirenDims dm;
    int rc = validate(desc, n_points, order, &dm);
    if (rc) return rc;
    if (desc->flags & INSR_FLAG_FORCE_GENERIC) return 0;
    if (insr_fused_supported(dm, order, 1) || !tiled_ok(desc, dm, order)) return 0;
    return insr_tiled_tape_fits(dm, n_points, order) ? 1 : 0;
}

size_t insr_siren_workspace_bytes(const insr_siren_desc *desc, int64_t n_points, int order, int backward) {
    SirenDims dm;
    if (validate(desc, n_points, order, &dm)) return 0;
    if (!(desc->flags & INSR_FLAG_FORCE_GENERIC)) {
        if (backward ? insr_fused_supported(dm, order, 1) : fused_fwd_ok(desc, dm, order)) return insr_fused_ws_bytes(dm, n_points, order, backward);
        if (tiled_ok(desc, dm, order)) return insr_tiled_ws_bytes(dm, n_points, order, backward);
    }
    return generic_ws_bytes(dm, n_points, order, backward);
}

int insr_siren_forward(const insr_siren_desc *desc, const float *theta, const float *x, int64_t n_points,
                       int order, float *y, float *jac, float *h2, void *workspace, size_t workspace_bytes,
                       void *stream) {
    SirenDims dm;
    int rc = validate(desc, n_points, order, &dm);
    if (rc) return rc;
    if (!theta || !x || !y) return fail(INSR_ERR_NULL, "theta, x and y must not be NULL");
    if (order >= INSR_ORDER_JAC && !jac) return fail(INSR_ERR_NULL, "jac must not be NULL for order >= 1");
    if (order >= INSR_ORDER_LAP && !h2) return fail(INSR_ERR_NULL, "h2 must not be NULL for order >= 2");
    if (!aligned16(theta) || !aligned16(x) || !aligned16(y) || !aligned16(jac) || !aligned16(h2))
        return fail(INSR_ERR_ALIGN, "all device buffers must be 16-byte aligned");
    if ((rc = check_device())) return rc;
    if (n_points == 0) return 0;
    const size_t need = insr_siren_workspace_bytes(desc, n_points, order, 0);
    if (need && (!workspace || workspace_bytes < need))
        return fail(INSR_ERR_WORKSPACE, "forward needs %zu workspace bytes, got %zu", need, workspace_bytes);
    if (fused_fwd_ok(desc, dm, order)) {
        rc = insr_fused_forward(dm, order, theta, x, n_points, y, jac, h2, (float *)workspace, stream, &g_launches,
                                (desc->flags & INSR_FLAG_NO_TENSOR) == 0);
        if (rc == INSR_ERR_UNSUPPORTED) return fail(rc, "fused forward dispatch failed for D=%d O=%d H=%d", dm.D, dm.O, dm.H);
        if (rc) return rc;
        return check_cuda("fused forward");
    }
    if (tiled_ok(desc, dm, order)) {
        const bool keep = (desc->flags & INSR_FLAG_KEEP_TAPE) != 0;
        if (keep) {
            if (!insr_tiled_tape_fits(dm, n_points, order))
                return fail(INSR_ERR_UNSUPPORTED, "KEEP_TAPE: %lld points do not fit one workspace chunk", (long long)n_points);
            const size_t need_b = insr_siren_workspace_bytes(desc, n_points, order, 1);
            if (workspace_bytes < need_b)
                return fail(INSR_ERR_WORKSPACE, "KEEP_TAPE forward needs the backward workspace (%zu bytes), got %zu", need_b, workspace_bytes);
        }
        rc = insr_tiled_forward(dm, order, theta, x, n_points, y, jac, h2, (float *)workspace, stream, &g_launches,
                                !(desc->flags & INSR_FLAG_NO_TENSOR), keep);
        if (rc) return fail(rc, "tiled forward dispatch failed for D=%d O=%d H=%d", dm.D, dm.O, dm.H);
        return check_cuda("tiled forward");
    }
    return generic_forward(dm, order, theta, x, n_points, y, jac, h2, (float *)workspace, stream);
}

int insr_siren_backward(const insr_siren_desc *desc, const float *theta, const float *x, int64_t n_points,
                        int order, const float *gy, const float *gjac, const float *gh2, float *gtheta,
                        float *gx, void *workspace, size_t workspace_bytes, void *stream) {
    SirenDims dm;
    int rc = validate(desc, n_points, order, &dm);
    if (rc) return rc;
    if (!theta || !x || !gtheta) return fail(INSR_ERR_NULL, "theta, x and gtheta must not be NULL");
    if (!aligned16(theta) || !aligned16(x) || !aligned16(gtheta) || !aligned16(gy) || !aligned16(gjac) ||
        !aligned16(gh2) || !aligned16(gx))
        return fail(INSR_ERR_ALIGN, "all device buffers must be 16-byte aligned");
    if ((rc = check_device())) return rc;
    if (n_points == 0) return 0;
    const size_t need = insr_siren_workspace_bytes(desc, n_points, order, 1);
    if (need && (!workspace || workspace_bytes < need))
        return fail(INSR_ERR_WORKSPACE, "backward needs %zu workspace bytes, got %zu", need, workspace_bytes);
    if (!(desc->flags & INSR_FLAG_FORCE_GENERIC) && insr_fused_supported(dm, order, 1)) {
        rc = insr_fused_backward(dm, order, theta, x, n_points, gy, gjac, gh2, gtheta, gx,
                                 (float *)workspace, stream, &g_launches,
                                 !(desc->flags & (INSR_FLAG_FFMA_BWD | INSR_FLAG_NO_TENSOR)));
        if (rc == INSR_ERR_UNSUPPORTED) return fail(rc, "fused backward dispatch failed for D=%d O=%d H=%d", dm.D, dm.O, dm.H);
        if (rc) return rc;
        return check_cuda("fused backward");
    }
    if (tiled_ok(desc, dm, order)) {
        const bool have_tape = (desc->flags & INSR_FLAG_KEEP_TAPE) != 0;
        if (have_tape && !insr_tiled_tape_fits(dm, n_points, order))
            return fail(INSR_ERR_UNSUPPORTED, "KEEP_TAPE: %lld points do not fit one workspace chunk", (long long)n_points);
        rc = insr_tiled_backward(dm, order, theta, x, n_points, gy, gjac, gh2, gtheta, gx, (float *)workspace, stream,
                                 &g_launches, !(desc->flags & INSR_FLAG_NO_TENSOR), have_tape);
        if (rc) return fail(rc, "tiled backward dispatch failed for D=%d O=%d H=%d", dm.D, dm.O, dm.H);
        return check_cuda("tiled backward");
    }
    return generic_backward(dm, order, theta, x, n_points, gy, gjac, gh2, gtheta, gx, (float *)workspace, stream);
}

int insr_siren_lsq_step(const insr_siren_desc *desc, const float *theta, const float *x, int64_t n_points,
                        int order, int n_res, const float *coef_host, const float *target, float scale,
                        float *loss_out, float *gtheta, void *workspace, size_t workspace_bytes,
                        void *stream) {
    SirenDims dm;
    int rc = validate(desc, n_points, order, &dm);
    if (rc) return rc;
    if (order > INSR_ORDER_LAP) return fail(INSR_ERR_ORDER, "lsq_step supports orders 0..2");
    if (n_res < 1 || n_res > 4) return fail(INSR_ERR_SHAPE, "n_res=%d outside [1,4]", n_res);
    if (!theta || !x || !coef_host || !loss_out || !gtheta)
        return fail(INSR_ERR_NULL, "theta, x, coef_host, loss_out and gtheta must not be NULL");
    if ((rc = check_device())) return rc;
    if (n_points == 0) return 0;
    const bool tensor = !(desc->flags & (INSR_FLAG_FFMA_BWD | INSR_FLAG_NO_TENSOR));
    if (tensor) {
        const size_t need = insr_siren_workspace_bytes(desc, n_points, order, 1);
        if (!workspace || workspace_bytes < need)
            return fail(INSR_ERR_WORKSPACE, "lsq_step needs %zu workspace bytes, got %zu", need, workspace_bytes);
    }
    rc = insr_fused_lsq_step(dm, order, n_res, coef_host, theta, x, n_points, target, scale, loss_out,
                             gtheta, (float *)workspace, workspace_bytes, stream, &g_launches, tensor);
    if (rc == INSR_ERR_UNSUPPORTED)
        return fail(rc, "lsq_step: no fused kernel for D=%d O=%d H=%d L=%d order=%d", dm.D, dm.O, dm.H, dm.L, order);
    if (rc) return rc;
    return check_cuda("fused lsq_step");
}

#ifndef INSR_CPU_EMU
namespace {
// pack cy / cj / cl (host, nullable) into [c][o][s] with the stream order value, tangents, (Laplacian)
void pack_coef(const insr_target_eval *e, int n_res, float *dst) {
    const int D = e->desc.in_features, O = e->desc.out_features, S = insr_nstreams(D, e->order);
    for (int c = 0; c < n_res; ++c)
        for (int o = 0; o < O; ++o) {
            float *q = dst + (c * O + o) * S;
            q[0] = e->coef_y ? e->coef_y[c * O + o] : 0.f;
            if (e->order >= 1)
                for (int d = 0; d < D; ++d) q[1 + d] = e->coef_jac ? e->coef_jac[(c * O + o) * D + d] : 0.f;
            if (e->order == 2) q[1 + D] = e->coef_lap ? e->coef_lap[c * O + o] : 0.f;
        }
}
}  // namespace
#endif

int insr_siren_target(const insr_target_eval *a, const insr_target_eval *b, int mode, float dt, float lo, float hi,
                      const float *x, int64_t n_points, int n_res, float *target, void *stream) {
#ifdef INSR_CPU_EMU
    return fail(INSR_ERR_UNSUPPORTED, "siren_target: tcgen05 kernels only");
#else
    if (!a || !x || !target || !a->theta) return fail(INSR_ERR_NULL, "siren_target: a, a->theta, x and target must not be NULL");
    if (mode < 0 || mode > 2) return fail(INSR_ERR_SHAPE, "siren_target: mode=%d (0, 1 or 2)", mode);
    if (mode != 0 && !b) return fail(INSR_ERR_NULL, "siren_target: mode %d needs b", mode);
    if (mode == 2 && !b->theta) return fail(INSR_ERR_NULL, "siren_target: b->theta must not be NULL in mode 2");
    if (n_res < 1 || n_res > 2) return fail(INSR_ERR_SHAPE, "siren_target: n_res=%d outside [1,2]", n_res);
    SirenDims dA, dB;
    int rc = validate(&a->desc, n_points, a->order, &dA);
    if (rc) return rc;
    dB = dA;
    if (mode == 2 && (rc = validate(&b->desc, n_points, b->order, &dB))) return rc;
    if (!insr_fused_supported(dA, a->order, 0) || !insr_fused_supported(dB, mode == 2 ? b->order : 0, 0) || dA.D != dB.D ||
        a->order > 1 || (mode != 0 && b->order > 1))
        return fail(INSR_ERR_UNSUPPORTED, "siren_target: both fields must belong to the H <= 32 resident-weights family, orders 0..1");
    if ((rc = check_device())) return rc;
    if (n_points == 0) return 0;
    insr_tc::TargetParams p{};
    p.dmA = dA; p.dmB = dB; p.thetaA = a->theta; p.thetaB = mode == 2 ? b->theta : a->theta;
    p.x = x; p.N = n_points; p.target = target; p.n_res = n_res; p.dt = dt; p.lo = lo; p.hi = hi;
    pack_coef(a, n_res, p.coefA);
    if (mode == 1) {
        insr_target_eval second = *b;
        second.desc = a->desc;
        pack_coef(&second, n_res, p.coefB);
    } else if (mode == 2) {
        pack_coef(b, n_res, p.coefB);
    }
    const int D = dA.D, OA = dA.O, OB = dB.O, oa = a->order, ob = mode == 0 ? 0 : b->order;
#define INSR_TGT(D_, OA_, ORDA_, OB_, ORDB_, MODE_)                                                          \
    if (D == D_ && OA == OA_ && oa == ORDA_ && mode == MODE_ && (MODE_ == 0 || (OB == OB_ && ob == ORDB_))) { \
        insr_tc::launch_tc_target<D_, OA_, ORDA_, OB_, ORDB_, MODE_>(p, stream, &g_launches);                 \
        return check_cuda("k_tc_target");                                                                     \
    }
    INSR_TGT(1, 1, 0, 1, 0, 0) INSR_TGT(1, 1, 1, 1, 0, 0) INSR_TGT(2, 1, 0, 1, 0, 0) INSR_TGT(2, 1, 1, 1, 0, 0)
    INSR_TGT(2, 2, 0, 2, 0, 0) INSR_TGT(2, 2, 1, 2, 0, 0)
    INSR_TGT(1, 1, 0, 1, 0, 1) INSR_TGT(2, 2, 0, 2, 0, 1)
    INSR_TGT(2, 2, 0, 1, 1, 2) INSR_TGT(2, 2, 0, 1, 0, 2) INSR_TGT(2, 2, 1, 1, 1, 2) INSR_TGT(1, 1, 0, 1, 1, 2) INSR_TGT(1, 1, 1, 1, 1, 2)
#undef INSR_TGT
    return fail(INSR_ERR_UNSUPPORTED, "siren_target: no kernel for D=%d O_a=%d order_a=%d O_b=%d order_b=%d mode=%d", D, OA, oa, OB, ob, mode);
#endif
}

int insr_adam_step(float *theta, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                   const float *sched, float beta1, float beta2, float eps, void *stream) {
    if (!theta || !grad || !exp_avg || !exp_avg_sq || !sched) return fail(INSR_ERR_NULL, "adam_step: NULL buffer");
    if (n < 0) return fail(INSR_ERR_SHAPE, "adam_step: n=%lld", (long long)n);
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 8 * n_sm()) blocks = 8 * n_sm();
    auto kfn = k_adam_step;
    INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(256), 0, stream, theta, grad, exp_avg, exp_avg_sq, n, sched, beta1,
                beta2, eps);
    ++g_launches;
    return check_cuda("k_adam_step");
}

int insr_plateau_step(const float *loss, float *sched, float factor, int patience, float threshold, float min_lr,
                      float eps, void *stream) {
    if (!loss || !sched) return fail(INSR_ERR_NULL, "plateau_step: NULL buffer");
    int rc = check_device();
    if (rc) return rc;
    auto kfn = k_plateau_step;
    INSR_LAUNCH(kfn, dim3(1), dim3(32), 0, stream, loss, sched, factor, patience, threshold, min_lr, eps);
    ++g_launches;
    return check_cuda("k_plateau_step");
}

int insr_iteration_update(int n_slots, float *const *theta, float *const *grad, float *const *exp_avg,
                          float *const *exp_avg_sq, const int64_t *n, float *sched, float *losses, int n_losses,
                          int main_index, float *hist, int64_t hist_capacity, int64_t *hist_idx, uint32_t *ticket, float beta1,
                          float beta2, float eps, float factor, int patience, float threshold, float min_lr, float eps_lr,
                          int zero_grad, int clear_losses, void *stream) {
    if (!theta || !grad || !exp_avg || !exp_avg_sq || !n || !sched || !losses || !ticket)
        return fail(INSR_ERR_NULL, "iteration_update: NULL argument");
    if (n_slots < 1 || n_slots > INSR_MAX_OPT_SLOTS)
        return fail(INSR_ERR_SHAPE, "iteration_update: n_slots=%d (1..%d)", n_slots, INSR_MAX_OPT_SLOTS);
    if (n_losses < 1 || main_index < 0 || main_index >= n_losses)
        return fail(INSR_ERR_SHAPE, "iteration_update: n_losses=%d main_index=%d", n_losses, main_index);
    if (hist && (!hist_idx || hist_capacity < 1)) return fail(INSR_ERR_NULL, "iteration_update: a loss log needs its index word and a capacity");
    insr_opt_slots sl{};
    sl.n_slots = n_slots;
    int64_t total = 0;
    for (int k = 0; k < n_slots; ++k) {
        if (!theta[k] || !grad[k] || !exp_avg[k] || !exp_avg_sq[k] || n[k] < 0)
            return fail(INSR_ERR_NULL, "iteration_update: slot %d has a NULL buffer or a negative size", k);
        sl.theta[k] = theta[k]; sl.grad[k] = grad[k]; sl.m[k] = exp_avg[k]; sl.v[k] = exp_avg_sq[k]; sl.n[k] = n[k];
        total += n[k];
    }
    int rc = check_device();
    if (rc) return rc;
    int64_t blocks = (total + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 8 * n_sm()) blocks = 8 * n_sm();
    auto kfn = k_iteration_update;
    INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(256), 0, stream, sl, sched, losses, n_losses, main_index, hist, hist_capacity,
                hist_idx, ticket, beta1, beta2, eps, factor, patience, threshold, min_lr, eps_lr, zero_grad, clear_losses);
    ++g_launches;
    return check_cuda("k_iteration_update");
}

// ---------------------------------------------------------------- peer memory (one box, one process per GPU)
namespace {
int peer_set(int world, int rank, void *const *bases, insr_peer_set *ps) {
    if (!bases) return fail(INSR_ERR_NULL, "peer: bases is NULL");
    if (world < 1 || world > INSR_PEER_MAX_WORLD || rank < 0 || rank >= world)
        return fail(INSR_ERR_SHAPE, "peer: world=%d rank=%d (world 1..%d)", world, rank, INSR_PEER_MAX_WORLD);
    ps->world = world; ps->rank = rank;
    for (int r = 0; r < INSR_PEER_MAX_WORLD; ++r) ps->base[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        if (!bases[r]) return fail(INSR_ERR_NULL, "peer: mapping of rank %d is NULL", r);
        if (reinterpret_cast<uintptr_t>(bases[r]) & 15u) return fail(INSR_ERR_ALIGN, "peer: mapping of rank %d is not 16-byte aligned", r);
        ps->base[r] = static_cast<unsigned char *>(bases[r]);
    }
    return 0;
}
}  // namespace

#ifndef INSR_CPU_EMU
int insr_peer_alloc(int64_t data_bytes, void **base, unsigned char *handle64) {
    if (!base || !handle64) return fail(INSR_ERR_NULL, "peer_alloc: NULL argument");
    if (data_bytes < 0) return fail(INSR_ERR_SHAPE, "peer_alloc: data_bytes=%lld", (long long)data_bytes);
    int rc = check_device();
    if (rc) return rc;
    void *p = nullptr;
    const size_t bytes = (size_t)INSR_PEER_HEADER_BYTES + (((size_t)data_bytes + 255) & ~(size_t)255);
    cudaError_t e = cudaMalloc(&p, bytes);              // a plain cudaMalloc: the only kind of allocation cudaIpcGetMemHandle exports
    if (e != cudaSuccess) return fail((int)e, "peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    const uint32_t one = 1u;
    if ((e = cudaMemset(p, 0, bytes)) != cudaSuccess ||
        (e = cudaMemcpy(static_cast<uint32_t *>(p) + INSR_PEER_EPOCH_WORD, &one, 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), p)) != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        return fail((int)e, "peer_alloc: %s", cudaGetErrorString(e));
    }
    *base = p;
    return 0;
}

int insr_peer_open(const unsigned char *handle64, void **base) {
    if (!handle64 || !base) return fail(INSR_ERR_NULL, "peer_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
    *base = p;
    return 0;
}

int insr_peer_close(void *base) {
    if (!base) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(base);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "peer_close: %s", cudaGetErrorString(e)); }
    return 0;
}

int insr_peer_free(void *base) {
    if (!base) return 0;
    cudaError_t e = cudaFree(base);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "peer_free: %s", cudaGetErrorString(e)); }
    return 0;
}

int insr_peer_status(void *base, int reset) {
    if (!base) return fail(INSR_ERR_NULL, "peer_status: NULL argument");
    uint32_t v = 0;
    uint32_t *w = static_cast<uint32_t *>(base) + INSR_PEER_STATUS_WORD;
    cudaError_t e = cudaMemcpy(&v, w, 4, cudaMemcpyDeviceToHost);           // synchronises: a host-side health check, never in a loop
    if (e == cudaSuccess && reset && v) { const uint32_t z = 0; e = cudaMemcpy(w, &z, 4, cudaMemcpyHostToDevice); }
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "peer_status: %s", cudaGetErrorString(e)); }
    return (int)v;                                      // 0 healthy, 1 a barrier timed out since the last reset
}

#else
// emulation build (tests/emu): "peer memory" is host memory of one process whose ranks are host threads; the handle is the address
int insr_peer_alloc(int64_t data_bytes, void **base, unsigned char *handle64) {
    if (!base || !handle64) return fail(INSR_ERR_NULL, "peer_alloc: NULL argument");
    if (data_bytes < 0) return fail(INSR_ERR_SHAPE, "peer_alloc: data_bytes=%lld", (long long)data_bytes);
    const size_t bytes = (size_t)INSR_PEER_HEADER_BYTES + (((size_t)data_bytes + 255) & ~(size_t)255);
    void *p = aligned_alloc(256, bytes);
    if (!p) return fail(INSR_ERR_WORKSPACE, "peer_alloc: out of memory");
    memset(p, 0, bytes);
    static_cast<uint32_t *>(p)[INSR_PEER_EPOCH_WORD] = 1u;
    memset(handle64, 0, 64);
    memcpy(handle64, &p, sizeof(p));
    *base = p;
    return 0;
}
int insr_peer_open(const unsigned char *handle64, void **base) {
    if (!handle64 || !base) return fail(INSR_ERR_NULL, "peer_open: NULL argument");
    memcpy(base, handle64, sizeof(void *));
    return 0;
}
int insr_peer_close(void *) { return 0; }
int insr_peer_free(void *base) { free(base); return 0; }
int insr_peer_status(void *base, int reset) {
    if (!base) return fail(INSR_ERR_NULL, "peer_status: NULL argument");
    uint32_t *w = static_cast<uint32_t *>(base) + INSR_PEER_STATUS_WORD;
    const uint32_t v = *w;
    if (reset) *w = 0;
    return (int)v;
}

#endif

int insr_peer_allreduce(int world, int rank, void *const *bases, int64_t offset_floats, int64_t n, float scale, float *out,
                        void *stream) {
    insr_peer_set ps;
    int rc = peer_set(world, rank, bases, &ps);
    if (rc) return rc;
    if (!out) return fail(INSR_ERR_NULL, "peer_allreduce: out is NULL");
    if (n < 0 || offset_floats < INSR_PEER_HEADER_BYTES / 4 || (offset_floats & 3))
        return fail(INSR_ERR_SHAPE, "peer_allreduce: n=%lld offset=%lld floats (offset must lie behind the %d-byte header, 16-byte aligned)",
                    (long long)n, (long long)offset_floats, INSR_PEER_HEADER_BYTES);
    if ((rc = check_device())) return rc;
    int64_t blocks = (n / 4 + 511) / 512;
    if (blocks < 1) blocks = 1;
    if (blocks > INSR_PEER_MAX_CTAS) blocks = INSR_PEER_MAX_CTAS;
    auto kfn = insr_peer::k_peer_allreduce;
    INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(512), 0, stream, ps, offset_floats, n, scale, out);
    ++g_launches;
    return check_cuda("k_peer_allreduce");
}

int insr_iteration_update_peer(int world, int rank, void *const *bases, int64_t peer_bytes, float scale, int n_slots,
                               float *const *theta, float *const *grad, float *const *exp_avg, float *const *exp_avg_sq,
                               const int64_t *n, float *sched, float *losses, int n_losses, int main_index, float *losses_red,
                               float *hist, int64_t hist_capacity, int64_t *hist_idx, float beta1, float beta2, float eps,
                               float factor, int patience, float threshold, float min_lr, float eps_lr, int zero_grad,
                               int clear_losses, void *stream) {
    insr_peer_set ps;
    int rc = peer_set(world, rank, bases, &ps);
    if (rc) return rc;
    if (!theta || !grad || !exp_avg || !exp_avg_sq || !n || !sched || !losses || !losses_red)
        return fail(INSR_ERR_NULL, "iteration_update_peer: NULL argument");
    if (n_slots < 1 || n_slots > INSR_MAX_OPT_SLOTS)
        return fail(INSR_ERR_SHAPE, "iteration_update_peer: n_slots=%d (1..%d)", n_slots, INSR_MAX_OPT_SLOTS);
    if (n_losses < 1 || n_losses > 32 || main_index < 0 || main_index >= n_losses)
        return fail(INSR_ERR_SHAPE, "iteration_update_peer: n_losses=%d main_index=%d", n_losses, main_index);
    if (hist && (!hist_idx || hist_capacity < 1)) return fail(INSR_ERR_NULL, "iteration_update_peer: a loss log needs its index word and a capacity");
    const unsigned char *own = ps.base[rank];
    auto inside = [&](const float *p, int64_t count) {
        const unsigned char *b = reinterpret_cast<const unsigned char *>(p);
        return b >= own + INSR_PEER_HEADER_BYTES && b + 4 * count <= own + peer_bytes;
    };
    if (!inside(losses, n_losses)) return fail(INSR_ERR_SHAPE, "iteration_update_peer: the loss slots are not inside this rank's peer allocation");
    insr_opt_slots sl{};
    sl.n_slots = n_slots;
    int64_t total = 0;
    for (int k = 0; k < n_slots; ++k) {
        if (!theta[k] || !grad[k] || !exp_avg[k] || !exp_avg_sq[k] || n[k] < 0)
            return fail(INSR_ERR_NULL, "iteration_update_peer: slot %d has a NULL buffer or a negative size", k);
        if (!inside(grad[k], n[k])) return fail(INSR_ERR_SHAPE, "iteration_update_peer: gradient slot %d is not inside this rank's peer allocation", k);
        sl.theta[k] = theta[k]; sl.grad[k] = grad[k]; sl.m[k] = exp_avg[k]; sl.v[k] = exp_avg_sq[k]; sl.n[k] = n[k];
        total += n[k];
    }
    if ((rc = check_device())) return rc;
    int64_t blocks = (total + 511) / 512;
    if (blocks < 1) blocks = 1;
    if (blocks > INSR_PEER_MAX_CTAS) blocks = INSR_PEER_MAX_CTAS;
    auto kfn = insr_peer::k_iteration_update_peer;
    INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(512), 0, stream, ps, sl, scale, sched, losses, n_losses, main_index, losses_red,
                hist, hist_capacity, hist_idx, beta1, beta2, eps, factor, patience, threshold, min_lr, eps_lr, zero_grad, clear_losses);
    ++g_launches;
    return check_cuda("k_iteration_update_peer");
}

int insr_svd_small(const float *F, int64_t n, int d, float *U, float *S, float *V, void *stream) {
    if (!F || !S) return fail(INSR_ERR_NULL, "svd_small: F and S must not be NULL");
    if (n < 0 || (d != 2 && d != 3)) return fail(INSR_ERR_SHAPE, "svd_small: n=%lld d=%d (d must be 2 or 3)", (long long)n, d);
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return 0;
    int64_t blocks = (n + 127) / 128;
    if (blocks > n_sm() * 16) blocks = n_sm() * 16;
    if (d == 2) { auto kfn = k_svd_small<2>; INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(128), 0, stream, F, n, U, S, V); }
    else        { auto kfn = k_svd_small<3>; INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(128), 0, stream, F, n, U, S, V); }
    ++g_launches;
    return check_cuda("k_svd_small");
}

int insr_elastic_energy(const float *F, int64_t n, int d, float ratio_arap, float ratio_volume, float *energy,
                        float *gF, void *stream) {
    if (!F || !energy) return fail(INSR_ERR_NULL, "elastic_energy: F and energy must not be NULL");
    if (n < 0 || (d != 2 && d != 3)) return fail(INSR_ERR_SHAPE, "elastic_energy: n=%lld d=%d (d must be 2 or 3)", (long long)n, d);
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return 0;
    int64_t blocks = (n + 127) / 128;
    if (blocks > n_sm() * 16) blocks = n_sm() * 16;
    if (d == 2) { auto kfn = k_elastic_energy<2>; INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(128), 0, stream, F, n, ratio_arap, ratio_volume, energy, gF); }
    else        { auto kfn = k_elastic_energy<3>; INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(128), 0, stream, F, n, ratio_arap, ratio_volume, energy, gF); }
    ++g_launches;
    return check_cuda("k_elastic_energy");
}

int insr_elastic_terms(const insr_elastic_terms_desc *t, int d, const float *y, const float *J, const float *x,
                       const float *y_prev, const float *y_pp, float *loss, float *gy, float *gJ, void *stream) {
    if (!t || !y || !loss || !gy) return fail(INSR_ERR_NULL, "elastic_terms: desc, y, loss and gy must not be NULL");
    if (d != 2 && d != 3) return fail(INSR_ERR_SHAPE, "elastic_terms: d=%d (must be 2 or 3)", d);
    if (t->n < 0 || t->n_left < 0 || t->n_right < 0) return fail(INSR_ERR_SHAPE, "elastic_terms: negative row count");
    if (t->n > 0 && (!x || !y_prev || !y_pp)) return fail(INSR_ERR_NULL, "elastic_terms: interior rows need x, y_prev, y_pp");
    if ((t->r_arap != 0.f || t->r_volume != 0.f) && (!J || !gJ))
        return fail(INSR_ERR_NULL, "elastic_terms: arap / volume need J and gJ");
    if (!(t->dt > 0.f)) return fail(INSR_ERR_SHAPE, "elastic_terms: dt must be positive");
    int rc = check_device();
    if (rc) return rc;
    insr_elastic_terms_k k{};
    k.n = t->n; k.n_left = t->n_left; k.n_right = t->n_right;
    k.dt = t->dt; k.r_arap = t->r_arap; k.r_volume = t->r_volume; k.r_kin = t->r_kinematics; k.r_left = t->r_left;
    k.r_right = t->r_right; k.r_plane = t->r_plane; k.plane_height = t->plane_height; k.r_sphere = t->r_sphere;
    k.radius = t->radius;
    for (int i = 0; i < 3; ++i) { k.ext[i] = t->external_force[i]; k.off[i] = t->offset_right[i]; k.center[i] = t->center[i]; }
    const int64_t n_all = k.n + k.n_left + k.n_right;
    if (n_all == 0) return 0;
    const bool energy = (k.r_arap != 0.f || k.r_volume != 0.f);
    int64_t blocks = (n_all + 127) / 128;
    if (blocks > n_sm() * 16) blocks = n_sm() * 16;
    if (d == 2) { auto kfn = k_elastic_terms<2>; INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(128), 0, stream, k, y, energy ? J : nullptr, x, y_prev, y_pp, loss, gy, energy ? gJ : nullptr); }
    else        { auto kfn = k_elastic_terms<3>; INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(128), 0, stream, k, y, energy ? J : nullptr, x, y_prev, y_pp, loss, gy, energy ? gJ : nullptr); }
    ++g_launches;
    return check_cuda("k_elastic_terms");
}

int insr_sample_boxes(int n_boxes, int dim, const int32_t *count, const float *lo, const float *hi, uint64_t seed,
                      int64_t *counter, uint32_t *ticket, int64_t point_offset, float *out, void *stream) {
    if (!count || !lo || !hi || !out) return fail(INSR_ERR_NULL, "sample_boxes: NULL argument");
    if (n_boxes < 1 || n_boxes > INSR_MAX_BOXES || dim < 1 || dim > 3)
        return fail(INSR_ERR_SHAPE, "sample_boxes: n_boxes=%d (1..%d) dim=%d (1..3)", n_boxes, INSR_MAX_BOXES, dim);
    if (counter && !ticket) return fail(INSR_ERR_NULL, "sample_boxes: a device counter needs a zeroed ticket word");
    insr_box_set bs{};
    bs.n_boxes = n_boxes; bs.dim = dim;
    int64_t total = 0;
    for (int b = 0; b < n_boxes; ++b) {
        if (count[b] < 0) return fail(INSR_ERR_SHAPE, "sample_boxes: count[%d]=%d", b, count[b]);
        bs.count[b] = count[b]; total += count[b];
        for (int d = 0; d < dim; ++d) { bs.lo[b][d] = lo[b * dim + d]; bs.hi[b][d] = hi[b * dim + d]; }
    }
    int rc = check_device();
    if (rc) return rc;
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    if (blocks > n_sm() * 8) blocks = n_sm() * 8;
    auto kfn = k_sample_boxes;
    INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(256), 0, stream, bs, seed, counter, ticket, point_offset, out);
    ++g_launches;
    return check_cuda("k_sample_boxes");
}

int insr_sample_mesh(const float *V, const int32_t *elem, const float *cdf, int n_elem, int verts_per_elem, int64_t n,
                     int dim_out, uint64_t seed, int64_t *counter, uint32_t *ticket, int64_t point_offset, float *out,
                     void *stream) {
    if (!V || !elem || !cdf || !out) return fail(INSR_ERR_NULL, "sample_mesh: NULL argument");
    if (n_elem < 1 || n < 0 || dim_out < 1 || dim_out > 3 || (verts_per_elem != 3 && verts_per_elem != 4))
        return fail(INSR_ERR_SHAPE, "sample_mesh: n_elem=%d n=%lld dim_out=%d (1..3) verts_per_elem=%d (3 or 4)", n_elem,
                    (long long)n, dim_out, verts_per_elem);
    if (counter && !ticket) return fail(INSR_ERR_NULL, "sample_mesh: a device counter needs a zeroed ticket word");
    int rc = check_device();
    if (rc) return rc;
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > n_sm() * 8) blocks = n_sm() * 8;
    if (verts_per_elem == 3) {
        auto kfn = k_sample_mesh<3>;
        INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(256), 0, stream, V, elem, cdf, n_elem, n, dim_out, seed, counter, ticket, point_offset, out);
    } else {
        auto kfn = k_sample_mesh<4>;
        INSR_LAUNCH(kfn, dim3((unsigned)blocks), dim3(256), 0, stream, V, elem, cdf, n_elem, n, dim_out, seed, counter, ticket, point_offset, out);
    }
    ++g_launches;
    return check_cuda("k_sample_mesh");
}

}  // extern "C"
