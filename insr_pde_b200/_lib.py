"""ctypes binding of libinsr_b200.so (the C ABI declared in include/insr_b200.h).

There is exactly one backend: the sm_100a CUDA library built in-tree by ``build.py``.  If it
is missing and cannot be built, importing the operators fails loudly -- there is no CPU or
PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libinsr_b200.so")

ORDER_VALUE, ORDER_JAC, ORDER_LAP, ORDER_HESS = 0, 1, 2, 3
FLAG_FORCE_GENERIC = 1
FLAG_NO_TENSOR = 2
FLAG_FFMA_BWD = 4
FLAG_KEEP_TAPE = 8

EXPORTS = [
    "insr_version", "insr_last_error", "insr_siren_theta_size", "insr_siren_workspace_bytes",
    "insr_siren_forward", "insr_siren_backward", "insr_siren_lsq_step", "insr_siren_kernel_family",
    "insr_launch_count", "insr_adam_step", "insr_plateau_step", "insr_svd_small", "insr_elastic_energy", "insr_sample_boxes", "insr_sample_mesh", "insr_siren_tape_supported", "insr_elastic_terms", "insr_iteration_update",
    "insr_siren_target",
    "insr_peer_alloc", "insr_peer_open", "insr_peer_close", "insr_peer_free", "insr_peer_status", "insr_peer_allreduce",
    "insr_iteration_update_peer",
]


class SirenDesc(ctypes.Structure):
    _fields_ = [("in_features", ctypes.c_int32), ("out_features", ctypes.c_int32),
                ("hidden_features", ctypes.c_int32), ("num_hidden_layers", ctypes.c_int32),
                ("omega", ctypes.c_float), ("flags", ctypes.c_int32)]


class TargetEval(ctypes.Structure):
    """insr_target_eval: one frozen field of insr_siren_target with its host coefficient arrays"""
    _fields_ = [("desc", SirenDesc), ("theta", ctypes.c_void_p), ("order", ctypes.c_int32),
                ("coef_y", ctypes.POINTER(ctypes.c_float)), ("coef_jac", ctypes.POINTER(ctypes.c_float)),
                ("coef_lap", ctypes.POINTER(ctypes.c_float))]


class InsrError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libinsr_b200 error {code}: {message}")
        self.code = code


_vp, _i64, _i32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
_dp = ctypes.POINTER(SirenDesc)


class Library:
    """thin typed wrapper over the shared object"""

    def __init__(self, path):
        self.path = path
        self.cdll = ctypes.CDLL(path)
        c = self.cdll
        c.insr_version.restype = _i32
        c.insr_last_error.restype = ctypes.c_char_p
        c.insr_siren_theta_size.restype = _i64
        c.insr_siren_theta_size.argtypes = [_dp]
        c.insr_siren_workspace_bytes.restype = _sz
        c.insr_siren_workspace_bytes.argtypes = [_dp, _i64, _i32, _i32]
        c.insr_siren_forward.restype = _i32
        c.insr_siren_forward.argtypes = [_dp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]
        c.insr_siren_backward.restype = _i32
        c.insr_siren_backward.argtypes = [_dp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
        c.insr_siren_lsq_step.restype = _i32
        c.insr_siren_lsq_step.argtypes = [_dp, _vp, _vp, _i64, _i32, _i32, ctypes.POINTER(ctypes.c_float),
                                          _vp, ctypes.c_float, _vp, _vp, _vp, _sz, _vp]
        c.insr_siren_tape_supported.restype = _i32
        c.insr_siren_tape_supported.argtypes = [_dp, _i64, _i32]
        c.insr_siren_kernel_family.restype = _i32
        c.insr_siren_kernel_family.argtypes = [_dp, _i32, _i32]
        _f = ctypes.c_float
        c.insr_adam_step.restype = _i32
        c.insr_adam_step.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _vp]
        c.insr_plateau_step.restype = _i32
        c.insr_plateau_step.argtypes = [_vp, _vp, _f, _i32, _f, _f, _f, _vp]
        c.insr_svd_small.restype = _i32
        c.insr_svd_small.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp]
        c.insr_elastic_energy.restype = _i32
        c.insr_elastic_energy.argtypes = [_vp, _i64, _i32, _f, _f, _vp, _vp, _vp]
        c.insr_iteration_update.restype = _i32
        c.insr_iteration_update.argtypes = [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp] + \
            [ctypes.c_float] * 4 + [_i32] + [ctypes.c_float] * 3 + [_i32, _i32, _vp]
        c.insr_peer_alloc.restype = _i32
        c.insr_peer_alloc.argtypes = [_i64, ctypes.POINTER(_vp), ctypes.c_char_p]
        c.insr_peer_open.restype = _i32
        c.insr_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(_vp)]
        c.insr_peer_close.restype = _i32
        c.insr_peer_close.argtypes = [_vp]
        c.insr_peer_free.restype = _i32
        c.insr_peer_free.argtypes = [_vp]
        c.insr_peer_status.restype = _i32
        c.insr_peer_status.argtypes = [_vp, _i32]
        c.insr_peer_allreduce.restype = _i32
        c.insr_peer_allreduce.argtypes = [_i32, _i32, _vp, _i64, _i64, ctypes.c_float, _vp, _vp]
        c.insr_iteration_update_peer.restype = _i32
        c.insr_iteration_update_peer.argtypes = [_i32, _i32, _vp, _i64, ctypes.c_float, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                                 _vp, _vp, _i64, _vp] + [ctypes.c_float] * 4 + [_i32] + [ctypes.c_float] * 3 + \
            [_i32, _i32, _vp]
        c.insr_elastic_terms.restype = _i32
        c.insr_elastic_terms.argtypes = [ctypes.POINTER(ElasticTermsDesc), _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        c.insr_sample_mesh.restype = _i32
        c.insr_sample_mesh.argtypes = [_vp, _vp, _vp, _i32, _i32, _i64, _i32, ctypes.c_uint64, _vp, _vp, _i64, _vp, _vp]
        c.insr_sample_boxes.restype = _i32
        c.insr_sample_boxes.argtypes = [_i32, _i32, _vp, _vp, _vp, ctypes.c_uint64, _vp, _vp, _i64, _vp, _vp]
        c.insr_launch_count.restype = _i64
        c.insr_launch_count.argtypes = [_i32]
        c.insr_siren_target.restype = _i32
        c.insr_siren_target.argtypes = [ctypes.POINTER(TargetEval), ctypes.POINTER(TargetEval), _i32, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_float, _vp, _i64, _i32, _vp, _vp]

    def check(self, rc):
        if rc != 0:
            raise InsrError(rc, (self.cdll.insr_last_error() or b"").decode())

    def version(self):
        return self.cdll.insr_version()

    def theta_size(self, desc):
        return int(self.cdll.insr_siren_theta_size(ctypes.byref(desc)))

    def workspace_bytes(self, desc, n, order, backward):
        return int(self.cdll.insr_siren_workspace_bytes(ctypes.byref(desc), n, order, int(backward)))

    def tape_supported(self, desc, n, order):
        return int(self.cdll.insr_siren_tape_supported(ctypes.byref(desc), n, order)) == 1

    def kernel_family(self, desc, order, backward):
        return int(self.cdll.insr_siren_kernel_family(ctypes.byref(desc), order, int(backward)))

    def launch_count(self, reset=False):
        return int(self.cdll.insr_launch_count(int(reset)))

    def forward(self, desc, theta, x, n, order, y, jac, h2, ws, ws_bytes, stream):
        self.check(self.cdll.insr_siren_forward(ctypes.byref(desc), theta, x, n, order, y, jac, h2,
                                                ws, ws_bytes, stream))

    def backward(self, desc, theta, x, n, order, gy, gjac, gh2, gtheta, gx, ws, ws_bytes, stream):
        self.check(self.cdll.insr_siren_backward(ctypes.byref(desc), theta, x, n, order, gy, gjac, gh2,
                                                 gtheta, gx, ws, ws_bytes, stream))

    def adam_step(self, theta, grad, m, v, n, sched, beta1, beta2, eps, stream):
        self.check(self.cdll.insr_adam_step(theta, grad, m, v, n, sched, beta1, beta2, eps, stream))

    def iteration_update(self, thetas, grads, ms, vs, sizes, sched, losses, n_losses, main_index, hist, hist_capacity,
                         hist_idx, ticket, beta1, beta2, eps, factor, patience, threshold, min_lr, eps_lr, zero_grad, stream,
                         clear_losses=False):
        k = len(thetas)
        arr = lambda ptrs: (ctypes.c_void_p * k)(*ptrs)
        self.check(self.cdll.insr_iteration_update(k, arr(thetas), arr(grads), arr(ms), arr(vs), (ctypes.c_int64 * k)(*sizes),
                                                   sched, losses, n_losses, main_index, hist, hist_capacity, hist_idx, ticket,
                                                   beta1, beta2, eps, factor, patience, threshold, min_lr, eps_lr,
                                                   int(zero_grad), int(clear_losses), stream))

    # ---- peer memory (one box, one process per GPU): see include/insr_b200.h
    def peer_alloc(self, data_bytes):
        """-> (base address, 64-byte IPC handle)"""
        base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self.check(self.cdll.insr_peer_alloc(int(data_bytes), ctypes.byref(base), handle))
        return int(base.value), handle.raw

    def peer_open(self, handle):
        base = ctypes.c_void_p()
        self.check(self.cdll.insr_peer_open(ctypes.create_string_buffer(bytes(handle), 64), ctypes.byref(base)))
        return int(base.value)

    def peer_close(self, base):
        self.check(self.cdll.insr_peer_close(base))

    def peer_free(self, base):
        self.check(self.cdll.insr_peer_free(base))

    def peer_status(self, base, reset=False):
        rc = self.cdll.insr_peer_status(base, int(reset))
        if rc < 0 or rc > 1:
            self.check(rc)
        return rc

    def peer_allreduce(self, world, rank, bases, offset_floats, n, scale, out, stream):
        self.check(self.cdll.insr_peer_allreduce(world, rank, (ctypes.c_void_p * world)(*bases), offset_floats, n, scale, out, stream))

    def iteration_update_peer(self, world, rank, bases, peer_bytes, scale, thetas, grads, ms, vs, sizes, sched, losses, n_losses,
                              main_index, losses_red, hist, hist_capacity, hist_idx, beta1, beta2, eps, factor, patience,
                              threshold, min_lr, eps_lr, zero_grad, clear_losses, stream):
        k = len(thetas)
        arr = lambda ptrs: (ctypes.c_void_p * k)(*ptrs)
        self.check(self.cdll.insr_iteration_update_peer(world, rank, (ctypes.c_void_p * world)(*bases), peer_bytes, scale, k,
                                                        arr(thetas), arr(grads), arr(ms), arr(vs), (ctypes.c_int64 * k)(*sizes),
                                                        sched, losses, n_losses, main_index, losses_red, hist, hist_capacity,
                                                        hist_idx, beta1, beta2, eps, factor, patience, threshold, min_lr, eps_lr,
                                                        int(zero_grad), int(clear_losses), stream))

    def plateau_step(self, loss, sched, factor, patience, threshold, min_lr, eps, stream):
        self.check(self.cdll.insr_plateau_step(loss, sched, factor, patience, threshold, min_lr, eps, stream))

    def sample_boxes(self, counts, lo, hi, dim, seed, counter, ticket, point_offset, out, stream):
        n = len(counts)
        c = (ctypes.c_int32 * n)(*counts)
        l = (ctypes.c_float * (n * dim))(*[v for row in lo for v in row])
        h = (ctypes.c_float * (n * dim))(*[v for row in hi for v in row])
        self.check(self.cdll.insr_sample_boxes(n, dim, c, l, h, seed, counter, ticket, point_offset, out, stream))

    def sample_mesh(self, V, elem, cdf, n_elem, k, n, dim_out, seed, counter, ticket, point_offset, out, stream):
        self.check(self.cdll.insr_sample_mesh(V, elem, cdf, n_elem, k, n, dim_out, seed, counter, ticket, point_offset,
                                              out, stream))

    def elastic_terms(self, tdesc, d, y, J, x, y_prev, y_pp, loss, gy, gJ, stream):
        self.check(self.cdll.insr_elastic_terms(ctypes.byref(tdesc), d, y, J, x, y_prev, y_pp, loss, gy, gJ, stream))

    def svd_small(self, F, n, d, U, S, V, stream):
        self.check(self.cdll.insr_svd_small(F, n, d, U, S, V, stream))

    def elastic_energy(self, F, n, d, ratio_arap, ratio_volume, energy, gF, stream):
        self.check(self.cdll.insr_elastic_energy(F, n, d, ratio_arap, ratio_volume, energy, gF, stream))

    def lsq_step(self, desc, theta, x, n, order, n_res, coef, target, scale, loss_out, gtheta, ws,
                 ws_bytes, stream):
        arr = (ctypes.c_float * len(coef))(*coef)
        self.check(self.cdll.insr_siren_lsq_step(ctypes.byref(desc), theta, x, n, order, n_res, arr, target,
                                                 scale, loss_out, gtheta, ws, ws_bytes, stream))


    def target(self, a, b, mode, dt, lo, hi, x, n, n_res, target, stream):
        """a, b: (desc, theta_ptr, order, cy, cj, cl) with flat python lists (or None) as coefficients; b may be None"""
        keep = []

        def ev(t):
            if t is None:
                return None
            desc, theta, order, cy, cj, cl = t
            e = TargetEval()
            e.desc, e.theta, e.order = desc, theta, int(order)
            for name, vals in (("coef_y", cy), ("coef_jac", cj), ("coef_lap", cl)):
                if vals is not None:
                    arr = (ctypes.c_float * len(vals))(*[float(v) for v in vals])
                    keep.append(arr)
                    setattr(e, name, ctypes.cast(arr, ctypes.POINTER(ctypes.c_float)))
            keep.append(e)
            return ctypes.byref(e)

        self.check(self.cdll.insr_siren_target(ev(a), ev(b), int(mode), float(dt), float(lo), float(hi), x, n, int(n_res),
                                               target, stream))


_LOCK = threading.Lock()
_LIB = None


def get_lib() -> Library:
    """the one and only backend; builds it in-tree on first use if nvcc is available."""
    global _LIB
    if _LIB is None:
        with _LOCK:
            if _LIB is None:
                if not os.path.exists(LIB_PATH):
                    from . import build
                    build.build_library()
                _LIB = Library(LIB_PATH)
    return _LIB


class ElasticTermsDesc(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("n_left", ctypes.c_int64), ("n_right", ctypes.c_int64), ("dt", ctypes.c_float),
                ("r_arap", ctypes.c_float), ("r_volume", ctypes.c_float), ("r_kinematics", ctypes.c_float),
                ("r_left", ctypes.c_float), ("r_right", ctypes.c_float), ("r_plane", ctypes.c_float),
                ("plane_height", ctypes.c_float), ("r_sphere", ctypes.c_float), ("radius", ctypes.c_float),
                ("external_force", ctypes.c_float * 3), ("offset_right", ctypes.c_float * 3), ("center", ctypes.c_float * 3)]


def make_desc(in_features, out_features, hidden_features, num_hidden_layers, omega=30.0, flags=0):
    return SirenDesc(int(in_features), int(out_features), int(hidden_features), int(num_hidden_layers),
                     float(omega), int(flags))
