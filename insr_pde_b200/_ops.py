"""Tensor-level wrappers over the C ABI: allocate outputs / workspace with the PyTorch caching
allocator, pass raw device pointers and the current CUDA stream.  PyTorch is plumbing here
(device memory, streams); all arithmetic happens in libinsr_b200.so.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ORDER_HESS, ORDER_JAC, ORDER_LAP, ORDER_VALUE  # noqa: F401


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError(
            "insr_pde_b200 runs on CUDA (sm_100a) tensors only; got a tensor on "
            f"'{t.device}'. There is no CPU fallback.")


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _DeviceGuard:
    def __init__(self, device):
        self.g = torch.cuda.device(device) if device.type == "cuda" else None

    def __enter__(self):
        if self.g is not None:
            self.g.__enter__()

    def __exit__(self, *a):
        if self.g is not None:
            self.g.__exit__(*a)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _workspace(nbytes, device):
    if nbytes == 0:
        return None, 0
    ws = torch.empty(nbytes + 16, dtype=torch.uint8, device=device)
    return ws, nbytes


def _check_input(t, name, shape=None):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:                 # the C ABI wants 16-byte aligned buffers: an offset view (a slice of a shard, a
        t = t.clone()                     # torch.split piece, a narrow()ed gradient) is copied, not refused
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def out_shapes(desc, n, order):
    D, O = desc.in_features, desc.out_features
    shapes = [(n, O)]
    if order >= ORDER_JAC:
        shapes.append((n, O, D))
    if order == ORDER_LAP:
        shapes.append((n, O))
    if order == ORDER_HESS:
        shapes.append((n, O, D, D))
    return shapes


def _tape_desc(desc):
    return _lib.make_desc(desc.in_features, desc.out_features, desc.hidden_features, desc.num_hidden_layers, desc.omega,
                          desc.flags | _lib.FLAG_KEEP_TAPE)


def siren_forward(desc, theta, x, order, keep_tape=False, out=None):
    """theta: flat (P,), x: (N, D) -> tuple (y[, jac[, h2]]).  ``out``: preallocated contiguous, 16-byte aligned fp32
    tensors to write the outputs into (e.g. row ranges of persistent buffers) instead of fresh ones.
    ``keep_tape=True`` (a forward whose backward will follow): returns ``(outs, tape)`` where ``tape`` is the workspace
    holding the activations of every layer, to be handed to ``siren_backward(..., tape=tape)``, or None where the kernel
    family keeps no tape (H <= 32: recomputed in registers; batches beyond one workspace chunk)."""
    lib = _lib.get_lib()
    _require_cuda(x)
    x = _check_input(x, "x")
    theta = _check_input(theta, "theta")
    n = x.shape[0]
    if x.dim() != 2 or x.shape[1] != desc.in_features:
        raise ValueError(f"x must be (N, {desc.in_features}), got {tuple(x.shape)}")
    if theta.numel() != lib.theta_size(desc):
        raise ValueError(f"theta has {theta.numel()} elements, expected {lib.theta_size(desc)}")
    if out is None:
        outs = [torch.empty(s, dtype=torch.float32, device=x.device) for s in out_shapes(desc, n, order)]
    else:
        outs = list(out)
        for t, s in zip(outs, out_shapes(desc, n, order)):
            if tuple(t.shape) != tuple(s) or not t.is_contiguous() or t.dtype != torch.float32 or t.data_ptr() % 16:
                raise ValueError(f"siren_forward: out must be contiguous, 16-byte aligned fp32 of shape {tuple(s)}")
    if n == 0:
        return (tuple(outs), None) if keep_tape else tuple(outs)
    tape = None
    with _DeviceGuard(x.device):
        taped = keep_tape and lib.tape_supported(desc, n, order)
        if taped:
            desc = _tape_desc(desc)
        ws, nb = _workspace(lib.workspace_bytes(desc, n, order, taped), x.device)
        y = outs[0]
        jac = outs[1] if order >= ORDER_JAC else None
        h2 = outs[2] if order >= ORDER_LAP else None
        lib.forward(desc, theta.data_ptr(), x.data_ptr(), n, order, y.data_ptr(), _ptr(jac), _ptr(h2),
                    _ptr(ws), nb, _stream(x.device))
        if taped:
            tape = (ws, nb)
    return (tuple(outs), tape) if keep_tape else tuple(outs)


def siren_backward(desc, theta, x, order, gy=None, gjac=None, gh2=None, need_gx=False, gtheta=None, tape=None):
    """returns (gtheta flat (P,), gx (N, D) or None).  If ``gtheta`` is given the parameter
    gradient is accumulated into it (it must be a flat fp32 buffer of P elements).  ``tape``: what
    ``siren_forward(..., keep_tape=True)`` returned for the same (desc, theta, x, order) -- skips the recomputation."""
    lib = _lib.get_lib()
    _require_cuda(x)
    x = _check_input(x, "x")
    theta = _check_input(theta, "theta")
    n = x.shape[0]
    shapes = out_shapes(desc, n, order)
    gy = _check_input(gy, "gy", shapes[0])
    gjac = _check_input(gjac, "gjac", shapes[1]) if order >= ORDER_JAC else None
    gh2 = _check_input(gh2, "gh2", shapes[2]) if order >= ORDER_LAP else None
    if gtheta is None:
        gtheta = torch.zeros(theta.numel(), dtype=torch.float32, device=x.device)
    gx = torch.empty_like(x) if need_gx else None
    if n == 0:
        return gtheta, gx
    with _DeviceGuard(x.device):
        if tape is not None:
            desc = _tape_desc(desc)
            ws, nb = tape
        else:
            ws, nb = _workspace(lib.workspace_bytes(desc, n, order, True), x.device)
        lib.backward(desc, theta.data_ptr(), x.data_ptr(), n, order, _ptr(gy), _ptr(gjac), _ptr(gh2),
                     gtheta.data_ptr(), _ptr(gx), _ptr(ws), nb, _stream(x.device))
    return gtheta, gx


def siren_lsq_step(desc, theta, x, order, coef_y, coef_jac, coef_lap, target, scale, loss_out=None,
                   gtheta=None):
    """fused residual + loss + backward (see include/insr_b200.h: insr_siren_lsq_step).
    coef_y: (R, O), coef_jac: (R, O, D), coef_lap: (R, O) python nested lists / tensors (host)."""
    lib = _lib.get_lib()
    _require_cuda(x)
    x = _check_input(x, "x")
    theta = _check_input(theta, "theta")
    n = x.shape[0]
    D, O = desc.in_features, desc.out_features
    cy = torch.as_tensor(coef_y, dtype=torch.float32).reshape(-1, O)
    R = cy.shape[0]
    cj = torch.zeros(R, O, D) if coef_jac is None else torch.as_tensor(coef_jac, dtype=torch.float32).reshape(R, O, D)
    cl = torch.zeros(R, O) if coef_lap is None else torch.as_tensor(coef_lap, dtype=torch.float32).reshape(R, O)
    coef = cy.flatten().tolist() + cj.flatten().tolist() + cl.flatten().tolist()
    target = _check_input(target, "target", (n, R)) if target is not None else None
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=torch.float32, device=x.device)
    if gtheta is None:
        gtheta = torch.zeros(theta.numel(), dtype=torch.float32, device=x.device)
    if n == 0:
        return loss_out, gtheta
    with _DeviceGuard(x.device):
        ws, nb = _workspace(lib.workspace_bytes(desc, n, order, True), x.device)
        lib.lsq_step(desc, theta.data_ptr(), x.data_ptr(), n, order, R, coef, _ptr(target), float(scale),
                     loss_out.data_ptr(), gtheta.data_ptr(), _ptr(ws), nb, _stream(x.device))
    return loss_out, gtheta


def siren_target(x, n_res, a, b=None, mode=0, dt=0.0, lo=-1.0, hi=1.0, out=None):
    """target (N, n_res) of a least-squares closure from one or two frozen fields in ONE kernel (include/insr_b200.h:
    insr_siren_target).  a, b: dicts with ``net`` (an MLP of the H <= 32 family), ``order`` and the coefficient lists
    ``cy`` (n_res x O), ``cj`` (n_res x O x D), ``cl`` (nested lists or None).  mode 0: a at x; 1: backtrace through a
    (b carries the coefficients of the second evaluation); 2: a and b at x.  Raises InsrError(-6) where no kernel exists."""
    lib = _lib.get_lib()
    _require_cuda(x)
    x = _check_input(x, "x")
    n = x.shape[0]

    def flat(v):
        if v is None:
            return None
        return torch.as_tensor(v, dtype=torch.float32).reshape(-1).tolist()

    def pack(t):
        if t is None:
            return None
        net = t["net"]
        return (net.desc, net.flat_theta().data_ptr(), t.get("order", 0), flat(t.get("cy")), flat(t.get("cj")), flat(t.get("cl")))

    if out is None:
        out = torch.empty(n, n_res, dtype=torch.float32, device=x.device)
    elif tuple(out.shape) != (n, n_res) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("siren_target: out must be a contiguous fp32 (N, n_res) tensor on the points' device")
    if n:
        with _DeviceGuard(x.device):
            lib.target(pack(a), pack(b), mode, dt, lo, hi, x.data_ptr(), n, n_res, out.data_ptr(), _stream(x.device))
    return out


def adam_step(theta, grad, exp_avg, exp_avg_sq, sched, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam semantics on a flat fp32 vector; lr and step index are read from the device
    tensor ``sched`` = [lr, best, num_bad_epochs, step] (include/insr_b200.h: insr_adam_step)"""
    lib = _lib.get_lib()
    _require_cuda(theta)
    with _DeviceGuard(theta.device):
        lib.adam_step(theta.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), theta.numel(),
                      sched.data_ptr(), beta1, beta2, eps, _stream(theta.device))


def plateau_step(loss, sched, factor=0.1, patience=500, threshold=1e-4, min_lr=1e-8, eps=1e-8):
    """ReduceLROnPlateau(mode='min', threshold_mode='rel').step(loss) on the device + step counter bump"""
    lib = _lib.get_lib()
    _require_cuda(sched)
    with _DeviceGuard(sched.device):
        lib.plateau_step(loss.data_ptr(), sched.data_ptr(), factor, int(patience), threshold, min_lr, eps,
                         _stream(sched.device))


def svd_small(F, compute_uv=True):
    """batched SVD of (..., d, d) fp32 matrices, d in {2, 3}: returns (U, S, V) with F = U diag(S) V^T, S descending
    (U, V are None when compute_uv is False).  See include/insr_b200.h: insr_svd_small."""
    lib = _lib.get_lib()
    _require_cuda(F)
    d = F.shape[-1]
    if F.shape[-2] != d or d not in (2, 3):
        raise ValueError(f"svd_small: expected (..., 2, 2) or (..., 3, 3), got {tuple(F.shape)}")
    Fc = _check_input(F.reshape(-1, d, d), "F")
    n = Fc.shape[0]
    S = torch.empty(n, d, dtype=torch.float32, device=F.device)
    U = torch.empty(n, d, d, dtype=torch.float32, device=F.device) if compute_uv else None
    V = torch.empty(n, d, d, dtype=torch.float32, device=F.device) if compute_uv else None
    if n:
        with _DeviceGuard(F.device):
            lib.svd_small(Fc.data_ptr(), n, d, _ptr(U), S.data_ptr(), _ptr(V), _stream(F.device))
    lead = F.shape[:-2]
    return (U.reshape(*lead, d, d) if compute_uv else None, S.reshape(*lead, d), V.reshape(*lead, d, d) if compute_uv else None)


def elastic_energy(F, ratio_arap, ratio_volume, need_grad=True):
    """E = ratio_arap sum (S - 1)^2 + ratio_volume sum (prod S - 1)^2 over all (d, d) matrices of F and dE/dF, in one
    kernel.  Returns (E (1,), gF or None).  See include/insr_b200.h: insr_elastic_energy."""
    lib = _lib.get_lib()
    _require_cuda(F)
    d = F.shape[-1]
    if F.shape[-2] != d or d not in (2, 3):
        raise ValueError(f"elastic_energy: expected (..., 2, 2) or (..., 3, 3), got {tuple(F.shape)}")
    Fc = _check_input(F.reshape(-1, d, d), "F")
    n = Fc.shape[0]
    E = torch.zeros(1, dtype=torch.float32, device=F.device)
    gF = torch.empty_like(Fc) if need_grad else None
    if n:
        with _DeviceGuard(F.device):
            lib.elastic_energy(Fc.data_ptr(), n, d, float(ratio_arap), float(ratio_volume), E.data_ptr(), _ptr(gF),
                               _stream(F.device))
    return E, (gF.reshape(F.shape) if need_grad else None)


def elastic_terms(y, J, x, y_prev, y_pp, n_left, n_right, *, dt, r_arap=0.0, r_volume=0.0, r_kinematics=0.0, r_left=0.0,
                  r_right=0.0, r_plane=0.0, plane_height=0.0, r_sphere=0.0, radius=0.0, external_force=None,
                  offset_right=None, center=None, loss_out=None):
    """every term of the elasticity closure and its cotangents in one kernel (include/insr_b200.h: insr_elastic_terms).
    y (n + n_left + n_right, d), J (same rows, d, d) or None, x / y_prev / y_pp (n, d).  Returns (loss (1,), gy, gJ);
    ``loss_out``: a 1-element fp32 view the loss is ACCUMULATED into instead of a fresh zero."""
    lib = _lib.get_lib()
    _require_cuda(y)
    d = y.shape[1]
    n_all = y.shape[0]
    n = n_all - n_left - n_right
    y = _check_input(y, "y")
    J = _check_input(J, "J", (n_all, d, d)) if J is not None else None
    x = _check_input(x, "x", (n, d))
    y_prev = _check_input(y_prev, "y_prev", (n, d))
    y_pp = _check_input(y_pp, "y_pp", (n, d))
    t = _lib.ElasticTermsDesc()
    t.n, t.n_left, t.n_right, t.dt = n, n_left, n_right, float(dt)
    t.r_arap, t.r_volume, t.r_kinematics = float(r_arap), float(r_volume), float(r_kinematics)
    t.r_left, t.r_right, t.r_plane, t.plane_height = float(r_left), float(r_right), float(r_plane), float(plane_height)
    t.r_sphere, t.radius = float(r_sphere), float(radius)
    for name, vals in (("external_force", external_force), ("offset_right", offset_right), ("center", center)):
        arr = getattr(t, name)
        for i in range(3):
            arr[i] = float(vals[i]) if vals is not None and i < len(vals) else 0.0
    loss = torch.zeros(1, dtype=torch.float32, device=y.device) if loss_out is None else loss_out
    gy = torch.empty_like(y)
    gJ = torch.empty_like(J) if J is not None else None
    if n_all:
        with _DeviceGuard(y.device):
            lib.elastic_terms(t, d, y.data_ptr(), _ptr(J), x.data_ptr(), y_prev.data_ptr(), y_pp.data_ptr(), loss.data_ptr(),
                              gy.data_ptr(), _ptr(gJ), _stream(y.device))
    return loss, gy, gJ
