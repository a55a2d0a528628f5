"""Fused loss closures: the mean-square residual losses of the advection
and fluid models evaluated by ``insr_siren_lsq_step`` -- forward streams, residual, loss partial
and reverse sweep in ONE kernel per term, no autograd graph, no activation recompute, no
y/J/lap round trip through HBM.  Gradients are accumulated straight into a flat buffer that the
parameters' ``.grad`` alias, so ``torch.optim.Adam`` (one fused update per net) consumes them
unchanged.

Each function mirrors one ``@_training_loop`` closure of the reference on EXPLICIT samples and
returns the same ``loss_dict`` (device scalars):

  advect_initialize      advection/model.py:43-52      advect_step          advection/model.py:68-91
  fluid_initialize       fluid/model.py:43-52          fluid_advect_velocity fluid/model.py:72-101
  fluid_solve_pressure   fluid/model.py:103-125        fluid_projection     fluid/model.py:127-151

``TrainingLoop`` restates the caller (base/baseModel.py:55-62, 73-81, 104-134): fresh Adam +
ReduceLROnPlateau(factor .1, patience 500, min_lr 1e-8) per loop, ``scheduler.step(main)``,
optional early stop at lr <= 1.1e-8.
"""
from __future__ import annotations

import gc
import math

import torch

from . import _ops
from ._ops import ORDER_JAC, ORDER_LAP, ORDER_VALUE


# ------------------------------------------------------------------------------------------------
# flat gradient plumbing
# ------------------------------------------------------------------------------------------------
def flat_grad(net):
    """One flat fp32 gradient buffer per net whose slices ARE the parameters' .grad (re-created
    if an optimizer's zero_grad(set_to_none=True) dropped them or the net moved device)."""
    theta = net.flat_theta()
    buf = getattr(net, "_flat_grad", None)
    ok = buf is not None and buf.device == theta.device and buf.numel() == theta.numel()
    if ok:
        for p, (off, numel, shape) in zip(net.parameters(), net.param_slices()):
            if p.grad is None or p.grad.data_ptr() != buf.data_ptr() + 4 * off:
                ok = False
                break
    if not ok:
        if buf is None or buf.device != theta.device or buf.numel() != theta.numel():
            buf = torch.zeros_like(theta)
            net._flat_grad = buf
        for p, (off, numel, shape) in zip(net.parameters(), net.param_slices()):
            if p.grad is None:                   # dropped by zero_grad(set_to_none=True): that means zero
                buf[off:off + numel].zero_()
            elif p.grad.data_ptr() != buf.data_ptr() + 4 * off:
                buf[off:off + numel].copy_(p.grad.reshape(-1))
            p.grad = buf[off:off + numel].view(shape)
    return buf


def zero_grads(*nets):
    for n in nets:
        flat_grad(n).zero_()


class Losses(dict):
    """a loss dict whose values are views of ONE small device vector (``.vector``, in key order): the kernels accumulate
    their loss terms straight into its slots, and the consumers (loss log, all-reduce, LR schedule) take the vector as it
    is -- no per-term zero fills, no ``bx + by``, no ``torch.stack`` in the iteration graph"""

    def __init__(self, vector, keys):
        super().__init__((k, vector[i]) for i, k in enumerate(keys))
        self.vector = vector


_ACC_PROVIDER = []          # innermost first-class owner of the loss slots (GraphedLoop._iteration pushes itself)


def _acc(n, device):
    """the ``n`` loss slots one closure call accumulates into.  Eager callers get a fresh zero vector; under a
    ``GraphedLoop`` the slots are a slice of the loop's persistent vector (with data parallelism: of the buffer the ranks
    exchange), which the iteration's update kernel leaves zeroed -- no fill node, no copy into the exchange buffer."""
    if _ACC_PROVIDER:
        return _ACC_PROVIDER[-1](n, device)
    return torch.zeros(n, dtype=torch.float32, device=device)


def lsq(net, x, order, cy, cj=None, cl=None, target=None, scale=None, out=None):
    """loss = scale * sum_{n,c} (sum_o cy[c,o] y + cj[c,o,:].J + cl[c,o] lap - target[n,c])^2 ;
    d loss / d theta is ACCUMULATED into flat_grad(net); returns the loss (1-element tensor).
    ``out``: a 1-element view the loss is ACCUMULATED into instead of a fresh zero (several terms may share a slot)."""
    x = x.detach().reshape(-1, net.in_features).contiguous()
    rows = len(cy)
    if scale is None:
        scale = 1.0 / (x.shape[0] * rows)
    if target is not None:
        target = target.detach().reshape(x.shape[0], rows).contiguous().float()
    loss = torch.zeros(1, dtype=torch.float32, device=x.device) if out is None else out
    if not lsq_kernel_available(net, order):
        return _lsq_by_layers(net, x, order, cy, cj, cl, target, scale, loss)
    _ops.siren_lsq_step(net.desc, net.flat_theta(), x, order, cy, cj, cl, target, scale,
                        loss_out=loss, gtheta=flat_grad(net))
    return loss[0]


def lsq_kernel_available(net, order):
    """insr_siren_lsq_step exists for the resident-weights family only (H <= 32, 1 <= L <= 3, (D, O) in {(1,1), (2,1), (2,2)})"""
    return order <= ORDER_LAP and _ops._lib.get_lib().kernel_family(net.desc, order, True) == 1


def _lsq_by_layers(net, x, order, cy, cj, cl, target, scale, loss):
    """the same least-squares term for every other shape (the reference's default hidden_features = 64, deeper nets, 3-D):
    one forward that keeps its tape, the residual / loss / cotangents as a few elementwise kernels, the reverse sweep from
    the tape into the flat gradient buffer.  Same arguments and results as the one-kernel path."""
    D, O = net.in_features, net.out_features
    dev = x.device
    cy_t = torch.as_tensor(cy, dtype=torch.float32, device=dev).reshape(-1, O)
    R = cy_t.shape[0]
    outs, tape = _ops.siren_forward(net.desc, net.flat_theta(), x, order, keep_tape=True)
    r = outs[0] @ cy_t.t()
    cj_t = cl_t = None
    if order >= ORDER_JAC and cj is not None:
        cj_t = torch.as_tensor(cj, dtype=torch.float32, device=dev).reshape(R, O, D)
        r = r + torch.einsum("nod,cod->nc", outs[1], cj_t)
    if order >= ORDER_LAP and cl is not None:
        cl_t = torch.as_tensor(cl, dtype=torch.float32, device=dev).reshape(R, O)
        r = r + outs[2] @ cl_t.t()
    if target is not None:
        r = r - target
    loss += scale * torch.sum(r * r)
    g = (2.0 * scale) * r
    gy = g @ cy_t
    gj = (torch.einsum("nc,cod->nod", g, cj_t) if cj_t is not None else torch.zeros_like(outs[1])) if order >= ORDER_JAC else None
    gl = (g @ cl_t if cl_t is not None else torch.zeros_like(outs[2])) if order >= ORDER_LAP else None
    _ops.siren_backward(net.desc, net.flat_theta(), x, order, gy, gj, gl, gtheta=flat_grad(net), tape=tape)
    return loss[0]


def evaluate(net, x, order):
    """(y[, J[, lap]]) of a (frozen) net without autograd"""
    x = x.detach().reshape(-1, net.in_features).contiguous()
    return _ops.siren_forward(net.desc, net.flat_theta(), x, order)


_TARGET_OFF = set()          # (mode, shapes) for which insr_siren_target has no kernel: fall back to separate evaluations


def _target(x, n_res, a, b=None, mode=0, dt=0.0, out=None):
    """the frozen-net side of a closure in one kernel where insr_siren_target serves the shapes; None otherwise"""
    def shape(t):
        n = t["net"]
        return (n.in_features, n.out_features, n.hidden_features, n.num_hidden_layers, t.get("order", 0))
    key = (mode, shape(a), shape(b) if (b is not None and mode == 2) else None)
    if key in _TARGET_OFF or not x.is_cuda:
        return None
    try:
        return _ops.siren_target(x, n_res, a, b, mode=mode, dt=dt, out=out)
    except _ops._lib.InsrError as e:
        if e.code != -6:
            raise
        _TARGET_OFF.add(key)
        return None


def _eye(n):
    return [[1.0 if i == j else 0.0 for j in range(n)] for i in range(n)]


# ------------------------------------------------------------------------------------------------
# parallel branches: the terms of one closure are independent kernels (they only meet in the flat gradient
# buffer, which every kernel updates with red.global); the boundary terms use a handful of CTAs, so they run
# beside the interior term on side streams.  Under CUDA-graph capture this becomes a fork / join in the graph.
# ------------------------------------------------------------------------------------------------
_side_streams = {}


def parallel(ref, *fns):
    """run fns[0] on the current stream and fns[1:] on side streams (fork after everything enqueued so far,
    join before returning); sequential for CPU tensors.  Returns the list of results."""
    if len(fns) == 1 or not ref.is_cuda:
        return [f() for f in fns]
    dev = ref.device
    cur = torch.cuda.current_stream(dev)
    streams = _side_streams.setdefault(dev.index, [])
    while len(streams) < len(fns) - 1:
        streams.append(torch.cuda.Stream(device=dev))
    fork = torch.cuda.Event()
    fork.record(cur)
    out = [None] * len(fns)
    for k, f in enumerate(fns[1:]):
        streams[k].wait_event(fork)
        with torch.cuda.stream(streams[k]):
            out[k + 1] = f()
    out[0] = fns[0]()
    for k in range(len(fns) - 1):
        done = torch.cuda.Event()
        done.record(streams[k])
        cur.wait_event(done)
    return out


# ------------------------------------------------------------------------------------------------
# advection (1-D, constant velocity)
# ------------------------------------------------------------------------------------------------
def advect_initialize(field, samples, init_values):
    acc = _acc(1, samples.device)
    lsq(field, samples, ORDER_VALUE, [[1.0]], target=init_values, out=acc[0:1])
    return Losses(acc, ("main",))


def _into(out, value):
    """``value`` (N, R), in ``out`` where the caller provides a persistent buffer"""
    if out is None:
        return value
    out.copy_(value.reshape(out.shape))
    return out


def advect_target(field_prev, samples, dt, vel, out=None):
    """the frozen side of the midpoint residual: u_prev / dt - vel / 2 * d u_prev / dx  (advection/model.py:78-84)"""
    target = _target(samples.detach().reshape(-1, 1), 1, dict(net=field_prev, order=ORDER_JAC, cy=[[1.0 / dt]], cj=[[[-0.5 * vel]]]), out=out)
    if target is None:
        u_prev, j_prev = evaluate(field_prev, samples, ORDER_JAC)
        target = _into(out, (u_prev / dt - (0.5 * vel) * j_prev[:, :, 0]).reshape(-1, 1))
    return target


def advect_step(field, field_prev, samples, boundary_samples, dt, vel, target=None):
    """midpoint residual  (u - u_prev)/dt + vel (u_x + u_prev_x)/2  and Dirichlet band; ``target``: the frozen side
    (``advect_target``) where the caller has prepared it ahead"""
    def interior():
        tgt = target if target is not None else advect_target(field_prev, samples, dt, vel)
        return lsq(field, samples, ORDER_JAC, [[1.0 / dt]], cj=[[[0.5 * vel]]], target=tgt, out=acc[0:1])

    acc = _acc(2, samples.device)
    parallel(samples, interior, lambda: lsq(field, boundary_samples, ORDER_VALUE, [[1.0]], out=acc[1:2]))
    return Losses(acc, ("main", "bc"))


# ------------------------------------------------------------------------------------------------
# fluid (2-D inviscid Euler, operator splitting)
# ------------------------------------------------------------------------------------------------
def _no_slip_terms(velocity, bc_x, bc_y, out):
    """mean(u_x^2) on the left / right bands + mean(u_y^2) on the bottom / top bands, both accumulated into ``out``"""
    return (lambda: lsq(velocity, bc_x, ORDER_VALUE, [[1.0, 0.0]], out=out),
            lambda: lsq(velocity, bc_y, ORDER_VALUE, [[0.0, 1.0]], out=out))


def fluid_initialize(velocity, samples, init_values):
    acc = _acc(1, samples.device)
    lsq(velocity, samples, ORDER_VALUE, _eye(2), target=init_values, out=acc[0:1])
    return Losses(acc, ("main",))


def fluid_advect_target(velocity_prev, samples, dt, out=None):
    """u_prev(clamp(x - u_prev(x) dt)): both evaluations, the clamp and nothing else in one kernel (fluid/model.py:78-87)"""
    x = samples.detach().reshape(-1, 2)
    u_adv = _target(x, 2, dict(net=velocity_prev, order=ORDER_VALUE), dict(net=velocity_prev, order=ORDER_VALUE, cy=_eye(2)),
                    mode=1, dt=dt, out=out)
    if u_adv is None:
        (u_prev,) = evaluate(velocity_prev, x, ORDER_VALUE)
        back = torch.clamp(x - u_prev * dt, min=-1.0, max=1.0)
        u_adv = _into(out, evaluate(velocity_prev, back, ORDER_VALUE)[0])
    return u_adv


def fluid_advect_velocity(velocity, velocity_prev, samples, bc_x, bc_y, dt, target=None):
    """semi-Lagrangian: u(x) = u_prev(clamp(x - u_prev(x) dt)); ``target``: the frozen side where prepared ahead"""
    x = samples.detach().reshape(-1, 2)

    def interior():
        u_adv = target if target is not None else fluid_advect_target(velocity_prev, x, dt)
        return lsq(velocity, x, ORDER_VALUE, _eye(2), target=u_adv, out=acc[0:1])

    acc = _acc(2, x.device)
    parallel(x, interior, *_no_slip_terms(velocity, bc_x, bc_y, acc[1:2]))
    return Losses(acc, ("main", "bc"))


def fluid_pressure_target(velocity, samples, out=None):
    """div u of the (detached) velocity (fluid/model.py:108-109)"""
    div_u = _target(samples.detach().reshape(-1, 2), 1, dict(net=velocity, order=ORDER_JAC, cj=[[[1.0, 0.0], [0.0, 1.0]]]), out=out)
    if div_u is None:
        _, jac_u = evaluate(velocity, samples, ORDER_JAC)
        div_u = _into(out, (jac_u[:, 0, 0] + jac_u[:, 1, 1]).reshape(-1, 1))
    return div_u


def fluid_solve_pressure(velocity, pressure, samples, bc_x, bc_y, target=None):
    """lap p = div u, Neumann band; ``target``: div u where prepared ahead"""
    def interior():
        div_u = target if target is not None else fluid_pressure_target(velocity, samples)
        return lsq(pressure, samples, ORDER_LAP, [[0.0]], cl=[[1.0]], target=div_u, out=acc[0:1])

    acc = _acc(2, samples.device)
    parallel(samples, interior,
             lambda: lsq(pressure, bc_x, ORDER_JAC, [[0.0]], cj=[[[1.0, 0.0]]], out=acc[1:2]),
             lambda: lsq(pressure, bc_y, ORDER_JAC, [[0.0]], cj=[[[0.0, 1.0]]], out=acc[1:2]))
    return Losses(acc, ("main", "bc"))


def fluid_projection_target(velocity_prev, pressure, samples, out=None):
    """u_prev - grad p (fluid/model.py:131-137)"""
    target = _target(samples.detach().reshape(-1, 2), 2, dict(net=velocity_prev, order=ORDER_VALUE, cy=_eye(2)),
                     dict(net=pressure, order=ORDER_JAC, cj=[[[-1.0, 0.0]], [[0.0, -1.0]]]), mode=2, out=out)
    if target is None:
        (u_prev,) = evaluate(velocity_prev, samples, ORDER_VALUE)
        _, jac_p = evaluate(pressure, samples, ORDER_JAC)
        target = _into(out, u_prev - jac_p[:, 0, :])
    return target


def fluid_projection(velocity, velocity_prev, pressure, samples, bc_x, bc_y, target=None):
    """u <- u_prev - grad p; ``target``: the frozen side where prepared ahead"""
    def interior():
        tgt = target if target is not None else fluid_projection_target(velocity_prev, pressure, samples)
        return lsq(velocity, samples, ORDER_VALUE, _eye(2), target=tgt, out=acc[0:1])

    acc = _acc(2, samples.device)
    parallel(samples, interior, *_no_slip_terms(velocity, bc_x, bc_y, acc[1:2]))
    return Losses(acc, ("main", "bc"))


# ------------------------------------------------------------------------------------------------
# elasticity (elasticity/model.py:127-189, elasticity/losses.py): the deformation-gradient energies in one kernel
# ------------------------------------------------------------------------------------------------
def _host_vec(v):
    """closure constants (forces, offsets, centres) as host floats, read back ONCE per tensor (they never change)"""
    if not torch.is_tensor(v):
        return [float(a) for a in v]
    cached = getattr(v, "_insr_host", None)
    if cached is None:
        cached = [float(a) for a in v.detach().reshape(-1).cpu()]
        try:
            v._insr_host = cached
        except AttributeError:
            pass
    return cached


def elasticity_solve_deformation(deformation, prev, prev_prev, samples, fixed_left, fixed_right, *, dt, timestep, energy,
                                 ratio_arap, ratio_volume, ratio_kinematics, ratio_constraint, ratio_collide,
                                 external_force, external_force_timesteps, constraint_offset_right, plane_height,
                                 circle_center, circle_radius, batch=None):
    """ElasticityModel._solve_deformation (elasticity/model.py:127-189) without an autograd graph: one order-1 evaluation
    of the trainable field over [interior | left face | right face] that keeps its tape, the two frozen fields on side
    streams, ONE kernel for every loss term and its cotangents (insr_elastic_terms: Jacobi SVD energies, kinematics,
    external force, constraints, collisions), and the reverse sweep from the tape straight into the flat gradient buffer.
    Returns detached values, like the fluid closures.  The reference's 3-D sphere collision (an (M,M,3) broadcast) and
    unknown terms fall back to ``elasticity_solve_deformation_autograd``."""
    known = {"arap", "volume", "kinematics", "external", "constraint", "constraint_right", "constraint_right_compress",
             "collision", "collision_sphere"}
    dim = deformation.in_features
    both_right = "constraint_right" in energy and "constraint_right_compress" in energy
    if (not set(energy) <= known or both_right or ("collision_sphere" in energy and dim != 2)
            or len(set(energy)) != len(list(energy))):
        return elasticity_solve_deformation_autograd(
            deformation, prev, prev_prev, samples, fixed_left, fixed_right, dt=dt, timestep=timestep, energy=energy,
            ratio_arap=ratio_arap, ratio_volume=ratio_volume, ratio_kinematics=ratio_kinematics,
            ratio_constraint=ratio_constraint, ratio_collide=ratio_collide, external_force=external_force,
            external_force_timesteps=external_force_timesteps, constraint_offset_right=constraint_offset_right,
            plane_height=plane_height, circle_center=circle_center, circle_radius=circle_radius)
    ra = ratio_arap if "arap" in energy else 0.0
    rv = ratio_volume if "volume" in energy else 0.0
    order = ORDER_JAC if (ra or rv) else ORDER_VALUE
    theta = deformation.flat_theta()
    if batch is not None:
        # a prepared batch (ElasticityBatch): [interior | left face | right face] already sits in ONE persistent buffer (no cat
        # kernels), and the previous-frame fields are only evaluated on the rows that change between iterations -- their
        # values on the constant rows (the uniform grid / the mesh vertices) were computed once for this time step
        x_all, samples, n, n_left, n_right = batch.x_all, batch.x_all[:batch.n], batch.n, batch.n_left, batch.n_right
        use_left, use_right = n_left > 0, n_right > 0
        x_var = batch.x_all[batch.n_const:batch.n]
        (outs, tape), _, _ = parallel(samples,
                                      lambda: _ops.siren_forward(deformation.desc, theta, x_all, order, keep_tape=True),
                                      lambda: _ops.siren_forward(prev.desc, prev.flat_theta(), x_var, ORDER_VALUE, out=[batch.y_prev[batch.n_const:]]),
                                      lambda: _ops.siren_forward(prev_prev.desc, prev_prev.flat_theta(), x_var, ORDER_VALUE, out=[batch.y_pp[batch.n_const:]]))
        y_prev, y_pp = batch.y_prev, batch.y_pp
    else:
        samples = samples.detach().reshape(-1, dim).contiguous()
        n = samples.shape[0]
        use_left = "constraint" in energy and torch.is_tensor(fixed_left)
        use_right = ("constraint_right" in energy or "constraint_right_compress" in energy) and torch.is_tensor(fixed_right)
        extra = ([fixed_left.detach()] if use_left else []) + ([fixed_right.detach()] if use_right else [])
        x_all = torch.cat([samples] + [e.reshape(-1, dim).to(samples.dtype) for e in extra], dim=0) if extra else samples
        n_left = extra[0].shape[0] if use_left else 0
        n_right = x_all.shape[0] - n - n_left
        (outs, tape), y_prev, y_pp = parallel(samples,
                                              lambda: _ops.siren_forward(deformation.desc, theta, x_all, order, keep_tape=True),
                                              lambda: evaluate(prev, samples, ORDER_VALUE)[0],
                                              lambda: evaluate(prev_prev, samples, ORDER_VALUE)[0])
    if batch is None and samples.is_cuda and not torch.cuda.is_current_stream_capturing():
        for t in (y_prev, y_pp):
            t.record_stream(torch.cuda.current_stream(samples.device))
    sign = -1.0 if "constraint_right_compress" in energy else 1.0
    forced = "external" in energy and timestep <= external_force_timesteps
    acc = _acc(1, x_all.device)              # under a graphed loop: the loop's own loss slot (no fill, no copy into the log)
    loss, gy, gJ = _ops.elastic_terms(
        outs[0], outs[1] if order == ORDER_JAC else None, samples, y_prev, y_pp, n_left, n_right, dt=dt, r_arap=ra, r_volume=rv,
        r_kinematics=ratio_kinematics if "kinematics" in energy else 0.0, r_left=ratio_constraint if use_left else 0.0,
        r_right=ratio_constraint if use_right else 0.0, r_plane=ratio_collide if "collision" in energy else 0.0,
        plane_height=plane_height, r_sphere=ratio_collide if "collision_sphere" in energy else 0.0, radius=circle_radius,
        external_force=_host_vec(external_force) if forced else None,
        offset_right=[sign * a for a in _host_vec(constraint_offset_right)], center=_host_vec(circle_center), loss_out=acc[0:1])
    _ops.siren_backward(deformation.desc, theta, x_all, order, gy, gJ, None, gtheta=flat_grad(deformation), tape=tape)
    return Losses(acc, ("main",))


def elasticity_solve_deformation_autograd(deformation, prev, prev_prev, samples, fixed_left, fixed_right, *, dt, timestep, energy,
                                          ratio_arap, ratio_volume, ratio_kinematics, ratio_constraint, ratio_collide,
                                          external_force, external_force_timesteps, constraint_offset_right, plane_height,
                                          circle_center, circle_radius):
    """the reference closure with ``jacobian -> torch.svd -> (S - 1)^2, (prod S - 1)^2`` replaced by the field's
    order-1 kernel and ONE ``insr_elastic_energy`` kernel (energy + adjoint, no SVD graph); the remaining terms are the
    reference's elementwise expressions under autograd (the caller runs ``backward``)."""
    from . import function, linalg
    samples = samples.detach().reshape(-1, deformation.in_features).contiguous()
    n = samples.shape[0]
    ra = ratio_arap if "arap" in energy else 0.0
    rv = ratio_volume if "volume" in energy else 0.0
    # the clamped faces ride along with the interior batch (+1..2 % points) instead of costing two more forward /
    # reverse kernel chains of their own: rows [n, n + n_left) and [n + n_left, ...) of the same evaluation
    use_left = "constraint" in energy and torch.is_tensor(fixed_left)
    use_right = ("constraint_right" in energy or "constraint_right_compress" in energy) and torch.is_tensor(fixed_right)
    extra = ([fixed_left.detach()] if use_left else []) + ([fixed_right.detach()] if use_right else [])
    x_all = torch.cat([samples] + [e.reshape(-1, samples.shape[1]).to(samples.dtype) for e in extra], dim=0) if extra else samples
    n_left = extra[0].shape[0] if use_left else 0

    # value and Jacobian of the trainable field from ONE order-1 kernel (and one reverse kernel for both cotangents);
    # jacobian(net(x) + x, x) = J + I, so the points need no gradient and no NaN-status host sync happens.  The two
    # frozen previous-frame evaluations are independent of it: side streams (a fork / join under graph capture).
    outs, y_prev, y_pp = parallel(samples,
                                  lambda: function.evaluate(deformation, x_all, ORDER_JAC if (ra or rv) else ORDER_VALUE),
                                  lambda: evaluate(prev, samples, ORDER_VALUE)[0],
                                  lambda: evaluate(prev_prev, samples, ORDER_VALUE)[0])
    if samples.is_cuda and not torch.cuda.is_current_stream_capturing():
        for t in (y_prev, y_pp):
            t.record_stream(torch.cuda.current_stream(samples.device))
    q_prev, q_pp = y_prev + samples, y_pp + samples
    y_all = outs[0]
    q = y_all[:n] + samples
    qdot = (q - q_prev) / dt
    qdot_prev = (q_prev - q_pp) / dt
    loss = 0
    if ra or rv:
        F = outs[1][:n] + torch.eye(samples.shape[1], device=samples.device, dtype=samples.dtype)
        loss = loss + linalg.elastic_energy(F, ra, rv)
    for term in energy:
        if term in ("arap", "volume"):
            continue
        if term == "kinematics":
            loss = loss + ratio_kinematics * torch.sum((qdot - qdot_prev) ** 2)
        elif term == "external":
            if timestep <= external_force_timesteps:
                loss = loss - dt * torch.sum(qdot * external_force.reshape(1, -1))
        elif term == "constraint":
            loss = loss + ratio_constraint * torch.sum(y_all[n:n + n_left] ** 2)
        elif term in ("constraint_right", "constraint_right_compress"):
            sign = 1.0 if term == "constraint_right" else -1.0
            loss = loss + ratio_constraint * torch.sum((y_all[n + n_left:] - sign * constraint_offset_right.reshape(1, -1)) ** 2)
        elif term == "collision":                                  # elasticity/losses.py:10-20, masked instead of gathered
            hit = (q[:, -1] < plane_height).to(q.dtype)
            loss = loss - dt * ratio_collide * torch.sum(hit * qdot[:, -1] * (plane_height - q[:, -1]))
        elif term == "collision_sphere":                           # elasticity/losses.py:22-39
            vec = q - circle_center.reshape(1, -1)
            dist = torch.sqrt(torch.sum(vec ** 2, dim=1))
            hit = (dist < circle_radius).to(q.dtype)
            if q.shape[1] == 2:        # force = ratio * dist * dir = ratio * vec  (dir = vec / dist)
                loss = loss - dt * ratio_collide * torch.sum(hit[:, None] * qdot * vec)
            else:                      # losses.py:37 broadcasts (M,1,1) * (M,3) -> (M,M,3): sum_i dist_i * sum_j qdot_j . dir_j
                dirs = vec / dist[:, None]
                loss = loss - dt * ratio_collide * torch.sum(hit * dist) * torch.sum(hit[:, None] * qdot * dirs)
        else:
            raise NotImplementedError(term)
    return {"main": loss}


# ------------------------------------------------------------------------------------------------
# the caller: one @_training_loop
# ------------------------------------------------------------------------------------------------
def _backward_if_needed(loss_dict):
    """the fluid / advection closures accumulate d loss / d theta inside their kernels and return detached values; a
    closure that still carries an autograd graph (elasticity) is differentiated here, summed like _update_network
    (base/baseModel.py:73-77) -- the accumulation lands in the same flat gradient buffers"""
    live = [v for v in loss_dict.values() if torch.is_tensor(v) and v.requires_grad]
    if live:
        sum(live).backward()
    if isinstance(loss_dict, Losses):
        return                                         # kernel-accumulated values: nothing to detach
    for k in loss_dict:
        if torch.is_tensor(loss_dict[k]):
            loss_dict[k] = loss_dict[k].detach()


class TrainingLoop:
    """fresh Adam + ReduceLROnPlateau per loop over the nets' parameters (base/baseModel.py:55-62);
    ``closure(i)`` must accumulate gradients via ``lsq`` and return the loss_dict."""

    def __init__(self, nets, lr, early_stop=False, reducer=None):
        self.nets = list(nets)
        self.reducer = reducer
        self.early_stop = early_stop
        for n in self.nets:
            flat_grad(n)
        self.opt = torch.optim.Adam([{"params": n.parameters(), "lr": lr} for n in self.nets])
        self.sched = torch.optim.lr_scheduler.ReduceLROnPlateau(self.opt, factor=0.1, min_lr=1e-8, patience=500)

    def run(self, closure, n_iters, on_step=None):
        hist = []
        for i in range(n_iters):
            zero_grads(*self.nets)
            loss_dict = closure(i)
            _backward_if_needed(loss_dict)
            if self.reducer is not None:
                # gradients AND loss values in the one collective: every rank must feed the scheduler (and the
                # early-stop test below) the same number, or the replicas' learning rates drift apart
                keys = list(loss_dict)
                red = self.reducer.allreduce(extra_scalars=torch.stack([loss_dict[k].detach().reshape(()).float() for k in keys]))
                if red is not None:
                    loss_dict = {k: red[j] for j, k in enumerate(keys)}
            self.opt.step()
            values = {k: float(v) for k, v in loss_dict.items()}        # host sync, as base/baseModel.py:116
            self.sched.step(values["main"])
            hist.append(values)
            if on_step is not None:
                on_step(i, values)
            if self.early_stop and self.opt.param_groups[0]["lr"] <= 1.1e-8:
                break
        return hist


class DeviceOptimizer:
    """Adam + ReduceLROnPlateau with ALL state on the device (insr_adam_step / insr_plateau_step):
    no host synchronisation per iteration, graph-capturable.  One schedule shared by every net,
    like the reference's single scheduler over all param groups (base/baseModel.py:55-62)."""

    def __init__(self, nets, lr, betas=(0.9, 0.999), eps=1e-8, factor=0.1, patience=500, threshold=1e-4,
                 min_lr=1e-8):
        self.nets = list(nets)
        self.betas, self.eps = betas, eps
        self.factor, self.patience, self.threshold, self.min_lr = factor, patience, threshold, min_lr
        dev = next(self.nets[0].parameters()).device
        self.sched = torch.tensor([lr, float("inf"), 0.0, 0.0], dtype=torch.float32, device=dev)
        self.state = []
        for n in self.nets:
            theta = n.flat_theta()
            self.state.append((n, theta, flat_grad(n), torch.zeros_like(theta), torch.zeros_like(theta)))

    def step(self, main_loss):
        for _, theta, grad, m, v in self.state:
            _ops.adam_step(theta, grad, m, v, self.sched, self.betas[0], self.betas[1], self.eps)
        _ops.plateau_step(main_loss.reshape(1), self.sched, self.factor, self.patience, self.threshold,
                          self.min_lr, 1e-8)

    def update(self, losses, main_index, hist=None, hist_idx=None, zero_grad=True, clear_losses=False):
        """the whole tail of an iteration in ONE kernel (insr_iteration_update): Adam for every net, gradients zeroed for
        the next iteration, plateau schedule on ``losses[main_index]``, the loss values appended to ``hist[hist_idx++]``.
        ``losses``: contiguous fp32 device vector of this iteration's loss terms (``clear_losses``: zeroed afterwards)."""
        if not hasattr(self, "_ticket"):
            self._ticket = torch.zeros(1, dtype=torch.int32, device=self.sched.device)
        lib = _ops._lib.get_lib()
        dev = self.sched.device
        with _ops._DeviceGuard(dev):
            lib.iteration_update([t.data_ptr() for _, t, _, _, _ in self.state], [g.data_ptr() for _, _, g, _, _ in self.state],
                                 [m.data_ptr() for _, _, _, m, _ in self.state], [v.data_ptr() for _, _, _, _, v in self.state],
                                 [t.numel() for _, t, _, _, _ in self.state], self.sched.data_ptr(), losses.data_ptr(),
                                 losses.numel(), main_index, _ops._ptr(hist), 0 if hist is None else hist.shape[0],
                                 _ops._ptr(hist_idx), self._ticket.data_ptr(), self.betas[0], self.betas[1], self.eps,
                                 self.factor, self.patience, self.threshold, self.min_lr, 1e-8, zero_grad, _ops._stream(dev),
                                 clear_losses=clear_losses)

    def update_peer(self, peer, losses, main_index, hist=None, hist_idx=None, zero_grad=True, clear_losses=False, scale=None):
        """``update`` with the data-parallel exchange folded in (insr_iteration_update_peer): the gradient buffers of the nets
        and ``losses`` live in this rank's peer allocation (``peer``: a PeerBuffer); the kernel reads every rank's copy over
        NVLink, averages (``scale`` = 1 / world) and applies Adam / schedule / log to the reduced values -- the iteration's
        all-reduce and its optimiser step are ONE kernel, and the reduced gradient is never written."""
        if not hasattr(self, "_losses_red"):
            self._losses_red = torch.zeros(32, dtype=torch.float32, device=self.sched.device)
        lib = _ops._lib.get_lib()
        dev = self.sched.device
        with _ops._DeviceGuard(dev):
            lib.iteration_update_peer(peer.world, peer.rank, peer.bases, peer.bytes, (1.0 / peer.world) if scale is None else scale,
                                      [t.data_ptr() for _, t, _, _, _ in self.state], [g.data_ptr() for _, _, g, _, _ in self.state],
                                      [m.data_ptr() for _, _, _, m, _ in self.state], [v.data_ptr() for _, _, _, _, v in self.state],
                                      [t.numel() for _, t, _, _, _ in self.state], self.sched.data_ptr(), losses.data_ptr(),
                                      losses.numel(), main_index, self._losses_red.data_ptr(), _ops._ptr(hist),
                                      0 if hist is None else hist.shape[0], _ops._ptr(hist_idx), self.betas[0], self.betas[1],
                                      self.eps, self.factor, self.patience, self.threshold, self.min_lr, 1e-8, zero_grad,
                                      clear_losses, _ops._stream(dev))

    def zero_grads(self):
        for _, _, grad, _, _ in self.state:
            grad.zero_()

    def reset(self, lr):
        """fresh optimiser + scheduler state (the reference builds new ones per training loop)"""
        self.sched.copy_(torch.tensor([lr, float("inf"), 0.0, 0.0]))
        for _, _, _, m, v in self.state:
            m.zero_()
            v.zero_()

    @property
    def lr(self):
        return float(self.sched[0])          # host sync


class SharedGradBuffer:
    """[g_theta(net 0) | g_theta(net 1) | ... | loss slots] in ONE fp32 buffer: the nets' flat gradient
    buffers ARE slices of it, so the fused kernels write there directly and the data-parallel exchange of an iteration --
    parameter gradients and the loss values every rank must agree on for the LR schedule -- moves one buffer with no
    flatten copy.  Where the ranks can map each other's memory the buffer is PEER memory (``peer.PeerBuffer``) and the
    exchange is the library's own one-shot reduction over NVLink, fused into the iteration's update kernel
    (``DeviceOptimizer.update_peer``); otherwise a single NCCL all-reduce.  Both record into CUDA graphs."""

    def __init__(self, nets, n_scalars=4, group=None):
        from . import peer as _peer
        sizes = [n.flat_theta().numel() for n in nets]
        padded = [(sz + 3) // 4 * 4 for sz in sizes]            # every slice 16-byte aligned (C ABI requirement)
        dev = nets[0].flat_theta().device
        total = sum(padded) + n_scalars
        self.peer = _peer.PeerBuffer.create(total, dev, group) if dev.type == "cuda" else None
        self.buf = self.peer.data if self.peer is not None else torch.zeros(total, dtype=torch.float32, device=dev)
        self.nets = list(nets)
        off = 0
        for n, sz, pd in zip(nets, sizes, padded):
            n._flat_grad = self.buf[off:off + sz]
            flat_grad(n)                                        # re-link the parameters' .grad views
            off += pd
        self.scalars = self.buf[off:]
        self.group = group

    def allreduce(self, values):
        """values: this rank's loss terms (a vector, or a list of 0-dim tensors; may BE ``scalars``) -> their means over the
        ranks as views of ``scalars``; the gradients are averaged in place by the same NCCL collective"""
        import torch.distributed as tdist
        k = len(values)
        if k > self.scalars.numel():
            raise ValueError(f"{k} loss terms, {self.scalars.numel()} slots")
        if not (torch.is_tensor(values) and values.data_ptr() == self.scalars.data_ptr()):
            self.scalars[:k].copy_(values if torch.is_tensor(values) else torch.stack([v.reshape(()) for v in values]))
        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(self.group) > 1:
            tdist.all_reduce(self.buf, group=self.group)
            self.buf.div_(tdist.get_world_size(self.group))
        return [self.scalars[i] for i in range(k)]

    def close(self):
        """collective: release the peer allocation; the nets go back to private gradient buffers"""
        if self.peer is None:
            return
        for n in self.nets:
            n._flat_grad = None
            for p in n.parameters():
                p.grad = None
        self.buf = self.scalars = None
        self.peer.close()
        self.peer = None


class GraphedLoop:
    """One @_training_loop with the whole iteration -- sampling, fused closures, Adam, LR schedule, loss
    logging -- captured once as a CUDA graph and replayed: the per-iteration host work is one
    cudaGraphLaunch.  Loss values are written to a device ring and read back in bulk; the early-stop test
    (lr <= 1.1e-8, base/baseModel.py:132-134) is evaluated every ``check_every`` iterations."""

    UNROLL = 4                                     # iterations per replayed graph (even: buffer sets alternate)

    def __init__(self, nets, lr, closure, capacity=20000, data_parallel=False, presample=None, prepare=None):
        self.nets, self.closure = list(nets), closure
        # prepare (a stronger form of presample): ``prepare(k)`` fills buffer set k in {0, 1} with an iteration's points AND
        # everything that depends only on them and on FROZEN networks (the closures' targets); ``closure(k)`` then consumes
        # set k.  An iteration runs closure(k) -> update on the main branch and prepare(1 - k) for the next iteration on a
        # parallel branch from its very start, so sampling and the frozen-net kernels leave the critical path altogether:
        # what remains is lsq -> update.  Two graphs (k = 0 / 1) are captured and replayed alternately.
        self.prepare = prepare
        self._k = 0
        self.graph_b = None
        self._prep_stream = None
        # graph_u: UNROLL consecutive iterations as ONE graph (the gap between two graph launches is longer than a dependency
        # edge inside a graph); the single-iteration graphs serve the remainders
        self.graph_u = None
        # presample: draws the NEXT iteration's points into the persistent buffers the closure reads.  It runs on a parallel
        # branch beside the update kernel (the closure's kernels have consumed the current points by then), so the sampling
        # kernel leaves the critical path of the iteration; run() draws the first set.  Iteration i still sees draw i.
        self.presample = presample
        # data_parallel: every rank runs the closure on its shard; ONE exchange per iteration inside the graph
        self.shared = SharedGradBuffer(self.nets) if data_parallel else None
        self.opt = DeviceOptimizer(self.nets, lr)
        self.capacity = capacity
        self.hist = None
        self.idx = torch.zeros(1, dtype=torch.long, device=self.opt.sched.device)
        # the loss slots of an iteration: the closures' kernels accumulate into them (``_acc``), the update kernel consumes
        # and re-zeroes them.  With data parallelism they are the tail of the exchanged buffer.
        self.slots = self.shared.scalars if self.shared is not None else torch.zeros(4, dtype=torch.float32, device=self.opt.sched.device)
        self._slot_off = 0
        self.graph = None
        self.graph_ready = False                   # True once the capture has completed

    def _provide(self, n, device):
        if device != self.slots.device or self._slot_off + n > self.slots.numel():
            return torch.zeros(n, dtype=torch.float32, device=device)
        out = self.slots[self._slot_off:self._slot_off + n]
        self._slot_off += n
        return out

    def _iteration(self, k=0):
        if self.prepare is None:
            return self._iteration_body(self.closure)
        # fork at the start of the iteration: the next iteration's buffer set is prepared beside this one's kernels
        dev = self.slots.device
        cur = torch.cuda.current_stream(dev)
        if self._prep_stream is None:
            self._prep_stream = torch.cuda.Stream(device=dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        self._prep_stream.wait_event(fork)
        with torch.cuda.stream(self._prep_stream):
            self.prepare(1 - k)
        keys = self._iteration_body(lambda: self.closure(k))
        done = torch.cuda.Event()
        done.record(self._prep_stream)
        cur.wait_event(done)
        return keys

    def _iteration_body(self, closure):
        # the gradient buffers and the loss slots are zero here: zeroed by run() before the first iteration, then by every
        # update kernel
        self._slot_off = 0
        _ACC_PROVIDER.append(self._provide)
        try:
            loss_dict = closure()
        finally:
            _ACC_PROVIDER.pop()
        _backward_if_needed(loss_dict)
        keys = list(loss_dict)
        k = len(keys)
        vector = loss_dict.vector if isinstance(loss_dict, Losses) else None
        own = vector is not None and vector.data_ptr() == self.slots.data_ptr() and vector.numel() == k
        if self.hist is None:
            self.hist = torch.zeros(self.capacity, k, device=self.slots.device)
        if self.shared is not None and self.shared.peer is not None:
            # exchange + Adam (all nets) + zero_grad + plateau + loss log: ONE kernel over peer memory
            if not own:
                self.slots[:k].copy_(vector if vector is not None else torch.stack([loss_dict[q].reshape(()) for q in keys]))
            self._tail(lambda: self.opt.update_peer(self.shared.peer, self.slots[:k], keys.index("main"), self.hist, self.idx,
                                                    clear_losses=own))
            return keys
        if self.shared is not None:
            self.shared.allreduce(self.slots[:k] if own else (vector if vector is not None else [loss_dict[q] for q in keys]))
            vals = self.shared.scalars[:k]                     # averaged over the ranks, already contiguous
        else:
            vals = vector if vector is not None else torch.stack([loss_dict[q].reshape(()) for q in keys])
        # Adam (all nets) + zero_grad + plateau + loss log: one kernel
        self._tail(lambda: self.opt.update(vals, keys.index("main"), self.hist, self.idx, clear_losses=own))
        return keys

    def _tail(self, update):
        if self.presample is None:
            update()
        else:
            parallel(self.slots, update, self.presample)

    def reset(self, lr):
        """reuse the captured graph for a new training loop: only the optimiser state and the log restart"""
        self.opt.reset(lr)
        self.idx.zero_()

    def close(self):
        """drop the captured graph NOW (it holds the kernels -- with data_parallel, the exchange -- of one iteration): a
        process group must not be destroyed, and peer memory not released, while a captured graph still references it.
        Collective when data_parallel (every rank closes its loops in the same order)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph = self.graph_b = self.graph_u = None
        if self.shared is not None:
            self.shared.close()

    def run(self, n_iters, early_stop=False, check_every=100):
        self.opt.zero_grads()                          # whatever ran on these nets in between may have left gradients
        self.slots.zero_()
        if self.presample is not None:
            self.presample()                           # the first iteration's points
        if self.prepare is not None:
            self.prepare(0)                            # the first iteration's buffer set
            self._k = 0
        if self.graph is None:
            keys = self._iteration(0)                  # iteration 0 eagerly (also warms everything up)
            self._k = 1
            self.keys = keys
            done = 1
            if self.shared is not None and self.shared.peer is not None:
                # the peer kernels give up (and say so) instead of hanging when a rank never shows up; find out now, on
                # every rank together, rather than replaying a graph whose every iteration would wait for the time-out
                import torch.distributed as tdist
                ok = torch.tensor([1.0 if self.shared.peer.healthy() else 0.0], device=self.slots.device)
                tdist.all_reduce(ok, op=tdist.ReduceOp.MIN, group=self.shared.group)
                if float(ok.item()) < 0.5:
                    raise RuntimeError("insr_pde_b200: the peer-memory exchange timed out on at least one rank "
                                       "(set INSR_PEER_ALLREDUCE=0 for the NCCL exchange)")
        else:
            keys, done = self.keys, 0
        if n_iters > done:
            if self.graph is None:
                # a garbage-collection pass that happens to free an OLD CUDAGraph (a dropped stepper is a reference
                # cycle) while this capture is open calls cudaGraphExecDestroy / cudaFree and invalidates the capture:
                # collect now, and keep the collector off until the capture has ended
                gc.collect()
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                gc_was_on = gc.isenabled()
                gc.disable()
                try:
                    with torch.cuda.graph(self.graph):
                        self._iteration(0)
                    if self.prepare is not None:       # the odd iterations: buffer set 1 consumed, set 0 prepared
                        self.graph_b = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(self.graph_b):
                            self._iteration(1)
                    if self.UNROLL > 1 and n_iters - done >= 2 * self.UNROLL:
                        self.graph_u = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(self.graph_u):
                            for u in range(self.UNROLL):
                                self._iteration(u & 1)
                    self.graph_ready = True
                finally:
                    if gc_was_on:
                        gc.enable()
            while done < n_iters:
                burst = min(check_every, n_iters - done)
                i = 0
                while i < burst:
                    if self.graph_u is not None and self._k == 0 and burst - i >= self.UNROLL:
                        self.graph_u.replay()
                        i += self.UNROLL
                    elif self.prepare is None:
                        self.graph.replay()
                        i += 1
                    else:
                        (self.graph_b if self._k else self.graph).replay()
                        self._k ^= 1
                        i += 1
                done += burst
                if early_stop and self.opt.lr <= 1.1e-8:
                    break
        h = self.hist[:done].cpu()
        return [{k: float(h[i, j]) for j, k in enumerate(keys)} for i in range(done)]


class _StepperBase:
    def close(self):
        """release every captured iteration graph of this stepper (see GraphedLoop.close)"""
        for lp in self.__dict__.get("_loops", {}).values():
            lp.close()
        self.__dict__.pop("_loops", None)


class FluidStepper(_StepperBase):
    """Fluid2DModel.step (fluid/model.py:61-70) on the fused closures, with the reference's sampling
    (base/sampling.py) and schedule; used for the seconds-per-timestep measurement."""

    def __init__(self, velocity, velocity_prev, pressure, dt=0.05, sample_resolution=128, lr=1e-4, reducer_factory=None,
                 graphed=False, device_sampler=False, seed=0):
        self.vel, self.prev, self.pres = velocity, velocity_prev, pressure
        self.dt, self.sr, self.lr = dt, sample_resolution, lr
        self.reducer_factory = reducer_factory
        self.graphed = graphed
        self.data_parallel = False                 # set by the caller for world > 1 with graphed=True
        self.device_sampler, self.seed, self._samplers, self._points = device_sampler, seed, {}, {}
        for p in self.prev.parameters():
            p.requires_grad_(False)

    def _sampler(self, n_shard_div):
        """the one-kernel sampler of the three point sets (insr_sample_boxes) and its two persistent point buffers"""
        from . import sampling
        if n_shard_div not in self._samplers:
            import torch.distributed as tdist
            dev = next(self.vel.parameters()).device
            n = self.sr ** 2
            rank = tdist.get_rank() if (tdist.is_available() and tdist.is_initialized()) else 0
            sets = sampling.fluid_sets(n // n_shard_div, n // 100)
            per_rank = sum(int(bx[0]) for st in sets for bx in st)
            # every rank draws its own slice of ONE global Philox stream (point index offset by the rank)
            self._samplers[n_shard_div] = sampling.BoxSampler(sets, 2, seed=self.seed, device=dev, point_offset=rank * per_rank)
            self._points[n_shard_div] = [torch.zeros(per_rank, 2, dtype=torch.float32, device=dev) for _ in range(2)]
        return self._samplers[n_shard_div], self._points[n_shard_div]

    def _samples(self, n_shard_div=1):
        from . import sampling
        dev = next(self.vel.parameters()).device
        n = self.sr ** 2
        if self.device_sampler:
            smp, bufs = self._sampler(n_shard_div)
            return tuple(smp.sample(out=bufs[0]))
        x = sampling.sample_random(n // n_shard_div, 2, device=dev)
        bx = sampling.sample_boundary2D_separate(n // 100, "horizontal", device=dev)
        by = sampling.sample_boundary2D_separate(n // 100, "vertical", device=dev)
        return x, bx, by                     # (same RNG stream order as the reference: not run in parallel)

    def _loop(self, nets, key, target, closure, n_iters, world=1):
        """one training loop.  ``target(x, out)``: the frozen side of the closure at the interior points (into ``out`` where
        given); ``closure(x, bx, by, tgt)``: the trainable side (``tgt`` None: compute the target inline)."""
        def eager(i):
            x, bx, by = self._samples(world)
            return closure(x, bx, by, None)
        if self.graphed:
            loops = self.__dict__.setdefault("_loops", {})
            if key not in loops:                       # capture once per closure kind, replay for every time step
                if self.device_sampler:
                    # buffer sets 0 / 1: an iteration consumes one while the next one's points AND frozen-net target are
                    # prepared into the other on a parallel branch (GraphedLoop.prepare)
                    smp, bufs = self._sampler(world)
                    tgts = [None, None]

                    def prepare(k):
                        pts = smp.sample(out=bufs[k])
                        tgts[k] = target(pts[0], tgts[k])

                    def consume(k):
                        x, bx, by = torch.split(bufs[k], smp.sizes, dim=0)
                        return closure(x, bx, by, tgts[k])
                    loops[key] = GraphedLoop(nets, self.lr, consume, data_parallel=self.data_parallel, prepare=prepare)
                else:
                    loops[key] = GraphedLoop(nets, self.lr, lambda: eager(0), data_parallel=self.data_parallel)
            else:
                loops[key].reset(self.lr)
            return loops[key].run(n_iters)
        if self.data_parallel and not self.reducer_factory:
            raise RuntimeError("FluidStepper: data_parallel needs graphed=True (all-reduce inside the iteration graph) or a "
                               "reducer_factory for the eager loop -- refusing to train unsynchronised replicas")
        red = self.reducer_factory(nets) if self.reducer_factory else None
        return TrainingLoop(nets, self.lr, reducer=red).run(eager, n_iters)

    def initialize(self, init_fn, n_iters, world=1):
        return self._loop([self.vel], "initialize", lambda x, out: _into(out, init_fn(x)) if out is not None else init_fn(x).contiguous(),
                          lambda x, bx, by, tgt: fluid_initialize(self.vel, x, tgt if tgt is not None else init_fn(x)),
                          n_iters, world=world)

    def step(self, n_iters, world=1):
        """advect -> pressure solve -> projection; returns the three loss histories"""
        nets = [self.vel, self.pres]
        self.prev.load_state_dict(self.vel.state_dict())
        h1 = self._loop(nets, "advect", lambda x, out: fluid_advect_target(self.prev, x, self.dt, out=out),
                        lambda x, bx, by, tgt: fluid_advect_velocity(self.vel, self.prev, x, bx, by, self.dt, target=tgt),
                        n_iters, world=world)
        h2 = self._loop(nets, "pressure", lambda x, out: fluid_pressure_target(self.vel, x, out=out),
                        lambda x, bx, by, tgt: fluid_solve_pressure(self.vel, self.pres, x, bx, by, target=tgt),
                        n_iters, world=world)
        self.prev.load_state_dict(self.vel.state_dict())
        h3 = self._loop(nets, "project", lambda x, out: fluid_projection_target(self.prev, self.pres, x, out=out),
                        lambda x, bx, by, tgt: fluid_projection(self.vel, self.prev, self.pres, x, bx, by, target=tgt),
                        n_iters, world=world)
        return h1, h2, h3


class AdvectionStepper(_StepperBase):
    """Advection1DModel.initialize / step (advection/model.py:37-91): the Gaussian fit, then per time step the previous-frame
    hand-over and one training loop of the midpoint-residual closure; points as the reference draws them
    (``sample_random(sr, 1) * length / 2`` and the 2 x eps boundary bands) -- from one Philox kernel when ``graphed``."""

    def __init__(self, field, field_prev, dt=0.05, vel=0.25, length=4.0, sample_resolution=5000, lr=1e-4, graphed=False, seed=0):
        self.field, self.prev = field, field_prev
        self.dt, self.vel, self.length, self.sr, self.lr = dt, vel, length, sample_resolution, lr
        self.graphed, self.seed, self._sampler = graphed, seed, None
        for p in self.prev.parameters():
            p.requires_grad_(False)

    def _device_sampler(self):
        """base/sampling.py:21-37 scaled by length / 2 (bands around -half and +half) as ONE Philox kernel, and the two
        persistent point buffers of the graphed loop"""
        from . import sampling
        if self._sampler is None:
            dev = next(self.field.parameters()).device
            nb, half, eps = max(self.sr // 100, 10), self.length / 2, 1e-4
            sets = [[(self.sr, (-half,), (half,))],
                    [(nb // 2, ((-1 - eps) * half,), ((-1 + eps) * half,)), (nb // 2, ((1 - eps) * half,), ((1 + eps) * half,))]]
            self._sampler = sampling.BoxSampler(sets, 1, seed=self.seed, device=dev)
            self._points = [torch.zeros(sum(self._sampler.sizes), 1, dtype=torch.float32, device=dev) for _ in range(2)]
        return self._sampler, self._points

    def _samples(self):
        from . import sampling
        dev = next(self.field.parameters()).device
        nb, half = max(self.sr // 100, 10), self.length / 2
        if self.graphed:
            smp, bufs = self._device_sampler()
            return tuple(smp.sample(out=bufs[0]))
        return (sampling.sample_random(self.sr, 1, device=dev) * half, sampling.sample_boundary(nb, 1, device=dev) * half)

    def _loop(self, key, target, closure, n_iters):
        """``target(x, out)``: the frozen side of the closure; ``closure(x, xb, tgt)``: the trainable side"""
        if self.graphed:
            loops = self.__dict__.setdefault("_loops", {})
            if key not in loops:
                # buffer sets 0 / 1: an iteration consumes one while the next one's points and frozen-field target are prepared
                # into the other on a parallel branch (GraphedLoop.prepare)
                smp, bufs = self._device_sampler()
                tgts = [None, None]

                def prepare(k):
                    pts = smp.sample(out=bufs[k])
                    tgts[k] = target(pts[0], tgts[k])

                def consume(k):
                    x, xb = torch.split(bufs[k], smp.sizes, dim=0)
                    return closure(x, xb, tgts[k])
                loops[key] = GraphedLoop([self.field], self.lr, consume, prepare=prepare)
            else:
                loops[key].reset(self.lr)
            return loops[key].run(n_iters)

        def eager(i):
            x, xb = self._samples()
            return closure(x, xb, None)
        return TrainingLoop([self.field], self.lr).run(eager, n_iters)

    def initialize(self, init_fn, n_iters):
        return self._loop("initialize", lambda x, out: _into(out, init_fn(x)) if out is not None else init_fn(x).contiguous(),
                          lambda x, xb, tgt: advect_initialize(self.field, x, tgt if tgt is not None else init_fn(x)), n_iters)

    def step(self, n_iters):
        self.prev.load_state_dict(self.field.state_dict())
        return self._loop("advect", lambda x, out: advect_target(self.prev, x, self.dt, self.vel, out=out),
                          lambda x, xb, tgt: advect_step(self.field, self.prev, x, xb, self.dt, self.vel, target=tgt), n_iters)


def gaussian_like(x, mu=-1.5, sigma=0.1):
    """advection/examples.py:14-16 with example1's mu"""
    return torch.exp(-0.5 * (x - mu) ** 2 / (sigma ** 2))


class ElasticityBatch:
    """The point batch of one elasticity iteration in ONE persistent buffer (graphed stepper):

        x_all = [ constant interior | random interior | uniform left | random left | random right | uniform right ]

    The constant part (the reference's 'uniform' grid, or every mesh vertex: elasticity/model.py:203-211) is written once;
    the random parts are drawn in place by the Philox samplers every iteration (no torch.cat in the iteration graph).
    ``y_prev`` / ``y_pp`` hold the two previous-frame fields on the interior rows: the constant rows are evaluated once
    per time step (``refresh``), the random rows every iteration."""

    def __init__(self, stepper, resolution, use_left, use_right):
        from . import sampling
        st, dev, dim = stepper, stepper._device(), stepper.dim
        self.dim = dim
        const, n_rand = [], 0
        for kind in st.pattern:
            if kind == "uniform":
                const.append(st.mesh[0][:, :dim].to(dev).float() if st.mesh is not None else sampling.sample_uniform(resolution, dim, device=dev))
            elif kind == "random":
                n_rand += resolution ** dim
            else:
                raise NotImplementedError(kind)
        self.n_const = sum(c.shape[0] for c in const)
        self.n = self.n_const + n_rand
        faces = st.mesh is None
        lu = ru = lr = rr = 0
        if faces:
            for kind in st.pattern:
                if kind == "uniform":
                    lu = ru = resolution ** (dim - 1)
                elif kind == "random":
                    lr = rr = resolution
        if not use_left:
            lu = lr = 0
        if not use_right:
            ru = rr = 0
        self.n_left, self.n_right = lu + lr, rr + ru
        self.x_all = torch.zeros(self.n + self.n_left + self.n_right, dim, dtype=torch.float32, device=dev)
        if const:
            self.x_all[:self.n_const] = torch.cat(const, dim=0)
        o = self.n
        if lu:
            face = sampling.sample_uniform(resolution, dim - 1, device=dev)
            self.x_all[o:o + lu] = torch.cat((-torch.ones(lu, 1, device=dev), face), dim=1)
        self._rand_faces = (o + lu, o + lu + lr + rr)
        if ru:
            face = sampling.sample_uniform(resolution, dim - 1, device=dev)
            self.x_all[o + lu + lr + rr:] = torch.cat((torch.ones(ru, 1, device=dev), face), dim=1)
        self.y_prev = torch.zeros(self.n, dim, dtype=torch.float32, device=dev)
        self.y_pp = torch.zeros(self.n, dim, dtype=torch.float32, device=dev)
        # samplers (one kernel for the interior, one for both faces)
        self._interior = None
        if n_rand:
            if st.mesh is not None:
                ms = sampling.MeshSampler(st.mesh[0].to(dev), st.mesh[1].to(dev), dim, st.seed)
                self._interior = lambda out: ms.sample(n_rand, out=out)
            else:
                bs = sampling.BoxSampler([[(n_rand, (-1.0,) * dim, (1.0,) * dim)]], dim, seed=st.seed, device=dev)
                self._interior = lambda out: bs.sample(out=out)
        self._faces = None
        if lr + rr:
            lo, hi = (-1.0,) * (dim - 1), (1.0,) * (dim - 1)
            boxes = ([[(lr, (-1.0,) + lo, (-1.0,) + hi)]] if lr else []) + ([[(rr, (1.0,) + lo, (1.0,) + hi)]] if rr else [])
            fs = sampling.BoxSampler(boxes, dim, seed=st.seed + 1, device=dev)
            self._faces = lambda out: fs.sample(out=out)

    @property
    def aligned(self):
        """the row ranges handed to the kernels start on 16-byte boundaries"""
        return (self.n_const * self.dim * 4) % 16 == 0

    def draw(self):
        if self._interior is not None:
            self._interior(self.x_all[self.n_const:self.n])
        if self._faces is not None:
            a, b = self._rand_faces
            self._faces(self.x_all[a:b])

    def refresh(self, prev, pp):
        """previous-frame fields on the constant rows: once per time step (they change only with the hand-over)"""
        if self.n_const:
            xc = self.x_all[:self.n_const]
            self.y_prev[:self.n_const] = evaluate(prev, xc, ORDER_VALUE)[0]
            self.y_pp[:self.n_const] = evaluate(pp, xc, ORDER_VALUE)[0]


class ElasticityStepper(_StepperBase):
    """ElasticityModel.initialize / step (elasticity/model.py:100-125) on the fused closure: previous-frame hand-over,
    the reference's sample pattern ('random' and / or 'uniform'; a mesh when ``mesh=(V, F)`` is given), one
    @_training_loop per time step.  ``graphed=True`` captures the whole iteration (sampling kernel, field kernels, energy
    kernel, autograd reverse sweep, device Adam + plateau) as one CUDA graph that every later time step replays."""

    def __init__(self, deformation, prev, prev_prev, dim, dt=0.05, sample_resolution=100, lr=1e-4,
                 sample_pattern=("random", "uniform"), mesh=None, graphed=False, seed=0, **closure_kw):
        self.defo, self.prev, self.pp = deformation, prev, prev_prev
        self.dim, self.dt, self.sr, self.lr = dim, dt, sample_resolution, lr
        self.pattern, self.mesh, self.graphed, self.seed = tuple(sample_pattern), mesh, graphed, seed
        self.kw = closure_kw
        self.timestep = 0
        self._samplers, self._const = {}, {}
        for net in (self.prev, self.pp):
            for p in net.parameters():
                p.requires_grad_(False)

    # --- sampling (elasticity/model.py:198-253) --------------------------------------------------
    def _device(self):
        return next(self.defo.parameters()).device

    def _interior(self, resolution):
        from . import sampling
        dev, parts = self._device(), []
        for kind in self.pattern:
            if kind == "random":
                if self.mesh is not None:
                    key = ("mesh",)
                    if key not in self._samplers:
                        self._samplers[key] = sampling.MeshSampler(self.mesh[0].to(dev), self.mesh[1].to(dev), self.dim, self.seed)
                    parts.append(self._samplers[key].sample(resolution ** self.dim))
                elif self.graphed:
                    key = ("box", resolution)
                    if key not in self._samplers:
                        box = [(resolution ** self.dim, (-1.0,) * self.dim, (1.0,) * self.dim)]
                        self._samplers[key] = sampling.BoxSampler([box], self.dim, seed=self.seed, device=dev)
                    parts.append(self._samplers[key].sample()[0])
                else:
                    parts.append(sampling.sample_random(resolution ** self.dim, self.dim, device=dev))
            elif kind == "uniform":
                key = ("uniform", resolution)
                if key not in self._const:
                    self._const[key] = (self.mesh[0][:, :self.dim].to(dev).float().contiguous() if self.mesh is not None
                                        else sampling.sample_uniform(resolution, self.dim, device=dev))
                parts.append(self._const[key])
            else:
                raise NotImplementedError(kind)
        return torch.cat(parts, dim=0).requires_grad_(True)

    def _fixed(self, resolution):
        from . import sampling
        dev, left, right = self._device(), [], []
        if self.mesh is not None:
            return [], []
        for kind in self.pattern:
            if kind == "random":
                if self.graphed:               # both faces from one kernel: a box that is flat in the first coordinate
                    key = ("fixed", resolution)
                    if key not in self._samplers:
                        lo, hi = (-1.0,) * (self.dim - 1), (1.0,) * (self.dim - 1)
                        sets = [[(resolution, (-1.0,) + lo, (-1.0,) + hi)], [(resolution, (1.0,) + lo, (1.0,) + hi)]]
                        self._samplers[key] = sampling.BoxSampler(sets, self.dim, seed=self.seed + 1, device=dev)
                    l, r = self._samplers[key].sample()
                else:
                    l = torch.cat((-torch.ones(resolution, 1, device=dev), sampling.sample_random(resolution, self.dim - 1, device=dev)), dim=1)
                    r = torch.cat((torch.ones(resolution, 1, device=dev), sampling.sample_random(resolution, self.dim - 1, device=dev)), dim=1)
                left.append(l); right.append(r)
            elif kind == "uniform":
                key = ("fixed_uniform", resolution)
                if key not in self._const:
                    face = sampling.sample_uniform(resolution, self.dim - 1, device=dev)
                    one = torch.ones(face.shape[0], 1, device=dev)
                    self._const[key] = (torch.cat((-one, face), dim=1), torch.cat((one, face), dim=1))
                left.append(self._const[key][0]); right.append(self._const[key][1])
        return torch.cat(left, dim=0), torch.cat(right, dim=0)

    def _batch(self):
        """the persistent batch of the graphed iteration, or None where the one-kernel closure does not apply"""
        if "_ebatch" not in self.__dict__:
            energy = list(self.kw.get("energy", ()))
            known = {"arap", "volume", "kinematics", "external", "constraint", "constraint_right", "constraint_right_compress",
                     "collision", "collision_sphere"}
            ok = (set(energy) <= known and len(set(energy)) == len(energy) and not ("collision_sphere" in energy and self.dim != 2)
                  and not ("constraint_right" in energy and "constraint_right_compress" in energy))
            b = None
            if ok:
                b = ElasticityBatch(self, self.sr, "constraint" in energy,
                                    "constraint_right" in energy or "constraint_right_compress" in energy)
                if not b.aligned:
                    b = None
            self._ebatch = b
        return self._ebatch

    # --- the two training loops -----------------------------------------------------------------
    def _loop(self, closure, n_iters, key, presample=None):
        if self.graphed:
            loops = self.__dict__.setdefault("_loops", {})
            if key not in loops:
                loops[key] = GraphedLoop([self.defo], self.lr, lambda: closure(0), presample=presample)
            else:
                loops[key].reset(self.lr)
            return loops[key].run(n_iters)
        return TrainingLoop([self.defo], self.lr).run(closure, n_iters)

    def initialize(self, n_iters, resolution=None):
        """fit the zero deformation (elasticity/model.py:108-117), then copy it into both previous frames (:100-105)"""
        res = resolution or self.sr

        def c(i):
            return {"main": torch.mean(self.defo(self._interior(res).detach()) ** 2)}      # widths above the lsq kernels
        hist = self._loop(c, n_iters, ("initialize", res))
        self.pp.load_state_dict(self.defo.state_dict())
        self.prev.load_state_dict(self.defo.state_dict())
        return hist

    def step(self, n_iters):
        self.timestep += 1
        self.pp.load_state_dict(self.prev.state_dict())
        self.prev.load_state_dict(self.defo.state_dict())
        forced = self.timestep <= self.kw.get("external_force_timesteps", 0)
        batch = self._batch() if self.graphed else None
        if batch is not None:
            batch.refresh(self.prev, self.pp)

        def c(i):
            if batch is not None:                      # its random rows are drawn ahead by the loop (presample = batch.draw)
                return elasticity_solve_deformation(self.defo, self.prev, self.pp, None, None, None, dt=self.dt,
                                                    timestep=self.timestep, batch=batch, **self.kw)
            left, right = self._fixed(self.sr)
            return elasticity_solve_deformation(self.defo, self.prev, self.pp, self._interior(self.sr), left, right,
                                                dt=self.dt, timestep=self.timestep, **self.kw)
        # the external-force term is a host-side branch on the time step: one captured graph per branch
        return self._loop(c, n_iters, ("solve", forced), presample=batch.draw if batch is not None else None)


def taylorgreen_velocity(samples):
    """fluid/examples.py:17-31 with rescale=True"""
    px = (samples[..., 0] + 1) * math.pi
    py = (samples[..., 1] + 1) * math.pi
    return torch.stack([torch.sin(px) * torch.cos(py) / math.pi, -torch.cos(px) * torch.sin(py) / math.pi], dim=-1)


def taylorgreen_multi_velocity(samples, scale=8, gap=0.05):
    """fluid/examples.py:34-52 (init_cond 'taylorgreen_multi', scripts/fluid2DtlgnM.sh): a Taylor-Green vortex squeezed into
    the lower-left quadrant and a small one into the upper-right corner, each faded out over a thin band.  Same values as
    the reference's masked assignments, written with ``torch.where`` (fixed shapes: graph-capturable)."""
    def tg(p):                                     # taylorgreen_velocity(..., rescale=False)
        px, py = (p[..., 0] + 1) * math.pi, (p[..., 1] + 1) * math.pi
        return torch.stack([torch.sin(px) * torch.cos(py), -torch.cos(px) * torch.sin(py)], dim=-1)

    x, y = samples[..., 0], samples[..., 1]
    in1 = torch.logical_and(x <= gap, y <= gap)
    w1 = 1.0 - samples.clamp(min=0, max=gap).norm(dim=-1) / gap
    v1 = tg(torch.clamp(samples * 2 + 1, min=-1, max=1)) * w1.unsqueeze(-1)
    p = 1 - 2 / scale
    gap2 = gap * 2 / scale
    in2 = torch.logical_and(x > p - gap2, y > p - gap2)
    w2 = 1.0 - (p - samples).clamp(min=0, max=gap2).norm(dim=-1) / gap2
    v2 = tg(torch.clamp(samples * scale + (1 - scale), min=-1, max=1)) * w2.unsqueeze(-1)
    vel = torch.where(in1.unsqueeze(-1), v1, torch.zeros_like(samples))
    return torch.where(in2.unsqueeze(-1), v2, vel)
