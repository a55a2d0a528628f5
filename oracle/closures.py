"""Restated loss closures of the three PDE families, on EXPLICIT sample tensors.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference draws its
collocation points inside each closure; here the points are arguments so that the oracle,
the real reference (goldens) and the CUDA path can be fed identical inputs.  ``ops`` is any
namespace exposing ``gradient / divergence / laplace / jacobian`` with the signatures of
``base/diff_ops.py`` (``oracle.torch_port`` for the oracle, ``insr_pde_b200.diff_ops`` for the
product under test).  Networks are callables ``net(x) -> y``.

Follows (relative to /root/reference):
  advection/examples.py:14-16   gaussian_like
  advection/model.py:43-52      _initialize
  advection/model.py:68-91      _advect
  fluid/examples.py:17-31       taylorgreen_velocity
  fluid/model.py:43-52          _initialize
  fluid/model.py:72-101         _advect_velocity
  fluid/model.py:103-125        _solve_pressure
  fluid/model.py:127-151        _projection
  elasticity/model.py:109-117   _initialize
  elasticity/model.py:127-189   _solve_deformation
  elasticity/losses.py:6-39     positional_constraint / collision_plane / collision_sphere
"""
from __future__ import annotations

import math

import torch


# ------------------------------------------------------------------ initial conditions
def gaussian_like(x, mu=-1.5, sigma=0.1):
    return torch.exp(-0.5 * (x - mu) ** 2 / (sigma ** 2))


def taylorgreen_velocity(samples, rescale=True):
    px = (samples[..., 0] + 1) * math.pi
    py = (samples[..., 1] + 1) * math.pi
    u = torch.sin(px) * torch.cos(py)
    v = -torch.cos(px) * torch.sin(py)
    if rescale:
        u, v = u / math.pi, v / math.pi
    return torch.stack([u, v], dim=-1)


# ------------------------------------------------------------------ advection
def advect_initialize(field, samples, init_fn=gaussian_like):
    return {"main": torch.nn.functional.mse_loss(field(samples), init_fn(samples))}


def advect_step(field, field_prev, ops, samples, boundary_samples, dt, vel):
    u_prev = field_prev(samples)
    u = field(samples)
    dudt = (u - u_prev) / dt
    du = ops.gradient(u, samples)
    du_prev = ops.gradient(u_prev, samples).detach()
    main = torch.mean((dudt + vel * (du + du_prev) / 2.0) ** 2)
    bc = torch.mean(field(boundary_samples) ** 2) * 1.0
    return {"main": main, "bc": bc}


# ------------------------------------------------------------------ fluid
def _no_slip_bc(velocity, bc_x, bc_y):
    vx = velocity(bc_x)[..., 0]
    vy = velocity(bc_y)[..., 1]
    return (torch.mean(vx ** 2) + torch.mean(vy ** 2)) * 1.0


def fluid_initialize(velocity, samples, init_fn=taylorgreen_velocity):
    return {"main": torch.nn.functional.mse_loss(velocity(samples), init_fn(samples))}


def fluid_advect_velocity(velocity, velocity_prev, samples, bc_x, bc_y, dt):
    with torch.no_grad():
        u_prev = velocity_prev(samples).detach()
    u = velocity(samples)
    back = torch.clamp(samples - u_prev * dt, min=-1.0, max=1.0)
    with torch.no_grad():
        u_adv = velocity_prev(back).detach()
    return {"main": torch.mean((u - u_adv) ** 2), "bc": _no_slip_bc(velocity, bc_x, bc_y)}


def fluid_solve_pressure(velocity, pressure, ops, samples, bc_x, bc_y):
    div_u = ops.divergence(velocity(samples), samples).detach()
    lap_p = ops.laplace(pressure(samples), samples)
    main = torch.mean((div_u - lap_p) ** 2)
    gpx = ops.gradient(pressure(bc_x), bc_x)[..., 0]
    gpy = ops.gradient(pressure(bc_y), bc_y)[..., 1]
    return {"main": main, "bc": torch.mean(gpx ** 2) + torch.mean(gpy ** 2)}


def fluid_projection(velocity, velocity_prev, pressure, ops, samples, bc_x, bc_y):
    with torch.no_grad():
        u_prev = velocity_prev(samples).detach()
    grad_p = ops.gradient(pressure(samples), samples).detach()
    target = u_prev - grad_p
    return {"main": torch.mean((velocity(samples) - target) ** 2),
            "bc": _no_slip_bc(velocity, bc_x, bc_y)}


# ------------------------------------------------------------------ elasticity
def positional_constraint_loss(q_fixed, q_target, ratio):
    return ratio * torch.sum((q_fixed - q_target) ** 2)


def collision_plane_loss(q, qdot, dt, ratio, plane_height):
    hit = q[:, -1] < plane_height
    if int(hit.sum()) == 0:
        return 0
    depth = plane_height - q[hit][:, -1]
    force = ratio * torch.column_stack((torch.zeros(depth.shape[0], q.shape[1] - 1, device=q.device), depth))
    return -dt * torch.sum(qdot[hit] * force)


def collision_sphere_loss(q, qdot, dt, ratio, center, radius):
    vec = q - center
    dist = torch.sqrt(torch.sum(vec ** 2, dim=1))
    direction = vec / dist[:, None]
    hit = dist < radius
    if int(hit.sum()) == 0:
        return 0
    if q.shape[1] == 2:
        force = ratio * dist[hit][:, None] * direction[hit]
    else:  # the reference broadcasts to (n, n, 3) in 3-D (elasticity/losses.py:38) -- kept
        force = ratio * dist[hit][:, None, None] * direction[hit]
    return -dt * torch.sum(qdot[hit] * force)


def elasticity_initialize(deformation, samples):
    return {"main": torch.mean(deformation(samples) ** 2)}


def elasticity_solve_deformation(deformation, prev, prev_prev, ops, samples, fixed_left, fixed_right,
                                 *, dt, timestep, energy, ratio_arap, ratio_volume, ratio_kinematics,
                                 ratio_constraint, ratio_collide, external_force, external_force_timesteps,
                                 constraint_offset_right, plane_height, circle_center, circle_radius):
    with torch.no_grad():
        q_prev = prev(samples) + samples
        q_pp = prev_prev(samples) + samples
    q = deformation(samples) + samples
    qdot = (q - q_prev) / dt
    qdot_prev = (q_prev - q_pp) / dt

    F, _ = ops.jacobian(q, samples)
    _, sing, _ = torch.svd(F)
    E_arap = ratio_arap * torch.sum((sing - 1.0) ** 2)
    E_volume = ratio_volume * torch.sum((torch.prod(sing, dim=1) - 1) ** 2)
    E_kin = ratio_kinematics * torch.sum((qdot - qdot_prev) ** 2)
    E_ext = -dt * torch.sum(qdot * external_force.repeat(samples.shape[0], 1))

    loss = 0
    for term in energy:
        if term == "arap":
            loss = loss + E_arap
        elif term == "volume":
            loss = loss + E_volume
        elif term == "kinematics":
            loss = loss + E_kin
        elif term == "external":
            if timestep <= external_force_timesteps:
                loss = loss + E_ext
        elif term == "constraint":
            loss = loss + positional_constraint_loss(deformation(fixed_left), 0, ratio_constraint)
        elif term == "constraint_right":
            tgt = constraint_offset_right.repeat(fixed_right.shape[0], 1)
            loss = loss + positional_constraint_loss(deformation(fixed_right), tgt, ratio_constraint)
        elif term == "constraint_right_compress":
            tgt = -constraint_offset_right.repeat(fixed_right.shape[0], 1)
            loss = loss + positional_constraint_loss(deformation(fixed_right), tgt, ratio_constraint)
        elif term == "collision":
            loss = loss + collision_plane_loss(q, qdot, dt, ratio_collide, plane_height)
        elif term == "collision_sphere":
            loss = loss + collision_sphere_loss(q, qdot, dt, ratio_collide, circle_center, circle_radius)
        else:
            raise NotImplementedError(term)
    return {"main": loss}
