"""Pin the oracle against the golden vectors generated from the REAL reference
(oracle/make_goldens.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import closures, siren_fwdmode as fm, torch_port as tp

OPS = ["advect1d", "fluid_vel", "fluid_pres", "elas2d", "bunny3d", "depth0", "deep5", "wide128"]


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-30))


@pytest.mark.parametrize("name", OPS)
def test_fwdmode_fp64_matches_reference_fp64(name):
    g = load_golden("op_" + name)
    D, O, H, L, N = (int(v) for v in g["shape"])
    th, x = g["theta"].astype(np.float64), g["x"].astype(np.float64)
    out = fm.forward(th, x, D, O, H, L, fm.ORDER_HESS)
    lap = fm.forward(th, x, D, O, H, L, fm.ORDER_LAP)["lap"]
    assert rel(out["y"], g["y_f64"]) < 1e-12
    assert rel(out["jac"], g["jacx_f64"]) < 1e-12
    assert rel(fm.gradient_from_jac(out["jac"]), g["grad_f64"]) < 1e-12
    assert rel(fm.laplace_from_lap(lap), g["lap_f64"]) < 1e-11
    assert rel(out["hess"], g["hess_f64"]) < 1e-6          # reference hessian buffer is fp32
    if O == D:
        assert rel(fm.divergence_from_jac(out["jac"]), g["div_f64"]) < 1e-12
    gth, gx = fm.backward(th, x, D, O, H, L, fm.ORDER_LAP, gy=g["gy"], gjac=g["gjac"],
                          glap=np.broadcast_to(g["glap"], (N, O)))
    assert rel(gth, g["gtheta_f64"]) < 1e-11
    assert rel(gx, g["gx_f64"]) < 1e-11
    # the fp32 reference sits within its own rounding of the fp64 restatement
    assert rel(g["y_f32"], out["y"]) < 2e-5
    assert rel(g["lap_f32"], fm.laplace_from_lap(lap)) < 1e-4
    assert rel(g["gtheta_f32"], gth) < 1e-4


@pytest.mark.parametrize("name", OPS)
def test_torch_port_matches_reference_fp32(name):
    g = load_golden("op_" + name)
    D, O, H, L, N = (int(v) for v in g["shape"])
    torch.manual_seed(0)
    net = tp.RefMLP(D, O, L, H).load_flat_theta(g["theta"])
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = net(x)
    jac, status = tp.jacobian(y, x)
    lap = tp.laplace(y, x)
    assert status == int(g["status_f32"])
    assert rel(y.detach(), g["y_f32"]) < 1e-6
    assert rel(jac.detach(), g["jac_f32"]) < 1e-6
    assert rel(tp.gradient(y, x).detach(), g["grad_f32"]) < 1e-6
    assert rel(lap.detach(), g["lap_f32"]) < 1e-5
    if O == D:
        assert rel(tp.divergence(y, x).detach(), g["div_f32"]) < 1e-6
    x3 = torch.from_numpy(g["x"])[None].clone().requires_grad_(True)
    h, _ = tp.hessian(net(x3), x3)
    assert rel(h[0].detach(), g["hess_f32"]) < 1e-5
    loss = (torch.from_numpy(g["gy"]) * y).sum() + (torch.from_numpy(g["gjac"]) * jac).sum() \
        + (torch.from_numpy(g["glap"]) * lap).sum()
    loss.backward()
    gth = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])
    assert rel(gth, g["gtheta_f32"]) < 2e-5


def _net(theta, D, O, H, L):
    return tp.RefMLP(D, O, L, H).load_flat_theta(theta)


def _flat_grad(net):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])


def _run(loss_dict, nets):
    for n in nets:
        n.zero_grad()
    sum(loss_dict.values()).backward()
    return {k: float(v) for k, v in loss_dict.items()}


def _t(a, grad=True):
    t = torch.from_numpy(np.asarray(a))
    return t.requires_grad_(True) if grad else t


def test_closures_advection_match_reference():
    g = load_golden("closure_advection")
    dt, vel, length, sr = (float(v) for v in g["cfg"])
    field, prev = _net(g["theta.field"], 1, 1, 20, 2), _net(g["theta.field_prev"], 1, 1, 20, 2)
    for p in prev.parameters():
        p.requires_grad_(False)
    x = _t(g["initialize.samples0.sample_random"]) * length / 2
    vals = _run(closures.advect_initialize(field, x), [field])
    assert abs(vals["main"] - float(g["initialize.loss.main"])) < 1e-6 * max(1, abs(vals["main"]))
    assert rel(_flat_grad(field), g["initialize.grad.field"]) < 2e-5
    x = _t(g["advect.samples0.sample_random"]) * length / 2
    xb = _t(g["advect.samples1.sample_boundary"], grad=False) * length / 2
    vals = _run(closures.advect_step(field, prev, tp, x, xb, dt, vel), [field])
    assert abs(vals["main"] - float(g["advect.loss.main"])) < 1e-5 * abs(vals["main"])
    assert abs(vals["bc"] - float(g["advect.loss.bc"])) < 1e-5 * abs(vals["bc"])
    assert rel(_flat_grad(field), g["advect.grad.field"]) < 5e-5


def test_closures_fluid_match_reference():
    g = load_golden("closure_fluid")
    dt = float(g["cfg"][0])
    vel, prev = _net(g["theta.velocity"], 2, 2, 32, 3), _net(g["theta.velocity_prev"], 2, 2, 32, 3)
    pres = _net(g["theta.pressure"], 2, 1, 32, 3)
    for p in prev.parameters():
        p.requires_grad_(False)

    def s(key, i, name):
        return _t(g[f"{key}.samples{i}.{name}"])

    def check(key, vals):
        for k, v in vals.items():
            ref = float(g[f"{key}.loss.{k}"])
            assert abs(v - ref) < 2e-5 * max(abs(ref), 1e-6), (key, k, v, ref)
        assert rel(_flat_grad(vel), g[f"{key}.grad.velocity"]) < 1e-4 or np.abs(g[f"{key}.grad.velocity"]).max() == 0
        assert rel(_flat_grad(pres), g[f"{key}.grad.pressure"]) < 1e-4 or np.abs(g[f"{key}.grad.pressure"]).max() == 0

    check("initialize", _run(closures.fluid_initialize(vel, s("initialize", 0, "sample_random")), [vel, pres]))
    check("advect_velocity", _run(closures.fluid_advect_velocity(
        vel, prev, s("advect_velocity", 0, "sample_random"), s("advect_velocity", 1, "sample_boundary2D_separate"),
        s("advect_velocity", 2, "sample_boundary2D_separate"), dt), [vel, pres]))
    check("solve_pressure", _run(closures.fluid_solve_pressure(
        vel, pres, tp, s("solve_pressure", 0, "sample_random"), s("solve_pressure", 1, "sample_boundary2D_separate"),
        s("solve_pressure", 2, "sample_boundary2D_separate")), [vel, pres]))
    check("projection", _run(closures.fluid_projection(
        vel, prev, pres, tp, s("projection", 0, "sample_random"), s("projection", 1, "sample_boundary2D_separate"),
        s("projection", 2, "sample_boundary2D_separate")), [vel, pres]))


ELAS = {
    "stretch2d": dict(dim=2, H=68, dt=0.05, energy=["arap", "constraint", "constraint_right", "volume"],
                      ratio_volume=1e3, ratio_arap=1e0, ratio_constraint=1e4, ratio_kinematics=1e0, ratio_collide=1e0,
                      ext=[0., 0., 0.], ext_T=5, off=[2.0, 0., 0.], plane=-2.0, center=[0., -2., 0.], radius=1.0),
    "collide2d": dict(dim=2, H=68, dt=0.1, energy=["arap", "kinematics", "collision_sphere", "external", "volume"],
                      ratio_volume=1e3, ratio_arap=2e1, ratio_constraint=1e3, ratio_kinematics=1e1, ratio_collide=1e4,
                      ext=[0., -2e2, 0.], ext_T=2, off=[1.0, 0., 0.], plane=-2.0, center=[0., -0.5, 0.], radius=1.0),
    "plane3d": dict(dim=3, H=66, dt=0.1, energy=["arap", "kinematics", "collision", "external", "volume"],
                    ratio_volume=1e3, ratio_arap=1e2, ratio_constraint=1e3, ratio_kinematics=1e0, ratio_collide=1e6,
                    ext=[0., 0., -1e2], ext_T=5, off=[1.0, 0., 0.], plane=-0.9, center=[0., -2., 0.], radius=1.0),
}


def elasticity_case(tag, g, make_net, ops):
    """shared by the oracle pin (here) and the CUDA parity tests"""
    c = ELAS[tag]
    dim = c["dim"]
    defo = make_net(g["theta.deformation"], dim, dim, c["H"], 3)
    prev = make_net(g["theta.prev"], dim, dim, c["H"], 3)
    pp = make_net(g["theta.prev_prev"], dim, dim, c["H"], 3)
    for n in (prev, pp):
        for p in n.parameters():
            p.requires_grad_(False)
    dev = next(defo.parameters()).device
    keys = sorted(k for k in g if k.startswith("solve_deformation.samples"))
    arrs = [torch.from_numpy(g[k]).to(dev) for k in sorted(keys, key=lambda k: int(k.split(".")[1][7:]))]
    # recorded order (elasticity/model.py:198-241): random, uniform | left-random, right-random, left-uniform, right-uniform
    samples = torch.cat([arrs[0].requires_grad_(True), arrs[1].requires_grad_(True)], dim=0)
    one = torch.ones
    left = torch.cat([torch.cat((-one(arrs[2].shape[0], 1, device=dev), arrs[2]), 1),
                      torch.cat((-one(arrs[4].shape[0], 1, device=dev), arrs[4]), 1)], 0)
    right = torch.cat([torch.cat((one(arrs[3].shape[0], 1, device=dev), arrs[3]), 1),
                       torch.cat((one(arrs[5].shape[0], 1, device=dev), arrs[5]), 1)], 0)
    defo.zero_grad()
    loss = closures.elasticity_solve_deformation(
        defo, prev, pp, ops, samples, left, right, dt=c["dt"], timestep=1, energy=c["energy"],
        ratio_arap=c["ratio_arap"], ratio_volume=c["ratio_volume"], ratio_kinematics=c["ratio_kinematics"],
        ratio_constraint=c["ratio_constraint"], ratio_collide=c["ratio_collide"],
        external_force=torch.tensor(c["ext"][:dim], device=dev), external_force_timesteps=c["ext_T"],
        constraint_offset_right=torch.tensor(c["off"][:dim], device=dev), plane_height=c["plane"],
        circle_center=torch.tensor(c["center"][:dim], device=dev), circle_radius=c["radius"])
    if loss["main"].requires_grad:                  # the fused CUDA closure accumulates its gradient itself
        loss["main"].backward()
    return float(loss["main"]), _flat_grad(defo).detach().cpu().numpy()


@pytest.mark.parametrize("tag", list(ELAS))
def test_closures_elasticity_match_reference(tag):
    g = load_golden("closure_elasticity_" + tag)
    val, grad = elasticity_case(tag, g, _net, tp)
    ref = float(g["solve_deformation.loss.main"])
    assert abs(val - ref) < 1e-4 * abs(ref)
    assert rel(grad, g["solve_deformation.grad.deformation"]) < 1e-5
    # the fp64 companion (oracle/make_goldens_fp64.py: the reference closure in double precision on the same weights and
    # samples) arbitrates: the reference's fp32 gradient -- and the port's -- sit ~1e-6 from it
    g64 = load_golden("closure_elasticity_" + tag + "_fp64")
    assert rel(grad, g64["solve_deformation.grad64.deformation"]) < 1e-5
    assert abs(val - float(g64["solve_deformation.loss64.main"])) < 1e-6 * abs(val)


def test_sampling_port_shapes_and_ranges():
    torch.manual_seed(0)
    assert tp.sample_uniform(4, 2).shape == (16, 2)
    assert tp.sample_uniform(4, 2, flatten=False).shape == (4, 4, 2)
    u = tp.sample_uniform(5, 1)[:, 0]
    assert torch.allclose(u, torch.tensor([-0.8, -0.4, 0.0, 0.4, 0.8]), atol=1e-7)
    r = tp.sample_random(100, 3)
    assert r.shape == (100, 3) and float(r.abs().max()) <= 1
    b = tp.sample_boundary(20, 1)
    assert b.shape == (20, 1) and float((b.abs() - 1).abs().max()) <= 1.0001e-4
    b2 = tp.sample_boundary(40, 2)
    assert b2.shape == (40, 2)
    h = tp.sample_boundary2D_separate(16, "horizontal")
    assert h.shape == (16, 2) and float((h[:, 0].abs() - 1).abs().max()) <= 1.0001e-4
    v = tp.sample_boundary2D_separate(16, "vertical")
    assert float((v[:, 1].abs() - 1).abs().max()) <= 1.0001e-4
    with pytest.raises(RuntimeError):
        tp.sample_boundary2D_separate(16, "diagonal")
