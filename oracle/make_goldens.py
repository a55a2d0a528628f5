"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist
on the GPU box):   python -m oracle.make_goldens

The reference has no golden vectors of its own, so these files *are* the
pin: operator-level outputs of ``base.MLP`` + ``base.diff_ops`` (fp32 as the reference
runs, plus the same code in fp64 as the arbiter), closure-level ``loss_dict`` values and
parameter gradients of every ``@_training_loop`` closure, and a short Adam trajectory.
Every random draw is seeded; collocation points are captured by wrapping the sampling
functions the model modules imported by name.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

OPERATOR_CASES = [
    # name,        D, O, H,   L, N
    ("advect1d",   1, 1, 20,  2, 96),
    ("fluid_vel",  2, 2, 32,  3, 96),
    ("fluid_pres", 2, 1, 32,  3, 96),
    ("elas2d",     2, 2, 68,  3, 64),
    ("bunny3d",    3, 3, 66,  3, 64),
    ("depth0",     2, 1, 8,   0, 33),
    ("deep5",      3, 1, 24,  5, 40),
    ("wide128",    2, 1, 128, 1, 48),
]


def flat(params):
    return torch.cat([p.detach().reshape(-1) for p in params])


def flat_grad(net):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                      for p in net.parameters()])


def operator_case(base, name, D, O, H, L, N, seed):
    torch.manual_seed(seed)
    net = base.MLP(D, O, L, H, nonlinearity="sine")
    theta = flat(net.parameters()).numpy().copy()
    g = torch.Generator().manual_seed(seed + 1000)
    x0 = torch.rand(N, D, generator=g) * 2 - 1
    gy = torch.randn(N, O, generator=g)
    gj = torch.randn(N, O, D, generator=g)
    gl = torch.randn(N, 1, generator=g)
    rec = dict(theta=theta, x=x0.numpy(), gy=gy.numpy(), gjac=gj.numpy(), glap=gl.numpy(),
               shape=np.array([D, O, H, L, N]))
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = base.MLP(D, O, L, H, nonlinearity="sine").to(dt)
        with torch.no_grad():
            for p, q in zip(m.parameters(), net.parameters()):
                p.copy_(q.to(dt))
        x = x0.to(dt).clone().requires_grad_(True)
        y = m(x)
        jac, st = base.jacobian(y, x)           # NOTE: the reference buffer is fp32 (torch.zeros)
        grad_sum = base.gradient(y, x)
        lap = base.laplace(y, x)
        if O == D:
            rec[f"div_{tag}"] = base.divergence(y, x).detach().numpy()
        x3 = x0.to(dt)[None].clone().requires_grad_(True)
        hess, _ = base.hessian(m(x3), x3)
        # exact-dtype Jacobian (the reference's jac buffer truncates fp64 to fp32)
        jac_exact = torch.stack([torch.autograd.grad(y[:, o].sum(), x, create_graph=True)[0]
                                 for o in range(O)], 1)
        loss = (gy.to(dt) * y).sum() + (gj.to(dt) * jac_exact).sum() + (gl.to(dt) * lap).sum()
        m.zero_grad()
        gx, = torch.autograd.grad(loss, x, retain_graph=True)
        loss.backward()
        rec.update({
            f"y_{tag}": y.detach().numpy(), f"jac_{tag}": jac.detach().numpy(),
            f"jacx_{tag}": jac_exact.detach().numpy(),
            f"grad_{tag}": grad_sum.detach().numpy(), f"lap_{tag}": lap.detach().numpy(),
            f"hess_{tag}": hess[0].detach().numpy(), f"gtheta_{tag}": flat_grad(m).numpy(),
            f"gx_{tag}": gx.numpy(), f"status_{tag}": np.array(st),
        })
    np.savez_compressed(os.path.join(OUT, f"op_{name}.npz"), **rec)
    print("operator", name, "theta", theta.size)


# ------------------------------------------------------------------------------ closures
class Recorder:
    """wraps the sampling functions a model module imported by name and records outputs."""

    def __init__(self, module, names):
        self.module, self.names, self.log, self.orig = module, names, [], {}

    def __enter__(self):
        for n in self.names:
            if hasattr(self.module, n):
                fn = getattr(self.module, n)
                self.orig[n] = fn

                def wrap(*a, __fn=fn, __n=n, **k):
                    out = __fn(*a, **k)
                    self.log.append((__n, out.detach().clone().numpy()))
                    return out

                setattr(self.module, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, fn in self.orig.items():
            setattr(self.module, n, fn)


def raw_closure(method):
    """the undecorated loss closure inside BaseModel._training_loop's ``loop`` wrapper."""
    for cell in method.__closure__:
        if callable(cell.cell_contents):
            return cell.cell_contents
    raise RuntimeError("closure not found")


SAMPLERS = ["sample_random", "sample_uniform", "sample_boundary", "sample_boundary2D_separate"]


def run_closure(model, module, closure_name, nets, rec, key, pre=None):
    fn = raw_closure(getattr(type(model), closure_name))
    for n in nets.values():
        n.zero_grad()
    with Recorder(module, SAMPLERS) as r:
        loss_dict = fn(model)
    total = sum(loss_dict.values())
    total.backward()
    for i, (name, arr) in enumerate(r.log):
        rec[f"{key}.samples{i}.{name}"] = arr
    for k, v in loss_dict.items():
        rec[f"{key}.loss.{k}"] = np.array(float(v))
    for nn_, net in nets.items():
        rec[f"{key}.grad.{nn_}"] = flat_grad(net).numpy()
    print("closure", key, {k: float(v) for k, v in loss_dict.items()})


def closures_advection(ref):
    torch.manual_seed(11)
    cfg = ref_loader.make_cfg("advection", sample_resolution=500)
    import advection.model as mod
    m = ref.advection.Advection1DModel(cfg)
    m.init_cond_func = ref.advection.model.get_examples(cfg.init_cond) if hasattr(ref.advection, "model") else None
    from advection.examples import get_examples
    m.init_cond_func = get_examples(cfg.init_cond)
    rec = {"theta.field": flat(m.field.parameters()).numpy(),
           "theta.field_prev": flat(m.field_prev.parameters()).numpy(),
           "cfg": np.array([cfg.dt, cfg.vel, cfg.length, cfg.sample_resolution])}
    m.timestep = 0
    run_closure(m, mod, "_initialize", {"field": m.field}, rec, "initialize")
    m.timestep = 1
    run_closure(m, mod, "_advect", {"field": m.field}, rec, "advect")
    np.savez_compressed(os.path.join(OUT, "closure_advection.npz"), **rec)


def closures_fluid(ref):
    torch.manual_seed(12)
    cfg = ref_loader.make_cfg("fluid", sample_resolution=24)
    import fluid.model as mod
    from fluid.examples import get_examples
    m = ref.fluid.Fluid2DModel(cfg)
    m.init_cond_func = get_examples(cfg.init_cond)
    nets = {"velocity": m.velocity_field, "pressure": m.pressure_field}
    rec = {"theta.velocity": flat(m.velocity_field.parameters()).numpy(),
           "theta.velocity_prev": flat(m.velocity_field_prev.parameters()).numpy(),
           "theta.pressure": flat(m.pressure_field.parameters()).numpy(),
           "cfg": np.array([cfg.dt, cfg.sample_resolution])}
    for name in ("_initialize", "_advect_velocity", "_solve_pressure", "_projection"):
        run_closure(m, mod, name, nets, rec, name.lstrip("_"))
    # write_output's curl on the visualisation grid (fluid/model.py:207-213)
    grid_u, grid_x = m.sample_field(8, return_samples=True)
    jaco, _ = ref.base.jacobian(grid_u, grid_x)
    rec["vis.grid_u"] = grid_u.detach().numpy()
    rec["vis.curl"] = (jaco[..., 1, 0] - jaco[..., 0, 1]).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "closure_fluid.npz"), **rec)


ELASTICITY_CASES = (
    ("stretch2d", dict(dim=2, energy=["arap", "constraint", "constraint_right", "volume"],
                       ratio_volume=1e3, ratio_arap=1e0, ratio_constraint=1e4,
                       constraint_right_offset_x=2.0, sample_resolution=10, hidden_features=68)),
    ("collide2d", dict(dim=2, energy=["arap", "kinematics", "collision_sphere", "external", "volume"],
                       ratio_volume=1e3, ratio_arap=2e1, ratio_collide=1e4, ratio_kinematics=1e1,
                       external_force_y=-2e2, external_force_timesteps=2, dt=0.1,
                       collide_circle_y=-0.5, sample_resolution=10, hidden_features=68)),
    ("plane3d", dict(dim=3, energy=["arap", "kinematics", "collision", "external", "volume"],
                     ratio_volume=1e3, ratio_arap=1e2, ratio_collide=1e6, ratio_kinematics=1e0,
                     external_force_z=-1e2, external_force_timesteps=5, dt=0.1,
                     plane_height=-0.9, sample_resolution=5, hidden_features=66)),
)


def closures_elasticity(ref):
    if ref.elasticity is None:
        print("elasticity import failed:", ref.elasticity_error)
        return
    import elasticity.model as mod
    for tag, over in ELASTICITY_CASES:
        torch.manual_seed(13)
        cfg = ref_loader.make_cfg("elasticity", **over)
        m = ref.elasticity.ElasticityModel(cfg)
        # make the three fields distinct so kinematics terms are non-trivial
        torch.manual_seed(14)
        for net in (m.deformation_field_prev, m.deformation_field_prev_prev):
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(0.02 * torch.randn_like(p) * p.abs().mean())
        nets = {"deformation": m.deformation_field}
        rec = {"theta.deformation": flat(m.deformation_field.parameters()).numpy(),
               "theta.prev": flat(m.deformation_field_prev.parameters()).numpy(),
               "theta.prev_prev": flat(m.deformation_field_prev_prev.parameters()).numpy(),
               "cfg": np.array([cfg.dim, cfg.dt, cfg.sample_resolution, cfg.hidden_features])}
        m.timestep = 0
        m.sample_resolution_init = 12 if cfg.dim == 2 else 5
        run_closure(m, mod, "_initialize", nets, rec, "initialize")
        m.timestep = 1
        run_closure(m, mod, "_solve_deformation", nets, rec, "solve_deformation")
        np.savez_compressed(os.path.join(OUT, f"closure_elasticity_{tag}.npz"), **rec)


def trajectory_advection(ref):
    """K Adam iterations of initialize() + one step(): final weights + loss history."""
    import advection.model as mod
    from advection.examples import get_examples
    K = 25
    torch.manual_seed(21)
    cfg = ref_loader.make_cfg("advection", sample_resolution=400, max_n_iters=K, lr=1e-4)
    m = ref.advection.Advection1DModel(cfg)
    theta0 = flat(m.field.parameters()).numpy().copy()
    hist = {"_initialize": [], "_advect": []}

    def spy(name):
        fn = raw_closure(getattr(type(m), name))

        def wrapped(self):
            d = fn(self)
            hist[name].append([float(v) for v in d.values()])
            return d
        return wrapped

    # re-decorate spying closures so that the reference's own _training_loop drives them
    BaseModel = ref.base.BaseModel
    type(m)._initialize = BaseModel._training_loop(spy("_initialize"))
    type(m)._advect = BaseModel._training_loop(spy("_advect"))
    torch.manual_seed(22)
    with Recorder(mod, SAMPLERS) as r:
        m.initialize()
        theta_init = flat(m.field.parameters()).numpy().copy()
        m.step()
    rec = {"theta0": theta0, "theta_after_init": theta_init,
           "theta_after_step": flat(m.field.parameters()).numpy().copy(),
           "hist_initialize": np.array(hist["_initialize"]), "hist_advect": np.array(hist["_advect"]),
           "cfg": np.array([cfg.dt, cfg.vel, cfg.length, cfg.sample_resolution, K, cfg.lr])}
    for i, (name, arr) in enumerate(r.log):
        rec[f"samples{i:03d}.{name}"] = arr.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "trajectory_advection.npz"), **rec)
    print("trajectory advection: last losses", hist["_initialize"][-1], hist["_advect"][-1])


def trajectory_fluid(ref):
    """fluid2Dtlgn (Taylor-Green): K Adam iterations of initialize() + one step() (advect -> pressure -> projection):
    weights, loss histories and the PER-FRAME velocity field on a uniform grid (the north star's parity clause)."""
    import fluid.model as mod
    K, SR, GRID = 12, 20, 12
    torch.manual_seed(31)
    cfg = ref_loader.make_cfg("fluid", sample_resolution=SR, max_n_iters=K, lr=1e-4)
    m = ref.fluid.Fluid2DModel(cfg)
    theta0 = {"velocity": flat(m.velocity_field.parameters()).numpy().copy(),
              "pressure": flat(m.pressure_field.parameters()).numpy().copy()}
    names = ["_initialize", "_advect_velocity", "_solve_pressure", "_projection"]
    hist = {n: [] for n in names}

    def spy(name):
        fn = raw_closure(getattr(type(m), name))

        def wrapped(self):
            d = fn(self)
            hist[name].append([float(v) for v in d.values()])
            return d
        return wrapped

    BaseModel = ref.base.BaseModel
    for n in names:
        setattr(type(m), n, BaseModel._training_loop(spy(n)))
    torch.manual_seed(32)
    with Recorder(mod, ["sample_random", "sample_boundary2D_separate"]) as r:
        m.initialize()
        frame0 = m.sample_field(GRID).detach().numpy().copy()
        th_init = flat(m.velocity_field.parameters()).numpy().copy()
        n_init = len(r.log)
        m.step()
        frame1 = m.sample_field(GRID).detach().numpy().copy()
    rec = {"theta0.velocity": theta0["velocity"], "theta0.pressure": theta0["pressure"],
           "theta_after_init.velocity": th_init,
           "theta_after_step.velocity": flat(m.velocity_field.parameters()).numpy().copy(),
           "theta_after_step.pressure": flat(m.pressure_field.parameters()).numpy().copy(),
           "frame0": frame0, "frame1": frame1,
           "cfg": np.array([cfg.dt, SR, K, cfg.lr, GRID, n_init])}
    for n in names:
        rec["hist" + n] = np.array(hist[n])
    for i, (name, arr) in enumerate(r.log):
        rec[f"samples{i:04d}.{name}"] = arr.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "trajectory_fluid.npz"), **rec)
    print("trajectory fluid: last losses", {n: hist[n][-1] for n in names}, "samples", len(r.log))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load(cpu=True)
    torch.set_num_threads(1)        # bit-reproducible reductions
    for i, case in enumerate(OPERATOR_CASES):
        operator_case(ref.base, *case, seed=100 + i)
    closures_advection(ref)
    closures_fluid(ref)
    closures_elasticity(ref)
    trajectory_advection(ref)
    trajectory_fluid(ref)


if __name__ == "__main__":
    main()
