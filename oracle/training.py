"""Restated optimisation runtime of the reference (the CALLER of the hot path).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows base/baseModel.py:55-62
(_reset_optimizer: Adam + ReduceLROnPlateau(factor .1, patience 500, min_lr 1e-8)),
:73-81 (_update_network: sum losses -> zero_grad -> backward -> Adam.step -> scheduler.step)
and :104-134 (_training_loop, without logging / visualisation / early stop).
Used to replay a recorded collocation-sample stream through either the oracle modules or
the fused modules, so that short trajectories can be compared with the reference's.
"""
from __future__ import annotations

import torch


def training_loop(closure, nets, n_iters, lr, on_step=None):
    """closure(i) -> loss_dict with key 'main'; nets: trainable modules.  Returns history."""
    opt = torch.optim.Adam([{"params": n.parameters(), "lr": lr} for n in nets])
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, min_lr=1e-8, patience=500)
    hist = []
    for i in range(n_iters):
        loss_dict = closure(i)
        loss = sum(loss_dict.values())
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step(loss_dict["main"])
        hist.append([float(v.detach()) for v in loss_dict.values()])
        if on_step is not None:
            on_step(i, loss_dict)
    return hist


def replay_advection(make_net, ops, golden, device="cpu"):
    """initialize() + one step() of Advection1DModel (advection/model.py:36-66) on the
    recorded sample stream of tests/golden/trajectory_advection.npz."""
    from . import closures
    dt, vel, length, sr, K, lr = (float(v) for v in golden["cfg"])
    K = int(K)
    field = make_net(golden["theta0"], 1, 1, 20, 2)
    prev = make_net(golden["theta0"], 1, 1, 20, 2)
    for p in prev.parameters():
        p.requires_grad_(False)
    keys = sorted(k for k in golden if k.startswith("samples"))
    stream = [torch.from_numpy(golden[k]).to(device) for k in keys]
    pos = [0]

    def nxt():
        t = stream[pos[0]]
        pos[0] += 1
        return t

    def c_init(i):
        x = nxt().requires_grad_(True) * length / 2
        return closures.advect_initialize(field, x)

    h_init = training_loop(c_init, [field], K, lr)
    theta_init = torch.cat([p.detach().reshape(-1) for p in field.parameters()]).cpu().numpy()
    prev.load_state_dict(field.state_dict())

    def c_step(i):
        x = nxt().requires_grad_(True) * length / 2
        xb = nxt() * length / 2
        return closures.advect_step(field, prev, ops, x, xb, dt, vel)

    h_step = training_loop(c_step, [field], K, lr)
    theta_step = torch.cat([p.detach().reshape(-1) for p in field.parameters()]).cpu().numpy()
    return dict(theta_after_init=theta_init, theta_after_step=theta_step, hist_initialize=h_init,
                hist_advect=h_step)


def replay_fluid(make_net, ops, golden, device="cpu"):
    """initialize() + one step() of Fluid2DModel (fluid/model.py:36-70: advect -> pressure -> projection) on the
    recorded sample stream of tests/golden/trajectory_fluid.npz; returns weights, loss histories and the per-frame
    velocity field on the golden's uniform grid (fluid/model.py:28-34 sample_field)."""
    from . import closures
    from .torch_port import sample_uniform
    dt, SR, K, lr, GRID, _ = (float(v) for v in golden["cfg"])
    K, GRID = int(K), int(GRID)
    vel = make_net(golden["theta0.velocity"], 2, 2, 32, 3)
    prev = make_net(golden["theta0.velocity"], 2, 2, 32, 3)
    pres = make_net(golden["theta0.pressure"], 2, 1, 32, 3)
    for p in prev.parameters():
        p.requires_grad_(False)
    keys = sorted(k for k in golden if k.startswith("samples"))
    stream = [torch.from_numpy(golden[k]).to(device) for k in keys]
    pos = [0]

    def nxt(grad=True):
        t = stream[pos[0]]
        pos[0] += 1
        return t.requires_grad_(True) if grad else t

    def flat(net):
        return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy()

    def frame():
        grid = sample_uniform(GRID, 2, device=device, flatten=False)
        with torch.no_grad():
            return vel(grid).cpu().numpy()

    out = {}
    out["hist_initialize"] = training_loop(lambda i: closures.fluid_initialize(vel, nxt()), [vel, pres], K, lr)
    out["theta_after_init.velocity"] = flat(vel)
    out["frame0"] = frame()
    prev.load_state_dict(vel.state_dict())
    out["hist_advect_velocity"] = training_loop(
        lambda i: closures.fluid_advect_velocity(vel, prev, nxt(), nxt(), nxt(), dt), [vel, pres], K, lr)
    out["hist_solve_pressure"] = training_loop(
        lambda i: closures.fluid_solve_pressure(vel, pres, ops, nxt(), nxt(), nxt()), [vel, pres], K, lr)
    prev.load_state_dict(vel.state_dict())
    out["hist_projection"] = training_loop(
        lambda i: closures.fluid_projection(vel, prev, pres, ops, nxt(), nxt(), nxt()), [vel, pres], K, lr)
    out["theta_after_step.velocity"] = flat(vel)
    out["theta_after_step.pressure"] = flat(pres)
    out["frame1"] = frame()
    return out
