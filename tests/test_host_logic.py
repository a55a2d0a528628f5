"""Host-side logic (drop-in MLP, autograd boundary, diff_ops fast/generic paths, parameter
aliasing) exercised on CPU tensors with the C-ABI calls routed to the emulation build.
The arithmetic checked here is the kernel sources' (under emulation) -- the product parity
tests against the oracle on the real device are in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import insr_pde_b200 as ib
from conftest import load_golden
from oracle import closures, torch_port as tp


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def pair(D, O, H, L, seed=0):
    torch.manual_seed(seed)
    ours = ib.MLP(D, O, L, H, nonlinearity="sine")
    ref = tp.RefMLP(D, O, L, H)
    ref.load_state_dict(ours.state_dict())
    return ours, ref


def flat_grad(net):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])


def test_state_dict_keys_and_init_stream_match_reference_layout():
    torch.manual_seed(5)
    ours = ib.MLP(2, 1, 3, 32, nonlinearity="sine")
    assert list(ours.state_dict().keys()) == [f"net.{i}.{k}" for i in (0, 2, 4, 6, 8) for k in ("weight", "bias")]
    g = load_golden("op_fluid_pres")     # reference MLP built under torch.manual_seed(102)
    torch.manual_seed(102)
    again = ib.MLP(2, 1, 3, 32, nonlinearity="sine")
    assert np.array_equal(again.flat_theta().numpy(), g["theta"])
    with pytest.raises(NotImplementedError):
        ib.MLP(2, 1, 3, 32)              # reference default nonlinearity='relu' is not on this path


def test_flat_theta_realiases_after_rehoming(emu_backend):
    net, ref = pair(2, 1, 8, 1)
    th0 = net.flat_theta()
    assert all(p.data_ptr() == th0.data_ptr() + 4 * off for p, (off, _, _) in zip(net.parameters(), net.param_slices()))
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    x = torch.rand(10, 2).requires_grad_(True)
    net(x).sum().backward()
    opt.step()
    assert net.flat_theta() is th0 and torch.equal(th0, torch.cat([p.detach().reshape(-1) for p in net.parameters()]))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.double().float()                 # re-homes every parameter like net.cpu()/net.cuda() does
    th1 = net.flat_theta()
    assert th1 is not th0
    net.load_state_dict(sd)
    assert torch.equal(net.flat_theta(), th1)
    ref.load_state_dict(net.state_dict())
    assert rel(net(x).detach(), ref(x).detach()) < 1e-5


@pytest.mark.parametrize("shape", [(1, 1, 20, 2), (2, 2, 32, 3), (2, 1, 32, 3), (3, 3, 12, 2)])
def test_dropin_diff_ops_fast_and_generic_paths(emu_backend, shape):
    D, O, H, L = shape
    net, ref = pair(D, O, H, L, seed=3)
    torch.manual_seed(1)
    x = (torch.rand(48, D) * 2 - 1).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = net(x), ref(xr)
    assert rel(y.detach(), yr.detach()) < 1e-5
    # fast paths (y is tagged with its source)
    assert rel(ib.gradient(y, x).detach(), tp.gradient(yr, xr).detach()) < 1e-5
    assert rel(ib.jacobian(y, x)[0].detach(), tp.jacobian(yr, xr)[0].detach()) < 1e-5
    assert rel(ib.laplace(y, x).detach(), tp.laplace(yr, xr).detach()) < 2e-5
    if O == D:
        assert rel(ib.divergence(y, x).detach(), tp.divergence(yr, xr).detach()) < 1e-5
    # generic paths: the reference's own autograd.grad calls, through SirenFn's differentiable backward
    y2 = net(x) * 1.0          # loses the tag
    assert rel(tp.gradient(y2, x).detach(), tp.gradient(yr, xr).detach()) < 1e-5
    assert rel(tp.jacobian(y2, x)[0].detach(), tp.jacobian(yr, xr)[0].detach()) < 1e-5
    assert rel(tp.laplace(y2, x).detach(), tp.laplace(yr, xr).detach()) < 2e-5
    # parameter gradients of a loss mixing value, first and second derivatives
    for fast in (True, False):
        net.zero_grad(); ref.zero_grad()
        yy = net(x) if fast else net(x) * 1.0
        ops = ib if fast else tp
        loss = (yy ** 2).mean() + (ops.gradient(yy, x) ** 2).mean() + 0.1 * (ops.laplace(yy, x) ** 2).mean()
        loss.backward()
        yr = ref(xr)
        lr = (yr ** 2).mean() + (tp.gradient(yr, xr) ** 2).mean() + 0.1 * (tp.laplace(yr, xr) ** 2).mean()
        lr.backward()
        assert abs(float(loss) - float(lr)) < 1e-5 * abs(float(lr))
        assert rel(flat_grad(net), flat_grad(ref)) < 5e-5, fast


def test_grid_shaped_coords_and_hessian(emu_backend):
    net, ref = pair(2, 2, 16, 1, seed=4)
    grid = ib.sample_uniform(6, 2, flatten=False).requires_grad_(True)      # (6, 6, 2) like fluid/model.py:30-31
    gr = grid.detach().clone().requires_grad_(True)
    u, ur = net(grid), ref(gr)
    assert u.shape == (6, 6, 2)
    j, st = ib.jacobian(u, grid)
    jr, _ = tp.jacobian(ur, gr)
    assert st == 0 and j.shape == (6, 6, 2, 2) and rel(j.detach(), jr.detach()) < 1e-5
    x3 = torch.rand(1, 20, 2).requires_grad_(True)
    h, _ = ib.hessian(net(x3), x3)
    hr, _ = tp.hessian(ref(x3.detach().clone().requires_grad_(True) if False else x3), x3)
    assert h.shape == (1, 20, 2, 2, 2) and rel(h.detach(), hr.detach()) < 2e-5


def test_elasticity_pattern_jacobian_of_net_plus_x(emu_backend):
    net, ref = pair(2, 2, 16, 2, seed=6)
    x = (torch.rand(40, 2) * 2 - 1).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)

    def energy(n, xx, ops):
        q = n(xx) + xx
        F, _ = ops.jacobian(q, xx)
        s = torch.linalg.svdvals(F)
        return ((s - 1) ** 2).sum() + 1e1 * ((s.prod(dim=1) - 1) ** 2).sum() + (q ** 2).sum()

    e, er = energy(net, x, ib), energy(ref, xr, tp)
    e.backward(); er.backward()
    assert abs(float(e) - float(er)) < 1e-5 * abs(float(er))
    assert rel(flat_grad(net), flat_grad(ref)) < 5e-5


def test_frozen_net_no_grad_and_weights_argument(emu_backend):
    net, ref = pair(2, 2, 16, 1, seed=7)
    for p in net.parameters():
        p.requires_grad_(False)
    x = torch.rand(9, 2)
    with torch.no_grad():
        out = net(x)
    assert not out.requires_grad and rel(out, ref(x).detach()) < 1e-5
    w = torch.rand(9, 2)
    assert rel(net(x, weights=w), (ref(x) * w).detach()) < 1e-5
    assert net(torch.zeros(0, 2)).shape == (0, 2)


def test_fluid_closures_with_dropin_modules(emu_backend):
    g = load_golden("closure_fluid")
    dt = float(g["cfg"][0])

    def mk(theta, D, O):
        n = ib.MLP(D, O, 3, 32, nonlinearity="sine")
        with torch.no_grad():
            n.flat_theta().copy_(torch.from_numpy(theta))
        return n

    vel, prev, pres = mk(g["theta.velocity"], 2, 2), mk(g["theta.velocity_prev"], 2, 2), mk(g["theta.pressure"], 2, 1)
    for p in prev.parameters():
        p.requires_grad_(False)

    def s(key, i, name):
        return torch.from_numpy(g[f"{key}.samples{i}.{name}"]).requires_grad_(True)

    def check(key, loss_dict):
        vel.zero_grad(); pres.zero_grad()
        sum(loss_dict.values()).backward()
        for k, v in loss_dict.items():
            ref = float(g[f"{key}.loss.{k}"])
            assert abs(float(v) - ref) < 5e-5 * max(abs(ref), 1e-6), (key, k)
        for name, net in (("velocity", vel), ("pressure", pres)):
            gr = g[f"{key}.grad.{name}"]
            if np.abs(gr).max() > 0:
                assert rel(flat_grad(net), gr) < 2e-4, (key, name)

    bn = "sample_boundary2D_separate"
    check("advect_velocity", closures.fluid_advect_velocity(
        vel, prev, s("advect_velocity", 0, "sample_random"), s("advect_velocity", 1, bn), s("advect_velocity", 2, bn), dt))
    check("solve_pressure", closures.fluid_solve_pressure(
        vel, pres, ib, s("solve_pressure", 0, "sample_random"), s("solve_pressure", 1, bn), s("solve_pressure", 2, bn)))
    check("projection", closures.fluid_projection(
        vel, prev, pres, ib, s("projection", 0, "sample_random"), s("projection", 1, bn), s("projection", 2, bn)))


def test_cpu_tensor_without_backend_patch_is_refused():
    net = ib.MLP(2, 1, 1, 8, nonlinearity="sine")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(4, 2))


def _mk_golden_net(theta, D, O, H, L, device="cpu"):
    n = ib.MLP(D, O, L, H, nonlinearity="sine").to(device)
    with torch.no_grad():
        n.flat_theta().copy_(torch.from_numpy(theta).to(device))
    return n


def check_fused_closures_against_goldens(device):
    """shared with the GPU test: fused lsq closures reproduce the reference closures' loss_dict and gradients"""
    from insr_pde_b200 import fused
    g = load_golden("closure_fluid")
    dt = float(g["cfg"][0])
    vel = _mk_golden_net(g["theta.velocity"], 2, 2, 32, 3, device)
    prev = _mk_golden_net(g["theta.velocity_prev"], 2, 2, 32, 3, device)
    pres = _mk_golden_net(g["theta.pressure"], 2, 1, 32, 3, device)

    def s(key, i, name):
        return torch.from_numpy(g[f"{key}.samples{i}.{name}"]).to(device)

    def check(key, fn, tol=2e-4):
        fused.zero_grads(vel, pres)
        loss_dict = fn()
        for k, v in loss_dict.items():
            ref = float(g[f"{key}.loss.{k}"])
            assert abs(float(v) - ref) < 1e-4 * max(abs(ref), 1e-6), (key, k, float(v), ref)
        for name, net in (("velocity", vel), ("pressure", pres)):
            gr = g[f"{key}.grad.{name}"]
            got = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu().numpy()
            if np.abs(gr).max() > 0:
                assert rel(got, gr) < tol, (key, name, rel(got, gr))
            else:
                assert np.abs(got).max() == 0

    bn = "sample_boundary2D_separate"
    x0 = s("initialize", 0, "sample_random")
    check("initialize", lambda: fused.fluid_initialize(vel, x0, fused.taylorgreen_velocity(x0)))
    check("advect_velocity", lambda: fused.fluid_advect_velocity(
        vel, prev, s("advect_velocity", 0, "sample_random"), s("advect_velocity", 1, bn), s("advect_velocity", 2, bn), dt))
    check("solve_pressure", lambda: fused.fluid_solve_pressure(
        vel, pres, s("solve_pressure", 0, "sample_random"), s("solve_pressure", 1, bn), s("solve_pressure", 2, bn)))
    check("projection", lambda: fused.fluid_projection(
        vel, prev, pres, s("projection", 0, "sample_random"), s("projection", 1, bn), s("projection", 2, bn)))

    g = load_golden("closure_advection")
    dt, velc, length, sr = (float(v) for v in g["cfg"])
    field = _mk_golden_net(g["theta.field"], 1, 1, 20, 2, device)
    fprev = _mk_golden_net(g["theta.field_prev"], 1, 1, 20, 2, device)
    x = torch.from_numpy(g["initialize.samples0.sample_random"]).to(device) * length / 2
    fused.zero_grads(field)
    ld = fused.advect_initialize(field, x, torch.exp(-0.5 * (x + 1.5) ** 2 / 0.01))
    assert abs(float(ld["main"]) - float(g["initialize.loss.main"])) < 1e-4 * float(g["initialize.loss.main"])
    assert rel(torch.cat([p.grad.reshape(-1) for p in field.parameters()]).cpu(), g["initialize.grad.field"]) < 2e-4
    x = torch.from_numpy(g["advect.samples0.sample_random"]).to(device) * length / 2
    xb = torch.from_numpy(g["advect.samples1.sample_boundary"]).to(device) * length / 2
    fused.zero_grads(field)
    ld = fused.advect_step(field, fprev, x, xb, dt, velc)
    assert abs(float(ld["main"]) - float(g["advect.loss.main"])) < 1e-4 * float(g["advect.loss.main"])
    assert abs(float(ld["bc"]) - float(g["advect.loss.bc"])) < 1e-4 * float(g["advect.loss.bc"])
    assert rel(torch.cat([p.grad.reshape(-1) for p in field.parameters()]).cpu(), g["advect.grad.field"]) < 2e-4


def test_fused_closures_match_reference_goldens_emulated(emu_backend):
    check_fused_closures_against_goldens("cpu")


def test_split_closures_and_loss_slots_emulated(emu_backend):
    """the closures split into a frozen side (``*_target``, what a graphed loop prepares one iteration ahead into a persistent
    buffer) and a trainable side (``target=`` override), and the loss slots handed out by a provider (``fused._acc``, what
    GraphedLoop does): same losses and gradients as the one-call closure; slots are accumulated into, not overwritten"""
    from insr_pde_b200 import fused
    torch.manual_seed(3)
    vel, prev = ib.MLP(2, 2, 3, 32, nonlinearity="sine"), ib.MLP(2, 2, 3, 32, nonlinearity="sine")
    pres = ib.MLP(2, 1, 3, 32, nonlinearity="sine")
    x = torch.rand(192, 2) * 2 - 1
    bx, by = ib.sample_boundary2D_separate(20, "horizontal"), ib.sample_boundary2D_separate(20, "vertical")

    def grads():
        return torch.cat([fused.flat_grad(n).clone() for n in (vel, pres)])

    cases = [
        (lambda out: fused.fluid_advect_target(prev, x, 0.05, out=out), 2,
         lambda t: fused.fluid_advect_velocity(vel, prev, x, bx, by, 0.05, target=t)),
        (lambda out: fused.fluid_pressure_target(vel, x, out=out), 1,
         lambda t: fused.fluid_solve_pressure(vel, pres, x, bx, by, target=t)),
        (lambda out: fused.fluid_projection_target(prev, pres, x, out=out), 2,
         lambda t: fused.fluid_projection(vel, prev, pres, x, bx, by, target=t)),
    ]
    for target, n_res, closure in cases:
        fused.zero_grads(vel, pres)
        ref = closure(None)
        g_ref, l_ref = grads(), {k: float(v) for k, v in ref.items()}
        buf = torch.full((x.shape[0], n_res), 7.0)                   # persistent buffer, stale content
        got = target(buf)
        assert got.data_ptr() == buf.data_ptr() and tuple(target(None).shape) == (x.shape[0], n_res)
        slots = torch.zeros(4)
        fused._ACC_PROVIDER.append(lambda n, device: slots[:n])
        try:
            fused.zero_grads(vel, pres)
            ld = closure(buf)
        finally:
            fused._ACC_PROVIDER.pop()
        assert isinstance(ld, fused.Losses) and ld.vector.data_ptr() == slots.data_ptr()
        assert torch.equal(grads(), g_ref) and {k: float(v) for k, v in ld.items()} == l_ref
        assert float(slots[0]) == l_ref["main"] and float(slots[1]) == l_ref["bc"] and float(slots[2:].abs().max()) == 0.0
        fused._ACC_PROVIDER.append(lambda n, device: slots[:n])      # a second call ACCUMULATES (the update kernel zeroes)
        try:
            closure(buf)
        finally:
            fused._ACC_PROVIDER.pop()
        assert abs(float(slots[0]) - 2 * l_ref["main"]) <= 1e-6 * abs(l_ref["main"])
    # advection (1-D): the same split
    field, fprev = ib.MLP(1, 1, 2, 20, nonlinearity="sine"), ib.MLP(1, 1, 2, 20, nonlinearity="sine")
    xa, xb = torch.rand(150, 1) * 4 - 2, torch.tensor([[-2.0], [2.0]])
    fused.zero_grads(field)
    ref = fused.advect_step(field, fprev, xa, xb, 0.05, 0.25)
    g_ref = fused.flat_grad(field).clone()
    fused.zero_grads(field)
    ld = fused.advect_step(field, fprev, xa, xb, 0.05, 0.25, target=fused.advect_target(fprev, xa, 0.05, 0.25, out=torch.empty(150, 1)))
    assert torch.equal(fused.flat_grad(field), g_ref) and float(ld["main"]) == float(ref["main"])


def test_graphed_loop_replay_schedule(monkeypatch):
    """GraphedLoop.run's bookkeeping without a GPU: CUDA graphs replaced by recorders that re-run the captured iteration
    indices.  With ``prepare`` the buffer sets must alternate 0, 1, 0, 1, ... across the eager first iteration, the unrolled
    graph (UNROLL iterations per replay) and the single-iteration remainders; every run() primes set 0 and starts there;
    exactly n_iters iterations happen, and iteration i consumes what prepare wrote i calls ago."""
    from insr_pde_b200 import fused

    state = {"capturing": None}

    class FakeGraph:
        def __init__(self):
            self.iters = []

        def replay(self):
            for k in self.iters:
                probe._execute(k)

    class FakeCapture:
        def __init__(self, g):
            self.g = g

        def __enter__(self):
            state["capturing"] = self.g

        def __exit__(self, *a):
            state["capturing"] = None

    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", FakeCapture)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    class Probe(fused.GraphedLoop):
        def __init__(self, with_prepare):
            self.draws = 0                      # counter of the "sampler"
            self.buf = [None, None]
            self.log = []                       # (buffer set, draw consumed) per executed iteration
            net = ib.MLP(1, 1, 1, 4, nonlinearity="sine")
            super().__init__([net], 1e-3, closure=None, prepare=self._prep if with_prepare else None)
            self.keys = ["main"]
            self.hist = torch.zeros(self.capacity, 1)

        def _prep(self, k):
            if state["capturing"] is None:
                self.buf[k] = self.draws
                self.draws += 1

        def _execute(self, k):
            if self.prepare is not None:
                self.log.append((k, self.buf[k]))
                self.buf[1 - k] = self.draws    # prepare(1 - k) of this iteration
                self.draws += 1
            else:
                self.log.append((0, None))
            self.hist[int(self.idx)] = float(len(self.log))
            self.idx += 1

        def _iteration(self, k=0):
            if state["capturing"] is not None:
                state["capturing"].iters.append(k)
            else:
                self._execute(k)
            return self.keys

    for with_prepare in (True, False):
        for n in (1, 2, 3, 8, 9, 10, 101):
            probe = Probe(with_prepare)
            h = probe.run(n, check_every=7)
            assert len(h) == n and len(probe.log) == n, (with_prepare, n, len(probe.log))
            if with_prepare:
                assert [k for k, _ in probe.log] == [i % 2 for i in range(n)]
                assert [d for _, d in probe.log] == list(range(n))             # iteration i consumes draw i
            # a second training loop on the same captured graphs: fresh prime, starts at set 0 again
            probe.reset(1e-3)
            probe.log.clear()
            first = probe.draws
            h = probe.run(n + 1, check_every=100)
            assert len(h) == n + 1 and len(probe.log) == n + 1
            if with_prepare:
                assert [k for k, _ in probe.log] == [i % 2 for i in range(n + 1)]
                assert [d for _, d in probe.log] == list(range(first, first + n + 1))
            if n >= 2 * fused.GraphedLoop.UNROLL + 1:
                assert probe.graph_u is not None and len(probe.graph_u.iters) == fused.GraphedLoop.UNROLL


def test_fused_training_loop_and_stepper_emulated(emu_backend):
    from insr_pde_b200 import fused
    torch.manual_seed(0)
    vel, prev, pres = (ib.MLP(2, o, 3, 32, nonlinearity="sine") for o in (2, 2, 1))
    stepper = fused.FluidStepper(vel, prev, pres, dt=0.05, sample_resolution=16, lr=1e-3)
    h0 = stepper.initialize(fused.taylorgreen_velocity, 6)
    assert h0[-1]["main"] < h0[0]["main"]                      # Adam on the flat gradient actually descends
    before = pres.flat_theta().clone()
    h1, h2, h3 = stepper.step(3)
    assert len(h1) == len(h2) == len(h3) == 3 and all(np.isfinite(list(d.values())).all() for d in h1 + h2 + h3)
    assert not torch.equal(before, pres.flat_theta())
    assert torch.equal(prev.flat_theta(), prev.flat_theta()) and not any(p.requires_grad for p in prev.parameters())


def test_advection_stepper_emulated(emu_backend):
    """fused.AdvectionStepper (advection/model.py:37-91): Gaussian fit descends, the step hands the frame over and trains"""
    from insr_pde_b200 import fused
    torch.manual_seed(0)
    field, prev = ib.MLP(1, 1, 2, 20, nonlinearity="sine"), ib.MLP(1, 1, 2, 20, nonlinearity="sine")
    st = fused.AdvectionStepper(field, prev, dt=0.05, vel=0.25, length=4.0, sample_resolution=300, lr=1e-3)
    x, xb = st._samples()
    assert x.shape == (300, 1) and float(x.abs().max()) <= 2.0 and xb.shape == (10, 1)
    assert float((xb.abs() - 2.0).abs().max()) <= 2.0001e-4            # the epsilon bands around -L/2 and +L/2
    h0 = st.initialize(fused.gaussian_like, 8)
    assert h0[-1]["main"] < h0[0]["main"]
    before = field.flat_theta().clone()
    h1 = st.step(3)
    assert torch.equal(prev.flat_theta(), before) and not torch.equal(field.flat_theta(), before)
    assert len(h1) == 3 and set(h1[0]) == {"main", "bc"} and all(np.isfinite(list(d.values())).all() for d in h1)


def test_elasticity_stepper_follows_reference_algorithm_emulated(emu_backend):
    """fused.ElasticityStepper (eager loop) under emulation: the reference's sample pattern (elasticity/model.py:198-253), the
    one-kernel closure with its kept tape on the tiled family, torch Adam -- against the reference algorithm (autograd
    jacobian + torch.svd + Adam, oracle port) from the same weights on the same 'uniform' samples"""
    from insr_pde_b200 import fused, sampling
    dim, H, sr, K, lr = 2, 40, 6, 3, 1e-3
    kw = dict(energy=["arap", "volume", "kinematics", "external", "constraint", "constraint_right", "collision_sphere"],
              ratio_arap=1.0, ratio_volume=20.0, ratio_kinematics=0.5, ratio_constraint=50.0, ratio_collide=4.0,
              external_force=torch.tensor([0.0, -1.0]), external_force_timesteps=5,
              constraint_offset_right=torch.tensor([0.3, 0.0]), plane_height=-0.7,
              circle_center=torch.tensor([0.1, -0.8]), circle_radius=0.6)
    torch.manual_seed(3)
    nets = [ib.MLP(dim, dim, 3, H, nonlinearity="sine") for _ in range(3)]
    ref = [tp.RefMLP(dim, dim, 3, H).load_flat_theta(n.flat_theta().detach().clone()) for n in nets]
    st = fused.ElasticityStepper(*nets, dim, dt=0.05, sample_resolution=sr, lr=lr, sample_pattern=("random", "uniform"), **kw)
    x = st._interior(sr)
    left, right = st._fixed(sr)
    assert x.shape == (2 * sr ** dim, dim) and x.requires_grad                  # random block, then the cell-centred grid
    assert torch.equal(x[sr ** dim:].detach(), sampling.sample_uniform(sr, dim))
    assert left.shape == (2 * sr, dim) and bool((left[:, 0] == -1).all()) and bool((right[:, 0] == 1).all())
    st = fused.ElasticityStepper(*nets, dim, dt=0.05, sample_resolution=sr, lr=lr, sample_pattern=("uniform",), **kw)
    ours = [h["main"] for h in st.step(K)]
    ref[2].load_state_dict(ref[1].state_dict()); ref[1].load_state_dict(ref[0].state_dict())
    for n in ref[1:]:
        for p_ in n.parameters():
            p_.requires_grad_(False)
    opt = torch.optim.Adam(ref[0].parameters(), lr=lr)
    xs = sampling.sample_uniform(sr, dim)
    face = sampling.sample_uniform(sr, dim - 1)
    one = torch.ones(face.shape[0], 1)
    theirs = []
    for _ in range(K):
        opt.zero_grad()
        loss = closures.elasticity_solve_deformation(ref[0], ref[1], ref[2], tp, xs.clone().requires_grad_(True),
                                                     torch.cat((-one, face), 1), torch.cat((one, face), 1), dt=0.05, timestep=1, **kw)
        loss["main"].backward()
        opt.step()
        theirs.append(float(loss["main"].detach()))
    assert rel(ours, theirs) < 1e-3
    assert rel(nets[0].flat_theta().detach().numpy(), torch.cat([p_.detach().reshape(-1) for p_ in ref[0].parameters()]).numpy()) < 2e-3


def test_device_optimizer_matches_torch_adam_and_plateau(emu_backend):
    """insr_adam_step / insr_plateau_step reproduce torch.optim.Adam + ReduceLROnPlateau (base/baseModel.py:55-81)"""
    from insr_pde_b200 import _ops
    torch.manual_seed(0)
    n = 257
    theta = torch.randn(n)
    ref_p = torch.nn.Parameter(theta.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-2)
    sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, min_lr=1e-8, patience=3)
    m, v = torch.zeros(n), torch.zeros(n)
    sched = torch.tensor([1e-2, float("inf"), 0.0, 0.0])
    losses = [1.0, 0.9, 0.95, 0.95, 0.96, 0.97, 0.98, 0.5, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6]
    for it, lv in enumerate(losses):
        g = torch.randn(n) * (1 + it)
        ref_p.grad = g.clone()
        opt.step()
        sch.step(lv)
        _ops.adam_step(theta, g, m, v, sched)
        _ops.plateau_step(torch.tensor([lv]), sched, factor=0.1, patience=3, min_lr=1e-8)
        assert abs(float(sched[0]) - opt.param_groups[0]["lr"]) < 1e-12 + 1e-6 * opt.param_groups[0]["lr"], it
        assert float((theta - ref_p.detach()).abs().max()) < 2e-6, it
    assert opt.param_groups[0]["lr"] < 1e-2          # the schedule actually fired
    assert float(sched[3]) == len(losses)


def test_one_kernel_iteration_update_matches_torch(emu_backend):
    """insr_iteration_update (DeviceOptimizer.update): Adam over TWO nets, gradients zeroed, ReduceLROnPlateau on the
    main loss, step counter and the loss log -- against torch.optim.Adam + ReduceLROnPlateau (base/baseModel.py:55-81)"""
    from insr_pde_b200 import fused
    torch.manual_seed(1)
    nets = [ib.MLP(2, 2, 3, 32, nonlinearity="sine"), ib.MLP(2, 1, 3, 32, nonlinearity="sine")]
    refs = [torch.nn.Parameter(n.flat_theta().detach().clone()) for n in nets]
    opt = torch.optim.Adam(refs, lr=1e-2)
    sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, min_lr=1e-8, patience=3)
    dev_opt = fused.DeviceOptimizer(nets, 1e-2, patience=3)
    hist = torch.full((32, 2), -1.0)
    idx = torch.zeros(1, dtype=torch.long)
    losses = [1.0, 0.9, 0.95, 0.95, 0.96, 0.97, 0.98, 0.5, 0.6, 0.6, 0.6, 0.6, 0.6, 0.6]
    for it, lv in enumerate(losses):
        for n, r in zip(nets, refs):
            g = torch.randn(r.numel()) * (1 + it)
            fused.flat_grad(n).copy_(g)
            r.grad = g.clone()
        opt.step()
        sch.step(lv)
        dev_opt.update(torch.tensor([7.0 + it, lv]), 1, hist, idx)          # 'main' is the second entry here
        assert abs(dev_opt.lr - opt.param_groups[0]["lr"]) < 1e-12 + 1e-6 * opt.param_groups[0]["lr"], it
        for n, r in zip(nets, refs):
            assert float((n.flat_theta() - r.detach()).abs().max()) < 2e-6, it
            assert not fused.flat_grad(n).any()                              # zero_grad for the next iteration
    assert opt.param_groups[0]["lr"] < 1e-2 and int(idx) == len(losses) and float(dev_opt.sched[3]) == len(losses)
    assert torch.equal(hist[:len(losses), 1], torch.tensor(losses)) and torch.equal(hist[:len(losses), 0], 7.0 + torch.arange(len(losses)))
    assert bool((hist[len(losses):] == -1).all()) and int(dev_opt._ticket) == 0
    # clear_losses: the loss slots are accumulators of the closures' kernels -- logged, then left zeroed for the next iteration
    slots = torch.tensor([3.0, 0.25, 9.0])
    dev_opt.update(slots[:2], 1, hist, idx, clear_losses=True)
    assert float(hist[len(losses), 0]) == 3.0 and float(hist[len(losses), 1]) == 0.25
    assert slots.tolist() == [0.0, 0.0, 9.0]                                 # only the n_losses consumed slots
    dev_opt.update(slots[:2], 1, hist, idx, clear_losses=False)
    slots[:2] = torch.tensor([1.0, 2.0])
    dev_opt.update(slots[:2], 1, hist, idx)
    assert slots.tolist() == [1.0, 2.0, 9.0]
