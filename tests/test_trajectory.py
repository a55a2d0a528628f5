"""Short seeded Adam trajectory of advect1D (initialize + one step, K iterations each) on the
collocation-sample stream recorded from the REAL reference (tests/golden/trajectory_advection.npz).

Tolerance: Adam divides by sqrt(v) so rounding-level gradient differences are amplified in the
first iterations; after 2 x 25 iterations at lr 1e-4 the fused path must stay within
5e-5 absolute (weights are O(1)) of the reference trajectory and reproduce the loss history
to 1e-3 relative."""
import numpy as np
import pytest
import torch

import insr_pde_b200 as ib
from conftest import load_golden
from oracle import torch_port as tp, training


def make_ours(device):
    def mk(theta, D, O, H, L):
        n = ib.MLP(D, O, L, H, nonlinearity="sine").to(device)
        with torch.no_grad():
            n.flat_theta().copy_(torch.from_numpy(theta).to(device))
        return n
    return mk


def check(out, g, atol, rtol_hist):
    for k in ("theta_after_init", "theta_after_step"):
        assert np.abs(out[k] - g[k]).max() < atol, (k, np.abs(out[k] - g[k]).max())
    for k in ("hist_initialize", "hist_advect"):
        a, b = np.array(out[k]), g[k]
        assert np.abs(a - b).max() <= rtol_hist * np.abs(b).max(), k


def test_port_replays_reference_trajectory_bit_exactly():
    g = load_golden("trajectory_advection")
    torch.set_num_threads(1)
    out = training.replay_advection(lambda th, D, O, H, L: tp.RefMLP(D, O, L, H).load_flat_theta(th), tp, g)
    check(out, g, 1e-7, 1e-6)


def test_fused_modules_follow_reference_trajectory_emulated(emu_backend):
    g = load_golden("trajectory_advection")
    out = training.replay_advection(make_ours("cpu"), ib, g)
    check(out, g, 5e-5, 1e-3)


@pytest.mark.gpu
def test_fused_modules_follow_reference_trajectory_on_device():
    g = load_golden("trajectory_advection")
    out = training.replay_advection(make_ours("cuda"), ib, g, device="cuda")
    print("max |dtheta| after init/step:", np.abs(out["theta_after_init"] - g["theta_after_init"]).max(),
          np.abs(out["theta_after_step"] - g["theta_after_step"]).max())
    check(out, g, 5e-5, 1e-3)


# ---------------------------------------------------------------------------------------------------------------
# fluid2Dtlgn: initialize() + one step() (advect -> pressure -> projection), 12 Adam iterations per loop (20^2 points) on the sample
# stream recorded from the REAL reference (tests/golden/trajectory_fluid.npz, oracle/make_goldens.py:trajectory_fluid).
# North-star clause: per-frame velocity fields within 1e-4 (relative to max |u|) of the reference.
# ---------------------------------------------------------------------------------------------------------------
def check_fluid(out, g, atol_theta, rtol_field, rtol_hist):
    for k in ("theta_after_init.velocity", "theta_after_step.velocity", "theta_after_step.pressure"):
        assert np.abs(out[k] - g[k]).max() < atol_theta, (k, np.abs(out[k] - g[k]).max())
    for k in ("frame0", "frame1"):
        err = np.abs(out[k] - g[k]).max() / np.abs(g[k]).max()
        assert err < rtol_field, (k, err)
    for k in ("initialize", "advect_velocity", "solve_pressure", "projection"):
        a, b = np.array(out["hist_" + k]), g["hist_" + k]
        assert a.shape == b.shape and np.abs(a - b).max() <= rtol_hist * np.abs(b).max(), k


def test_port_replays_reference_fluid_trajectory():
    g = load_golden("trajectory_fluid")
    torch.set_num_threads(1)
    out = training.replay_fluid(lambda th, D, O, H, L: tp.RefMLP(D, O, L, H).load_flat_theta(th), tp, g)
    check_fluid(out, g, 1e-6, 1e-6, 1e-5)


def test_fused_modules_follow_reference_fluid_trajectory_emulated(emu_backend):
    g = load_golden("trajectory_fluid")
    out = training.replay_fluid(make_ours("cpu"), ib, g)
    check_fluid(out, g, 5e-5, 1e-4, 1e-3)


@pytest.mark.gpu
def test_fused_modules_follow_reference_fluid_trajectory_on_device():
    g = load_golden("trajectory_fluid")
    out = training.replay_fluid(make_ours("cuda"), ib, g, device="cuda")
    for k in ("frame0", "frame1"):
        print(k, "max rel field error:", np.abs(out[k] - g[k]).max() / np.abs(g[k]).max())
    check_fluid(out, g, 5e-5, 1e-4, 1e-3)
