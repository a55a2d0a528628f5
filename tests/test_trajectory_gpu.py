"""Per-frame trajectories at the SCRIPTS' own configurations (scripts/fluid2Dtlgn.sh, advect1D.sh, elasticity2Dstretch.sh)
on the B200: the reference's unmodified main.py on the fused layer (patch.run_main) against the reference itself as stock
PyTorch on the same GPU (the shipped copy oracle/_ref), same seed -> same initial weights and the same torch sample
stream, 200 Adam iterations per training loop, several frames.  What is compared is what main.py writes per frame.

Tolerances (relative to the frame's max |value|; Adam divides by sqrt(v), so rounding-level gradient differences are
amplified over the hundreds of iterations of a frame, and every frame starts from the previous one):
    fluid2Dtlgn velocity field      frame 0 (after initialize): 1e-4; frames 1, 2: 5e-4
    advect1D field                  every frame: 2e-4
    elasticity2Dstretch deformation frame 0: 1e-5; frame 1: 3x the reference's own sensitivity to a 1-ulp perturbation of its
                                    initial weights (see the test)
The measured errors are printed; they are recorded in DESIGN.md."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="no reference tree / shipped copy")]

K = "200"
CASES = {
    "fluid2Dtlgn": ["fluid", "--init_cond", "taylorgreen", "--num_hidden_layers", "3", "--hidden_features", "32", "-sr", "128",
                    "-vr", "32", "--dt", "0.05", "-T", "2", "--max_n_iters", K, "--no-early_stop"],
    "advect1D": ["advection", "--init_cond", "example1", "--num_hidden_layers", "2", "--hidden_features", "20", "-sr", "5000",
                 "--dt", "0.05", "-T", "4", "--max_n_iters", K, "--no-early_stop"],
    "elasticity2Dstretch": ["elasticity", "--num_hidden_layers", "3", "--hidden_features", "68", "-sr", "100", "-vr", "50", "-T", "1",
                            "--max_n_iters", K, "--lr", "1e-4", "--dim", "2", "--energy", "arap", "constraint", "constraint_right",
                            "volume", "--ratio_volume", "1e3", "--ratio_arap", "1e0", "--ratio_constraint", "1e4",
                            "--constraint_right_offset_x", "2.0", "--no-early_stop"],
}
_RUNS = {}


def run(case, device, tmp, repeat=0):
    key = (case, device, repeat)
    if key not in _RUNS:
        argv = CASES[case] + ["--proj_dir", str(tmp), "--tag", f"{case}_{device}_{repeat}"]
        res = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "tests", "run_main_dropin.py"), device, "0", *argv],
                             capture_output=True, text=True, timeout=1500)
        assert res.returncode == 0, res.stderr[-3000:]
        _RUNS[key] = json.loads(res.stdout.strip().splitlines()[-1])
    return _RUNS[key]


def frames(out, suffix):
    d = out["results_dir"]
    fs = sorted(f for f in out["files"] if f.endswith(suffix))
    arrs = []
    for f in fs:
        a = np.load(os.path.join(d, f))
        arrs.append(np.asarray(a["arr_0"] if hasattr(a, "files") else a, dtype=np.float64))
    return fs, arrs


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_fluid2dtlgn_frames(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("traj_fluid")
    ours, ref = run("fluid2Dtlgn", "cuda", tmp), run("fluid2Dtlgn", "cuda-reference", tmp)
    (fa, a), (fb, b) = frames(ours, ".npy"), frames(ref, ".npy")
    assert fa == fb == ["t000.npy", "t001.npy", "t002.npy"]
    errs = [rel(x, y) for x, y in zip(a, b)]
    print("fluid2Dtlgn per-frame max relative field error:", errs)
    print(f"fluid2Dtlgn wall clock of main.py (initialize + 2 time steps, 200 iterations per loop, incl. output writing): "
          f"reference's own loop on the fused layer {ours['seconds']} s, reference as stock PyTorch on the same GPU {ref['seconds']} s")
    assert errs[0] < 1e-4 and max(errs[1:]) < 5e-4, errs
    assert len(ours["hist"]) == len(ref["hist"]) == 200 * 7


def test_advect1d_frames(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("traj_adv")
    ours, ref = run("advect1D", "cuda", tmp), run("advect1D", "cuda-reference", tmp)
    (fa, a), (fb, b) = frames(ours, ".npz"), frames(ref, ".npz")
    assert fa == fb and len(fa) == 5
    errs = [rel(x, y) for x, y in zip(a, b)]
    print("advect1D per-frame max relative field error:", errs)
    assert max(errs) < 2e-4, errs


def test_elasticity2dstretch_frames(tmp_path_factory):
    """frame 0 = the zero-deformation fit (500^2 points per iteration), frame 1 = the first stretch step.  The stretch step
    is a stiff, ill-conditioned minimisation (ratio_constraint 1e4 against ratio_arap 1): 200 Adam iterations amplify
    rounding-level differences by orders of magnitude.  The yardstick for frame 1 is therefore measured, not guessed: the
    reference against ITSELF with every initial weight moved by 1e-7 relative (one fp32 ulp).  Ours may sit at most 3x
    that far from the reference (floor 1e-3).  The closure itself is pinned to 1e-4 against the reference in fp32 AND fp64
    (test_gpu_parity.py, closure goldens)."""
    tmp = tmp_path_factory.mktemp("traj_ela")
    ours = run("elasticity2Dstretch", "cuda", tmp)
    ref_a, ref_b = run("elasticity2Dstretch", "cuda-reference", tmp), run("elasticity2Dstretch", "cuda-reference-perturbed", tmp)
    (fa, a), (fb, b), (fc, c) = frames(ours, ".ply.npy"), frames(ref_a, ".ply.npy"), frames(ref_b, ".ply.npy")
    assert fa == fb == fc and len(fa) == 2
    errs = [rel(x, y) for x, y in zip(a, b)]
    spread = [rel(x, y) for x, y in zip(c, b)]
    print("elasticity2Dstretch per-frame max relative deformation error: ours vs reference", errs,
          " reference vs reference with initial weights perturbed by 1e-7", spread)
    assert errs[0] < 1e-5, errs
    assert errs[1] < max(3 * spread[1], 1e-3), (errs, spread)
