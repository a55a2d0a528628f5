"""The torchrun twin of main.py (python -m insr_pde_b200.patch --insr-dp): the reference's scripts data-parallel over the
GPUs of one box, with the CUDA-graphed iteration (--insr-graphed).  Needs >= 2 GPUs (skipped otherwise): the 2-rank run of
scripts/fluid2Dtlgn.sh's arguments must reproduce the single-GPU run -- same seed, same global sample stream, shards of
it on the ranks, ONE all-reduce of [gradients | losses] per iteration inside the graph -- within the tolerance of a
reordered floating-point sum."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="no reference tree / shipped copy"),
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]

ARGS = ["fluid", "--init_cond", "taylorgreen", "--num_hidden_layers", "3", "--hidden_features", "32", "-sr", "128", "-vr", "32",
        "--dt", "0.05", "-T", "1", "--max_n_iters", "100", "--no-early_stop"]


def _run(tmp, tag, launcher, extra):
    env = dict(os.environ, INSR_REFERENCE_ROOT=ref_loader.REF_ROOT, PYTHONPATH=ROOT)
    cmd = launcher + ["-m", "insr_pde_b200.patch", *extra, "--insr-seed", "5", *ARGS, "--proj_dir", str(tmp), "--tag", tag]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    d = os.path.join(str(tmp), tag, "results")
    return [np.load(os.path.join(d, f)) for f in ("t000.npy", "t001.npy")]


@pytest.mark.parametrize("graphed", [True, False])
def test_two_rank_run_reproduces_single_gpu_frames(tmp_path, graphed):
    mode = ["--insr-graphed"] if graphed else []
    single = _run(tmp_path, "single", [sys.executable, "-W", "ignore"], mode)
    twin = _run(tmp_path, "twin", [sys.executable, "-W", "ignore", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29533"], mode + ["--insr-dp"])
    errs = [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(twin, single)]
    print("2-rank vs 1-rank per-frame max relative field error (graphed =", graphed, "):", errs)
    # frame 0 (the initial-condition fit, 100 iterations) isolates the mechanics -- shards, all-reduce, identical Adam steps:
    # a reordered sum, nothing else; frame 1 adds three loops of 100 iterations each on top of it
    assert errs[0] < 1e-5 and errs[1] < 5e-4, errs
