"""Tiny driver for ncu: a few launches of the hot-path kernels on one workload."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import insr_pde_b200 as ib
from insr_pde_b200 import _ops

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="fluid2Dtlgn.pressure")
ap.add_argument("--points", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lsq", action="store_true")
a = ap.parse_args()
D, O, H, L, order, _ = bench.WORKLOADS[a.workload]
torch.manual_seed(0)
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
x = torch.rand(a.points, D, device="cuda") * 2 - 1
cots = [torch.randn(s, device="cuda") / a.points for s in _ops.out_shapes(net.desc, a.points, order)]
target = torch.randn(a.points, 1, device="cuda")
for _ in range(a.reps):
    outs = _ops.siren_forward(net.desc, theta, x, order)
    g, _ = _ops.siren_backward(net.desc, theta, x, order, *cots)
    if a.lsq and _ops._lib.get_lib().kernel_family(net.desc, order, True) == 1:
        cy = [[0.0] * O]
        cl = [[1.0] * O] if order == 2 else None
        _ops.siren_lsq_step(net.desc, theta, x, order, cy if order == 2 else [[1.0] * O], None, cl, target, 1.0 / a.points)
torch.cuda.synchronize()
print("ok", float(g.abs().sum()))
