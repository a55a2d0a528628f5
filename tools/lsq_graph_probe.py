"""GPU time of insr_siren_lsq_step per launch against the batch size, measured through CUDA-graph replays (no host overhead)"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import _ops, fused
torch.manual_seed(0)
vel, pres = ib.MLP(2, 2, 3, 32, nonlinearity="sine").cuda(), ib.MLP(2, 1, 3, 32, nonlinearity="sine").cuda()
K = 10
for n in (128, 1024, 4096, 8192, 12288, 16384, 18944, 37888):
    x = torch.rand(n, 2, device="cuda") * 2 - 1
    tv, tp = torch.randn(n, 2, device="cuda"), torch.randn(n, 1, device="cuda")
    lv, lp = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    gv, gp = fused.flat_grad(vel), fused.flat_grad(pres)
    res = []
    for fn in (lambda: _ops.siren_lsq_step(vel.desc, vel.flat_theta(), x, 0, [[1.0, 0.0], [0.0, 1.0]], None, None, tv, 1.0 / n, loss_out=lv, gtheta=gv),
               lambda: _ops.siren_lsq_step(pres.desc, pres.flat_theta(), x, 2, [[0.0]], None, [[1.0]], tp, 1.0 / n, loss_out=lp, gtheta=gp),
               lambda: _ops.siren_target(x, 2, dict(net=vel, order=0), dict(net=vel, order=0, cy=[[1.0, 0.0], [0.0, 1.0]]), mode=1, dt=0.05)):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(K):
                fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            g.replay()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / (20 * K) * 1e3)
    print(f"N {n:6d} tiles {(n + 127) // 128:4d}: lsq velocity S=1 {res[0]:6.1f} us  lsq pressure S=4 {res[1]:6.1f} us  target backtrace {res[2]:6.1f} us", flush=True)
