"""insr_siren_lsq_step (k_tc_bwd<LSQ>) and insr_siren_target launch time against the batch size (CUDA events, 200 reps)"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import _ops, fused
torch.manual_seed(0)
vel, pres = ib.MLP(2, 2, 3, 32, nonlinearity="sine").cuda(), ib.MLP(2, 1, 3, 32, nonlinearity="sine").cuda()


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for n in (1, 128, 384, 1024, 4096, 8192, 16384, 18944, 37888, 75776, 1 << 20):
    x = torch.rand(n, 2, device="cuda") * 2 - 1
    tv, tp = torch.randn(n, 2, device="cuda"), torch.randn(n, 1, device="cuda")
    lv, lp = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    gv, gp = fused.flat_grad(vel), fused.flat_grad(pres)
    wsv = None
    t_v = timed(lambda: _ops.siren_lsq_step(vel.desc, vel.flat_theta(), x, 0, [[1.0, 0.0], [0.0, 1.0]], None, None, tv, 1.0 / n, loss_out=lv, gtheta=gv))
    t_p = timed(lambda: _ops.siren_lsq_step(pres.desc, pres.flat_theta(), x, 2, [[0.0]], None, [[1.0]], tp, 1.0 / n, loss_out=lp, gtheta=gp))
    t_t = timed(lambda: _ops.siren_target(x, 2, dict(net=vel, order=0), dict(net=vel, order=0, cy=[[1.0, 0.0], [0.0, 1.0]]), mode=1, dt=0.05))
    t_f = timed(lambda: _ops.siren_forward(vel.desc, vel.flat_theta(), x, 0))
    print(f"N {n:8d} tiles {(n + 127) // 128:6d}:  lsq velocity (S=1) {t_v:8.1f} us   lsq pressure (S=4) {t_p:8.1f} us   target backtrace {t_t:7.1f} us   forward value {t_f:7.1f} us", flush=True)
