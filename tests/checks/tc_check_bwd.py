"""tcgen05 backward / lsq kernels vs the FFMA kernels and the fp64 oracle (run on the B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import insr_pde_b200 as ib
from insr_pde_b200 import _lib, _ops
from oracle import siren_fwdmode as fm

def rel(a, b):
    a = a.detach().double().cpu().numpy(); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

torch.manual_seed(0)
cases = [(2, 1, 32, 3, 2, 128), (2, 1, 32, 3, 2, 1000), (2, 2, 32, 3, 1, 777), (1, 1, 20, 2, 1, 300), (2, 1, 32, 1, 0, 130), (2, 2, 32, 2, 2, 5000)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (D, O, H, L, order, N) in cases:
    net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
    theta = net.flat_theta()
    x = torch.rand(N, D, device="cuda") * 2 - 1
    shapes = _ops.out_shapes(net.desc, N, order)
    cots = [torch.randn(s, device="cuda") for s in shapes]
    d_tc = _lib.make_desc(D, O, H, L, flags=0)
    d_ff = _lib.make_desc(D, O, H, L, flags=_lib.FLAG_NO_TENSOR)
    g_tc, gx_tc = _ops.siren_backward(d_tc, theta, x, order, *cots, need_gx=True)
    torch.cuda.synchronize()
    g_ff, gx_ff = _ops.siren_backward(d_ff, theta, x, order, *cots, need_gx=True)
    kw = dict(gy=cots[0].double().cpu().numpy())
    if order >= 1: kw["gjac"] = cots[1].double().cpu().numpy()
    if order == 2: kw["glap"] = cots[2].double().cpu().numpy()
    gref, gxref = fm.backward(theta.double().cpu().numpy(), x.double().cpu().numpy(), D, O, H, L, order, **kw)
    print((D, O, H, L, order, N), f"gtheta tc {rel(g_tc, gref):.1e} ffma {rel(g_ff, gref):.1e} | gx tc {rel(gx_tc, gxref):.1e} ffma {rel(gx_ff, gxref):.1e}", flush=True)
    P = theta.numel()
    offs = net.param_slices()
    worst = max(((rel(g_tc[o:o+n], gref[o:o+n]), i) for i, (o, n, _) in enumerate(offs)))
    print("    per-tensor rel err:", [f"{rel(g_tc[o:o+n], gref[o:o+n]):.1e}" for (o, n, _) in offs], flush=True)

if len(sys.argv) > 2:
    D, O, H, L, order, N = 2, 1, 32, 3, 2, 1 << 22
    net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
    theta = net.flat_theta()
    x = torch.rand(N, D, device="cuda") * 2 - 1
    cots = [torch.randn(s, device="cuda") / N for s in _ops.out_shapes(net.desc, N, order)]
    tgt = torch.randn(N, 1, device="cuda")
    gth = torch.zeros_like(theta); loss = torch.zeros(1, device="cuda")
    for flags, name in ((_lib.FLAG_NO_TENSOR, "ffma"), (0, "tcgen05")):
        desc = _lib.make_desc(D, O, H, L, flags=flags)
        for what, fn in (("bwd", lambda: _ops.siren_backward(desc, theta, x, order, *cots, gtheta=gth)),
                         ("lsq", lambda: _ops.siren_lsq_step(desc, theta, x, order, [[0.0]], None, [[1.0]], tgt, 1.0 / N, loss_out=loss, gtheta=gth))):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): fn()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(name, what, f"{ms:.3f} ms  {N / ms / 1e3:.1f} Mpts/s", flush=True)
