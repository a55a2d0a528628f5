"""tcgen05 forward kernel vs the FFMA fused kernel and the fp64 oracle (run on the B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import insr_pde_b200 as ib
from insr_pde_b200 import _lib, _ops
from oracle import siren_fwdmode as fm

def rel(a, b):
    a = a.detach().double().cpu().numpy(); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

torch.manual_seed(0)
for (D, O, H, L, order, N) in [(2, 1, 32, 3, 2, 1000), (2, 2, 32, 3, 1, 777), (1, 1, 20, 2, 1, 300), (2, 1, 32, 1, 0, 128), (2, 2, 32, 2, 2, 5000)]:
    net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
    theta = net.flat_theta()
    x = torch.rand(N, D, device="cuda") * 2 - 1
    d_tc = _lib.make_desc(D, O, H, L)
    outs_tc = _ops.siren_forward(d_tc, theta, x, order)
    torch.cuda.synchronize()
    outs_ff = _ops.siren_forward(_lib.make_desc(D, O, H, L, flags=_lib.FLAG_NO_TENSOR), theta, x, order)
    ref = fm.forward(theta.double().cpu().numpy(), x.double().cpu().numpy(), D, O, H, L, order)
    names = ["y", "jac", "lap"][:len(outs_tc)]
    print((D, O, H, L, order, N), {k: (f"tc {rel(a, ref[k]):.1e}", f"ffma {rel(b, ref[k]):.1e}") for k, a, b in zip(names, outs_tc, outs_ff)}, flush=True)

# timing at size
D, O, H, L, order, N = 2, 1, 32, 3, 2, 1 << 22
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
x = torch.rand(N, D, device="cuda") * 2 - 1
for flags, name in ((_lib.FLAG_NO_TENSOR, "ffma"), (0, "tcgen05")):
    desc = _lib.make_desc(D, O, H, L, flags=flags)
    for _ in range(3):
        _ops.siren_forward(desc, theta, x, order)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        _ops.siren_forward(desc, theta, x, order)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(name, f"{ms:.3f} ms  {N / ms / 1e3:.1f} Mpts/s  {24960 * N / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)

# small-batch latency (script sizes)
for N in (16384, 324, 162):
    x = torch.rand(N, D, device="cuda") * 2 - 1
    for flags, name in ((_lib.FLAG_NO_TENSOR, "ffma"), (0, "tcgen05")):
        desc = _lib.make_desc(D, O, H, L, flags=flags)
        for order_ in (0, 2):
            for _ in range(5):
                _ops.siren_forward(desc, theta, x, order_)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50):
                _ops.siren_forward(desc, theta, x, order_)
            b.record(); torch.cuda.synchronize()
            print(f"N={N} order={order_} {name}: {a.elapsed_time(b) / 50 * 1e3:.1f} us/call", flush=True)
