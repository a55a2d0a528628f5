"""kernel-time breakdown (torch.profiler / CUPTI) of fwd + bwd steps of one workload at a given batch:
python tools/step_kernels.py <workload> <points> [keep_tape]"""
import os, sys, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import insr_pde_b200 as ib
from insr_pde_b200 import _ops

wl = sys.argv[1] if len(sys.argv) > 1 else "elasticity2Dstretch"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
keep = len(sys.argv) > 3 and sys.argv[3] == "1"
D, O, H, L, order, _ = bench.WORKLOADS[wl]
torch.manual_seed(0)
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
x = torch.rand(n, D, device="cuda") * 2 - 1
cots = [torch.randn(s, device="cuda") / n for s in _ops.out_shapes(net.desc, n, order)]
g = torch.zeros_like(theta)


def step():
    if keep:
        outs, tape = _ops.siren_forward(net.desc, theta, x, order, keep_tape=True)
        _ops.siren_backward(net.desc, theta, x, order, *cots, gtheta=g, tape=tape)
    else:
        _ops.siren_forward(net.desc, theta, x, order)
        _ops.siren_backward(net.desc, theta, x, order, *cots, gtheta=g)


for _ in range(3):
    step()
torch.cuda.synchronize()
R = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
for e in ev:
    a = agg[e.name[:90]]; a[0] += 1; a[1] += e.time_range.end - e.time_range.start
tot = sum(a[1] for a in agg.values())
print(f"{wl} N={n} keep_tape={keep}: span per step {(t1 - t0) / R:.1f} us, kernel time per step {tot / R:.1f} us, kernels per step {sum(a[0] for a in agg.values()) / R:.1f}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{t / R:9.1f} us {c / R:5.1f}x  {k}")
