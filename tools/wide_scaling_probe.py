"""k_wide_tc duration against the number of 128-point tiles (torch.profiler kernel times, back-to-back launches):
separates the per-launch fixed cost from the per-tile cost and shows the wave steps.  usage: [H] [D] [order]"""
import os, sys, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
import insr_pde_b200 as ib
from insr_pde_b200 import _ops
H = int(sys.argv[1]) if len(sys.argv) > 1 else 68
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2
order = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.manual_seed(0)
net = ib.MLP(D, D, 3, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
for tiles in (1, 2, 74, 148, 157, 296, 297, 444, 592, 1184, 4736):
    x = torch.rand(tiles * 128, D, device="cuda") * 2 - 1
    for _ in range(5):
        _ops.siren_forward(net.desc, theta, x, order)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            _ops.siren_forward(net.desc, theta, x, order)
        torch.cuda.synchronize()
    agg = collections.defaultdict(list)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            agg[e.name.split("(")[0][-40:]].append(e.time_range.end - e.time_range.start)
    line = "  ".join(f"{k.split('::')[-1]} {sum(v) / len(v):.1f}us" for k, v in sorted(agg.items()))
    print(f"tiles {tiles:5d}: {line}", flush=True)
