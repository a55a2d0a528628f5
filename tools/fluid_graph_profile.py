"""kernel-time breakdown of the graphed fluid2Dtlgn iterations (torch.profiler / CUPTI over graph replays), per training loop"""
import os, sys, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
import insr_pde_b200 as ib
from insr_pde_b200 import fused
torch.manual_seed(0)
vel, prev, pres = (ib.MLP(2, o, 3, 32, nonlinearity="sine").cuda() for o in (2, 2, 1))
st = fused.FluidStepper(vel, prev, pres, dt=0.05, sample_resolution=128, lr=1e-4, graphed=True, device_sampler=True)
st.initialize(fused.taylorgreen_velocity, 20)
st.step(5)
torch.cuda.synchronize()
R = 50
for key, lp in st._loops.items():
    g = lp.graph
    if g is None:
        continue
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(R):                      # loops that prepare ahead alternate two graphs (buffer sets 0 / 1)
            (lp.graph_b if (lp.graph_b is not None and i % 2) else g).replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0, 0.0])
    t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
    for e in ev:
        a = agg[e.name[:90]]; a[0] += 1; a[1] += e.time_range.end - e.time_range.start
    tot = sum(a[1] for a in agg.values())
    print(f"loop {key}: span per iteration {(t1 - t0) / R:.1f} us, sum of kernel times {tot / R:.1f} us, kernels per iteration {sum(a[0] for a in agg.values()) / R:.1f}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"{t / R:8.1f} us {c / R:5.1f}x  {k}")
