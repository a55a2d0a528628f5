"""kernel-time breakdown of ONE graphed elasticity iteration (torch.profiler / CUPTI over graph replays)"""
import os, sys, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
import insr_pde_b200 as ib
from insr_pde_b200 import fused
case = sys.argv[1] if len(sys.argv) > 1 else "2d"
if case == "2d":
    dim, H, sr = 2, 68, 100
    kw = dict(energy=["arap", "constraint", "constraint_right", "volume"], ratio_arap=1.0, ratio_volume=1e3, ratio_kinematics=1.0,
              ratio_constraint=1e4, ratio_collide=1.0, external_force=torch.zeros(2, device="cuda"), external_force_timesteps=5,
              constraint_offset_right=torch.tensor([2.0, 0.0], device="cuda"), plane_height=-2.0,
              circle_center=torch.tensor([0.0, -2.0], device="cuda"), circle_radius=1.0)
else:
    dim, H, sr = 3, 66, 24
    kw = dict(energy=["arap", "kinematics", "collision", "external", "volume"], ratio_arap=1e2, ratio_volume=1e3, ratio_kinematics=1.0,
              ratio_constraint=1e3, ratio_collide=1e6, external_force=torch.tensor([0., 0., -1e2], device="cuda"), external_force_timesteps=5,
              constraint_offset_right=torch.tensor([1.0, 0.0, 0.0], device="cuda"), plane_height=-0.9,
              circle_center=torch.tensor([0.0, -2.0, 0.0], device="cuda"), circle_radius=1.0)
torch.manual_seed(0)
nets = [ib.MLP(dim, dim, 3, H, nonlinearity="sine").cuda() for _ in range(3)]
st = fused.ElasticityStepper(*nets, dim, dt=0.05, sample_resolution=sr, graphed=True, **kw)
st.initialize(3)
st.step(20)
g = st._loops[("solve", True)].graph
torch.cuda.synchronize()
R = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R):
        g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
for e in ev:
    a = agg[e.name[:80]]; a[0] += 1; a[1] += e.time_range.end - e.time_range.start
tot = sum(a[1] for a in agg.values())
print(f"case {case}: span per iteration {(t1 - t0) / R:.1f} us, sum of kernel times per iteration {tot / R:.1f} us, kernels per iteration {sum(a[0] for a in agg.values()) / R:.1f}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t / R:8.1f} us {c / R:5.1f}x  {k}")
