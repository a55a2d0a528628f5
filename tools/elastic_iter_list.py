"""three eager iterations of the elasticity2Dstretch closure (for an ncu launch list)"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import fused
dim, H = 2, 68
kw = dict(energy=["arap", "constraint", "constraint_right", "volume"], ratio_arap=1.0, ratio_volume=1e3, ratio_kinematics=1.0,
          ratio_constraint=1e4, ratio_collide=1.0, external_force=torch.zeros(2, device="cuda"), external_force_timesteps=5,
          constraint_offset_right=torch.tensor([2.0, 0.0], device="cuda"), plane_height=-2.0,
          circle_center=torch.tensor([0.0, -2.0], device="cuda"), circle_radius=1.0)
torch.manual_seed(0)
nets = [ib.MLP(dim, dim, 3, H, nonlinearity="sine").cuda() for _ in range(3)]
st = fused.ElasticityStepper(*nets, dim, dt=0.05, sample_resolution=100, graphed=False, **kw)
st.initialize(1)
st.step(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
torch.cuda.synchronize()
print("done")
