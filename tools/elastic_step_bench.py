"""seconds per elasticity training iteration / time step on one B200 (elasticity2Dstretch and bunny-sized configs):
ElasticityStepper eager and CUDA-graphed, beside the reference closure restated in stock PyTorch on the same GPU
(bench.reference_elasticity_iteration: the reference's own ElasticityModel).  Usage: python tools/elastic_step_bench.py [iters]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import fused

K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
import bench
CASES = bench.ELASTIC_CASES
for name, c in CASES.items():
    dim = c["dim"]
    kw = dict(energy=c["energy"], ratio_arap=c["ratio_arap"], ratio_volume=c["ratio_volume"], ratio_kinematics=c["ratio_kinematics"],
              ratio_constraint=c["ratio_constraint"], ratio_collide=c["ratio_collide"],
              external_force=torch.tensor(c["ext"][:dim], device="cuda"), external_force_timesteps=c["ext_T"],
              constraint_offset_right=torch.tensor(c["off"][:dim], device="cuda"), plane_height=c["plane"],
              circle_center=torch.tensor(c["center"][:dim], device="cuda"), circle_radius=c["radius"])
    res = {}
    for graphed in (False, True):
        torch.manual_seed(0)
        nets = [ib.MLP(dim, dim, 3, c["H"], nonlinearity="sine").cuda() for _ in range(3)]
        mesh = None
        if c.get("mesh"):
            from insr_pde_b200 import medit
            mesh = medit.load_normalized(bench.find_mesh(c["mesh"]), dim, device="cuda")
        st = fused.ElasticityStepper(*nets, dim, dt=c["dt"], sample_resolution=c["sr"], graphed=graphed, mesh=mesh, **kw)
        st.initialize(5)
        st.step(10)                                   # warm-up (graph capture happens here)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h = st.step(K)
        torch.cuda.synchronize(); res[graphed] = (time.perf_counter() - t0) / K
        npts = st._interior(c["sr"]).shape[0]
    # the reference's own ElasticityModel._solve_deformation loop as stock PyTorch on the same GPU (autograd jacobian + torch.svd + Adam)
    tref = bench.reference_elasticity_iteration("cuda", name, iters=max(5, K // 10))
    print(f"{name}: {npts} points/iter  eager {res[False]*1e3:.3f} ms/iter  graphed {res[True]*1e3:.3f} ms/iter "
          f"({npts / res[True] / 1e6:.1f} Mpts/s)  stock PyTorch on the same GPU {tref*1e3:.2f} ms/iter  "
          f"-> {tref / res[True]:.1f}x", flush=True)
