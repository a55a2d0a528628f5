import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import fused
torch.manual_seed(0)
vel, prev, pres = (ib.MLP(2, o, 3, 32, nonlinearity="sine").cuda() for o in (2, 2, 1))
st = fused.FluidStepper(vel, prev, pres, dt=0.05, sample_resolution=128, lr=1e-4, graphed=True)
st.initialize(fused.taylorgreen_velocity, 20)
for name, closure in (("advect", lambda: fused.fluid_advect_velocity(vel, prev, *st._samples(1), 0.05)),
                      ("pressure", lambda: fused.fluid_solve_pressure(vel, pres, *st._samples(1))),
                      ("project", lambda: fused.fluid_projection(vel, prev, pres, *st._samples(1)))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loop = fused.GraphedLoop([vel, pres], 1e-4, closure)
    loop.run(2)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(200):
        loop.graph.replay()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: setup+capture {1e3*(t1-t0):.1f} ms, replay {1e6*(t2-t1)/200:.1f} us/iter, tensor={'off' if os.environ.get('INSR_NO_TENSOR')=='1' else 'on'}", flush=True)
