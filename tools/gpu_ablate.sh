#!/bin/bash
# timing of the tcgen05 backward kernel (library possibly built with an INSR_ABL_* ablation macro: results are wrong then)
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import _lib, _ops
D, O, H, L, order, N = 2, 1, 32, 3, 2, 1 << 22
torch.manual_seed(0)
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta(); desc = net.desc
x = torch.rand(N, D, device="cuda") * 2 - 1
cots = [torch.randn(s, device="cuda") / N for s in _ops.out_shapes(desc, N, order)]
gth = torch.zeros_like(theta)
fn = lambda: _ops.siren_backward(desc, theta, x, order, *cots, gtheta=gth)
for _ in range(3): fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): fn()
b.record(); torch.cuda.synchronize()
print(f"bwd {a.elapsed_time(b) / 10:.3f} ms", flush=True)
PY
