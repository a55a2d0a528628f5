#!/bin/bash
# ncu --set full capture of the tcgen05 forward and backward kernels (262144 points)
mkdir -p gpurun_out
TAG=${1:-tc}
cat > /tmp/prof_tc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import _lib, _ops
D, O, H, L, order, N = 2, 1, 32, 3, 2, 262144
torch.manual_seed(0)
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
x = torch.rand(N, D, device="cuda") * 2 - 1
cots = [torch.randn(s, device="cuda") / N for s in _ops.out_shapes(net.desc, N, order)]
desc = _lib.make_desc(D, O, H, L, flags=0)
for _ in range(2):
    _ops.siren_forward(desc, theta, x, order)
    g, _ = _ops.siren_backward(desc, theta, x, order, *cots)
torch.cuda.synchronize()
print("ok", float(g.abs().sum()))
PY
python /tmp/prof_tc.py > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 2 -c 2 -f -o gpurun_out/prof_$TAG \
    python /tmp/prof_tc.py > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_full_$TAG.log
