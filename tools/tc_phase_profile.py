"""Phase timing of the tcgen05 backward kernel (library built with -DINSR_TC_PROFILE): cycles per phase of warp 0,
averaged per tile.  usage: python tools/tc_phase_profile.py [points]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import insr_pde_b200 as ib
from insr_pde_b200 import _lib, _ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
D, O, H, L, order = 2, 1, 32, 3, 2
torch.manual_seed(0)
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
desc = net.desc
lib = _lib.get_lib()
x = torch.rand(N, D, device="cuda") * 2 - 1
cots = [torch.randn(s, device="cuda") / N for s in _ops.out_shapes(desc, N, order)]
nb = lib.workspace_bytes(desc, N, order, True)
ws = torch.zeros(nb + 16, dtype=torch.uint8, device="cuda")
gth = torch.zeros_like(theta)
stream = torch.cuda.current_stream().cuda_stream
S = 4
tape_bytes = 148 * (L + 1) * 2 * (S + 1) * 512 * 16
names = ["layer 0 fwd", "fwd: sync + issue", "fwd: wait MMA", "fwd: epilogue", "output layer / cotangents", "rev: adjoint",
         "rev: reduce + dgrad operands", "rev: a_{l-1} from tape", "rev: wait slots + wgrad operands", "rev: sync + dgrad issue", "rev: wait dgrad",
         "rev: wgrad issue + tcgen05.ld", "first layer + tile end"]
for rep in range(2):
    ws.zero_()
    lib.backward(desc, theta.data_ptr(), x.data_ptr(), N, order, cots[0].data_ptr(), cots[1].data_ptr(), cots[2].data_ptr(),
                 gth.data_ptr(), None, ws.data_ptr(), nb, stream)
    torch.cuda.synchronize()
off = (ws.data_ptr() + 15) // 16 * 16 - ws.data_ptr() + tape_bytes
prof = ws[off:off + 128].view(torch.int64).cpu().tolist()
tiles = prof[15]
tot = sum(prof[:13])
print(f"tiles {tiles}, cycles per tile {tot / tiles:.0f}")
for i, nme in enumerate(names):
    print(f"  {nme:32s} {prof[i] / tiles:9.0f} cycles/tile  {100 * prof[i] / tot:5.1f}%")
