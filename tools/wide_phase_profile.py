"""Phase timing of k_wide_tc (library built with -DINSR_WIDE_PROFILE): cycles per phase of warp 1, per 128-point tile.
usage: python tools/wide_phase_profile.py [workload] [points] [order]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import insr_pde_b200 as ib
from insr_pde_b200 import _lib, _ops

wl = sys.argv[1] if len(sys.argv) > 1 else "elasticity2Dstretch"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
D, O, H, L, order, _ = bench.WORKLOADS[wl]
if len(sys.argv) > 3:
    order = int(sys.argv[3])
torch.manual_seed(0)
net = ib.MLP(D, O, L, H, nonlinearity="sine").cuda()
theta = net.flat_theta()
x = torch.rand(N, D, device="cuda") * 2 - 1
lib = _lib.get_lib()
fn = lib.cdll.insr_debug_wide_prof
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_uint64 * 16)()
names = ["first slab loads issued", "split + store slab", "fence + barrier", "issue next loads (+ MMA issue on warp 0)", "wait MMA", "epilogue"]
for what in ("forward", "backward"):
    for rep in range(2):
        fn(buf, 1)
        if what == "forward":
            _ops.siren_forward(net.desc, theta, x, order)
        else:
            outs = _ops.siren_forward(net.desc, theta, x, order)
            fn(buf, 1)
            _ops.siren_backward(net.desc, theta, x, order, *[torch.randn_like(o) for o in outs])
        torch.cuda.synchronize()
    fn(buf, 0)
    tiles = buf[7]
    tot = sum(buf[i] for i in range(6))
    print(f"{wl} {what}: {tiles} CTA tiles, {tot / max(tiles, 1):.0f} cycles per tile")
    for i, nme in enumerate(names):
        print(f"   {nme:44s} {buf[i] / max(tiles, 1):9.0f}  {100 * buf[i] / max(tot, 1):5.1f}%")
