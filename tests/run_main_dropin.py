"""Helper (run as a subprocess): the reference's UNMODIFIED main.py through ``insr_pde_b200.patch.run_main`` -- the
launcher INTEGRATION.md describes -- with the given main.py arguments.  ``cuda`` as first argument: the real library on
cuda:0 (GPU box); ``cpu``: C-ABI calls routed to the emulation build and the reference's hard-coded cuda:0 proxied to
the CPU (build container); ``cuda-reference`` / ``cuda-reference64``: the reference alone (no patch), stock PyTorch on
cuda:0 in fp32 / fp64 -- the other side of the trajectory comparisons.  Prints one JSON object: the files main.py wrote
under <exp_dir>/results and the loss history."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

device, fused_closures, graphed, argv = sys.argv[1], sys.argv[2] in ("1", "2"), sys.argv[2] == "2", sys.argv[3:]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from insr_pde_b200 import patch  # noqa: E402
from oracle import ref_loader  # noqa: E402

root = ref_loader.REF_ROOT
if device == "cpu":
    import build_emu
    from insr_pde_b200 import _lib, _ops
    torch.set_num_threads(1)
    _lib._LIB = _lib.Library(build_emu.build_emu())
    _ops._require_cuda = lambda t: None
    _ops._stream = lambda dev: None
    patch.install(root)
    ref_loader._patch_torch(True)
    import base.baseModel as bm
    bm.torch = ref_loader._TorchCpuProxy(torch)

reference_only = "reference" in device       # cuda-reference[64|-perturbed]; cpu-reference (build container)
if reference_only:
    if device.endswith("64"):
        torch.set_default_dtype(torch.float64)
    ref_loader.load(cpu=device.startswith("cpu"))    # stubs + scheduler shim only; base / fluid / ... stay the reference's own
else:
    patch.install(root)
import base.baseModel as bm  # noqa: E402
import elasticity.model as _ela  # noqa: E402

if device.endswith("-perturbed"):
    # sensitivity yardstick of a trajectory: the same run with every initial weight moved by ~1e-7 relative (about one
    # fp32 ulp) -- how far apart two runs end up that differ by rounding only
    _orig_create = bm.BaseModel._create_network

    def _create_network(self, *a, **k):
        net = _orig_create(self, *a, **k)
        g = torch.Generator(device="cpu").manual_seed(99)
        with torch.no_grad():
            for p_ in net.parameters():
                p_.mul_(1.0 + 1e-7 * torch.randn(p_.shape, generator=g).to(p_.device))
        return net
    bm.BaseModel._create_network = _create_network

# elasticity writes its frames through open3d (absent here): keep the sampled deformation field as .npy instead
_ela.write_pointcloud_to_file = lambda path, values: np.save(path + ".npy", np.asarray(values))

hist = []
ref_update = None


def spy(self, loss_dict, __orig=bm.BaseModel._update_network):
    hist.append([float(v) for v in loss_dict.values()])
    return __orig(self, loss_dict)


if not fused_closures:
    bm.BaseModel._update_network = spy
torch.manual_seed(123)
np.random.seed(123)
import time  # noqa: E402
_t0 = time.perf_counter()
if reference_only:
    import runpy
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [os.path.join(root, "main.py"), *argv]
    os.chdir(root)
    try:
        runpy.run_path(sys.argv[0], run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
else:
    patch.run_main(argv, root, fused_closures=fused_closures, graphed=graphed)
if torch.cuda.is_available():
    torch.cuda.synchronize()
_elapsed = time.perf_counter() - _t0
proj = argv[argv.index("--proj_dir") + 1]
tag = argv[argv.index("--tag") + 1]
res = os.path.join(proj, tag, "results")
files = sorted(os.listdir(res))
out = {"seconds": round(_elapsed, 3), "files": files, "hist": hist, "ckpts": sorted(os.listdir(os.path.join(proj, tag, "model"))), "results_dir": res}
for f in files:
    if f.endswith(".npy"):
        out.setdefault("npy", {})[f] = [float(np.abs(np.load(os.path.join(res, f))).max())]
print(json.dumps(out))
