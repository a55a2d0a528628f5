"""Helper (run as a subprocess by test_reference_dropin.py): executes the reference's UNMODIFIED
model classes for a few optimisation iterations, either as they are ("reference") or with the
fused field/operator layer patched in ("fused"; "fused_closures" additionally swaps the loss closures for the one-kernel ones; C-ABI calls routed to the
emulation build because the build container has no GPU).  Prints one JSON object with the per-iteration loss history."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

mode, pde = sys.argv[1], sys.argv[2]
on_gpu = len(sys.argv) > 3 and sys.argv[3] == "cuda"     # GPU box: the real library, the reference on cuda:0 as it is
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402

if not on_gpu:
    torch.set_num_threads(1)
if mode in ("fused", "fused_closures"):
    from insr_pde_b200 import _lib, _ops, patch
    if not on_gpu:
        import build_emu
        _lib._LIB = _lib.Library(build_emu.build_emu())
        _ops._require_cuda = lambda t: None
        _ops._stream = lambda device: None
    patch.install(ref_loader.REF_ROOT)          # rebind BEFORE the PDE packages import the names
ref = ref_loader.load(cpu=not on_gpu)           # stubs (+ cpu device proxy), then imports advection/fluid/elasticity
if mode == "fused_closures":                    # second stage: the loss closures themselves (needs the PDE packages imported)
    patch.install_fused_closures()

hist = []


def spy_training_loop(model_cls, names):
    BaseModel = ref.base.BaseModel
    for name in names:
        method = getattr(model_cls, name)
        fn = next(c.cell_contents for c in method.__closure__ if callable(c.cell_contents))

        def wrapped(self, __fn=fn, __name=name):
            d = __fn(self)
            hist.append([__name] + [float(v) for v in d.values()])
            return d
        wrapped.__name__ = name
        setattr(model_cls, name, BaseModel._training_loop(wrapped))


K = 3
torch.manual_seed(123)
np.random.seed(123)                              # torchgp/sample_volume.py:37 draws its barycentric weights with numpy
if pde == "fluid":                               # GPU box: the script's own sizes (scripts/fluid2Dtlgn.sh, advect1D.sh, ...)
    cfg = ref_loader.make_cfg("fluid", sample_resolution=128 if on_gpu else 16, max_n_iters=K)
    Model = ref.fluid.Fluid2DModel
    spy_training_loop(Model, ["_initialize", "_advect_velocity", "_solve_pressure", "_projection"])
elif pde == "advection":
    cfg = ref_loader.make_cfg("advection", sample_resolution=5000 if on_gpu else 300, max_n_iters=K)
    Model = ref.advection.Advection1DModel
    spy_training_loop(Model, ["_initialize", "_advect"])
elif pde == "bunny":                             # scripts/elasticity3Dbunny.sh: the real mesh, 20^3 volume samples + 18 592 vertices
    cfg = ref_loader.make_cfg("elasticity", sample_resolution=20, max_n_iters=K, dim=3, hidden_features=66, dt=0.1,
                              energy=["arap", "kinematics", "collision", "external", "volume"], ratio_volume=1e3, ratio_arap=1e2,
                              ratio_collide=1e6, ratio_kinematics=1e0, external_force_z=-1e2, plane_height=-2.0,
                              use_mesh=True, mesh_path=os.path.join(ref_loader.REF_ROOT, "elasticity", "data", "bunny.mesh"),
                              vis_resolution=100)
    Model = ref.elasticity.ElasticityModel
    spy_training_loop(Model, ["_initialize", "_solve_deformation"])
else:
    big = dict(sample_resolution=100, hidden_features=68) if on_gpu else dict(sample_resolution=8, hidden_features=24)
    cfg = ref_loader.make_cfg("elasticity", max_n_iters=K, dim=2, **big,
                              energy=["arap", "kinematics", "external", "constraint", "volume", "collision_sphere"],
                              external_force_y=-1.0, collide_circle_y=-0.5)
    Model = ref.elasticity.ElasticityModel
    spy_training_loop(Model, ["_initialize", "_solve_deformation"])
model = Model(cfg)
if pde == "elasticity":
    model.sample_resolution_init = 100 if on_gpu else 12       # the reference hard-codes 500 (250k+250k points) for the zero-fit
net = next(iter(model._trainable_networks.values()))
model.initialize()
model.step()
out_dir = os.path.join(cfg.exp_dir, "results")
os.makedirs(out_dir, exist_ok=True)
extra = {}
if pde == "bunny":
    extra["n_points"] = int(model._sample_in_training(model.sample_resolution).shape[0])
    extra["mesh"] = [int(model.mesh_V.shape[0]), int(model.mesh_F.shape[0])]
if pde == "advection":
    model.write_output(out_dir)
elif pde == "fluid":                             # write_output's curl (fluid/model.py:207-213) without the plotting
    import fluid.model as fm_mod
    grid_u, grid_x = model.sample_field(8, return_samples=True)
    jaco, _ = fm_mod.jacobian(grid_u, grid_x)    # the name fluid/model.py bound at import time
    extra["curl"] = (jaco[..., 1, 0] - jaco[..., 0, 1]).detach().reshape(-1).tolist()
    extra["jacobian_fn"] = fm_mod.jacobian.__module__
ck = torch.load(os.path.join(cfg.model_dir, "ckpt_step_t001.pth"))
print(json.dumps({"net_class": type(net).__module__ + "." + type(net).__name__, "hist": hist,
                  "ckpt_keys": sorted(k for k in ck if k.startswith("net_")),
                  "state_keys": sorted(next(v for k, v in ck.items() if k.startswith("net_")).keys()), "extra": extra}))
