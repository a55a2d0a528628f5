"""Import the REAL reference: the read-only tree at /root/reference in the build container, or
the verbatim copy ``oracle/_ref`` that ``oracle/build_ref.py`` makes from it (git-ignored, shipped
to the GPU box with the snapshot -- /root/reference does not exist there).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by ``oracle/make_goldens.py``, the
pinning tests, the GPU drop-in tests and the baseline legs of ``bench.py``; never by the product.

The reference cannot be imported unmodified here: ``pytorch3d``,
``tensorboardX``, ``matplotlib``, ``meshio``, ``open3d`` are missing, torch>=2.7 dropped
``ReduceLROnPlateau(verbose=)`` (base/baseModel.py:61-62) and ``base/baseModel.py:25``
hard-codes ``cuda:0``.  The recipe below leaves every reference file untouched:
  1. ``sys.modules`` stubs for the missing third-party modules,
  2. a ``ReduceLROnPlateau`` subclass that swallows ``verbose=``,
  3. (cpu=True) a proxy for the ``torch`` global of ``base.baseModel`` whose ``device()``
     returns cpu, and no-op ``.cuda()`` on modules / tensors,
  4. ``make_cfg`` builds an argparse.Namespace instead of ``config.Config`` (which parses
     sys.argv, prompts and rmtree's, config.py:44-48).
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import tempfile
import types

_SHIPPED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _resolve_root():
    env = os.environ.get("INSR_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", _SHIPPED]:
        if cand and os.path.isfile(os.path.join(cand, "base", "networks.py")):
            return cand
    return env or "/root/reference"


REF_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "base", "networks.py"))


class _Anything:
    """Attribute/call sink used for the plotting / logging stubs."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):  # ``fig, ax = plt.subplots()``
        return iter((_Anything(), _Anything()))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    def _missing(key):
        if key.startswith("__"):  # inspect.getmodule() probes __file__ etc. on every module
            raise AttributeError(key)
        return _Anything()

    m.__getattr__ = _missing  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


def _install_stubs():
    if "pytorch3d" not in sys.modules:
        p3 = _stub("pytorch3d")
        p3.ops = _stub("pytorch3d.ops", knn_points=None, knn_gather=None)
    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=_Anything)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.colors = _stub("matplotlib.colors")
        # fluid/visualize.py:31-43 does arithmetic on ``cm.bwr(field)``: the colormap stand-in returns a real RGBA array
        cm = _stub("matplotlib.cm")

        def _colormap(name):
            if name.startswith("__"):
                raise AttributeError(name)

            def ramp(values, *a, **k):
                import numpy as np
                v = np.clip(np.asarray(values, dtype=np.float64), 0.0, 1.0)
                return np.stack([v, v, v, np.ones_like(v)], axis=-1)
            return ramp

        cm.__getattr__ = _colormap  # type: ignore[attr-defined]
        mpl.cm = cm
    if "meshio" not in sys.modules:               # elasticity/model.py:77 reads MEDIT .mesh files through meshio.read
        from insr_pde_b200 import medit
        _stub("meshio", read=medit.read)
    for name in ("open3d", "cupy", "cupyx"):
        if name not in sys.modules:
            _stub(name)


def _patch_torch(cpu: bool):
    import torch

    sched = torch.optim.lr_scheduler
    if not getattr(sched.ReduceLROnPlateau, "_insr_shim", False):
        base = sched.ReduceLROnPlateau

        class ReduceLROnPlateau(base):  # swallows the removed ``verbose`` kwarg
            _insr_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        sched.ReduceLROnPlateau = ReduceLROnPlateau
    if cpu:
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self


class _TorchCpuProxy:
    """stands in for the ``torch`` global of base.baseModel: device(...) -> cpu."""

    def __init__(self, torch):
        self._t = torch

    def device(self, *a, **k):
        return self._t.device("cpu")

    def __getattr__(self, name):
        return getattr(self._t, name)


_LOADED = {}


def load(cpu: bool = True):
    """Returns a namespace with the reference packages: .base .advection .fluid .elasticity"""
    key = bool(cpu)
    if key in _LOADED:
        return _LOADED[key]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    _install_stubs()
    _patch_torch(cpu)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    ns.base = importlib.import_module("base")
    if cpu:
        bm = importlib.import_module("base.baseModel")
        bm.torch = _TorchCpuProxy(torch)
    ns.advection = importlib.import_module("advection")
    ns.fluid = importlib.import_module("fluid")
    try:
        ns.elasticity = importlib.import_module("elasticity")
        if cpu:     # torchgp/per_vertex_areas.py:14 picks "cuda" whenever a GPU is visible (the CPU legs also run on the GPU box)
            pva = importlib.import_module("elasticity.torchgp.per_vertex_areas")
            pva.torch = _TorchCpuProxy(torch)
    except Exception as exc:  # pragma: no cover - optional (needs numpy-only torchgp bits)
        ns.elasticity = None
        ns.elasticity_error = exc
    _LOADED[key] = ns
    return ns


def make_cfg(pde: str, **over):
    """argparse.Namespace with the defaults of config.py:86-168 (+ per-script overrides)."""
    tmp = over.pop("exp_dir", None) or tempfile.mkdtemp(prefix="insr_ref_")
    d = dict(
        pde=pde, proj_dir=tmp, tag="run", gpu_ids=0, exp_dir=tmp,
        log_dir=os.path.join(tmp, "log"), model_dir=os.path.join(tmp, "model"),
        network="siren", num_hidden_layers=3, hidden_features=64, nonlinearity="sine",
        ckpt=None, vis_frequency=10 ** 9, max_n_iters=3, lr=1e-4, sample_resolution=128,
        vis_resolution=32, early_stop=False,
        init_cond=None, dt=0.05, n_timesteps=2, fps=10,
    )
    if pde == "advection":
        d.update(length=4.0, vel=0.25, init_cond="example1", num_hidden_layers=2,
                 hidden_features=20, sample_resolution=5000)
    elif pde == "fluid":
        d.update(init_cond="taylorgreen", hidden_features=32, sample_resolution=128)
    elif pde == "elasticity":
        d.update(dim=2, sample_pattern=["random", "uniform"],
                 energy=["arap", "kinematics", "external", "constraint"],
                 ratio_constraint=1e3, ratio_volume=1e1, ratio_arap=1e0, ratio_collide=1e0,
                 ratio_kinematics=1e0, use_mesh=False, mesh_path="",
                 external_force_timesteps=5, external_force_x=0.0, external_force_y=0.0,
                 external_force_z=0.0, constraint_right_offset_x=1.0,
                 constraint_right_offset_y=0.0, constraint_right_offset_z=0.0,
                 plane_height=-2.0, collide_circle_x=0.0, collide_circle_y=-2.0,
                 collide_circle_z=0.0, collide_circle_radius=1.0,
                 hidden_features=68, sample_resolution=100)
    d.update(over)
    os.makedirs(d["log_dir"], exist_ok=True)
    os.makedirs(d["model_dir"], exist_ok=True)
    return argparse.Namespace(**d)
