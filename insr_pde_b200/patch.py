"""Run the UNMODIFIED reference (main.py, */model.py, base/baseModel.py) on the fused kernels.

``install(reference_root)`` puts the reference tree on sys.path, provides stub modules for
its optional third-party imports that are absent in this image (plotting / logging / mesh
I/O -- none is on the hot path), imports ``base`` and rebinds

    base.networks.get_network, base.baseModel.get_network, base.get_network, base.MLP
    base.diff_ops.{gradient,divergence,laplace,jacobian,hessian} and their re-exports in base
    elasticity.model's ``torch.svd`` (a module-level proxy: batches of 2x2 / 3x3 fp32 CUDA matrices -> insr_svd_small,
    elasticity/model.py:144; torch.svd itself is left alone)

to this package *before* ``advection`` / ``fluid`` / ``elasticity`` bind them by name
(``from base import gradient, ...``: fluid/model.py:5-6, advection/model.py:5,
elasticity/model.py:11).  ``run_main(argv)`` then executes the reference's own main.py.

Nothing here touches the kernels; it is the integration shim a maintainer of the reference
would replace by two import lines.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types

from . import diff_ops, networks

_DIFF_NAMES = ("gradient", "divergence", "laplace", "jacobian", "hessian")


class _Sink:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Sink()

    def __getattr__(self, name):
        return _Sink()

    def __iter__(self):
        return iter((_Sink(), _Sink()))


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)

    def _missing(key):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Sink()

    mod.__getattr__ = _missing
    sys.modules[name] = mod
    return mod


def _ensure_optional_modules():
    """the reference imports these at module import time; none is used by the hot path"""
    def have(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not have("pytorch3d"):
        p3 = _stub_module("pytorch3d")
        p3.ops = _stub_module("pytorch3d.ops", knn_points=None, knn_gather=None)
    if not have("tensorboardX"):
        _stub_module("tensorboardX", SummaryWriter=_Sink)
    if not have("matplotlib"):
        mpl = _stub_module("matplotlib")
        for sub in ("pyplot", "colors"):
            setattr(mpl, sub, _stub_module(f"matplotlib.{sub}"))
        # fluid/visualize.py:31-43 does arithmetic on ``cm.bwr(field)``: a colormap stand-in must return a real RGBA
        # array (a grey ramp); the images it ends up in are plotting output, not part of the hot path
        cm = _stub_module("matplotlib.cm")

        def _colormap(name):
            if name.startswith("__"):
                raise AttributeError(name)

            def ramp(values, *a, **k):
                import numpy as np
                v = np.clip(np.asarray(values, dtype=np.float64), 0.0, 1.0)
                return np.stack([v, v, v, np.ones_like(v)], axis=-1)
            return ramp

        cm.__getattr__ = _colormap
        mpl.cm = cm
    if not have("meshio"):                        # elasticity/model.py:77: meshio.read(cfg.mesh_path) of a MEDIT .mesh
        from . import medit
        _stub_module("meshio", read=medit.read)
    if not have("open3d"):
        _stub_module("open3d")


def _shim_torch():
    """torch >= 2.7 removed ReduceLROnPlateau(verbose=) used at base/baseModel.py:61-62"""
    import inspect
    import torch

    sched = torch.optim.lr_scheduler
    if "verbose" not in inspect.signature(sched.ReduceLROnPlateau.__init__).parameters:
        base_cls = sched.ReduceLROnPlateau

        class ReduceLROnPlateau(base_cls):
            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        sched.ReduceLROnPlateau = ReduceLROnPlateau



class _TorchWithDeviceSvd:
    """stands in for the ``torch`` global of elasticity/model.py ONLY: ``torch.svd`` of the (N, D, D) deformation
    gradients (elasticity/model.py:144) goes to the one-kernel SVD (insr_svd_small; differentiable through S, which is
    all that file uses), everything else -- and every other module's ``torch.svd`` -- is torch's own."""

    def __init__(self, torch_module):
        self._t = torch_module

    def svd(self, A, *args, **kwargs):
        from . import linalg
        if linalg.supports(A) and not args and not kwargs:
            return linalg.svd(A)
        return self._t.svd(A, *args, **kwargs)

    def __getattr__(self, name):
        return getattr(self._t, name)


def _scope_device_svd():
    """bind the proxy inside elasticity.model (imported here, after ``base`` has been rebound)"""
    import torch
    try:
        ela = importlib.import_module("elasticity.model")
    except Exception:                    # a tree without the elasticity package: nothing to scope
        return
    if not isinstance(ela.torch, _TorchWithDeviceSvd):
        ela.torch = _TorchWithDeviceSvd(ela.torch if ela.torch is not None else torch)


def install(reference_root: str):
    """rebind the reference's field + operator layer to the fused implementation; returns the
    imported ``base`` package."""
    reference_root = os.path.abspath(reference_root)
    if not os.path.isfile(os.path.join(reference_root, "base", "networks.py")):
        raise FileNotFoundError(f"no INSR-PDE tree at {reference_root}")
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    _ensure_optional_modules()
    _shim_torch()
    base = importlib.import_module("base")
    ref_networks = importlib.import_module("base.networks")
    ref_model = importlib.import_module("base.baseModel")
    ref_diff = importlib.import_module("base.diff_ops")
    for holder in (ref_networks, ref_model, base):
        holder.get_network = networks.get_network
    for holder in (ref_networks, base):
        holder.MLP = networks.MLP
    for name in _DIFF_NAMES:
        fn = getattr(diff_ops, name)
        setattr(ref_diff, name, fn)
        setattr(ref_networks, name, fn)     # networks.py does ``from .diff_ops import *``
        setattr(base, name, fn)
    _scope_device_svd()
    return base


def install_fused_closures():
    """second stage, after ``install``: the reference's loss closures -- the bodies of the ``@_training_loop`` methods of
    ``Advection1DModel`` (advection/model.py:43-52, 68-91), ``Fluid2DModel`` (fluid/model.py:43-52, 72-151) and
    ``ElasticityModel._solve_deformation`` (elasticity/model.py:127-189) -- are replaced by the one-kernel closures of
    ``insr_pde_b200.fused``.  The replacements draw their samples through the model's own sampling methods in the reference's
    order (same random stream), leave their gradient in ``.grad`` and return detached loss values, so
    ``BaseModel._update_network`` (base/baseModel.py:73-81) is taught to skip ``zero_grad`` / ``backward`` for such a
    loss dict and go straight to ``optimizer.step()`` / ``scheduler.step()``.  Everything else -- ``main.py``, the training
    loop, time stepping, checkpoints, visualisation -- stays the reference's code."""
    import torch
    from . import fused
    base_model = importlib.import_module("base.baseModel")
    BaseModel = base_model.BaseModel
    if getattr(BaseModel._update_network, "_insr_fused", False):
        return
    ref_update = BaseModel._update_network

    def _update_network(self, loss_dict):
        loss = sum(loss_dict.values())
        if torch.is_tensor(loss) and loss.requires_grad:
            return ref_update(self, loss_dict)
        self.optimizer.step()                      # the closure has already accumulated d loss / d theta into .grad
        if self.scheduler is not None:
            self.scheduler.step(loss_dict['main'])

    _update_network._insr_fused = True
    BaseModel._update_network = _update_network

    def zero(self):
        fused.zero_grads(*self._trainable_networks.values())

    def loop(cls, name, fn):
        fn.__name__ = name                         # the tag of the progress bar / tensorboard scalars
        setattr(cls, name, BaseModel._training_loop(fn))

    sampling_ref = importlib.import_module("base.sampling")

    adv = importlib.import_module("advection.model").Advection1DModel

    def adv_initialize(self):
        samples = self._sample_in_training()
        ref = self.init_cond_func(samples)
        zero(self)
        return fused.advect_initialize(self.field, samples, ref)

    def adv_advect(self):
        samples = self._sample_in_training()
        boundary = sampling_ref.sample_boundary(max(self.sample_resolution // 100, 10), 1, device=self.device) * self.length / 2
        zero(self)
        return fused.advect_step(self.field, self.field_prev, samples, boundary, self.dt, self.vel)

    loop(adv, "_initialize", adv_initialize)
    loop(adv, "_advect", adv_advect)

    flu = importlib.import_module("fluid.model").Fluid2DModel

    def flu_sets(self):
        samples = self._sample_in_training()
        n_bc = samples.shape[0] // 100
        bc_x = sampling_ref.sample_boundary2D_separate(n_bc, side='horizontal', device=self.device)
        bc_y = sampling_ref.sample_boundary2D_separate(n_bc, side='vertical', device=self.device)
        return samples, bc_x, bc_y

    def flu_initialize(self):
        samples = self._sample_in_training()
        ref = self.init_cond_func(samples)
        zero(self)
        return fused.fluid_initialize(self.velocity_field, samples, ref)

    def flu_advect(self):
        sets = flu_sets(self)
        zero(self)
        return fused.fluid_advect_velocity(self.velocity_field, self.velocity_field_prev, *sets, self.cfg.dt)

    def flu_pressure(self):
        sets = flu_sets(self)
        zero(self)
        return fused.fluid_solve_pressure(self.velocity_field, self.pressure_field, *sets)

    def flu_projection(self):
        sets = flu_sets(self)
        zero(self)
        return fused.fluid_projection(self.velocity_field, self.velocity_field_prev, self.pressure_field, *sets)

    loop(flu, "_initialize", flu_initialize)
    loop(flu, "_advect_velocity", flu_advect)
    loop(flu, "_solve_pressure", flu_pressure)
    loop(flu, "_projection", flu_projection)

    ela = importlib.import_module("elasticity.model").ElasticityModel

    def ela_solve(self):
        samples = self._sample_in_training(self.sample_resolution)
        fixed, fixed_right = self._sample_fixed_in_training(self.sample_resolution)
        zero(self)
        return fused.elasticity_solve_deformation(
            self.deformation_field, self.deformation_field_prev, self.deformation_field_prev_prev, samples, fixed, fixed_right,
            dt=self.dt, timestep=self.timestep, energy=self.energy, ratio_arap=self.ratio_arap, ratio_volume=self.ratio_volume,
            ratio_kinematics=self.ratio_kinematics, ratio_constraint=self.ratio_constraint, ratio_collide=self.ratio_collide,
            external_force=self.external_force, external_force_timesteps=self.external_force_timesteps,
            constraint_offset_right=self.constraint_offset_right, plane_height=self.plane_height,
            circle_center=self.circle_center, circle_radius=self.circle_radius)

    loop(ela, "_solve_deformation", ela_solve)


def run_main(argv, reference_root: str, fused_closures: bool = None):
    """python main.py <argv>  of the reference, unchanged, on the fused kernels.  ``fused_closures`` (default: the
    environment variable INSR_FUSED_CLOSURES=1) additionally swaps the loss closures for the one-kernel ones."""
    install(reference_root)
    if fused_closures is None:
        fused_closures = os.environ.get("INSR_FUSED_CLOSURES", "0") == "1"
    if fused_closures:
        install_fused_closures()
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [os.path.join(reference_root, "main.py"), *argv]
    os.chdir(reference_root)     # config.py:55-57 copies *.py relative to cwd; mesh paths are relative
    try:
        runpy.run_path(sys.argv[0], run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)


if __name__ == "__main__":
    root = os.environ.get("INSR_REFERENCE_ROOT")
    if not root:
        raise SystemExit("set INSR_REFERENCE_ROOT to the INSR-PDE checkout; usage: "
                         "python -m insr_pde_b200.patch fluid --tag ... (main.py arguments)")
    run_main(sys.argv[1:], root)
