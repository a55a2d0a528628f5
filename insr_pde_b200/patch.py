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

    flu_mod = importlib.import_module("fluid.model")
    flu = flu_mod.Fluid2DModel

    def flu_sets(self):
        samples = self._sample_in_training()
        n_bc = samples.shape[0] // 100
        # through the name fluid/model.py itself uses (the data-parallel twin rebinds it there)
        bc_x = flu_mod.sample_boundary2D_separate(n_bc, side='horizontal', device=self.device)
        bc_y = flu_mod.sample_boundary2D_separate(n_bc, side='vertical', device=self.device)
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


def install_graphed_loops(data_parallel: bool = False):
    """third stage, after ``install`` (implies ``install_fused_closures``): the BODY of the reference's training loop
    (base/baseModel.py:104-134: closure -> _update_network -> .item() logging, once per iteration from Python) is replaced
    by ``fused.GraphedLoop`` -- the whole iteration (the model's own sampling, the one-kernel closures, Adam of every
    trainable net, ReduceLROnPlateau, loss log; with ``data_parallel`` also the ONE all-reduce of [gradients | losses]) is
    captured once per (model, loop) as a CUDA graph and replayed ``max_n_iters`` times; the loss history is read back in
    bulk and handed to tensorboard, early stop (lr <= 1.1e-8, :132-134) is tested every 100 iterations.  main.py, time
    stepping, checkpoints and output writing stay the reference's code."""
    import torch
    from . import fused
    install_fused_closures()
    base_model = importlib.import_module("base.baseModel")
    BaseModel = base_model.BaseModel
    if getattr(BaseModel, "_insr_graphed", False):
        return
    BaseModel._insr_graphed = True

    def graphed(cls, name):
        method = getattr(cls, name)
        body = next(c.cell_contents for c in method.__closure__ if callable(c.cell_contents))     # the fused closure
        widths_ok = getattr(body, "_insr_graph_ok", True)

        def loop(self):
            nets = list(self._trainable_networks.values())
            off = self.__dict__.setdefault("_insr_graph_off", set())
            if name in off or not widths_ok or not all(p.is_cuda for n in nets for p in n.parameters()):
                return method(self)                              # CPU tensors: the reference's own loop
            loops = self.__dict__.setdefault("_insr_loops", {})
            # host-side branches of a closure are frozen into its captured graph: one graph per branch.  The elasticity
            # closure applies the external force only while timestep <= external_force_timesteps (elasticity/model.py:151-155)
            forced = bool("external" in getattr(self, "energy", ()) and getattr(self, "timestep", 0) <= getattr(self, "external_force_timesteps", -1))
            key = (name, forced) if hasattr(self, "external_force_timesteps") else name
            closure, presample = (lambda: body(self)), None
            batch = _elastic_batch(self) if (name == "_solve_deformation" and not data_parallel) else None
            if batch is not None:
                # the stepper's persistent batch (fused.ElasticityBatch) under the reference's own model: the constant rows
                # (uniform grid / mesh vertices) are written once, the random rows drawn in place by the Philox samplers one
                # iteration ahead, the previous-frame fields evaluated on the constant rows once per time step -- instead of
                # the model's ~40 torch sampling kernels and two full frozen-field evaluations per iteration
                batch.refresh(self.deformation_field_prev, self.deformation_field_prev_prev)
                closure, presample = (lambda: _elastic_closure(self, batch)), batch.draw
            if key not in loops:
                loops[key] = fused.GraphedLoop(nets, self.cfg.lr, closure, capacity=max(int(self.max_n_iters), 1),
                                               data_parallel=data_parallel, presample=presample)
            else:
                loops[key].reset(self.cfg.lr)
            try:
                hist = loops[key].run(int(self.max_n_iters), early_stop=bool(self.cfg.early_stop))
            except RuntimeError as e:
                if loops[key].graph is not None and loops[key].graph_ready or "capture" not in str(e).lower() or data_parallel:
                    raise
                # the closure does something a CUDA graph cannot record (host-side sampling with a pageable copy, a
                # synchronising call): this loop stays on the eager fused closures (one iteration has been taken already)
                sys.stderr.write(f"insr_pde_b200: {type(self).__name__}.{name} cannot be captured as a CUDA graph ({str(e)[:120]}); "
                                 "running it eagerly on the fused closures\n")
                loops.pop(key).graph = None
                off.add(name)
                torch.cuda.synchronize()
                return method(self)
            self.train_step = len(hist)
            for i, values in enumerate(hist):                   # the .item() logging of :116-118, in bulk
                self.tb.add_scalars(name, values, global_step=i)
            if hasattr(self, f"_vis{name}"):
                getattr(self, f"_vis{name}")()
            self._insr_last_hist = hist
        loop.__name__ = name
        setattr(cls, name, loop)

    class _StepperView:
        """what fused.ElasticityBatch / ElasticityStepper._batch read from a stepper, taken from the reference's model"""

        def __init__(self, m):
            self.dim, self.sr, self.pattern = m.dim, m.sample_resolution, tuple(m.sample_pattern)
            self.mesh = (m.mesh_V, m.mesh_F) if m.use_mesh else None
            self.seed = int(torch.initial_seed()) & 0x7FFFFFFF
            self.kw = {"energy": list(m.energy)}
            self._dev = next(m.deformation_field.parameters()).device

        def _device(self):
            return self._dev

    def _elastic_batch(m):
        if "_insr_view" not in m.__dict__:
            try:
                m._insr_view = _StepperView(m)
            except Exception:                                    # an attribute this view does not know: the model's own sampling
                m._insr_view = None
        view = m._insr_view
        if view is None or not all(k in ("random", "uniform") for k in view.pattern):
            return None
        return fused.ElasticityStepper._batch(view)

    def _elastic_closure(m, batch):
        return fused.elasticity_solve_deformation(
            m.deformation_field, m.deformation_field_prev, m.deformation_field_prev_prev, None, None, None,
            dt=m.dt, timestep=m.timestep, energy=m.energy, ratio_arap=m.ratio_arap, ratio_volume=m.ratio_volume,
            ratio_kinematics=m.ratio_kinematics, ratio_constraint=m.ratio_constraint, ratio_collide=m.ratio_collide,
            external_force=m.external_force, external_force_timesteps=m.external_force_timesteps,
            constraint_offset_right=m.constraint_offset_right, plane_height=m.plane_height,
            circle_center=m.circle_center, circle_radius=m.circle_radius, batch=batch)

    adv = importlib.import_module("advection.model").Advection1DModel
    flu = importlib.import_module("fluid.model").Fluid2DModel
    ela_mod = importlib.import_module("elasticity.model")
    ela = ela_mod.ElasticityModel
    # elasticity/sampling.py:4-9 draws the tetrahedron samples with numpy on the host and copies them (not capturable, and a
    # host round trip per iteration): under the graphed loop the mesh is sampled by insr_sample_mesh on the device -- the
    # same distribution from the library's Philox stream (keyed by torch's seed), one sampler per (V, F)
    from .sampling import MeshSampler
    samplers = {}

    def sample_mesh(V, F, N, distrib=None):
        if not V.is_cuda:
            return ela_mod._insr_sample_mesh_ref(V, F, N, distrib)
        key = (V.data_ptr(), F.data_ptr(), tuple(V.shape), tuple(F.shape))
        if key not in samplers:
            samplers[key] = MeshSampler(V, F, dim_out=3, seed=int(torch.initial_seed()) & 0x7FFFFFFF)
        return samplers[key].sample(int(N))
    if not hasattr(ela_mod, "_insr_sample_mesh_ref"):
        ela_mod._insr_sample_mesh_ref = ela_mod.sample_mesh
        ela_mod.sample_mesh = sample_mesh
    for cls, names in ((adv, ("_initialize", "_advect")), (flu, ("_initialize", "_advect_velocity", "_solve_pressure", "_projection")),
                       (ela, ("_solve_deformation",))):
        for name in names:
            graphed(cls, name)


_DP_MODELS = []          # lists of the models the data-parallel twin has seen (their captured graphs must go before NCCL does)


def close_graphed_loops():
    """drop every captured iteration graph of the models run so far (``GraphedLoop.close``): a process group must not be
    destroyed -- and a process must not exit -- while a captured graph still references its communicator"""
    for models in _DP_MODELS:
        for m in models:
            for lp in m.__dict__.get("_insr_loops", {}).values():
                lp.close()
            m.__dict__.pop("_insr_loops", None)
        del models[:]


def install_data_parallel(pde: str):
    """one process per GPU (torchrun): every rank draws the SAME global sample sets (same seed, same generator state) and
    keeps its contiguous shard of the interior points; gradients -- and the scheduler's loss -- are
    all-reduced before every optimizer step (``dist.install_global``).  advection / fluid losses are means over points
    (average over the ranks); elasticity's are sums over points: every point set is sharded and the reduction is a sum.
    Returns (rank, world)."""
    from . import dist as idist
    rank, world = idist.init_from_env()
    if world == 1:
        return rank, world
    module = {"advection": "advection.model", "fluid": "fluid.model", "elasticity": "elasticity.model"}[pde]
    cls_name = {"advection": "Advection1DModel", "fluid": "Fluid2DModel", "elasticity": "ElasticityModel"}[pde]
    cls = getattr(importlib.import_module(module), cls_name)
    trainable = []

    def shard_method(name):
        orig = getattr(cls, name)

        def sharded(self, *a, **k):
            if self not in trainable:
                trainable.append(self)
            out = orig(self, *a, **k)
            if isinstance(out, tuple):
                return tuple(idist.shard_points(o) if hasattr(o, "shape") else o for o in out)
            return idist.shard_points(out)
        sharded.__name__ = name
        setattr(cls, name, sharded)

    shard_method("_sample_in_training")
    if pde == "elasticity":
        shard_method("_sample_fixed_in_training")
    if pde == "fluid":
        # fluid/model.py:94-95 etc. size the boundary sets from the (now sharded) interior batch: samples.shape[0] // 100.
        # Scale the request back so that every rank draws the single-GPU boundary sets (same count, same random stream);
        # they are evaluated redundantly on every rank, which the averaging all-reduce leaves unchanged.
        mod = importlib.import_module(module)
        orig_bc = mod.sample_boundary2D_separate

        def sample_boundary2D_separate(N, *a, **k):
            return orig_bc(N * world, *a, **k)
        mod.sample_boundary2D_separate = sample_boundary2D_separate
    idist.install_global(lambda: [n for m in trainable for n in m._trainable_networks.values()], average=(pde != "elasticity"))
    _DP_MODELS.append(trainable)
    return rank, world


def _silence_non_zero_rank(pde: str):
    """rank > 0 of a data-parallel run: no frame output, no checkpoints (rank 0 writes them; the replicas are identical)"""
    module = {"advection": "advection.model", "fluid": "fluid.model", "elasticity": "elasticity.model"}[pde]
    cls_name = {"advection": "Advection1DModel", "fluid": "Fluid2DModel", "elasticity": "ElasticityModel"}[pde]
    cls = getattr(importlib.import_module(module), cls_name)
    cls.write_output = lambda self, folder: None
    cls.save_ckpt = lambda self, *a, **k: None


def run_main(argv, reference_root: str, fused_closures: bool = None, graphed: bool = False, data_parallel: bool = False,
             seed: int = None):
    """python main.py <argv>  of the reference, unchanged, on the fused kernels.  ``fused_closures`` (default: the
    environment variable INSR_FUSED_CLOSURES=1) additionally swaps the loss closures for the one-kernel ones,
    ``graphed`` the loop body for the CUDA-graphed iteration; ``data_parallel`` is the torchrun twin of main.py
    (one process per GPU, RANK / LOCAL_RANK / WORLD_SIZE from the environment)."""
    argv = list(argv)
    rank, world = 0, 1
    import time
    t_start = time.perf_counter()

    def note(msg):
        if os.environ.get("INSR_PATCH_VERBOSE", "0") == "1":
            sys.stderr.write(f"[insr-dp rank {os.environ.get('RANK', '0')}] {time.perf_counter() - t_start:7.1f} s  {msg}\n")
            sys.stderr.flush()
    if data_parallel:
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        local = os.environ.get("LOCAL_RANK", "0")
        # the reference hard-codes cuda:0 (base/baseModel.py:25) and config.py:28-29 exports CUDA_VISIBLE_DEVICES from -g:
        # give every rank its own GPU as device 0, before CUDA is initialised
        os.environ["CUDA_VISIBLE_DEVICES"] = local
        argv += ["-g", local]
        os.environ["LOCAL_RANK"] = "0"
        if rank > 0 and "--proj_dir" in argv:                    # config.py creates / wipes the experiment directory
            k = argv.index("--proj_dir") + 1
            argv[k] = os.path.join(argv[k], f".rank{rank}")
        elif rank > 0:
            argv += ["--proj_dir", os.path.join("checkpoints", f".rank{rank}")]
    install(reference_root)
    if fused_closures is None:
        fused_closures = os.environ.get("INSR_FUSED_CLOSURES", "0") == "1"
    pde = next((a for a in argv if a in ("advection", "fluid", "elasticity")), None)
    note("field / operator layer installed")
    if data_parallel and pde:
        install_data_parallel(pde)
        note("process group up")
        if rank > 0:
            _silence_non_zero_rank(pde)
    if graphed:
        install_graphed_loops(data_parallel=data_parallel and world > 1)
    elif fused_closures:
        install_fused_closures()
    if seed is not None:
        import numpy as np
        import torch
        torch.manual_seed(seed)
        np.random.seed(seed)
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [os.path.join(reference_root, "main.py"), *argv]
    os.chdir(reference_root)     # config.py:55-57 copies *.py relative to cwd; mesh paths are relative
    try:
        note("starting main.py")
        runpy.run_path(sys.argv[0], run_name="__main__")
        note("main.py done")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    if data_parallel and world > 1:
        # orderly teardown: captured iteration graphs hold NCCL kernels -- drop them before the communicator goes
        import gc
        import torch
        import torch.distributed as tdist
        close_graphed_loops()
        gc.collect()
        torch.cuda.synchronize()
        if tdist.is_initialized():
            tdist.barrier()
            tdist.destroy_process_group()
        note("process group destroyed")


if __name__ == "__main__":
    root = os.environ.get("INSR_REFERENCE_ROOT")
    if not root:
        raise SystemExit("set INSR_REFERENCE_ROOT to the INSR-PDE checkout; usage: "
                         "python -m insr_pde_b200.patch [--insr-closures] [--insr-graphed] [--insr-dp] [--insr-seed S] "
                         "fluid --tag ... (main.py arguments);  data-parallel: torchrun --nproc-per-node G -m insr_pde_b200.patch --insr-dp ...")
    args = sys.argv[1:]
    flags = {"--insr-closures": False, "--insr-graphed": False, "--insr-dp": False}
    for f in list(flags):
        if f in args:
            flags[f] = True
            args.remove(f)
    seed = None
    if "--insr-seed" in args:
        k = args.index("--insr-seed")
        seed = int(args[k + 1])
        del args[k:k + 2]
    run_main(args, root, fused_closures=flags["--insr-closures"] or None, graphed=flags["--insr-graphed"],
             data_parallel=flags["--insr-dp"], seed=seed)
