#!/usr/bin/env python
"""bench.py -- headline benchmark of the INSR-PDE hot path on B200.

Metric (BASELINE.json): collocation points / s for SIREN forward + gradient + Laplacian +
backward.  A "step" is one pass of the hot path over one batch of synthetic collocation
points: evaluate the field with its Jacobian and Laplacian streams (insr_siren_forward,
order=LAP), then the reverse sweep for given output cotangents (insr_siren_backward) producing
the flat parameter gradient; with N > 1 ranks the points are sharded (weights replicated) and
the per-step parameter gradient is all-reduced with NCCL.

Workload (config.workload): the fluid2Dtlgn pressure network of scripts/fluid2Dtlgn.sh
(2 -> 1, hidden 32, 3 hidden layers = the `_solve_pressure` closure's operator,
fluid/model.py:103-125) on a FIXED GLOBAL batch of `--points` synthetic points per step (default
2^24, split over the ranks: strong scaling -- the north star's ">= 7x at 8 GPUs"; the per-GPU
buffers exceed L2 at every N <= 8); "weak" in the line is the same step at 2^22 points per GPU,
"sweep" the other BASELINE.json shapes on their own fixed global batches, "script_size" the script's
own 128^2-point batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (D, O, H, L, order, description)
    "fluid2Dtlgn.pressure": (2, 1, 32, 3, 2, "fluid2Dtlgn pressure net 2->1 H=32 L=3: y+J+Laplacian fwd + bwd"),
    "fluid2Dtlgn.velocity": (2, 2, 32, 3, 1, "fluid2Dtlgn velocity net 2->2 H=32 L=3: y+J fwd + bwd"),
    "advect1D": (1, 1, 20, 2, 1, "advect1D field 1->1 H=20 L=2: y+J fwd + bwd"),
    "elasticity2Dstretch": (2, 2, 68, 3, 1, "elasticity2Dstretch deformation 2->2 H=68 L=3: y+J fwd + bwd"),
    "elasticity3Dbunny": (3, 3, 66, 3, 1, "elasticity3Dbunny deformation 3->3 H=66 L=3: y+J fwd + bwd"),
    "sweep.h64": (2, 1, 64, 3, 2, "synthetic sweep 2->1 H=64 L=3: y+J+Laplacian fwd + bwd"),
    "sweep.h128": (2, 1, 128, 3, 2, "synthetic sweep 2->1 H=128 L=3: y+J+Laplacian fwd + bwd"),
    "sweep.h256": (2, 1, 256, 3, 2, "synthetic sweep 2->1 H=256 L=3: y+J+Laplacian fwd + bwd"),
    "sweep.h512": (2, 1, 512, 5, 2, "synthetic sweep 2->1 H=512 L=5: y+J+Laplacian fwd + bwd"),
    "sweep.h512.l3": (2, 1, 512, 3, 2, "synthetic sweep 2->1 H=512 L=3: y+J+Laplacian fwd + bwd"),
    # further points of SURVEY.md 8d's sweep (depth 4 / 5, the 3 -> 3 value + Jacobian mode)
    "sweep.h128.l5": (2, 1, 128, 5, 2, "synthetic sweep 2->1 H=128 L=5: y+J+Laplacian fwd + bwd"),
    "sweep.h256.l4": (2, 1, 256, 4, 2, "synthetic sweep 2->1 H=256 L=4: y+J+Laplacian fwd + bwd"),
    "sweep.3d.h128": (3, 3, 128, 3, 1, "synthetic sweep 3->3 H=128 L=3: y+J fwd + bwd (elasticity3Dlucy's network)"),
    "sweep.3d.h256": (3, 3, 256, 3, 1, "synthetic sweep 3->3 H=256 L=3: y+J fwd + bwd"),
}


_REAL_STDOUT = None


def keep_stdout_clean():
    """ONE JSON line on stdout is the contract: park the real stdout and point fd 1 at stderr, so that library chatter
    (the NCCL version banner, warnings printed by child threads) cannot land in front of the line"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def flops_fwd_per_point(D, O, H, L, order):
    """MAC(S) = D*H + S*(L*H^2 + H*O); F_fwd = 2*MAC; F_fwd+bwd = 3*F_fwd"""
    S = 1 + (D if order >= 1 else 0) + (1 if order == 2 else (D * (D + 1) // 2 if order == 3 else 0))
    return 2 * (D * H + S * (L * H * H + H * O))


def sm_count():
    """cudaDeviceProp.multiProcessorCount of the current device (148 on the B200)"""
    try:
        return int(torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count)
    except Exception:
        return 148


def fp32_peak_tflops():
    sm_mhz, how = 1965.0, "fallback sm_max 1965 MHz"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        sm_mhz, how = float(peaks["sm_max_mhz"]), "MEASURED_PEAKS.json sm_max_mhz"
    except Exception:
        pass
    return sm_count() * 128 * 2 * sm_mhz * 1e6 / 1e12, f"{sm_count()} SM (cudaDeviceProp) x 128 lanes x 2 x {how}"


def bf16_peak_tflops():
    """measured dense bf16 throughput: the sustained figure (a kernel timed inside a long step), else the recipe's fallback"""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained"
    except Exception:
        return 1384.0, "fallback 1384 TFLOP/s"


def source_sha16(*rel):
    """hash of kernel sources: committed ncu captures carry it, so a capture of an older kernel is never quoted"""
    import hashlib
    h = hashlib.sha256()
    for r in rel:
        h.update(open(os.path.join(ROOT, r), "rb").read())
    return h.hexdigest()[:16]


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p, self.idx = None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # "under load": the upper half of the samples (idle samples before/after the region are lower)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(smax) if smax else None),
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_theta(net_shape, device, seed=0):
    """reference init (sine_init / first_layer_sine_init + default bias init) under manual_seed"""
    import insr_pde_b200 as ib
    D, O, H, L = net_shape
    torch.manual_seed(seed)
    net = ib.MLP(D, O, L, H, nonlinearity="sine").to(device)
    return net


def fluid_timestep_ours(dev, iters, world=1, sample_resolution=128):
    """seconds per PDE time step of fluid2Dtlgn (fluid/model.py:61-70: advect -> pressure -> projection,
    `iters` Adam iterations per loop, sample_resolution^2 points (the script: 128^2) + 2 x sr^2/100 boundary points per
    iteration -- split over the ranks --, early stop off) on the fused closures (insr_pde_b200.fused)."""
    import insr_pde_b200 as ib
    from insr_pde_b200 import dist as idist, fused
    torch.manual_seed(0)
    vel, prev, pres = (ib.MLP(2, o, 3, 32, nonlinearity="sine").to(dev) for o in (2, 2, 1))
    # world > 1: the iteration graph contains the ONE all-reduce of [gradients | loss values] (fused.SharedGradBuffer);
    # INSR_GRAPH_NCCL=0 falls back to the eager loop (torch Adam, all-reduce and a host sync per iteration)
    graph_dp = world == 1 or os.environ.get("INSR_GRAPH_NCCL", "1") != "0"
    factory = (lambda nets: idist.GradAllReducer(nets)) if (world > 1 and not graph_dp) else None
    st = fused.FluidStepper(vel, prev, pres, dt=0.05, sample_resolution=sample_resolution, lr=1e-4, reducer_factory=factory,
                            graphed=graph_dp, device_sampler=True)
    st.data_parallel = world > 1 and graph_dp
    st.initialize(fused.taylorgreen_velocity, 20, world)
    st.step(3, world)                                   # warm-up
    sec = float("inf")
    for _ in range(2):                                  # best of two time steps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h1, h2, h3 = st.step(iters, world)
        torch.cuda.synchronize()
        sec = min(sec, time.perf_counter() - t0)
    peer_exchange = world > 1 and graph_dp and all(lp.shared is not None and lp.shared.peer is not None
                                                   for lp in st.__dict__.get("_loops", {}).values())
    st.close()                                          # release the captured graphs (and the exchange recorded in them) now
    return {"sec_per_timestep": round(sec, 4), "iters_per_loop": iters, "loops": 3,
            "us_per_iteration": round(sec / (3 * iters) * 1e6, 1), "points_per_iteration": sample_resolution ** 2,
            "final_losses": [round(h[-1]["main"], 8) for h in (h1, h2, h3)],
            "note": ("fluid2Dtlgn step on fused lsq closures; one CUDA graph per iteration (one-kernel Philox sampling of the "
                     "three point sets, closures with interior / boundary terms on parallel branches, device Adam, device "
                     "ReduceLROnPlateau), losses read back in bulk" + (("; points sharded over the ranks, the exchange of [gradients | loss values] "
                     "fused into the update kernel over NVLink peer memory (insr_iteration_update_peer; no NCCL in the graph)" if peer_exchange else
                     "; points sharded over the ranks, one NCCL all-reduce of [gradients | loss values] per iteration inside the graph")
                     if world > 1 else "")) if graph_dp else
                    "fluid2Dtlgn step on fused lsq closures + torch Adam + flat-gradient all-reduce, host sync per iteration"}


ELASTIC_CASES = {
    # scripts/elasticity2Dstretch.sh: 100^2 uniform + 100^2 random interior points, clamped faces, H = 68
    "elasticity2Dstretch": dict(dim=2, H=68, sr=100, dt=0.05, energy=["arap", "constraint", "constraint_right", "volume"],
                                ratio_volume=1e3, ratio_arap=1e0, ratio_constraint=1e4, ratio_kinematics=1e0, ratio_collide=1e0,
                                ext=[0., 0., 0.], ext_T=5, off=[2.0, 0., 0.], plane=-2.0, center=[0., -2., 0.], radius=1.0),
    # scripts/elasticity3Dbunny.sh on the real mesh (elasticity/data/bunny.mesh: 18 592 vertices, 76 854 tetrahedra):
    # 20^3 volume samples + every vertex = 26 592 points per iteration, H = 66
    "elasticity3Dbunny": dict(dim=3, H=66, sr=20, dt=0.1, energy=["arap", "kinematics", "collision", "external", "volume"],
                              ratio_volume=1e3, ratio_arap=1e2, ratio_constraint=1e3, ratio_kinematics=1e0, ratio_collide=1e6,
                              ext=[0., 0., -1e2], ext_T=5, off=[1.0, 0., 0.], plane=-2.0, center=[0., -2., 0.], radius=1.0,
                              mesh="bunny.mesh"),
}


def find_mesh(name):
    """elasticity/data/<name> of the reference tree (data, not source): /root/reference here, the shipped copy on the GPU box"""
    for root in (os.environ.get("INSR_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if root and os.path.isfile(os.path.join(root, "elasticity", "data", name)):
            return os.path.join(root, "elasticity", "data", name)
    return None


def elasticity_timestep_ours(dev, iters, case="elasticity2Dstretch"):
    """seconds per elasticity time step (elasticity/model.py:119-189: one @_training_loop of `iters` Adam iterations) on
    ElasticityStepper: the whole iteration -- Philox sampling, order-1 field kernels with the tape kept, the frozen
    previous frames on side streams, insr_elastic_terms, reverse sweep, device Adam / plateau -- as one CUDA graph."""
    import insr_pde_b200 as ib
    from insr_pde_b200 import fused
    c = ELASTIC_CASES[case]
    dim = c["dim"]
    kw = dict(energy=c["energy"], ratio_arap=c["ratio_arap"], ratio_volume=c["ratio_volume"], ratio_kinematics=c["ratio_kinematics"],
              ratio_constraint=c["ratio_constraint"], ratio_collide=c["ratio_collide"],
              external_force=torch.tensor(c["ext"][:dim], device=dev), external_force_timesteps=c["ext_T"],
              constraint_offset_right=torch.tensor(c["off"][:dim], device=dev), plane_height=c["plane"],
              circle_center=torch.tensor(c["center"][:dim], device=dev), circle_radius=c["radius"])
    torch.manual_seed(0)
    nets = [ib.MLP(dim, dim, 3, c["H"], nonlinearity="sine").to(dev) for _ in range(3)]
    mesh, n_vert = None, 0
    if c.get("mesh"):
        from insr_pde_b200 import medit
        path = find_mesh(c["mesh"])
        if path is None:
            return {"case": case, "error": f"{c['mesh']} not found (neither /root/reference nor oracle/_ref)"}
        mesh = medit.load_normalized(path, dim, device=dev)
        n_vert = int(mesh[0].shape[0])
    st = fused.ElasticityStepper(*nets, dim, dt=c["dt"], sample_resolution=c["sr"], graphed=True, mesh=mesh, **kw)
    st.initialize(5)
    st.step(5)                                          # warm-up: the graph is captured here
    sec = float("inf")
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = st.step(iters)
        torch.cuda.synchronize()
        sec = min(sec, time.perf_counter() - t0)
    npts = c["sr"] ** dim + (n_vert if mesh is not None else c["sr"] ** dim)
    st.close()
    return {"case": case, "sec_per_timestep": round(sec, 4), "iters_per_loop": iters, "us_per_iteration": round(sec / iters * 1e6, 1),
            "points_per_iteration": npts, "points_per_s": round(npts * iters / sec, 1), "final_loss": round(h[-1]["main"], 6)}


def advection_timestep_ours(dev, iters):
    """seconds per advect1D time step (advection/model.py:62-91: one loop of `iters` Adam iterations, 5000 points + 50
    boundary points) on AdvectionStepper, one CUDA graph per iteration"""
    import insr_pde_b200 as ib
    from insr_pde_b200 import fused
    torch.manual_seed(0)
    field, prev = (ib.MLP(1, 1, 2, 20, nonlinearity="sine").to(dev) for _ in range(2))
    st = fused.AdvectionStepper(field, prev, dt=0.05, vel=0.25, length=4.0, sample_resolution=5000, lr=1e-4, graphed=True)
    st.initialize(fused.gaussian_like, 10)
    st.step(5)
    sec = float("inf")
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = st.step(iters)
        torch.cuda.synchronize()
        sec = min(sec, time.perf_counter() - t0)
    st.close()
    return {"case": "advect1D", "sec_per_timestep": round(sec, 4), "iters_per_loop": iters, "us_per_iteration": round(sec / iters * 1e6, 1),
            "points_per_iteration": 5000, "final_loss": round(h[-1]["main"], 8)}


# ------------------------------------------------------------------------------------------------
# reference legs (run in a subprocess: ref_loader's CPU shims must not live in the process that drives the GPU kernels)
# ------------------------------------------------------------------------------------------------
def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def reference_operator(workload, n_points, steps, warmup, device="cpu", budget_s=120.0):
    """the REAL reference (base/networks.py MLP + base/diff_ops.py jacobian / laplace + loss.backward(), imported from
    /root/reference or the shipped copy oracle/_ref) on the workload's operator: y (+ J (+ Laplacian)) and the reverse
    sweep for given output cotangents; `steps` timed steps of `n_points` points after `warmup` untimed ones"""
    from oracle import ref_loader
    D, O, H, L, order, _ = WORKLOADS[workload]
    ns = ref_loader.load(cpu=(device == "cpu"))
    torch.manual_seed(0)
    on_gpu = str(device).startswith("cuda")
    net = ns.base.MLP(D, O, L, H, nonlinearity="sine").to(device)
    x = (torch.rand(n_points, D, device=device) * 2 - 1).requires_grad_(True)
    gy = torch.randn(n_points, O, device=device) / n_points
    gj = torch.randn(n_points, O, D, device=device) / n_points
    gl = torch.randn(n_points, 1, device=device) / n_points

    def step():
        net.zero_grad()
        y = net(x)
        loss = (gy * y).sum()
        if order >= 1:
            jac, _ = ns.base.jacobian(y, x)
            loss = loss + (gj * jac).sum()
        if order >= 2:
            loss = loss + (gl * ns.base.laplace(y, x)).sum()
        loss.backward()

    for _ in range(max(warmup, 1)):
        step()
    if on_gpu:
        torch.cuda.synchronize()
    t0, n = time.perf_counter(), 0
    while n < steps:
        step()
        n += 1
        if on_gpu:
            torch.cuda.synchronize()
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    where = f"stock PyTorch eager on {torch.cuda.get_device_name()}" if on_gpu else f"CPU ({cpu_model_name()})"
    return {"value": round(n_points * n / el, 1), "unit": "points/s", "cores": torch.get_num_threads(), "kind": "reference",
            "cpu_model": cpu_model_name(),
            "sample": f"{n} timed steps of {n_points} points each (a bounded sample of the step's global batch) on the {workload} "
                      f"operator; the reference's own base.networks.MLP + base.diff_ops + loss.backward() from {ref_loader.REF_ROOT}, "
                      f"torch {torch.__version__}, {where}, fp32",
            "ms_per_step": round(el / n * 1e3, 3), "points_per_step": n_points}


def reference_fluid_timestep(device="cpu", iters_measured=3, iters_per_loop=100):
    """seconds per fluid2Dtlgn time step of the REAL reference: Fluid2DModel(cfg).step() (fluid/model.py:61-70, its own
    training loop, Adam, scheduler, sampling) with max_n_iters = iters_measured, scaled to iters_per_loop"""
    from oracle import ref_loader
    ns = ref_loader.load(cpu=(device == "cpu"))
    torch.manual_seed(0)
    cfg = ref_loader.make_cfg("fluid", sample_resolution=128, max_n_iters=iters_measured)
    model = ns.fluid.Fluid2DModel(cfg)
    model.initialize()
    model.step()                                        # warm-up
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.step()
    if device != "cpu":
        torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    return {"sec_per_timestep": round(sec / iters_measured * iters_per_loop, 3), "iters_per_loop": iters_per_loop,
            "ms_per_iteration": round(sec / (3 * iters_measured) * 1e3, 2), "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"Fluid2DModel.step() with max_n_iters = {iters_measured} (3 loops, 128^2 + 2 x 162 points), scaled to {iters_per_loop}; "
                      "includes the per-step checkpoint write"}


def reference_elasticity_iteration(device="cpu", case="elasticity2Dstretch", iters=5):
    """seconds per _solve_deformation iteration of the REAL reference ElasticityModel (elasticity/model.py:127-189 under
    base/baseModel.py:104-134) at the script's sizes; the bunny case reads the real mesh"""
    from oracle import ref_loader
    ns = ref_loader.load(cpu=(device == "cpu"))
    torch.manual_seed(0)
    c = ELASTIC_CASES[case]
    kw = dict(dim=c["dim"], hidden_features=c["H"], sample_resolution=c["sr"], dt=c["dt"], energy=c["energy"],
              ratio_arap=c["ratio_arap"], ratio_volume=c["ratio_volume"], ratio_kinematics=c["ratio_kinematics"],
              ratio_constraint=c["ratio_constraint"], ratio_collide=c["ratio_collide"], external_force_timesteps=c["ext_T"],
              external_force_x=c["ext"][0], external_force_y=c["ext"][1], external_force_z=c["ext"][2],
              constraint_right_offset_x=c["off"][0], plane_height=c["plane"], max_n_iters=iters, vis_resolution=50)
    if c.get("mesh"):
        kw.update(use_mesh=True, mesh_path=os.path.join(ref_loader.REF_ROOT, "elasticity", "data", c["mesh"]))
    model = ns.elasticity.ElasticityModel(ref_loader.make_cfg("elasticity", **kw))
    model.timestep = 1
    model._create_tb("bench")
    model._solve_deformation()                          # warm-up loop
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    model._solve_deformation()
    if device != "cpu":
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


def run_reference(args):
    """the reference arm: the reference's own implementation of the path (kind "reference": the tree itself, from
    /root/reference here or from the shipped copy oracle/_ref on the GPU box) on the box's host cores with every thread
    this process may use -- or, --device cuda (internal, the like-for-like leg), as stock PyTorch on the GPU"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every core this process may run on explicitly
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    D, O, H, L, order, desc_txt = WORKLOADS[args.workload]
    res = reference_operator(args.workload, args.ref_points, args.steps, args.warmup, device=args.device,
                             budget_s=args.ref_budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_block(args, world),
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.ref_legs == "all":                          # the in-line legs of our arm: time step and elasticity iteration too
        try:
            line["timestep"] = reference_fluid_timestep(args.device, 3, args.timestep_iters or 100)
        except Exception as e:
            line["timestep"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        el = {}
        for case in ELASTIC_CASES:
            try:
                el[case] = round(reference_elasticity_iteration(args.device, case, 3 if args.device == "cpu" else 10), 5)
            except Exception as e:
                el[case] = f"{type(e).__name__}: {str(e)[:200]}"
        line["elasticity_sec_per_iteration"] = el
    emit(line)


def reference_subprocess(args, device, legs="all", steps=8, warmup=2, budget=14.0, points=None):
    """run a reference leg in its own process (all host threads) and return its JSON line"""
    cmd = [sys.executable, "-W", "ignore", os.path.abspath(__file__), "--impl", "reference", "--device", device, "--workload",
           args.workload, "--steps", str(steps), "--warmup", str(warmup), "--ref-budget", str(budget), "--ref-legs", legs,
           "--ref-points", str(points or args.ref_points), "--points", str(args.points), "--scaling", args.scaling,
           "--timestep-iters", str(args.timestep_iters)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
        lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"error": (res.stderr or "no output")[-300:]}
        return json.loads(lines[-1])
    except Exception as e:
        return {"error": f"{type(e).__name__}: {str(e)[:200]}"}


METRIC = "collocation points/s (SIREN fwd + grad + Laplacian + bwd)"


def config_block(args, world):
    """the same for both arms: what one step is"""
    D, O, H, L, order, desc_txt = WORKLOADS[args.workload]
    n_local = args.points // world if args.scaling == "strong" else args.points
    return {"workload": f"{args.workload}: {desc_txt}", "global_points_per_step": n_local * world,
            "points_per_step_per_gpu": n_local,
            "parallelism": f"dp{world} (points sharded, weights replicated, one all-reduce of the flat gradient per step)",
            "l2": "per-GPU inputs + outputs + cotangents of a step exceed the 126 MB L2 (40 B/point in, out and cotangents each)"
                  if n_local >= (1 << 21) else "per-GPU working set may fit L2 at this size",
            "init": "reference sine init, torch.manual_seed(0); points U[-1,1]^D seed 1234; cotangents randn/N seed 4321"}


class OperatorCase:
    """one workload at one per-rank batch: synthetic points, cotangents and buffers; step() = forward streams + reverse
    sweep (+ the all-reduce of the flat gradient when world > 1)"""

    def __init__(self, workload, n_local, dev, rank, world):
        import torch.distributed as dist
        from insr_pde_b200 import _ops
        self.dist, self._ops = dist, _ops
        self.workload, self.N, self.world, self.dev = workload, n_local, world, dev
        self.D, self.O, self.H, self.L, self.order, self.text = WORKLOADS[workload]
        self.net = synth_theta((self.D, self.O, self.H, self.L), dev)
        self.theta, self.desc = self.net.flat_theta(), self.net.desc
        self.P = self.theta.numel()
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.x = (torch.rand(n_local, self.D, generator=gen, device=dev) * 2 - 1).contiguous()
        self.cg = torch.Generator(device=dev).manual_seed(4321 + rank)
        self.shapes = _ops.out_shapes(self.desc, n_local, self.order)
        self.cots = [torch.randn(sh, generator=self.cg, device=dev) / n_local for sh in self.shapes]
        self.f_fwd = flops_fwd_per_point(self.D, self.O, self.H, self.L, self.order)
        # the exchange of a step.  Where the ranks can map each other's memory the flat gradient lives in PEER memory and
        # is reduced by the library's own one-shot kernel over NVLink (insr_peer_allreduce: every rank reads all copies
        # and adds them in rank order) into gsum; otherwise -- or if the self-check below fails -- NCCL all-reduces it in
        # place.  The choice is made collectively: every rank takes the same path.
        self.peer, self.collective = None, None
        if world > 1:
            from insr_pde_b200 import peer as _peer
            # one-shot means every rank reads all W copies: right for the latency-bound payloads of the script networks
            # (3.6-56 KB), wrong for megabytes (measured at 8 GPUs: H = 512, 3-5 MB per exchange, 5 % slower per step than
            # NCCL's ring) -- above PEER_MAX_FLOATS the flat gradient goes through NCCL
            if self.P <= PEER_MAX_FLOATS:
                self.peer = _peer.PeerBuffer.create(self.P, dev)
            self.collective = "nccl all-reduce of the flat gradient"
        self.gtheta = self.peer.data if self.peer is not None else torch.zeros(self.P, device=dev)
        self.gsum = torch.empty(self.P, device=dev) if self.peer is not None else self.gtheta
        if self.peer is not None:
            self._check_peer()

    def _check_peer(self):
        """one exchange of a known pattern against NCCL's result, plus the kernels' own time-out flag; any rank's failure
        sends every rank to the NCCL path"""
        dist = self.dist
        rank = dist.get_rank()
        pattern = torch.arange(self.P, device=self.dev, dtype=torch.float32) * 1e-3 + (rank + 1)
        self.gtheta.copy_(pattern)
        want = pattern.clone()
        dist.all_reduce(want)
        dist.barrier()
        self.peer.allreduce_into(self.gsum)
        torch.cuda.synchronize()
        ok = self.peer.healthy()
        if ok:                                          # twice more: the barrier epochs keep cycling
            self.peer.allreduce_into(self.gsum)
            self.peer.allreduce_into(self.gsum)
            torch.cuda.synchronize()
            ok = self.peer.healthy() and bool(torch.allclose(self.gsum, want, rtol=1e-6, atol=1e-6))
        flag = torch.tensor([1.0 if ok else 0.0], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) < 0.5:
            self.peer.close()
            self.peer = None
            os.environ["INSR_PEER_ALLREDUCE"] = "0"     # (on every rank: the verdict is collective) no further attempts in this run
            self.gtheta = torch.zeros(self.P, device=self.dev)
            self.gsum = self.gtheta
        else:
            self.collective = ("one-shot all-reduce over NVLink peer memory (own kernel k_peer_allreduce: flag barrier, every rank "
                               "reads all ranks' gradient buffers and adds them in rank order; no NCCL on the step)")
        self.gtheta.zero_()

    def step(self):
        outs = self._ops.siren_forward(self.desc, self.theta, self.x, self.order)
        self.gtheta.zero_()
        self._ops.siren_backward(self.desc, self.theta, self.x, self.order, *self.cots, gtheta=self.gtheta)
        if self.peer is not None:
            self.peer.allreduce_into(self.gsum)
        elif self.world > 1:
            self.dist.all_reduce(self.gtheta)
        return outs

    def close(self):
        """collective: release the peer allocation"""
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, warmup):
        """ms for `steps` steps: CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks"""
        for _ in range(warmup):
            self.step()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(steps):
            self.step()
        ev1.record()
        self.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


PEER_MAX_FLOATS = 1 << 18        # 1 MB: largest flat gradient exchanged by the one-shot peer-memory kernel in this bench

# the other BASELINE.json shapes, each on its own fixed GLOBAL batch (split over the ranks like the headline)
SWEEP = [("advect1D", 1 << 22), ("fluid2Dtlgn.velocity", 1 << 22), ("elasticity2Dstretch", 1 << 20), ("elasticity3Dbunny", 1 << 20),
         ("sweep.h64", 1 << 20), ("sweep.h128", 1 << 20), ("sweep.h128.l5", 1 << 19), ("sweep.3d.h128", 1 << 19),
         ("sweep.h256", 1 << 18), ("sweep.h256.l4", 1 << 18), ("sweep.h512.l3", 1 << 16), ("sweep.h512", 1 << 16)]


def time_call(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run_ours(args):
    import gc
    import math
    import threading
    import torch.distributed as dist
    import insr_pde_b200 as ib
    from insr_pde_b200 import _lib, _ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.get_lib()
    N = args.points // world if args.scaling == "strong" else args.points
    case = OperatorCase(args.workload, N, dev, rank, world)
    D, O, H, L, order, desc_txt = case.D, case.O, case.H, case.L, case.order, case.text
    net, theta, desc, P, x, cots, gtheta, shapes = case.net, case.theta, case.desc, case.P, case.x, case.cots, case.gtheta, case.shapes
    cg = case.cg
    step, barrier = case.step, case.barrier

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.launch_count(True)
    ms = case.timed(args.steps, 0)
    launches = lib.launch_count(True) + args.steps * (1 + (1 if (world > 1 and case.peer is None) else 0))   # + zero-fill (+ NCCL; the peer-memory all-reduce is one of ours and counted by the library)
    # sustain the load long enough for the clock sampler when the timed region is short.  The step
    # contains a collective for N > 1, so EVERY rank runs the same (deterministic) number of extra steps.
    n_extra = int(max(0.0, 1000.0 - ms) / max(ms / args.steps, 1e-3)) + 1
    for _ in range(min(n_extra, 2000)):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    pts_per_s = world * N * args.steps / (ms / 1e3)

    # ---- the same step with a fixed batch PER GPU (weak scaling), 2^22 points each
    weak = None
    if args.weak_points > 0:
        wc = case if args.weak_points == N else OperatorCase(args.workload, args.weak_points, dev, rank, world)
        wms = wc.timed(args.steps, 3)
        weak = {"points_per_gpu": args.weak_points, "ms_per_step": round(wms / args.steps, 4),
                "value": round(world * args.weak_points * args.steps / (wms / 1e3), 1), "unit": "points/s", "scaling": "weak"}
        if wc is not case:
            wc.close()
        del wc

    # ---- dominant kernel alone (backward), CUDA events on the launch stream
    ws_b = torch.empty(lib.workspace_bytes(desc, N, order, True) + 16, dtype=torch.uint8, device=dev)
    ws_f = torch.empty(lib.workspace_bytes(desc, N, order, False) + 16, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    outs = [torch.empty(s, device=dev) for s in shapes]
    optr = [o.data_ptr() for o in outs] + [None] * (3 - len(outs))
    cptr = [c.data_ptr() for c in cots] + [None] * (3 - len(cots))

    def bwd_only():
        lib.backward(desc, theta.data_ptr(), x.data_ptr(), N, order, cptr[0], cptr[1], cptr[2], gtheta.data_ptr(),
                     None, ws_b.data_ptr(), ws_b.numel() - 16, stream)

    def fwd_only():
        lib.forward(desc, theta.data_ptr(), x.data_ptr(), N, order, optr[0], optr[1], optr[2], ws_f.data_ptr(),
                    ws_f.numel() - 16, stream)

    reps = max(3, min(args.steps, 20))
    ms_bwd, ms_fwd = time_call(bwd_only, reps), time_call(fwd_only, reps)
    del ws_b, ws_f, outs
    f_fwd = case.f_fwd
    peak, peak_how = fp32_peak_tflops()
    bf16_peak, bf16_how = bf16_peak_tflops()
    fam_b = lib.kernel_family(desc, order, True)
    tensor_path = fam_b == 1 and not (desc.flags & (_lib.FLAG_NO_TENSOR | _lib.FLAG_FFMA_BWD))
    no_tensor = bool(desc.flags & _lib.FLAG_NO_TENSOR)
    wide_tensor = fam_b == 2 and not no_tensor and len(shapes) <= 3 and order <= 2
    tensor_pipe_txt = ("tcgen05: 3xTF32 forward / data gradient (hi operand in TMEM), 2-level bf16 weight gradient, FP32 accumulators in TMEM"
                       if tensor_path else
                       ("tcgen05 hidden-layer GEMMs: 3xTF32 forward + data gradient, 2-level bf16 weight gradient with TMEM-resident accumulators"
                        if wide_tensor else "fp32 ffma"))
    ach_bwd = 2 * f_fwd * N / (ms_bwd / 1e3) / 1e12
    ach_fwd = f_fwd * N / (ms_fwd / 1e3) / 1e12
    bytes_pt = 4 * (D + 2 * sum(math.prod(s[1:]) for s in shapes))   # x + outputs written + cotangents read
    hbm, hbm_how = hbm_peak_gbs()
    step_frac = 3 * f_fwd * N * args.steps / (ms / 1e3) / 1e12 / peak
    fp32_view = {"bound": "fp32", "achieved": round(ach_bwd, 3), "peak": round(peak, 2), "unit": "TFLOP/s",
                 "frac": round(ach_bwd / peak, 4), "peak_source": peak_how, "step_frac": round(step_frac, 4),
                 "fwd_kernel": {"achieved": round(ach_fwd, 3), "frac": round(ach_fwd / peak, 4)},
                 "note": "algorithmic FP32 flops (SURVEY.md 8d: 2 MAC(S) per point forward, twice that backward) against the FP32 FFMA "
                         "peak -- the north star's FP32 roofline; not a bound for a kernel that executes on the tensor pipe"}
    roofline = {
        "kernel": "siren backward (recompute + dgrad + wgrad)",
        "traffic": None, "algorithmic_flops_per_point": {"fwd": f_fwd, "bwd": 2 * f_fwd},
        "ms_per_launch": {"bwd": round(ms_bwd, 4), "fwd": round(ms_fwd, 4)},
        "hbm": {"algorithmic_bytes_per_point": bytes_pt,
                "achieved_gbs": round(bytes_pt * N / ((ms_bwd + ms_fwd) / 1e3) / 1e9, 1), "peak_gbs": hbm, "peak_source": hbm_how},
        "kernel_family": {"fwd": lib.kernel_family(desc, order, False), "bwd": fam_b, "pipe": tensor_pipe_txt},
        "fp32": fp32_view, "step_frac": round(step_frac, 4),
    }
    if tensor_path:
        # the kernel executes on the tensor pipe: EXECUTED work of the backward kernel in bf16-equivalent flops (a TF32
        # flop costs two bf16 flops of pipe time): forward recompute + data gradient = 2 x 3 TF32 products, weight
        # gradient = 4 bf16 products, each 2 S L HP^2 flops per point (HP = 32: the padded width the MMAs run at)
        S_ = (f_fwd // 2 - D * H) // (L * H * H + H * O)
        hp = 32
        exec_bf16eq = (2 * (2 * 3) + 4) * 2 * S_ * L * hp * hp
        ach_t = exec_bf16eq * N / (ms_bwd / 1e3) / 1e12
        roofline.update({"bound": "tensor", "achieved": round(ach_t, 1), "peak": bf16_peak, "unit": "TFLOP/s",
                         "frac": round(ach_t / bf16_peak, 4), "peak_source": bf16_how,
                         "bound_note": "tensor pipe: executed bf16-equivalent flops of the backward kernel (3xTF32 = 6, 2-level bf16 weight "
                                       "gradient = 4 bf16-equivalents per algorithmic flop) / its launch time against the measured dense bf16 "
                                       "throughput; 'fp32' holds the same launch against the FP32 FFMA peak (algorithmic flops)"})
    else:
        roofline.update({k: fp32_view[k] for k in ("bound", "achieved", "peak", "unit", "frac", "peak_source")})
    # DRAM traffic and pipe utilisation of the dominant kernel: one ncu capture at the bench size, committed under profiles/
    # and keyed to a hash of the kernel source -- a capture of an older kernel is not quoted
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_k_tc_bwd.json")))
        if tr.get("source_sha16") != source_sha16("insr_pde_b200/csrc/siren_tc.cuh"):
            roofline["traffic_note"] = "committed ncu capture is of an older siren_tc.cuh: not quoted"
        elif tensor_path and args.workload == "fluid2Dtlgn.pressure":
            k = tr["kernel"]
            per_point = (k["dram_read_bytes"] + k["dram_write_bytes"]) / tr["points"]
            roofline["traffic"] = int(per_point * N)
            roofline["traffic_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum of {k['name']} measured at {tr['points']} points "
                                        f"({tr['source']}), scaled per point to this launch; algorithmic bytes of the backward = "
                                        f"{4 * (D + sum(math.prod(sh[1:]) for sh in shapes)) * N}")
            if "pipe_tensor_pct" in k:
                roofline["ncu"] = {"sm__pipe_tensor_cycles_active_pct": k["pipe_tensor_pct"], "issue_slots_pct": k.get("issue_pct"),
                                   "warps_active_pct": k.get("warps_active_pct")}
    except Exception:
        pass

    # ---- fused closure step (insr_siren_lsq_step): forward streams + residual + loss + backward in ONE kernel
    fused_closure = None
    if fam_b == 1:
        tgt = torch.randn(N, 1, generator=cg, device=dev)
        loss_buf = torch.zeros(1, device=dev)
        cyc = [[0.0] * O] if order == 2 else [[1.0] * O]
        clc = [[1.0] * O] if order == 2 else None

        def lsq_only():
            _ops.siren_lsq_step(desc, theta, x, order, cyc, None, clc, tgt, 1.0 / N, loss_out=loss_buf, gtheta=gtheta)

        ms_lsq = time_call(lsq_only, reps)
        fused_closure = {"ms_per_step": round(ms_lsq, 4), "points_per_s": round(world * N / (ms_lsq / 1e3), 1),
                         "achieved_tflops": round(3 * f_fwd * N / (ms_lsq / 1e3) / 1e12, 3),
                         "frac_fp32": round(3 * f_fwd * N / (ms_lsq / 1e3) / 1e12 / peak, 4),
                         "kernel": ("k_tc_bwd" if tensor_path else "k_fused_bwd") + "<..., LSQ=true>: loss = mean((lap - target)^2) and d loss/d theta in one kernel, no output round trip"}
        del tgt

    # ---- end to end through the public API with HOST buffers (rank-local shard)
    Ne = args.e2e_points or N
    xh = (torch.rand(Ne, D) * 2 - 1).pin_memory()
    th = torch.randn(Ne, 1).pin_memory()
    grad_host = torch.empty(P).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)

    def fetch():
        """host -> device copy of ONE step's inputs from pinned memory, on the copy stream"""
        with torch.cuda.stream(copy_stream):
            xd = xh.to(dev, non_blocking=True)
            td = th.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xd, td, ev

    def e2e_step(staged, prefetch):
        """one step on inputs staged by fetch(); the NEXT step's H2D copy is enqueued first so that it overlaps this
        step's kernels (every step still pays its own copy inside the timed region)"""
        xd, td, ev = staged
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        xd.record_stream(cur)
        td.record_stream(cur)
        nxt = fetch() if prefetch else None
        xd.requires_grad_(True)
        for p in net.parameters():
            p.grad = None
        y = net(xd)
        loss = torch.mean((td - ib.laplace(y, xd)) ** 2) + torch.mean(ib.gradient(y, xd) ** 2) if order == 2 else \
            torch.mean((y - td) ** 2) + (torch.mean(ib.gradient(y, xd) ** 2) if order >= 1 else 0.0)
        loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        if world > 1:
            dist.all_reduce(g)
        grad_host.copy_(g, non_blocking=True)
        return float(loss.detach()), nxt            # device -> host read of the step's result (syncs)

    staged = fetch()
    for i in range(3):
        _, staged = e2e_step(staged, True)
    barrier()
    k_e = max(2, min(args.steps, 10))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    staged = fetch()                                    # k_e copies inside the timed region: this one is exposed,
    for i in range(k_e):                                # the others overlap the previous step's kernels
        _, staged = e2e_step(staged, i + 1 < k_e)
    b.record()
    barrier()
    ms_e = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e = {"value": round(world * Ne * k_e / (float(ms_e.item()) / 1e3), 1), "unit": "points/s",
           "h2d_bytes_per_step": int(xh.numel() * 4 + th.numel() * 4), "d2h_bytes_per_step": int(P * 4 + 4),
           "points_per_step_per_gpu": Ne, "steps": k_e,
           "api": "MLP.forward + diff_ops.laplace/gradient + loss.backward() (autograd boundary); inputs in pinned host buffers, each step's H2D copy issued on a copy stream one step ahead; gradient + loss read back every step",
           "note": None if world == 1 else f"the {world} ranks share the host's PCIe / memory bandwidth for their H2D copies"}
    del xh, th, staged

    # ---- the script's own batch size (128^2 points / iteration), launch-latency bound
    Ns = min(16384, N)
    xs = x[:Ns].contiguous()
    cs = [c[:Ns].contiguous() for c in cots]

    def small_step():
        _ops.siren_forward(desc, theta, xs, order)
        gtheta.zero_()
        _ops.siren_backward(desc, theta, xs, order, *cs, gtheta=gtheta)

    ms_small = time_call(small_step, 50)
    script = {"points": Ns, "us_per_step": round(ms_small * 1e3, 2), "points_per_s": round(Ns / (ms_small / 1e3), 1)}

    # ---- the other BASELINE.json shapes: each a fixed global batch split over the ranks, gradient all-reduce included
    sweep = None
    if not args.no_sweep:
        sweep = {}
        for wl, g_pts in SWEEP:
            try:
                n_loc = g_pts // world
                sc = OperatorCase(wl, n_loc, dev, rank, world)
                k_s = 3
                sms = sc.timed(k_s, 2)
                v = world * n_loc * k_s / (sms / 1e3)
                sweep[wl] = {"global_points": n_loc * world, "ms_per_step": round(sms / k_s, 4), "points_per_s": round(v, 1),
                             "frac_fp32_step": round(3 * sc.f_fwd * v / 1e12 / (peak * world), 4),
                             "family": lib.kernel_family(sc.desc, sc.order, True), "params": sc.P}
                if world > 1:
                    sweep[wl]["exchange"] = "peer" if sc.peer is not None else "nccl"
                sc.close()
                del sc
            except Exception as e:                      # a secondary measurement must not take the headline line down
                sweep[wl] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            torch.cuda.empty_cache()

    # ---- seconds per PDE time step (second half of BASELINE.json's metric): fluid2Dtlgn, fixed iterations
    timestep = None
    if args.timestep_iters > 0:
        try:
            timestep = fluid_timestep_ours(dev, args.timestep_iters, world)
        except Exception as e:                          # the second half of the metric must not take the first half down
            timestep = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        try:                                            # the same time step at a batch that CAN be sharded: 1024^2 points / iteration
            timestep["large_batch"] = fluid_timestep_ours(dev, max(args.timestep_iters // 5, 5), world, sample_resolution=1024)
        except Exception as e:
            timestep["large_batch"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        if world == 1:                                  # the 32 < H <= 512 family's closure, same metric
            for key, fn in (("elasticity", lambda: [elasticity_timestep_ours(dev, args.timestep_iters, c) for c in ELASTIC_CASES]),
                            ("advection", lambda: advection_timestep_ours(dev, args.timestep_iters))):
                try:                                    # secondary measurements must not take the headline line down
                    timestep[key] = fn()
                except Exception as e:
                    timestep[key] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---- baselines: the reference itself, in its own process.  cpu_baseline: rank 0 at N = 1 only (the contract); the
    # like-for-like leg (the same reference code as stock PyTorch on this GPU) likewise.
    cpu, torch_gpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        ref_cpu = reference_subprocess(args, "cpu", legs="all", steps=16, warmup=2, budget=args.cpu_budget)
        cpu = ref_cpu.get("cpu_baseline") or {"error": ref_cpu.get("error", "no line")}
        if isinstance(timestep, dict) and "error" not in timestep:
            timestep["cpu_reference"] = ref_cpu.get("timestep")
            timestep["cpu_reference_elasticity_sec_per_iteration"] = ref_cpu.get("elasticity_sec_per_iteration")
        ref_gpu = reference_subprocess(args, "cuda", legs="all", steps=10, warmup=3, budget=8.0, points=min(N, 1 << 20))
        torch_gpu = ref_gpu.get("cpu_baseline") or {"error": ref_gpu.get("error", "no line")}
        if isinstance(torch_gpu, dict) and "error" not in torch_gpu:
            torch_gpu["timestep"] = ref_gpu.get("timestep")
            torch_gpu["elasticity_sec_per_iteration"] = ref_gpu.get("elasticity_sec_per_iteration")

    if rank == 0:
        cfg = config_block(args, world)
        cfg["params"] = P
        line = {
            "metric": METRIC, "value": round(pts_per_s, 1),
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "weak": weak, "sweep": sweep, "script_size": script,
            "timestep": timestep, "fused_closure": fused_closure,
            "collective": None if world == 1 else f"{case.collective} (sum of the flat fp32 gradient, one per step, on the compute stream)",
        }
        emit(line)
    case.close()
    if world > 1:
        # Orderly teardown: every captured iteration graph (they hold NCCL kernels) has been dropped by the stepper
        # helpers above; collect, drain the device, then destroy the communicator.  A watchdog turns a teardown that
        # does not return (seen once at N = 2 in round 1, before the graphs were released first) into a clean exit.
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        done = threading.Event()

        def watchdog():
            if not done.wait(30.0):
                sys.stderr.write("bench.py: destroy_process_group did not return within 30 s; leaving without it\n")
                sys.stderr.flush()
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        dist.destroy_process_group()
        done.set()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fluid2Dtlgn.pressure", choices=sorted(WORKLOADS))
    ap.add_argument("--points", type=int, default=1 << 24,
                    help="collocation points per step: in total (--scaling strong, the default) or per GPU (--scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default, the driver's scaling run): --points in total, split over the GPUs; weak: --points per GPU")
    ap.add_argument("--weak-points", type=int, default=1 << 22, help="per-GPU batch of the extra weak-scaling measurement (0 = skip)")
    ap.add_argument("--e2e-points", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=14.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--timestep-iters", type=int, default=100, help="Adam iterations per training loop of the time-step measurements (0 = skip)")
    # reference arm
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"], help="(--impl reference) cpu = the arm the driver runs; cuda = the like-for-like leg")
    ap.add_argument("--ref-points", type=int, default=1 << 16, help="(--impl reference) points per timed step: a bounded sample of the global batch")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="(--impl reference) wall-clock cap of the timed steps, seconds")
    ap.add_argument("--ref-legs", default="operator", choices=["operator", "all"])
    args = ap.parse_args()
    keep_stdout_clean()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
