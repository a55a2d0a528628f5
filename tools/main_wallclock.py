"""Wall clock of the reference's UNMODIFIED main.py at the scripts' configurations, per time step, in four modes on the same GPU:
    reference   the reference alone (stock PyTorch, oracle/_ref)
    dropin      python -m insr_pde_b200.patch main.py ...                    (field / operator layer only)
    closures    ... --insr-closures                                          (+ one-kernel loss closures)
    graphed     ... --insr-graphed                                           (+ CUDA-graphed iteration under @_training_loop)
Every mode is its own process; the clock runs around model.initialize() / model.step() inside main.py's time loop (hooked through
the model class, main.py itself unchanged), outputs and checkpoints included as main.py writes them.
usage: python tools/main_wallclock.py [iters_per_loop] [case ...]      (child: --child <mode> <case> <iters> <dir>)"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # scripts/*.sh argument lists; -T and --max_n_iters are set by this tool
    "fluid2Dtlgn": ("fluid", ["--init_cond", "taylorgreen", "--num_hidden_layers", "3", "--hidden_features", "32", "-sr", "128",
                              "-vr", "32", "--dt", "0.05"], 3),
    "advect1D": ("advection", ["--init_cond", "example1", "--num_hidden_layers", "2", "--hidden_features", "20", "-sr", "5000",
                               "--dt", "0.05"], 3),
    "elasticity2Dstretch": ("elasticity", ["--num_hidden_layers", "3", "--hidden_features", "68", "-sr", "100", "-vr", "100",
                                           "--lr", "1e-4", "--dim", "2", "--energy", "arap", "constraint", "constraint_right", "volume",
                                           "--ratio_volume", "1e3", "--ratio_arap", "1e0", "--ratio_constraint", "1e4",
                                           "--constraint_right_offset_x", "2.0"], 2),      # (the script has -T 1; 2 so that the last step shows the steady state, not the graph capture)
    "elasticity3Dbunny": ("elasticity", ["--num_hidden_layers", "3", "--hidden_features", "66", "-sr", "20", "-vr", "1000", "--dt", "0.1",
                                         "--lr", "1e-4", "--dim", "3", "--energy", "arap", "kinematics", "collision", "external", "volume",
                                         "--ratio_volume", "1e3", "--ratio_arap", "1e2", "--ratio_collide", "1e6", "--ratio_kinematics", "1e0",
                                         "-f_ext_x", "0", "-f_ext_y", "0", "-f_ext_z", " -1e2", "-T_ext", "5", "--plane_height", "-2",
                                         "--use_mesh", "1", "--mesh_path", "./elasticity/data/bunny.mesh"], 2),
}
MODES = tuple(os.environ.get("INSR_WALLCLOCK_MODES", "reference,dropin,closures,graphed").split(","))


def child(mode, case, iters, tmp):
    import numpy as np
    import torch
    from insr_pde_b200 import patch
    from oracle import ref_loader
    root = ref_loader.REF_ROOT
    pde, args, T = CASES[case]
    argv = [pde, *args, "-T", str(T), "--max_n_iters", str(iters), "--no-early_stop", "--proj_dir", tmp, "--tag", f"{case}_{mode}"]
    if mode == "reference":
        ref_loader.load(cpu=False)
    else:
        patch.install(root)
    import importlib
    mod = importlib.import_module({"advection": "advection.model", "fluid": "fluid.model", "elasticity": "elasticity.model"}[pde])
    cls = getattr(mod, {"advection": "Advection1DModel", "fluid": "Fluid2DModel", "elasticity": "ElasticityModel"}[pde])
    if pde == "elasticity":
        mod.write_pointcloud_to_file = lambda path, values: np.save(path + ".npy", np.asarray(values))
    marks = []

    def clocked(name):
        orig = getattr(cls, name)

        def f(self, *a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = orig(self, *a, **k)
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter() - t0))
            return out
        setattr(cls, name, f)
    for name in ("initialize", "step", "write_output"):
        clocked(name)
    torch.manual_seed(123)
    np.random.seed(123)
    t0 = time.perf_counter()
    if mode == "reference":
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
        patch.run_main(argv, root, fused_closures=(mode == "closures") or None, graphed=(mode == "graphed"))
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    steps = [t for n, t in marks if n == "step"]
    print(json.dumps({"case": case, "mode": mode, "iters_per_loop": iters, "main_py_seconds": round(total, 3),
                      "initialize_s": round(sum(t for n, t in marks if n == "initialize"), 3),
                      "step_s": [round(t, 4) for t in steps], "write_output_s": round(sum(t for n, t in marks if n == "write_output"), 3)}))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    cases = sys.argv[2:] or list(CASES)
    rows = []
    for case in cases:
        for mode in MODES:
            with tempfile.TemporaryDirectory() as tmp:
                res = subprocess.run([sys.executable, "-W", "ignore", os.path.abspath(__file__), "--child", mode, case, str(iters), tmp],
                                     capture_output=True, text=True, timeout=1200)
            if res.returncode != 0:
                print(f"{case:22s} {mode:10s} FAILED: {res.stderr[-600:]}")
                continue
            r = json.loads(res.stdout.strip().splitlines()[-1])
            rows.append(r)
            last = r["step_s"][-1] if r["step_s"] else float("nan")
            print(f"{case:22s} {mode:10s} iters/loop {iters:5d}  initialize {r['initialize_s']:8.3f} s  last time step {last:8.4f} s  "
                  f"main.py {r['main_py_seconds']:8.3f} s  (outputs {r['write_output_s']:.3f} s)", flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
    else:
        main()
