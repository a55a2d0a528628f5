#!/bin/bash
# 1 GPU: parity of the changed tcgen05 kernels, a short bench line, the iteration kernel lists, main.py wall clock in four modes
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_parity_$TAG.log 2>&1; echo "parity rc=$?" >> gpurun_out/pytest_parity_$TAG.log
tail -3 gpurun_out/pytest_parity_$TAG.log
timeout 600 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "bwd/fwd ms", d["roofline"]["ms_per_launch"], "script", d["script_size"])
print("timestep", {k: d["timestep"].get(k) for k in ("sec_per_timestep", "us_per_iteration")}, "adv", d["timestep"].get("advection"), "el", d["timestep"].get("elasticity"))
PY
timeout 300 python tools/fluid_graph_profile.py > gpurun_out/fluid_iteration_kernels_$TAG.txt 2>&1
grep -E "^loop|k_tc|k_iter|k_sample|Fill|memcpy" gpurun_out/fluid_iteration_kernels_$TAG.txt | cut -c1-150
timeout 300 python tools/lsq_graph_probe.py > gpurun_out/lsq_graph_probe_$TAG.txt 2>&1; tail -9 gpurun_out/lsq_graph_probe_$TAG.txt
if [ -z "$2" ]; then
  timeout 1200 python tools/main_wallclock.py 200 > gpurun_out/main_wallclock_$TAG.txt 2>&1; grep -v "^\[" gpurun_out/main_wallclock_$TAG.txt | cut -c1-220
fi
