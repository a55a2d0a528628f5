#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stepper or graphed or loop or prepared or timestep" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python tools/fluid_graph_profile.py > gpurun_out/fluid_iteration_kernels_$TAG.txt 2>&1
grep -v "profiler.py\|_warn_once" gpurun_out/fluid_iteration_kernels_$TAG.txt | cut -c1-150 | head -40
timeout 600 python bench.py --no-cpu-baseline --no-sweep --steps 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"])
print("timestep", d["timestep"].get("sec_per_timestep"), d["timestep"].get("us_per_iteration"), d["timestep"].get("final_losses"), "large", d["timestep"].get("large_batch", {}).get("sec_per_timestep"), "adv", d["timestep"].get("advection"))
PY
