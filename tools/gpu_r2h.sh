#!/bin/bash
# 1 GPU: steppers / graphed loops after the sample-ahead change, iteration kernel lists, short bench, bunny main.py graphed
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_dropin_gpu.py -m gpu -q -x -k "stepper or graphed or loop or timestep or main_py or sampler" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python tools/fluid_graph_profile.py > gpurun_out/fluid_iteration_kernels_$TAG.txt 2>&1
grep -E "^loop" gpurun_out/fluid_iteration_kernels_$TAG.txt | cut -c1-150
timeout 300 python tools/elastic_graph_profile.py > gpurun_out/elastic_iteration_kernels_$TAG.txt 2>&1
grep -E "^case" gpurun_out/elastic_iteration_kernels_$TAG.txt | cut -c1-150
timeout 600 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("timestep", {k: d["timestep"].get(k) for k in ("sec_per_timestep", "us_per_iteration")}, "adv", d["timestep"].get("advection", {}).get("us_per_iteration"), "el", [e.get("us_per_iteration") for e in d["timestep"].get("elasticity", [])])
PY
timeout 900 python tools/main_wallclock.py 200 elasticity3Dbunny > gpurun_out/main_wallclock_bunny_$TAG.txt 2>&1; grep -v "^\[" gpurun_out/main_wallclock_bunny_$TAG.txt | cut -c1-220
