#!/bin/bash
# round 2, run A: GPU tests (incl. the drop-in tests on the shipped reference copy), the new bench line, kernel breakdowns
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
tail -15 gpurun_out/pytest_gpu_r2a.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r2a.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
for k in ('value','ms_per_step','e2e','weak','cpu_baseline','torch_gpu_baseline','script_size','fused_closure'): print(k, json.dumps(d.get(k))[:400])
print('roofline', json.dumps(d['roofline'])[:900])
print('timestep', json.dumps(d['timestep'])[:1500])
for k,v in (d.get('sweep') or {}).items(): print(k, v)
"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_r2a.json
for wl in elasticity2Dstretch elasticity3Dbunny sweep.h64; do
  timeout 120 python tools/step_kernels.py $wl 1048576 0 2>&1 | grep -v Warn >> gpurun_out/step_kernels_r2a.txt
  timeout 120 python tools/step_kernels.py $wl 1048576 1 2>&1 | grep -v Warn >> gpurun_out/step_kernels_r2a.txt
done
cat gpurun_out/step_kernels_r2a.txt
