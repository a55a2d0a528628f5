#!/bin/bash
# round 2, run B: the full GPU test-suite (incl. script-config trajectories), then the bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b.log
grep -E "per-frame|passed|failed|rc=|Error" gpurun_out/pytest_gpu_r2b.log | cut -c1-400 | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2b.json'))
print('value',d['value'],'e2e',d['e2e']['value'])
print('timestep', json.dumps(d['timestep'])[:2500])
for k,v in (d.get('sweep') or {}).items(): print(k, v)
"
