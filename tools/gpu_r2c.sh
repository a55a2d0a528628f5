#!/bin/bash
# round 2, run C (2 GPUs): trajectory + data-parallel twin tests, bench under torchrun at 2 GPUs
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_trajectory_gpu.py tests/test_dp_main_gpu.py -m gpu -x -q -s > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
grep -E "per-frame|passed|failed|rc=|Error" gpurun_out/pytest_gpu_r2c.log | cut -c1-600 | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/bench_r2c_n2.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_r2c_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'weak',d['weak'])
print('timestep', json.dumps(d['timestep'])[:1200])
for k,v in (d.get('sweep') or {}).items(): print(k, v)
"
