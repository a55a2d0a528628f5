#!/bin/bash
# bench.py under torchrun at N GPUs (the driver's scaling run): tools/gpu_scale.sh N TAG
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench n$N rc=$?"
tail -2 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}_n$N.json'))
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'weak', d['weak'])
t=d['timestep']; print('timestep', t.get('sec_per_timestep'), t.get('us_per_iteration'), 'large', t.get('large_batch',{}).get('sec_per_timestep'), t.get('large_batch',{}).get('us_per_iteration'))
for k,v in (d.get('sweep') or {}).items(): print(k, v.get('points_per_s'), v.get('frac_fp32_step'), v.get('error'))
"
