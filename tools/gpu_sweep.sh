#!/bin/bash
# per-workload single-GPU bench lines (fused + generic families) and a 2-GPU run
mkdir -p gpurun_out
TAG=${1:-sweep}
NG=${2:-1}
: > gpurun_out/sweep_$TAG.jsonl
for wl in ${WORKLOADS:-advect1D fluid2Dtlgn.velocity fluid2Dtlgn.pressure elasticity2Dstretch elasticity3Dbunny sweep.h64 sweep.h128 sweep.h256 sweep.h512}; do
  pts=1048576
  case $wl in sweep.h256*|sweep.3d.h256) pts=131072;; sweep.h512) pts=32768;; sweep.h128*|sweep.3d.h128) pts=262144;; esac
  timeout 300 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/sweep_$TAG.jsonl 2>> gpurun_out/sweep_$TAG.err
done
if [ "$NG" -gt 1 ]; then
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n$NG.json 2> gpurun_out/bench_${TAG}_n$NG.err
  echo "multi rc=$?"; tail -n 3 gpurun_out/bench_${TAG}_n$NG.err; cat gpurun_out/bench_${TAG}_n$NG.json
fi
python -c "
import json
for l in open('gpurun_out/sweep_$TAG.jsonl'):
    d=json.loads(l); r=d['roofline']
    print(d['config']['workload'].split(':')[0], 'N', d['config']['points_per_step_per_gpu'], 'Mpts/s %.1f'%(d['value']/1e6), 'fam', r['kernel_family'], 'fwd frac', r['fwd_kernel']['frac'], 'bwd frac', r['frac'], 'step frac', r['step_frac'], 'e2e %.1f'%(d['e2e']['value']/1e6))
"
