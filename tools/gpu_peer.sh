#!/bin/bash
# 2 (or N) GPUs: the peer-memory exchange against NCCL, the data-parallel twin of main.py, and the bench line under torchrun
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-peer}
timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q -s -x > gpurun_out/pytest_peer_$TAG.log 2>&1; echo "peer rc=$?" >> gpurun_out/pytest_peer_$TAG.log
grep -E "peer exchange|passed|failed|skipped|rror|rc=" gpurun_out/pytest_peer_$TAG.log | cut -c1-1800 | tail -12
if [ -z "$3" ]; then
  timeout 900 python -m pytest tests/test_dp_main_gpu.py -m gpu -q -s -x > gpurun_out/pytest_dp_$TAG.log 2>&1; echo "dp rc=$?" >> gpurun_out/pytest_dp_$TAG.log
  grep -E "per-frame|passed|failed|rror|rc=" gpurun_out/pytest_dp_$TAG.log | cut -c1-500 | tail -6
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "collective:", d["collective"][:90])
    print("timestep", {k: d["timestep"].get(k) for k in ("sec_per_timestep", "us_per_iteration")}, d["timestep"]["note"][-160:])
    print("large", d["timestep"].get("large_batch", {}).get("sec_per_timestep"))
    print("weak", d["weak"]); print("e2e", d["e2e"]["value"])
except Exception as e:
    print("no line:", e)
PY
