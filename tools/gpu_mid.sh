#!/bin/bash
# fused mid-width kernels: parity, then timings
mkdir -p gpurun_out
TAG=${1:-mid}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_width" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest mid rc=$?" >> gpurun_out/pytest_$TAG.log
tail -25 gpurun_out/pytest_$TAG.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity_$TAG.log 2>&1; echo "pytest parity rc=$?" >> gpurun_out/pytest_parity_$TAG.log
tail -6 gpurun_out/pytest_parity_$TAG.log | cut -c1-400
: > gpurun_out/step_kernels_$TAG.txt
for wl in elasticity2Dstretch elasticity3Dbunny; do
  timeout 120 python tools/step_kernels.py $wl 1048576 0 2>&1 | grep -v "Warn\|warn" >> gpurun_out/step_kernels_$TAG.txt
done
cat gpurun_out/step_kernels_$TAG.txt
timeout 300 python tools/elastic_step_bench.py 200 2>&1 | grep -v "Warn\|return float\|warn" > gpurun_out/elastic_step_$TAG.txt; cat gpurun_out/elastic_step_$TAG.txt
