#!/bin/bash
# next hardware run: the K16 variant of k_wide_tc (INSR_WIDE_K16=1: 16-wide K slabs, two CTAs per SM for S >= 3) against the
# default, on the shapes it targets.  Parity first, then the workload sweep, the elasticity iteration and the tile scaling.
mkdir -p gpurun_out
for k in 0 1; do
  export INSR_WIDE_K16=$k
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or tiled or elasticity or tape" 2>&1 | tail -1
  WORKLOADS="elasticity2Dstretch elasticity3Dbunny sweep.h64" bash tools/gpu_sweep.sh k16_$k 2>&1 | tail -3 | cut -c1-70
  timeout 200 python tools/elastic_step_bench.py 200 2>&1 | grep -v "Warn\|return float" | tail -2 > gpurun_out/elastic_step_k16_$k.txt; cat gpurun_out/elastic_step_k16_$k.txt
  timeout 120 python tools/wide_scaling_probe.py 68 2 1 2>&1 | grep tiles > gpurun_out/wide_scaling_k16_$k.txt; tail -4 gpurun_out/wide_scaling_k16_$k.txt
done
