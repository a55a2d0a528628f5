#!/bin/bash
# one gpurun call: tcgen05 kernels vs FFMA kernels and the fp64 oracle (+ timings)
mkdir -p gpurun_out
TAG=${1:-tc}
timeout 300 python tests/checks/tc_check.py > gpurun_out/tc_check_$TAG.log 2>&1; echo "tc_check rc=$?"
tail -n 22 gpurun_out/tc_check_$TAG.log
timeout 300 python tests/checks/tc_check_bwd.py 6 time > gpurun_out/tc_check_bwd_$TAG.log 2>&1; echo "tc_check_bwd rc=$?"
tail -n 22 gpurun_out/tc_check_bwd_$TAG.log
