#!/bin/bash
# ncu --set full capture of the fused mid-width kernels (elasticity2Dstretch's network, 2^18 points; plain forward, then
# the backward that tapes its own forward: k_mid_fwd<tape>, k_mid_dgrad, k_mid_wgrad, k_mid_edge)
mkdir -p gpurun_out
TAG=${1:-mid}
WL=${2:-elasticity2Dstretch}
timeout 120 python tools/step_kernels.py $WL 262144 0 > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mid_ -s 10 -c 5 -f -o gpurun_out/prof_$TAG \
    python tools/step_kernels.py $WL 262144 0 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
