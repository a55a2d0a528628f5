#!/bin/bash
# one gpurun call: tcgen05 probe + ncu capture of the tcgen05 forward kernel
mkdir -p gpurun_out
timeout 120 tools/probe/umma_probe2 > gpurun_out/probe2.txt 2>&1; echo "probe rc=$?"
tail -n 16 gpurun_out/probe2.txt
python tools/prof_one.py --points 262144 --reps 1 > gpurun_out/prof_plain_tcf.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_fwd -c 1 -f -o gpurun_out/prof_tcfwd \
    python tools/prof_one.py --points 262144 --reps 1 > gpurun_out/ncu_full_tcfwd.log 2>&1
tail -n 3 gpurun_out/ncu_full_tcfwd.log
