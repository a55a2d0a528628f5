#!/bin/bash
# round-2 evidence: launch list of the default bench command, ncu --set full of the top kernels, DRAM traffic at the bench size
mkdir -p gpurun_out
TAG=${1:-r2}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --timestep-iters 0 > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --timestep-iters 0 > gpurun_out/ncu_list_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_list_$TAG.log | cut -c1-200
python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/prof_plain_$TAG.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k_tc_" --csv --log-file gpurun_out/traffic_$TAG.csv python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/ncu_traffic_$TAG.log 2>&1
tail -n 4 gpurun_out/traffic_$TAG.csv | cut -c1-250
python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/prof_plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_tc_" -c 3 -f -o gpurun_out/prof_tc_$TAG \
    python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/ncu_full_tc_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_full_tc_$TAG.log
bash tools/gpu_ncu_mid.sh mid_$TAG elasticity2Dstretch
