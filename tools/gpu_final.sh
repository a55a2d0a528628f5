#!/bin/bash
# round-2 final evidence on one GPU (everything judged lands in gpurun_out/, summaries are made ON the box so that the pull stays
# under 64 MiB): default bench line, launch list, DRAM traffic at the bench size, ncu --set full of the tcgen05 and mid-width kernels,
# elasticity iteration timings.   usage: bash tools/gpu_final.sh [tag] [skip_bench]
mkdir -p gpurun_out
TAG=${1:-r2final}
if [ -z "$2" ]; then
  timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
  tail -c 400 gpurun_out/bench_$TAG.err
fi
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --timestep-iters 0 > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --timestep-iters 0 > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k_tc_" --csv --log-file gpurun_out/traffic_$TAG.csv python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "traffic rc=$?"
python tools/make_traffic_json.py gpurun_out/traffic_$TAG.csv 4194304 gpurun_out/traffic_k_tc_bwd.json | cut -c1-300
timeout 300 python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/prof_plain2_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_" -c 3 -f -o gpurun_out/prof_tc_$TAG \
    python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/ncu_full_tc_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_tc_$TAG.ncu-rep > gpurun_out/ncu_full_tc_$TAG.txt 2>&1
# mid-width kernels: the report is ~55 MB, so it is summarised here and only the summary + the source-page csv travel
timeout 120 python tools/step_kernels.py elasticity2Dstretch 262144 0 > gpurun_out/step_kernels_mid_$TAG.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mid_ -s 10 -c 4 -f -o /tmp/prof_mid_$TAG \
    python tools/prof_one.py --workload elasticity2Dstretch --points 262144 --reps 4 > gpurun_out/ncu_full_mid_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_mid_$TAG.ncu-rep > gpurun_out/ncu_full_mid_$TAG.txt 2>&1
ncu -i /tmp/prof_mid_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_source_mid_$TAG.csv.gz
timeout 300 python tools/step_kernels.py elasticity2Dstretch 1048576 0 >> gpurun_out/step_kernels_mid_$TAG.txt 2>&1
timeout 300 python tools/step_kernels.py elasticity3Dbunny 1048576 0 >> gpurun_out/step_kernels_mid_$TAG.txt 2>&1
timeout 300 python tools/step_kernels.py sweep.h64 1048576 0 >> gpurun_out/step_kernels_mid_$TAG.txt 2>&1
timeout 600 python tools/elastic_step_bench.py > gpurun_out/elastic_step_$TAG.txt 2>&1
timeout 300 python tools/elastic_graph_profile.py > gpurun_out/elastic_iteration_kernels_$TAG.txt 2>&1
timeout 300 python tools/fluid_graph_profile.py > gpurun_out/fluid_iteration_kernels_$TAG.txt 2>&1
du -sh gpurun_out; ls gpurun_out | head -40
