#!/bin/bash
# end of round 2 on one GPU: the whole -m gpu suite, smoke(), the default bench line, launch list, DRAM traffic at the bench size
# (keyed to the kernel source), ncu --set full of the tcgen05 kernels, iteration kernel lists (the mid-width captures of
# tools/gpu_final.sh stand: those kernels did not change)
mkdir -p gpurun_out
TAG=${1:-r2end}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "per-frame|wall clock|passed|failed|rc=|Error" gpurun_out/pytest_gpu_$TAG.log | cut -c1-400 | tail -12
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep -i smoke | tail -2
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --timestep-iters 0 > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --timestep-iters 0 > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k_tc_" --csv --log-file gpurun_out/traffic_$TAG.csv python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "traffic rc=$?"
python tools/make_traffic_json.py gpurun_out/traffic_$TAG.csv 4194304 gpurun_out/traffic_k_tc_bwd.json | cut -c1-400
timeout 300 python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/prof_plain2_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_" -c 3 -f -o gpurun_out/prof_tc_$TAG \
    python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/ncu_full_tc_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_tc_$TAG.ncu-rep > gpurun_out/ncu_full_tc_$TAG.txt 2>&1
timeout 300 python tools/elastic_step_bench.py 2>&1 | grep -a "points/iter" > gpurun_out/elastic_step_$TAG.txt
timeout 300 python tools/fluid_graph_profile.py 2>&1 | grep -v "profiler.py\|_warn_once" > gpurun_out/fluid_iteration_kernels_$TAG.txt
cat gpurun_out/elastic_step_$TAG.txt; grep "^loop" gpurun_out/fluid_iteration_kernels_$TAG.txt
du -sh gpurun_out
