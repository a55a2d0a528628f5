#!/bin/bash
# one gpurun call: tests, bench, launch list, one full capture of the top kernels, DRAM traffic at the bench size
mkdir -p gpurun_out
TAG=${1:-run}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
if [ "$2" == "ncu" ]; then
  python tools/prof_one.py --points 1048576 --reps 2 --lsq > gpurun_out/prof_plain_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
      python tools/prof_one.py --points 1048576 --reps 2 --lsq > gpurun_out/ncu_list_$TAG.log 2>&1
  python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/prof_plain2_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k_tc_|k_fused" -c 3 -f -o gpurun_out/prof_$TAG \
      python tools/prof_one.py --points 262144 --reps 1 --lsq > gpurun_out/ncu_full_$TAG.log 2>&1
  # DRAM traffic of the hot kernels at the bench size (2^22 points), second repetition (warm)
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_tc_|k_fused" \
      --csv --log-file gpurun_out/traffic_$TAG.csv python tools/prof_one.py --points 4194304 --reps 2 --lsq > gpurun_out/ncu_traffic_$TAG.log 2>&1
  tail -n 3 gpurun_out/ncu_list_$TAG.log; tail -n 3 gpurun_out/ncu_full_$TAG.log; tail -n 7 gpurun_out/traffic_$TAG.csv | cut -c1-300
fi
if [ "$3" == "elastic" ]; then
  timeout 300 python tools/elastic_step_bench.py 200 2>&1 | grep -v "Warn\|return float" > gpurun_out/elastic_step_$TAG.txt; cat gpurun_out/elastic_step_$TAG.txt
  (timeout 200 python tools/elastic_graph_profile.py 2d; timeout 200 python tools/elastic_graph_profile.py 3d) 2>&1 | grep -v "Warn\|_warn" > gpurun_out/elastic_kernels_$TAG.txt
  (timeout 200 python tools/wide_scaling_probe.py 68 2 0; timeout 200 python tools/wide_scaling_probe.py 68 2 1) 2>&1 | grep tiles > gpurun_out/wide_scaling_$TAG.txt
fi
