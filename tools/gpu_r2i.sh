#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_dropin_gpu.py -m gpu -q -x -k "elastic or stepper or graphed or loop" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python tools/elastic_graph_profile.py > gpurun_out/elastic_iteration_kernels_$TAG.txt 2>&1
grep -v "profiler.py\|_warn_once" gpurun_out/elastic_iteration_kernels_$TAG.txt | cut -c1-150
timeout 300 python tools/elastic_graph_profile.py 3d >> gpurun_out/elastic_iteration_kernels_$TAG.txt 2>&1
grep -E "^case 3d" gpurun_out/elastic_iteration_kernels_$TAG.txt | cut -c1-150
