#!/bin/bash
# 2 GPUs: the data-parallel twin of main.py against the single-GPU frames
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_main_gpu.py -m gpu -q -s -x > gpurun_out/pytest_dp_twin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp_twin.log
grep -E "per-frame|passed|failed|rror|rc=" gpurun_out/pytest_dp_twin.log | cut -c1-500 | tail
