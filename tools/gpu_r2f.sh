#!/bin/bash
# final checks on one GPU: the whole -m gpu suite, smoke(), evidence captures, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
grep -E "per-frame|wall clock|passed|failed|rc=|Error" gpurun_out/pytest_gpu_final.log | cut -c1-500 | tail -14
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep smoke
