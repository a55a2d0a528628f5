#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2l}
timeout 600 python -m pytest tests/test_reference_dropin_gpu.py -m gpu -q -x -k "main_py and elasticity" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log | cut -c1-300
INSR_WALLCLOCK_MODES=graphed timeout 600 python tools/main_wallclock.py 200 elasticity2Dstretch elasticity3Dbunny > gpurun_out/main_wallclock_ela_$TAG.txt 2>&1; grep -v "^\[" gpurun_out/main_wallclock_ela_$TAG.txt | cut -c1-400
