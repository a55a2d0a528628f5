#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the hot kernels of every family, small batches; logs kept in gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r2}
SAN=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/sanitizer_$TAG.txt
for tool in memcheck racecheck; do
  for wl in fluid2Dtlgn.pressure fluid2Dtlgn.velocity advect1D elasticity2Dstretch elasticity3Dbunny sweep.h128; do
    echo "=== $tool $wl" >> gpurun_out/sanitizer_$TAG.txt
    timeout 240 $SAN --tool $tool --print-limit 20 python tools/prof_one.py --workload $wl --points 20000 --reps 1 --lsq 2>&1 \
      | grep -v "^ok\|Warn" | tail -n 12 >> gpurun_out/sanitizer_$TAG.txt
    echo "rc=$?" >> gpurun_out/sanitizer_$TAG.txt
  done
done
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitizer_$TAG.txt
grep "ERROR SUMMARY\|===" gpurun_out/sanitizer_$TAG.txt
