#!/bin/bash
# diagnostic: the torchrun twin of main.py on 2 GPUs with timestamps (single run first for the warm-up of the box)
mkdir -p gpurun_out; rm -rf /tmp/dpd; mkdir -p /tmp/dpd
export INSR_PATCH_VERBOSE=1 INSR_REFERENCE_ROOT=$PWD/oracle/_ref PYTHONPATH=$PWD
ARGS="fluid --init_cond taylorgreen --num_hidden_layers 3 --hidden_features 32 -sr 128 -vr 32 --dt 0.05 -T 1 --max_n_iters 100 --no-early_stop"
date +%T
for mode in "--insr-graphed"; do
  date +%T
  TORCH_DISTRIBUTED_DEBUG=OFF timeout 240 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      -m insr_pde_b200.patch $mode --insr-dp --insr-seed 5 $ARGS --proj_dir /tmp/dpd --tag twin$RANDOM > gpurun_out/dp_twin_"${mode:-eager}".log 2>&1; echo "twin [$mode] rc=$?"
  date +%T
  grep -v "it/s" gpurun_out/dp_twin_"${mode:-eager}".log | grep -E "insr-dp|rror|Traceback|time step|rank" | tail -20 | cut -c1-300
done
