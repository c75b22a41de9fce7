#!/bin/bash
# ncu evidence for the round (run under gpurun, ONE GPU).  Produces, under gpurun_out/:
#   launches.csv   every kernel launch of a short bench run with its device time (cold, serialised: SHARES only)
#   edge.ncu-rep   --set full capture of 3 launches of the fused edge-layer kernel
# Numbers printed by bench.py under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
export DFB_NCU=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_tc -s 14 -c 3 -o gpurun_out/edge \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
ls -la gpurun_out
