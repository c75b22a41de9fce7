#!/bin/bash
# one --set full capture of the CTA-pair edge kernel (full-size launch, E = 400 000) with source-level sampling
mkdir -p gpurun_out
export DFB_PAIR_KERNEL=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_pair -s 14 -c 1 \
  -o gpurun_out/pair_full -f python scripts/probe_tc.py fwd > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
ls -la gpurun_out/pair_full.ncu-rep
