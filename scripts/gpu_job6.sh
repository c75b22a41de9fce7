#!/bin/bash
mkdir -p gpurun_out
DFB_TC_WPQ=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tc and (forward_tsp_categorical_golden or traj_mis_cat or forward_tsp_vs_oracle)" 2>&1 | tail -3
PROBE_B=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_tc -s 13 -c 1 \
   -o gpurun_out/edge_v2 python scripts/probe_tc.py fwd > gpurun_out/ncu_edge2.log 2>&1
tail -2 gpurun_out/ncu_edge2.log
