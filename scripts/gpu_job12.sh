#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe12.log
for W in 4 2 1; do
  DFB_TC_WPQ=$W timeout 300 python scripts/probe_tc.py fwd 2>&1 | grep -E "probe=" | sed "s/^/WPQ=$W /" | tee -a gpurun_out/probe12.log
  DFB_TC_WPQ=$W timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tc_gemm or (tc and (forward_tsp_categorical_golden or traj_tsp_cat or traj_mis_cat or forward_mis_golden or forward_tsp_vs_oracle or tsp500_full)) or config4" 2>&1 | tail -2
done
