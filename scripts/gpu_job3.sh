#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe3.log
for W in 1 2 4; do
  echo "=== WPQ=$W parity subset ==="
  DFB_TC_WPQ=$W timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tc_gemm or (tc and (forward_tsp_categorical_golden or traj_tsp_cat or traj_mis_cat or forward_mis_golden or forward_tsp_vs_oracle))" 2>&1 | tail -4 | tee -a gpurun_out/t3_w$W.log
  for p in 0 16 1; do
    DFB_TC_WPQ=$W DFB_TC_PROBE=$p timeout 300 python scripts/probe_tc.py all 2>&1 | grep -E "probe=" | sed "s/^/WPQ=$W /" | tee -a gpurun_out/probe3.log
  done
done
