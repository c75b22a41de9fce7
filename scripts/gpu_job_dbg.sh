#!/bin/bash
mkdir -p gpurun_out
for E in 128 256 1000; do
  DFB_PAIR_KERNEL=1 timeout 120 python scripts/debug_pair_gemm.py tc $E > gpurun_out/dbg_tc_$E.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_tc_$E.log
done
DFB_PAIR_KERNEL=1 timeout 120 python scripts/debug_pair_gemm.py tc1 256 > gpurun_out/dbg_tc1_256.log 2>&1
DFB_PAIR_KERNEL=1 timeout 120 python scripts/debug_pair_gemm.py tc 400000 > gpurun_out/dbg_tc_400000.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_tc_400000.log
