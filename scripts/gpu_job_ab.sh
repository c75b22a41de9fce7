#!/bin/bash
# A/B of tuning variants built into variants/*.so (shipped build first), then the core parity subset on the shipped build
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do
  echo "== shipped" >> gpurun_out/ab.log
  timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep -v Warn | tail -n 3 >> gpurun_out/ab.log
  for v in variants/*.so; do
    [ -f "$v" ] || continue
    echo "== $v" >> gpurun_out/ab.log
    DFB_LIB=$PWD/$v timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep -v Warn | tail -n 3 >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or vs_oracle or bitwise or gemm" > gpurun_out/t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/t_parity.log
tail -n 3 gpurun_out/t_parity.log
