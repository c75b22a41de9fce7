#!/bin/bash
# timing experiments: variants/*.so against the shipped build (results of the experiment builds are wrong by construction)
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do
  echo "== shipped" >> gpurun_out/ab.log
  timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "forward" >> gpurun_out/ab.log
  for v in variants/*.so; do
    echo "== $v" >> gpurun_out/ab.log
    DFB_LIB=$PWD/$v timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep -v Warn | tail -n 3 >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
