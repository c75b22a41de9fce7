#!/bin/bash
# E1 ablation in the tuning build: which resource bounds the phase (results of these runs are wrong by construction)
mkdir -p gpurun_out
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
for pb in 128; do
  DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=$pb timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "pair kernel\|forward"
done > gpurun_out/e1exp.log 2>&1
cat gpurun_out/e1exp.log
