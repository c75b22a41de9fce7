#!/bin/bash
# quick A/B: timing probe + phase profile + core parity subset
mkdir -p gpurun_out
timeout 200 python scripts/probe_tc.py fwd > gpurun_out/probe.log 2>&1
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=128 timeout 200 python scripts/probe_tc.py fwd > gpurun_out/prof_pair.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or vs_oracle or bitwise or config" > gpurun_out/t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/t_parity.log
grep -v Warn gpurun_out/probe.log; grep "pair kernel" gpurun_out/prof_pair.log; tail -n 3 gpurun_out/t_parity.log
