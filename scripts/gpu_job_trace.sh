#!/bin/bash
# tuning: shipped build's probe timing, time line + phase counters of the pair kernel (--prof build), core parity subset
mkdir -p gpurun_out
timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "forward\|checksum" > gpurun_out/trace.log
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=128 timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "pair kernel\|forward\|trace" >> gpurun_out/trace.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or vs_oracle or bitwise" > gpurun_out/t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/t_parity.log
grep -v "trace" gpurun_out/trace.log; tail -n 3 gpurun_out/t_parity.log
