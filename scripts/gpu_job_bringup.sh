#!/bin/bash
# GPU bring-up job for the CTA-pair edge kernel: timing probe (+ phase profile), parity tests.
mkdir -p gpurun_out
export DFB_PAIR_KERNEL=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 200 python scripts/probe_tc.py all > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/probe.log
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=128 timeout 200 python scripts/probe_tc.py fwd > gpurun_out/prof_pair.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/t_parity.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -q -m gpu -rs > gpurun_out/t_full.log 2>&1
echo "full rc=$?" >> gpurun_out/t_full.log
timeout 300 python scripts/parity_margin.py > gpurun_out/margin.log 2>&1
tail -n 4 gpurun_out/probe.log gpurun_out/prof_pair.log gpurun_out/t_parity.log gpurun_out/t_full.log gpurun_out/margin.log
