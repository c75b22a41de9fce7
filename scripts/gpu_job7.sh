#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu test suite ==="
timeout 1200 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/t7_all.log
DFB_TC_PROBE=128 timeout 300 python scripts/probe_tc.py fwd 2>&1 | grep -E "probe=" | tee gpurun_out/probe7.log
echo "=== launch list (ncu, shares only) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/launches.csv python scripts/probe_tc.py fwd > gpurun_out/ncu_launches.log 2>&1
echo "=== bench ==="
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tc.log | cut -c1-400
