#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe4.log
for W in 1 2; do
  for p in 128 384; do
    DFB_TC_WPQ=$W DFB_TC_PROBE=$p timeout 300 python scripts/probe_tc.py fwd 2>&1 | grep -E "probe=" | sed "s/^/WPQ=$W /" | tee -a gpurun_out/probe4.log
  done
done
echo "=== full gpu test suite (default WPQ=2) ==="
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 | tee gpurun_out/t4_all.log
echo "=== smoke ==="
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "=== bench ==="
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_tc.log
