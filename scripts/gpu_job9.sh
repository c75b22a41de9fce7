#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu test suite ==="
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee gpurun_out/t9_all.log
DFB_TC_PROBE=128 timeout 300 python scripts/probe_tc.py fwd 2>&1 | grep -E "probe=" | tee gpurun_out/probe9.log
echo "=== bench ==="
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_tc.log | cut -c1-300
echo "=== ncu launch list on the denoise loop ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "=== ncu full, one edge-kernel launch at full size ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_tc16w -s 30 -c 1 -o gpurun_out/edge_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -4
