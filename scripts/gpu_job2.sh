#!/bin/bash
mkdir -p gpurun_out
echo "=== mean/max fix check ==="
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mis_golden" 2>&1 | tail -3 | tee gpurun_out/t_mis.log
echo "=== probes ==="
for p in 0 32 1 2 64 4 8 16 127; do
  DFB_TC_PROBE=$p timeout 300 python scripts/probe_tc.py all 2>&1 | grep -E "probe=" | tee -a gpurun_out/probe.log
done
echo "=== ncu full on one edge-kernel launch (B=4 to keep replay short) ==="
PROBE_B=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_tc -s 13 -c 2 \
   -o gpurun_out/edge_v1 python scripts/probe_tc.py fwd > gpurun_out/ncu_edge.log 2>&1
tail -3 gpurun_out/ncu_edge.log
ls -la gpurun_out | tail -5
