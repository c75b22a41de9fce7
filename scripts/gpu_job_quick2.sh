#!/bin/bash
# quick A/B: timing probe, DRAM bytes of a few launches, core parity subset
mkdir -p gpurun_out
timeout 200 python scripts/probe_tc.py fwd > gpurun_out/probe.log 2>&1
DFB_GRAPH_CAPTURE=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k regex:k_edge_layer_pair -s 4 -c 4 --csv --log-file gpurun_out/quick_dram.csv python scripts/probe_tc.py fwd > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or vs_oracle or bitwise" > gpurun_out/t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/t_parity.log
grep -v Warn gpurun_out/probe.log; python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/quick_dram.csv")) if len(r)>10][1:]
for r in rows: print(r[0], r[-3], r[-1])
PY
tail -n 2 gpurun_out/t_parity.log
