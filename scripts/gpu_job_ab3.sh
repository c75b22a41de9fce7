#!/bin/bash
# DRAM bytes of middle-layer launches with the real cache state (ncu --cache-control none), runtime switches
mkdir -p gpurun_out
for cfg in "1 0" "1 48" "1 96" "0 0"; do
  set -- $cfg
  DFB_SERPENTINE=$1 DFB_L2_KEEP_MB=$2 DFB_GRAPH_CAPTURE=0 timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:k_edge_layer_pair -s 14 -c 4 --csv --log-file gpurun_out/quick_dram_$1_$2.csv python scripts/probe_tc.py fwd > /dev/null 2>&1
done
python - <<'PY'
import csv
for s in ("1_0", "1_48", "1_96", "0_0"):
  rows=[r for r in csv.reader(open(f"gpurun_out/quick_dram_{s}.csv")) if len(r)>10][1:]
  d={}
  for r in rows: d.setdefault(r[0],{})[r[-3]]=r[-1]
  for k,v in d.items(): print("serp_keep", s, k, v)
PY
