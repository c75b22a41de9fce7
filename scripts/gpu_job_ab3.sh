#!/bin/bash
# A/B variants + DRAM bytes with the real cache state (ncu --cache-control none)
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do
  echo "== shipped" >> gpurun_out/ab.log
  timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "forward" >> gpurun_out/ab.log
  for v in variants/*.so; do
    echo "== $v" >> gpurun_out/ab.log
    DFB_LIB=$PWD/$v timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep -v Warn | tail -n 1 >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
for v in shipped variants/*.so; do
  L=""; [ "$v" != shipped ] && L=$PWD/$v
  DFB_LIB=$L DFB_GRAPH_CAPTURE=0 timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:k_edge_layer_pair -s 14 -c 4 --csv --log-file gpurun_out/quick_dram.csv python scripts/probe_tc.py fwd > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, sys
rows=[r for r in csv.reader(open("gpurun_out/quick_dram.csv")) if len(r)>10][1:]
d={}
for r in rows: d.setdefault(r[0],{})[r[-3]]=float(r[-1])
for k,v in d.items(): print(sys.argv[1], k, "rd %.0f MB wr %.0f MB t %.1f us" % (v['dram__bytes_read.sum']/1e6, v['dram__bytes_write.sum']/1e6, v['gpu__time_duration.sum']/1e3))
PY
done
