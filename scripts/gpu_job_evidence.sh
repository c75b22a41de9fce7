#!/bin/bash
# round-2 evidence: ncu launch list + DRAM traffic of one denoise step, full capture of the pair kernel, phase counters,
# compute-sanitizer memcheck of the small forwards.  Nothing printed under ncu / sanitizer is a bench value.
mkdir -p gpurun_out
# (a) one denoise step of the headline workload: per-launch time + DRAM bytes (cold-cache, serialised: SHARES only)
DFB_GRAPH_CAPTURE=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -s 120 -c 80 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# (b) full capture of one full-size pair-kernel launch (middle layer)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_pair -s 14 -c 1 \
  -o gpurun_out/r02_pair_full -f python scripts/probe_tc.py fwd > gpurun_out/ncu_pair.log 2>&1
# (c) phase counters of the tuning build
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=128 timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "pair kernel\|forward" > gpurun_out/r02_phase_cycles.txt
# (d) memcheck
DFB_GRAPH_CAPTURE=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -n 3 gpurun_out/r02_sanitizer_memcheck.log; cat gpurun_out/r02_phase_cycles.txt
