#!/bin/bash
# round-2 evidence: ncu launch list + DRAM traffic of one denoise step, full capture of the pair kernel,
# compute-sanitizer memcheck / racecheck of the small forwards.  Nothing printed under ncu / sanitizer is a bench value.
mkdir -p gpurun_out
export DFB_GRAPH_CAPTURE=0
# (a) one denoise step of the headline workload: per-launch time + DRAM bytes (cold-cache, serialised: SHARES only)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -s 120 -c 80 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# (b) full capture of one full-size pair-kernel launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_pair -s 14 -c 1 \
  -o gpurun_out/r02_pair_full -f python scripts/probe_tc.py fwd > gpurun_out/ncu_pair.log 2>&1
# (c) compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -n 6 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
ls -la gpurun_out | tail -8
