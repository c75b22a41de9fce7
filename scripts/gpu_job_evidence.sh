#!/bin/bash
# round-2 evidence: ncu launch list + DRAM traffic of one denoise step, full capture of the pair kernel, phase counters + time line,
# compute-sanitizer memcheck and racecheck of the small forwards.  Nothing printed under ncu / sanitizer is a bench value.
mkdir -p gpurun_out
# (a) one denoise step of the headline workload: per-launch time + DRAM bytes (cold-cache, serialised: SHARES only)
DFB_GRAPH_CAPTURE=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -s 120 -c 80 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# (b) full capture of one full-size pair-kernel launch (middle layer)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_pair -s 14 -c 1 \
  -o gpurun_out/r02_pair_full -f python scripts/probe_tc.py fwd > gpurun_out/ncu_pair.log 2>&1
# (c) phase counters + time line of the tuning build; timing experiments (results wrong by construction)
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=128 timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "pair kernel\|forward\|trace" > gpurun_out/r02_phase_cycles.txt
{
  echo "shipped build:"; timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep forward
  for e in 2 8; do
    python -m difusco_b200.build -DDFB_EXP=$e --out /tmp/libexp$e.so > /dev/null 2>&1
    echo "DFB_EXP=$e (2 = no result stores, 8 = every CTA stores to its own first tile, L2-resident):"
    DFB_LIB=/tmp/libexp$e.so timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep forward
  done
  for pb in 1 2 3; do
    echo "tuning build, DFB_TC_PROBE=$((128 + pb)) (bit 0 = no gathers, bit 1 = no sigmoid):"
    DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=$((128 + pb)) timeout 200 python scripts/probe_tc.py fwd 2>&1 | grep "pair kernel cycles\|forward"
  done
} > gpurun_out/r02_store_gather_ablation.txt 2>&1
# (d) sanitizer
DFB_GRAPH_CAPTURE=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
DFB_GRAPH_CAPTURE=0 timeout 1200 compute-sanitizer --tool racecheck --print-limit 12 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -n 3 gpurun_out/r02_sanitizer_memcheck.log; tail -n 4 gpurun_out/r02_sanitizer_racecheck.log; grep -v trace gpurun_out/r02_phase_cycles.txt; cat gpurun_out/r02_store_gather_ablation.txt
