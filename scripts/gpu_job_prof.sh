#!/bin/bash
# phase-cycle profile of the edge kernels (tuning build with clock64 counters, built on the box)
mkdir -p gpurun_out
export DFB_PAIR_KERNEL=1
python -m difusco_b200.build --prof --out /tmp/libprof.so > /dev/null 2>&1
DFB_LIB=/tmp/libprof.so DFB_TC_PROBE=128 timeout 200 python scripts/probe_tc.py fwd > gpurun_out/prof_pair.log 2>&1
timeout 200 python scripts/probe_tc.py all > gpurun_out/probe_pair.log 2>&1
DFB_PAIR_KERNEL=0 timeout 200 python scripts/probe_tc.py all > gpurun_out/probe_old.log 2>&1
cat gpurun_out/prof_pair.log gpurun_out/probe_pair.log gpurun_out/probe_old.log
