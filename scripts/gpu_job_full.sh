#!/bin/bash
# full GPU test suite + timing probe
mkdir -p gpurun_out
timeout 200 python scripts/probe_tc.py fwd > gpurun_out/probe.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all rc=$?" >> gpurun_out/t_all.log
grep -v Warn gpurun_out/probe.log; tail -n 4 gpurun_out/t_all.log
