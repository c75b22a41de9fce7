#!/bin/bash
mkdir -p gpurun_out
export DFB_GRAPH_CAPTURE=0
timeout 900 compute-sanitizer --tool racecheck --print-limit 12 python scripts/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
grep -c "Race reported" gpurun_out/r02_sanitizer_racecheck.log; grep "Race reported" gpurun_out/r02_sanitizer_racecheck.log | sed 's/.*in edge/edge/' | sort | uniq -c; tail -3 gpurun_out/r02_sanitizer_racecheck.log
