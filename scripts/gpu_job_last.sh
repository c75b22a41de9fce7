#!/bin/bash
# last check of the tree: full GPU suite + default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all rc=$?" >> gpurun_out/t_all.log
tail -n 3 gpurun_out/t_all.log
timeout 600 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
tail -n 1 gpurun_out/bench_n1.log | cut -c1-300
