#!/bin/bash
# First GPU bring-up: isolate stages in separate processes so a tcgen05 fault cannot mask the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== fp32-path parity (validation kernel) ===" 
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32" -x 2>&1 | tail -15 | tee gpurun_out/t_fp32.log
echo "=== tcgen05 GEMM building block ==="
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tc_gemm" 2>&1 | tail -25 | tee gpurun_out/t_gemm.log
echo "=== full tc parity ==="
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not fp32 and not tc_gemm" 2>&1 | tail -40 | tee gpurun_out/t_tc.log
echo "=== smoke ==="
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench fp32 impl (1 step) ==="
DFB_EDGE_IMPL=fp32 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_fp32.log
echo "=== bench tc ==="
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_tc.log
