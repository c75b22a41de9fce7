#!/bin/bash
# GPU bring-up job for the CTA-pair edge kernel: GEMM1 parity, forward parity, timing probe.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gemm" > gpurun_out/t_gemm.log 2>&1
echo "gemm rc=$?" >> gpurun_out/t_gemm.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or vs_oracle or bitwise" > gpurun_out/t_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/t_parity.log
timeout 300 python scripts/probe_tc.py all > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/probe.log
tail -5 gpurun_out/t_gemm.log gpurun_out/t_parity.log gpurun_out/probe.log
