#!/bin/bash
# N-GPU evidence: bitwise per-instance equality across world sizes + bench at N = all visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  scripts/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_n$N.txt 2>&1
echo "rc=$?" >> gpurun_out/r02_multi_gpu_check_n$N.txt
grep -v "Warn\|warn" gpurun_out/r02_multi_gpu_check_n$N.txt | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.log
