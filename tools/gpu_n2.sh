#!/bin/bash
# GPU box with N >= 2 GPUs: multi-GPU equivalence check, then the bench at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/check_multi_gpu_n$N.log 2>&1
grep -v "^\s*$" gpurun_out/check_multi_gpu_n$N.log | grep -B2 -A12 "Traceback\|Error\|ok:" | head -60
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-500
grep -B2 -A12 "Traceback" gpurun_out/bench_n$N.log | head -40
