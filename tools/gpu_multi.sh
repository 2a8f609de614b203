#!/bin/bash
# GPU box with N GPUs (gpurun --gpus N -- bash tools/gpu_multi.sh N): multi-GPU equivalence check (tools/check_multi_gpu.py), then the bench at N
N=${1:-4}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/check_multi_gpu_n$N.log 2>&1
grep -v "^\s*$" gpurun_out/check_multi_gpu_n$N.log | grep -A8 "Traceback\|Error\|ok:" | head -30
for K in ${KS:-20 100}; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/bench_n${N}_k$K.log 2>&1
grep '^{' gpurun_out/bench_n${N}_k$K.log | tail -1 | tee gpurun_out/bench_n${N}_k$K.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('N',j['n_gpus'],'K',j['steps'],'value',round(j['value']),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value']),'sc8',round(j['e2e_sc8']['value']),'frames',j['config']['frames_per_step'],'xchg',j['config']['frame_exchange'][:40],'c5',j['c5'] and round(j['c5']['value']))"
grep -A8 "Traceback" gpurun_out/bench_n${N}_k$K.log | head -20
done
