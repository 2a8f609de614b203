#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for K in 20 200; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/bench_n${N}_k$K.log 2>&1
grep '^{' gpurun_out/bench_n${N}_k$K.log | tail -1 | tee gpurun_out/bench_n${N}_k$K.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('N',j['n_gpus'],'K',j['steps'],'value',round(j['value']),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value']),'frames',j['config']['frames_per_step'])"
done
timeout 150 python bench.py --steps 200 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('N',j['n_gpus'],'K',j['steps'],'value',round(j['value']),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value']))"
