#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo.txt
for CH in 4 1; do
SNRX_NCCL_CHANNELS=$CH timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_n${N}_ch$CH.log 2>&1
grep '^{' gpurun_out/bench_n${N}_ch$CH.log | tail -1 | tee gpurun_out/bench_n${N}_ch$CH.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('N',j['n_gpus'],'K',j['steps'],'value',round(j['value']),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value']),'sc8',round(j['e2e_sc8']['value']))"
grep -A8 "Traceback" gpurun_out/bench_n${N}_ch$CH.log | head -20
done
