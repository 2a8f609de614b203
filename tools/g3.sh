mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 6 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_n1.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('N1 value',round(j['value']),'ms',round(j['ms_per_step'],4),'e2e',round(j['e2e']['value']),'cpu',j.get('cpu_baseline'),'c5',j['c5'] and (round(j['c5']['value']), j['c5']['ms_per_step'], j['c5']['config']['halo_overhead']))"
KS="20" bash tools/gpu_multi.sh 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1 | cut -c1-300
