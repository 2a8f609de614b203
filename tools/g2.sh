mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zb_rx -s 4 -c 1 -o gpurun_out/prof_zbrx -f \
    python bench.py --workload zb_wb16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_zbrx.log 2>&1
echo "full capture rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_time" 2>&1 | tail -5
ls -la gpurun_out
