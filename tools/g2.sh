mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zb_rx -s 4 -c 1 -o gpurun_out/prof_zbrx -f \
    python bench.py --workload zb_wb16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_zbrx.log 2>&1
echo "full capture rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_zb_wb16_t40.csv python bench.py --workload zb_wb16 --tiles 40 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
