mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py c4u 2>&1 | grep '"exp"' | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; tail -3 gpurun_out/launches.csv | cut -c1-200
