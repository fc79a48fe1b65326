set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
cat gpurun_out/bench_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mf -s 3 -c 1 -o gpurun_out/kmf_v5_100M python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mf -s 3 -c 1 -o gpurun_out/kmf_v5_5M python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --rows-per-step 5000000 > gpurun_out/b_ncu3.log 2>&1
