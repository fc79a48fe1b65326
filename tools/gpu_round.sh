mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/bench_configs.py c5 --scale 0.5 2>&1 | grep '"exp"' | cut -c1-250
timeout 600 python tools/bench_configs.py c5 --scale 0.5 --opt stream_tile=64 2>&1 | grep '"exp"' | cut -c1-250
timeout 600 python tools/bench_configs.py c5 --scale 0.5 --opt stream_tile=8 2>&1 | grep '"exp"' | cut -c1-250
timeout 300 python tools/bench_rank.py 18000 4000 64 10 2>&1 | tail -1 | tee gpurun_out/bench_rank_top10.json | cut -c1-400
