mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/bench_configs.py c5 --scale 0.5 2>&1 | grep '"exp"' | cut -c1-250
timeout 600 python tools/bench_configs.py c5 --scale 0.5 --opt pass1=2 2>&1 | grep '"exp"' | cut -c1-250
timeout 600 python tools/bench_configs.py c2 c4 --scale 0.5 2>&1 | grep '"exp"' | cut -c1-250
