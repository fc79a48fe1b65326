mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 900 python tools/bench_configs.py c5 --scale 0.5 2>&1 | grep '"exp"' | cut -c1-200
