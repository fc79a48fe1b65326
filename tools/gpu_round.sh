mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/bench_configs.py c4u --scale 0.5 2>&1 | grep '"exp"\|Error\|error' | cut -c1-400
