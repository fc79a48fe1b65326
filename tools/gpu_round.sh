mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_abi.py -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
