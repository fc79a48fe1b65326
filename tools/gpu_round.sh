mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -o gpurun_out/kstream_c4 python tools/bench_configs.py c4 --scale 0.3 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
