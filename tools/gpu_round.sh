mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -k "ugroup or svdpp or golden or trainer or ingest or side or pairs" 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 1200 python tools/hogwild_parity.py 20000 200000 --svdpp 2>&1 | tail -1 | tee gpurun_out/hogwild_parity_svdpp.json | cut -c1-900
timeout 900 python tools/bench_configs.py c3 --scale 0.2 2>&1 | grep '"exp"' | cut -c1-200
