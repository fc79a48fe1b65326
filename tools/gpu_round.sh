mkdir -p gpurun_out
timeout 600 python tools/bench_ingest.py 20000000 2>&1 | tail -3 | tee gpurun_out/bench_ingest.json
