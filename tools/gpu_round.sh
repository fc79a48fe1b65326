mkdir -p gpurun_out
timeout 900 python tools/hogwild_parity.py 2>&1 | tail -1 | tee gpurun_out/hogwild_parity.json
timeout 900 python tools/hogwild_parity.py 480000 10000000 2>&1 | tail -1 | tee gpurun_out/hogwild_parity_480k.json
