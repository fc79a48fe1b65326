mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -k "ugroup or svdpp" 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 900 python tools/bench_configs.py c3 --scale 0.5 2>&1 | grep '"exp"' | cut -c1-200
timeout 1200 python tools/hogwild_parity.py 20000 200000 --svdpp 2>&1 | tail -1 | tee gpurun_out/hogwild_parity_svdpp.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print([round(e['rmse_pred_vs_sequential'], 4) for e in d['epochs']], d['gpu_epoch_s'])"
