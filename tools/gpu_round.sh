mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
grep -E "rank|value" gpurun_out/bench_2gpu.err | tail; python -c "
import json; l=json.load(open('gpurun_out/bench_2gpu.json')); print(l['value']/1e9, l['ms_per_step'], l['roofline']['launch_ms'])"
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --rows-per-step 5000000 > gpurun_out/bench_2gpu_5M.json 2> gpurun_out/bench_2gpu_5M.err
grep -E "rank [01]\]|NVLS|P2P|SHM|via" gpurun_out/bench_2gpu_5M.err | tail -12
