"""Multi-GPU convergence sweep (run under torchrun): held-out RMSE of the user-hash sharded ordered run
on the planted-signal stream against the single-GPU ordered run, for several exchange frequencies and
both ways of combining the item-side deltas (sum / mean).  One JSON line per setting on rank 0.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/conv_sweep.py [rows]
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from svdfeature_b200 import api  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
stream = torch.cuda.Stream(device=dev)
rng = np.random.default_rng(10)
W0 = (rng.standard_normal((bench.NUM_USER + bench.NUM_ITEM, bench.K)) * 0.01).astype(np.float32)
out = open(os.path.join(ROOT, "gpurun_out", "conv_sweep_n%d.jsonl" % world), "a") if rank == 0 else None
for scale in (1.0, 0.0):  # 0 = mean
    for E in (4, 16, 64, 256):
        args = types.SimpleNamespace(convergence_rows=rows, exchanges_per_step=E, allreduce_scale=scale)
        with torch.cuda.stream(stream):
            res = bench.convergence_leg(args, api, torch, dist, dev, stream, rank, world, local, W0)
        if rank == 0:
            res["world"] = world
            line = json.dumps(res)
            print(line, flush=True)
            out.write(line + "\n")
            out.flush()
dist.destroy_process_group()
