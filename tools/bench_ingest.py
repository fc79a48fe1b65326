"""GPU experiment: training instances/s from a BINARY_BUFFER file (page cache) through
svdgpu_update_buffer_file, beside the same rows through svdgpu_update_csr (pinned arrays)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svdfeature_b200 import api, buffer_io  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
data = bench.gen_rows_numpy(N, seed=10)
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(d, "train.buffer")
t0 = time.perf_counter()
buffer_io.write_feature_buffer(path, data, batch_size=1000)
print("wrote %s: %.1f MB in %.1f s" % (path, os.path.getsize(path) / 1e6, time.perf_counter() - t0), file=sys.stderr)
g = api.SvdGpu(bench.NUM_USER, bench.NUM_ITEM, bench.K)
g.set_hparams(**bench.HP)
g.set_mode(api.MODE_HOGWILD)
rng = np.random.default_rng(10)
rows = bench.NUM_USER + bench.NUM_ITEM
g.upload(np.zeros(rows, np.float32), (rng.standard_normal((rows, bench.K)) * 0.01).astype(np.float32), np.zeros(1, np.float32))
res = {"rows": N, "file_mb": os.path.getsize(path) / 1e6}
for rep in range(3):
    t0 = time.perf_counter()
    n = g.update_buffer_file(path)
    g.sync()
    dt = time.perf_counter() - t0
    assert n == N
    res["file_pass_%d_minst_s" % rep] = N / dt / 1e6
    res["file_pass_%d_read_call_ms" % rep] = (g.counter("ingest_read_us") / 1e3, g.counter("ingest_call_us") / 1e3)
t0 = time.perf_counter()
s, c = g.eval_buffer_file(path)
res["file_eval_minst_s"] = N / (time.perf_counter() - t0) / 1e6
res["rmse"] = float(np.sqrt(s / c))
os.unlink(path)
os.rmdir(d)
print(json.dumps(res))
