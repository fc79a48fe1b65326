"""GPU, two devices (skipped on a one-GPU box; run with `gpurun --gpus 2`): the multi-GPU path of the
C++ trainer -- one process per GPU behind the ISVDTrainer seam, every process fed the same input
(gpu:world / gpu:rank / gpu:nccl_id), rows split by user id mod world, item side all-reduced by the
library over NCCL (svdgpu_comm.cu), user side completed before the model is saved."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, id_file, out, tmp):
    try:
        import sys

        here = os.path.dirname(os.path.abspath(__file__))
        for p in (os.path.dirname(here), here):
            if p not in sys.path:
                sys.path.insert(0, p)
        import _ml100k
        from svdfeature_b200 import api

        train, test, truth, gold = _ml100k.load()
        params = dict(gold["params"])
        params.update({"gpu:device": rank, "gpu:mode": "exact", "gpu:world": world, "gpu:rank": rank,
                       "gpu:nccl_id": id_file, "gpu:allreduce_rows": 20000})
        t = api.GpuTrainer(0, 0, 0, params, bulk=(rank == 0))  # one rank per-row, one bulk: same split either way
        t.init(gold["seed"])
        for r in range(40):
            t.set_round(r)
            t.update_csr(train)  # EVERY process is handed the whole input and keeps its users' rows
            t.finish_round()
        blob = t.model_bytes(os.path.join(tmp, "r%d" % rank))  # save_model completes the user side (collective)
        p = t.predict_csr(test).astype(np.float64)
        t.close()
        out.put((rank, hashlib.sha256(blob).hexdigest(), float(np.sqrt(np.mean((p - truth) ** 2))), None))
    except Exception as e:  # pragma: no cover
        out.put((rank, None, None, repr(e)))


def test_two_processes_train_one_model_through_the_seam(native, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import _ml100k

    gold = _ml100k.load()[3]
    for r in range(2):
        os.makedirs(tmp_path / ("r%d" % r))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, str(tmp_path / "nccl.id"), out, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[3] is None for r in res), res
    # both processes hold -- and would write -- the same complete model
    assert res[0][1] == res[1][1]
    assert abs(res[0][2] - res[1][2]) < 1e-12
    # and it has learned what the sequential loop learns (0.9327 after round 40), to the band a
    # delayed item side allows
    assert abs(res[0][2] - gold["test_rmse_after_round"]["40"]) < 0.02, res
