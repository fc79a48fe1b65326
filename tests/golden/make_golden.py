"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden.py

Writes, next to this script:
  ua.base.buffer           make_feature_buffer on demo/basicMF/ua.base.example -- asserted
                           byte-identical to the reference's own demo/basicMF/ua.base.buffer
  ua.base.example.txt, eg.pred.txt
                           the 4-row example and the reference's shipped predictions after 40 rounds
  ref_cli_pred.txt, ref_cli_0040.model.json
                           predictions (and size + sha256 of the model file) of the reference CLI rebuilt
                           here (glibc rand), same 40-round run
  implicit.buffer.svdpp    make_ugroup_buffer on demo/implicitFeedback (group + feedback example)
  neighborhood.buffer      make_feature_buffer on demo/neighborhoodModel/ua.base.example
  hotpath_<case>.npz       inputs + model bytes + predictions of the compiled reference trainer
                           (ISVDTrainer driven through oracle/_ref/libsvdf_ref.so) on seeded cases
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")

import _cases  # noqa: E402
from _oracle import RefTrainer, build_oracle  # noqa: E402
from svdfeature_b200 import synth  # noqa: E402

build_oracle()


def run(cmd, cwd):
    subprocess.check_call(cmd, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def cli_fixtures():
    tmp = tempfile.mkdtemp()
    demo = os.path.join(REF, "demo", "basicMF")
    for f in ("ua.base.example", "ua.test.example", "basicMF.conf"):
        shutil.copy(os.path.join(demo, f), tmp)
    run([os.path.join(BIN, "make_feature_buffer"), "ua.base.example", "ua.base.buffer"], tmp)
    run([os.path.join(BIN, "make_feature_buffer"), "ua.test.example", "ua.test.buffer"], tmp)
    ours = open(os.path.join(tmp, "ua.base.buffer"), "rb").read()
    theirs = open(os.path.join(demo, "ua.base.buffer"), "rb").read()
    assert ours == theirs, "rebuilt make_feature_buffer output differs from the shipped golden buffer"
    shutil.copy(os.path.join(tmp, "ua.base.buffer"), os.path.join(HERE, "ua.base.buffer"))
    shutil.copy(os.path.join(demo, "ua.base.example"), os.path.join(HERE, "ua.base.example.txt"))
    shutil.copy(os.path.join(demo, "ua.test.example"), os.path.join(HERE, "ua.test.example.txt"))
    shutil.copy(os.path.join(demo, "eg.pred.txt"), os.path.join(HERE, "eg.pred.txt"))
    run([os.path.join(BIN, "svd_feature"), "basicMF.conf", "num_round=40"], tmp)
    run([os.path.join(BIN, "svd_feature_infer_eval"), "basicMF.conf", "pred=40"], tmp)
    shutil.copy(os.path.join(tmp, "pred.txt"), os.path.join(HERE, "ref_cli_pred.txt"))
    import hashlib
    import json

    blob = open(os.path.join(tmp, "0040.model"), "rb").read()  # 683588 bytes: keep size + digest only
    json.dump({"size": len(blob), "sha256": hashlib.sha256(blob).hexdigest()},
              open(os.path.join(HERE, "ref_cli_0040.model.json"), "w"))
    # user-group buffer
    demo = os.path.join(REF, "demo", "implicitFeedback")
    for f in ("ua.base.group.example", "ua.base.feedbackexample"):
        shutil.copy(os.path.join(demo, f), tmp)
    run([os.path.join(BIN, "make_ugroup_buffer"), "ua.base.group.example", "implicit.buffer.svdpp", "-fd",
         "ua.base.feedbackexample"], tmp)
    shutil.copy(os.path.join(tmp, "implicit.buffer.svdpp"), HERE)
    demo = os.path.join(REF, "demo", "neighborhoodModel")
    shutil.copy(os.path.join(demo, "ua.base.example"), os.path.join(tmp, "nb.example"))
    run([os.path.join(BIN, "make_feature_buffer"), "nb.example", "neighborhood.buffer"], tmp)
    shutil.copy(os.path.join(tmp, "neighborhood.buffer"), HERE)
    shutil.copy(os.path.join(demo, "ua.base.example"), os.path.join(HERE, "neighborhood.example.txt"))


def small_cases():
    nu, ni, ng = 40, 30, 12
    base = dict(num_user=nu, num_item=ni, num_factor=8, learning_rate=0.01, wd_user=0.004, wd_item=0.004,
                base_score=3.6)
    gen = dict(base, num_factor=12, wd_user_bias=0.01, wd_item_bias=0.02, num_global=ng, wd_global=0.001)
    pp = dict(base, num_ufeedback=ni, wd_ufeedback=0.004, wd_ufeedback_bias=0.001, scale_lr_ufeedback=0.5,
              ufeedback_init_sigma=0.01)
    out = {
        "basic": (0, 0, base, synth.basic_mf(300, nu, ni, seed=1), "csr"),
        "general": (0, 0, gen, synth.random_general(200, nu, ni, ng, seed=2, allow_dup=True), "csr"),
        "hinge": (0, 5, dict(gen, base_score=0.6), _cases._binary(synth.random_general(200, nu, ni, ng, seed=3)), "csr"),
        "sigmoid": (0, 2, dict(gen, base_score=0.6), _cases._binary(synth.random_general(200, nu, ni, ng, seed=4)), "csr"),
        "pairwise": (1, 3, dict(base, no_user_bias=1, base_score=0.5), _cases.as_ugroup(synth.pairwise(200, nu, ni, seed=5)), "ug"),
        "svdpp_tags": (1, 0, pp, _cases.split_tags(synth.user_grouped(300, nu, ni, avg_fb=6, seed=6)), "ug"),
    }
    return out


def hotpath_fixtures():
    tmp = tempfile.mkdtemp()
    for name, (fmt, act, params, data, kind) in small_cases().items():
        t = RefTrainer(fmt, act, 0, params)
        t.init(10)
        for r in range(2):
            t.set_round(r)
            (t.update_csr if kind == "csr" else t.update_ugroup)(data)
        pred = (t.predict_csr if kind == "csr" else t.predict_ugroup)(data)
        model = np.frombuffer(t.model_bytes(tmp), np.uint8)
        arrays = {"d%d" % i: a for i, a in enumerate(data)}
        np.savez_compressed(os.path.join(HERE, "hotpath_%s.npz" % name), fmt=fmt, act=act, kind=kind,
                            params=np.array(sorted((k, str(v)) for k, v in params.items())), model=model,
                            pred=pred, **arrays)


if __name__ == "__main__":
    cli_fixtures()
    hotpath_fixtures()
    print("golden fixtures written to", HERE)
