"""CPU-only: every reference citation (file:line[-line]) in the interface header, the design
documents and the sources points inside an existing file of the reference tree (skipped where
/root/reference is absent)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
ALIAS = {"base.h": "solvers/base-solver/apex_svd_base.h", "model.h": "apex_svd_model.h",
         "sse.h": "apex-tensor/apex_tensor_sse.h", "data.cpp": "apex_svd_data.cpp", "data.h": "apex_svd_data.h"}
DOCS = ["include/svdgpu.h", "DESIGN.md", "INTEGRATION.md", "oracle/svdf_oracle.c", "svdfeature_b200/csrc/gpu_trainer.cpp",
        "svdfeature_b200/csrc/svdgpu_device.cuh", "svdfeature_b200/csrc/svdgpu_ordered.cu",
        "svdfeature_b200/csrc/svdgpu_rank.cu", "svdfeature_b200/csrc/svdgpu_pairs.cu",
        "svdfeature_b200/csrc/svdgpu_ingest.cu", "svdfeature_b200/buffer_io.py"]
CITE = re.compile(r"([A-Za-z0-9_\-/]+\.(?:h|cpp|py)):(\d+)(?:-(\d+))?((?:,\d+(?:-\d+)?)*)")


def _index():
    by_name = {}
    for d, _, files in os.walk(REF):
        if "node_modules" in d:
            continue
        for f in files:
            by_name.setdefault(f, []).append(os.path.join(d, f))
    return by_name


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_citations_point_into_the_reference():
    by_name = _index()
    lengths = {}
    checked, bad = 0, []
    own = {f for _, _, fs in os.walk(os.path.join(ROOT, "svdfeature_b200")) for f in fs} | {"svdgpu.h", "svdf_oracle.h"}
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for m in CITE.finditer(text):
            name = m.group(1)
            base = os.path.basename(name)
            if base in own and base not in ALIAS:
                continue  # a citation of this repo's own file
            path = None
            if name in ALIAS:
                path = os.path.join(REF, ALIAS[name])
            elif os.path.exists(os.path.join(REF, name)):
                path = os.path.join(REF, name)
            elif len(by_name.get(base, [])) >= 1:
                path = sorted(by_name[base], key=len)[0]
            else:  # an abbreviated name ("cpu_inline_common.h" for apex_tensor_cpu_inline_common.h)
                tails = sorted(f for f in by_name if f.endswith("_" + base))
                if len(tails) == 1:
                    path = by_name[tails[0]][0]
            if path is None:
                bad.append((doc, m.group(0), "no such file in the reference"))
                continue
            if path not in lengths:
                with open(path, "rb") as f:
                    lengths[path] = f.read().count(b"\n") + 1
            nums = [int(m.group(2))] + ([int(m.group(3))] if m.group(3) else [])
            nums += [int(x) for x in re.findall(r"\d+", m.group(4) or "")]
            checked += 1
            if max(nums) > lengths[path] or min(nums) < 1 or (m.group(3) and int(m.group(3)) < int(m.group(2))):
                bad.append((doc, m.group(0), "beyond the file's %d lines" % lengths[path]))
    assert checked > 150, checked
    assert not bad, bad[:20]
