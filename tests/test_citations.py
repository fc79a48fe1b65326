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


ANCHORS = [  # (file, first line, last line, text that must occur inside)
    ("solvers/base-solver/apex_svd_base.h", 33, 75, "class ParameterSet"),
    ("solvers/base-solver/apex_svd_base.h", 188, 210, "reg_global"),
    ("solvers/base-solver/apex_svd_base.h", 211, 250, "reg_user"),
    ("solvers/base-solver/apex_svd_base.h", 251, 283, "reg_item"),
    ("solvers/base-solver/apex_svd_base.h", 286, 311, "regularize"),
    ("solvers/base-solver/apex_svd_base.h", 313, 353, "calc_bias"),
    ("solvers/base-solver/apex_svd_base.h", 354, 381, "prepare_tmp"),
    ("solvers/base-solver/apex_svd_base.h", 383, 427, "update_no_decay"),
    ("solvers/base-solver/apex_svd_base.h", 445, 454, "float pred("),
    ("solvers/base-solver/apex_svd_base.h", 456, 462, "update_inner"),
    ("solvers/base-solver/apex_svd_base.h", 470, 478, "set_round"),
    ("solvers/base-solver/apex_svd_base.h", 512, 520, "update_svdpp"),
    ("solvers/base-solver/apex_svd_base.h", 523, 538, "prepare_ufeedback"),
    ("solvers/base-solver/apex_svd_base.h", 539, 554, "update_ufeedback"),
    ("solvers/base-solver/apex_svd_base.h", 568, 582, "update( const SVDPlusBlock"),
    ("solvers/base-solver/apex_svd_base.h", 583, 591, "predict("),
    ("solvers/base-solver/apex_svd_base.h", 666, 685, "init_ranker"),
    ("solvers/base-solver/apex_svd_base.h", 687, 710, "prepare_ifactor"),
    ("solvers/base-solver/apex_svd_base.h", 712, 717, "proc_item"),
    ("solvers/base-solver/apex_svd_base.h", 719, 740, "proc_user"),
    ("solvers/base-solver/apex_svd_base.h", 741, 749, "proc_tag"),
    ("solvers/base-solver/apex_svd_base.h", 750, 758, "proc_spec"),
    ("solvers/base-solver/apex_svd_base.h", 759, 782, "proc_rank"),
    ("solvers/base-solver/apex_svd_base.h", 795, 797, "process("),
    ("solvers/base-solver/apex_svd_base.h", 798, 812, "const SVDPlusBlock &data"),
    ("apex_svd_model.h", 112, 123, "map_active"),
    ("apex_svd_model.h", 132, 156, "cal_grad"),
    ("apex_svd_model.h", 220, 237, "calc_base_score"),
    ("apex_svd_model.h", 511, 556, "alloc_space"),
    ("apex_svd_model.h", 638, 660, "save_to_file"),
    ("apex_svd_model.h", 665, 705, "rand_init"),
    ("apex_svd.h", 33, 107, "class ISVDTrainer"),
    ("apex_svd.h", 160, 197, "class ISVDRanker"),
    ("apex_svd.cpp", 32, 47, "create_svd_trainer"),
    ("apex-tensor/apex_tensor_sse.h", 289, 317, "sdot"),
    ("apex_svd_data.cpp", 812, 1025, "class PairwiseRankGenerator"),
    ("apex_svd_data.cpp", 886, 915, "genpair"),
    ("apex_svd_data.cpp", 920, 944, "sample_cmp"),
    ("apex_svd_data.cpp", 946, 965, "sample_posneg"),
    ("svd_feature.cpp", 231, 247, "update("),
    ("apex-utils/apex_utils.h", 141, 196, "SparseFeatureArray"),
]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("path,first,last,text", ANCHORS)
def test_cited_ranges_hold_what_they_are_cited_for(path, first, last, text):
    lines = open(os.path.join(REF, path), errors="replace").read().split("\n")
    assert text in "\n".join(lines[first - 1:last]), (path, first, last, text)
