"""CPU-only: the C-ABI library loads and exports exactly what include/svdgpu.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "svdgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svdgpu_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree(native):
    assert header_symbols() == sorted(native.SVDGPU_SYMBOLS)


def test_library_exports_every_declared_symbol(native):
    lib = native.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_trainer_library_exports_seam(native):
    lib = native.load_trainer_library()
    for name in ("svdtr_create", "svdtr_set_param", "svdtr_init_model", "svdtr_init_trainer",
                 "svdtr_update_csr", "svdtr_predict_csr", "svdtr_update_ugroup", "svdtr_predict_ugroup",
                 "svdtr_save_model", "svdtr_load_model", "svdtr_set_round", "svdtr_finish_round"):
        assert hasattr(lib, name), name
    # the C++ seam itself: apex_svd::create_svd_trainer(apex_svd::SVDTypeParam)
    raw = ctypes.CDLL(native.LIB_TRAINER)
    assert hasattr(raw, "_ZN8apex_svd18create_svd_trainerENS_12SVDTypeParamE")


def test_no_cpu_fallback_without_gpu(native):
    """On a box without CUDA the product refuses to create a trainer."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.SvdGpuError):
        native.SvdGpu(10, 10, 8)


def test_product_does_not_link_the_oracle(native):
    import subprocess

    for so in (native.LIB_GPU, native.LIB_TRAINER):
        needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
        assert "oracle" not in needed and "svdf_ref" not in needed
    for root, _, files in os.walk(os.path.join(ROOT, "svdfeature_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                for pat in ("liboracle.so", "import _oracle", "from _oracle", '#include "svdf_oracle'):
                    assert pat not in src, (f, pat)
