"""GPU: the reference's own CLI drivers linked against the GPU trainer (integration/Makefile)
reproduce the reference CLI run on demo/basicMF byte for byte."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "svd_feature_gpu")


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/_build not built (needs the reference tree)")
def test_reference_cli_with_gpu_trainer(tmp_path):
    out = subprocess.run([os.path.join(ROOT, "integration", "run_demo.sh"), str(tmp_path)], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "IDENTICAL predictions" in out.stdout
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == lines[-2], "0040.model differs from the reference CLI's model file"
