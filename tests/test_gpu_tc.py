"""Pairwise contraction kernels through the C ABI: the SIMT kernel and the tcgen05 / TMEM 3xTF32 kernel
against numpy einsum (complex128), forced one at a time (the choice is made once per process)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["simt", "tc", "auto"])
def test_pairwise_contraction_kernels(cuda, kernel):
    env = dict(os.environ, TCB_TN_KERNEL=kernel)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tc_gemm_check.py")], env=env, capture_output=True,
                       text=True, timeout=300)  # fmt: skip
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok" in r.stdout
