"""The drop-in boundary without torch: tests/abi_plain_run.py drives libtinycarlo_b200.so through ctypes with device memory from
libcudart alone, in a fresh interpreter that must never import torch, and checks the results against the CPU oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_c_abi_without_torch():
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "abi_plain_run.py")], capture_output=True, text=True, timeout=600)
    if r.returncode == 77:
        pytest.skip(r.stdout.strip())
    assert r.returncode == 0 and "OK: C ABI driven without torch" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
