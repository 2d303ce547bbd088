"""CPU check of kernel index arithmetic: the column-walker neighbourhood-sum / diffusion kernels and the
k-nearest ring search are extracted textually from csrc/*.cu, compiled with g++ against host shims
(tools/emu/cuda_shim.h) and compared with brute force on ragged shapes (tools/emulate_kernels.py).
Not a substitute for the device parity tests -- it catches indexing mistakes where no GPU is available."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_walker_and_knn_kernels_agree_with_brute_force_on_the_host():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emulate_kernels.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "all cases agree" in r.stdout
