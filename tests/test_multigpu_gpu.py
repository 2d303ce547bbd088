"""Multi-GPU parity (identical results for every GPU count): spawns torchrun over the visible GPUs and runs
tools/check_multigpu_parity.py + the sharded check scripts.  Needs >= 2 GPUs (`gpurun --gpus N`); on a single-GPU box
the same code paths run with a 1-rank group in tests/test_sharded_gpu.py and inside bench.py (parity_ok)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, *args, timeout=900):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_parity_is_identical_for_every_gpu_count(world):
    if world not in _world_sizes():
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun("check_multigpu_parity.py", world)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("script", ["check_sharded_schelling.py", "check_sharded_bits.py", "check_sharded_layer.py"])
def test_sharded_check_scripts_two_gpus(script):
    if 2 not in _world_sizes():
        pytest.skip("needs 2 GPUs")
    args = ["64"] if script == "check_sharded_bits.py" else []
    r = _torchrun(script, 2, *args)
    assert r.returncode == 0 and ("ALL OK" in r.stdout), r.stdout[-3000:] + r.stderr[-3000:]
