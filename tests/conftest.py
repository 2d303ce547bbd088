"""pytest configuration: registers the `gpu` marker and shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class HostCompactionPlan:
    """Test double of ops.CompactionPlan (the stable compaction of csrc/agents.cu) on CPU tensors."""

    def __init__(self, keep):
        self.keep = keep.to(bool)
        self.n, self.kept = keep.numel(), int(self.keep.sum())

    def apply(self, col):
        return col[self.keep].clone()


def host_sort_pairs(keys, vals, begin_bit=0, end_bit=None):
    """Test double of ops.sort_pairs (the stable radix sort of csrc/sort.cu: keys compared as UNSIGNED integers)."""
    import torch

    order = torch.argsort(keys ^ (-2**63 if keys.dtype == torch.int64 else -2**31), stable=True)
    return keys[order], vals[order]


@pytest.fixture()
def host_kernels(monkeypatch):
    """CPU stand-ins for the handful of device calls the HOST-SIDE logic of the drop-in classes makes (sort, compaction,
    reduction, fill), so that this logic can be tested on CPU tensors without a GPU.  Test infrastructure only: the
    product rejects CPU tensors (tests/test_host_logic.py checks that too)."""
    import torch

    from mesa_b200 import ops

    monkeypatch.setattr(ops, "sort_pairs", host_sort_pairs)
    monkeypatch.setattr(ops, "CompactionPlan", HostCompactionPlan)
    monkeypatch.setattr(ops, "layer_fill", lambda layer, value: layer.fill_(value))
    monkeypatch.setattr(ops, "layer_reduce", lambda layer, op="sum": {"sum": torch.sum, "min": torch.min, "max": torch.max,
                                                                      "count_nonzero": torch.count_nonzero}[op](layer).reshape(1))
    return ops
