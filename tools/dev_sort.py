"""Developer check + micro-benchmark of the radix sort (csrc/sort.cu): one-sweep passes against the three-launch passes
(MB_SORT_LEGACY=1 in a child process) and against torch's stable sort, sizes from one partial tile to 10^7 pairs.

    python tools/dev_sort.py            # prints one JSON line per case
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mesa_b200 import _native as N


def run_sort(keys, vals, b0, b1, reps=0):
    n = keys.numel()
    wide = keys.dtype == torch.int64
    nbytes = ctypes.c_size_t(0)
    N.check(N.lib().mb_sort_workspace_bytes(n, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    fn = N.lib().mb_sort_pairs_u64 if wide else N.lib().mb_sort_pairs_u32
    k0, v0 = keys.clone(), vals.clone()
    k1, v1 = torch.empty_like(k0), torch.empty_like(v0)
    which = ctypes.c_int(0)
    N.check(fn(N.ptr(k0), N.ptr(v0), N.ptr(k1), N.ptr(v1), n, b0, b1, N.ptr(ws), nbytes.value, ctypes.byref(which),
               N.stream_ptr(keys.device)))
    out = (k1.clone(), v1.clone()) if which.value else (k0.clone(), v0.clone())
    us = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):   # (re-sorting the partly sorted buffers: same work per pass for a radix sort)
            k0.copy_(keys); v0.copy_(vals)
            fn(N.ptr(k0), N.ptr(v0), N.ptr(k1), N.ptr(v1), n, b0, b1, N.ptr(ws), nbytes.value, ctypes.byref(which),
               N.stream_ptr(keys.device))
        e1.record()
        torch.cuda.synchronize()
        t_all = e0.elapsed_time(e1)
        e0.record()
        for _ in range(reps):
            k0.copy_(keys); v0.copy_(vals)
        e1.record()
        torch.cuda.synchronize()
        us = (t_all - e0.elapsed_time(e1)) / reps * 1e3
    return out, us


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    ok_all = True
    cases = [(1, torch.int32, 0, 32), (100, torch.int64, 0, 64), (4096, torch.int32, 0, 32), (4097, torch.int64, 32, 56),
             (70_001, torch.int32, 0, 16), (1_000_000, torch.int64, 32, 56), (1_000_000, torch.int64, 0, 64),
             (10_000_000, torch.int32, 0, 32), (10_000_000, torch.int64, 0, 64), (3_000_000, torch.int32, 8, 24)]
    for n, dt, b0, b1 in cases:
        if dt == torch.int64:
            keys = torch.randint(-2**62, 2**62, (n,), device="cuda", generator=g, dtype=torch.int64)
            if n == 1_000_000 and b0 == 32:
                keys = (torch.randint(0, 4096 * 4096, (n,), device="cuda", generator=g, dtype=torch.int64) << 32) | (keys & 0xffffffff)
        else:
            keys = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
        if n == 70_001:
            keys = keys & 0x3ff          # many equal keys: stability
        vals = torch.arange(n, device="cuda", dtype=torch.int32)
        (ko, vo), us = run_sort(keys, vals, b0, b1, reps=20 if n >= 1_000_000 else 0)
        bits = 64 if dt == torch.int64 else 32
        passes = (b1 - b0 + 7) // 8
        hi = min(bits, b0 + 8 * passes)
        # reference: stable sort on the bits the passes cover, [b0, hi)
        u = keys.to(torch.int64)
        if dt == torch.int32:
            u = u & 0xffffffff
        field = (u >> b0) & ((1 << (hi - b0)) - 1) if hi - b0 < 64 else None
        if field is None:
            field = (u ^ (-2**63))       # unsigned order of all 64 bits as signed order
        _, perm = torch.sort(field, stable=True)
        ok = bool(torch.equal(vo.long(), perm) and torch.equal(ko, keys[perm]))
        ok_all &= ok
        print(json.dumps({"n": n, "key_bits": bits, "sort_bits": [b0, b1], "ok": ok, "us": us,
                          "legacy": os.environ.get("MB_SORT_LEGACY", "0")}), flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
