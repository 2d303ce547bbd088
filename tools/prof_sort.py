"""One radix sort of 10^7 (uint32 key, uint32 value) pairs -- the target of `ncu --set full -k regex:sort_onesweep`."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import _native as N

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
wide = len(sys.argv) > 2 and sys.argv[2] == "u64"
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", generator=g, dtype=torch.int64)
if not wide:
    keys = keys.to(torch.int32)
else:
    keys = keys << 32
vals = torch.arange(n, device="cuda", dtype=torch.int32)
nbytes = ctypes.c_size_t(0)
N.check(N.lib().mb_sort_workspace_bytes(n, ctypes.byref(nbytes)))
ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
which = ctypes.c_int(0)
fn = N.lib().mb_sort_pairs_u64 if wide else N.lib().mb_sort_pairs_u32
for _ in range(2):
    N.check(fn(N.ptr(keys), N.ptr(vals), N.ptr(k1), N.ptr(v1), n, 32 if wide else 0, 56 if wide else 32, N.ptr(ws), nbytes.value,
               ctypes.byref(which), N.stream_ptr(keys.device)))
torch.cuda.synchronize()
print("ok")
