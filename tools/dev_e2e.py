"""Dev: host-buffer GoL step (H2D + stencil + D2H) vs chunk count."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200.host_api import GolHostStepper
n = 16384
h_in = (torch.rand((n, n)) < 0.2).to(torch.uint8).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
for chunks, ramp in [(1, False), (8, False), (16, False), (32, False), (8, True), (12, True), (16, True), (24, True), (32, True)]:
    st = GolHostStepper(n, n, "cuda", chunks=chunks, ramp=ramp)
    for _ in range(2):
        st.step(h_in, h_out); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        st.step(h_in, h_out); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(json.dumps({"chunks": chunks, "ramp": ramp, "blocks": st.chunks, "ms": dt * 1e3, "GBps_each_way": n * n / dt / 1e9}))
# raw copies for reference
d = torch.empty((n, n), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h_in, non_blocking=True)
torch.cuda.synchronize(); print("h2d GB/s", n * n * 10 / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter()
for _ in range(10):
    h_out.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("d2h GB/s", n * n * 10 / (time.perf_counter() - t0) / 1e9)
