"""One launch of every property-layer kernel on a 16384^2 layer (after one warm launch) -- the target of
`ncu --set full -k regex:"k_box3|k_nsum|k_reduce1|k_affine|k_binary|k_fill|k_diffuse"`."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mesa_b200 import ops  # noqa: E402

if __name__ == "__main__":
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    u8 = (torch.rand((grid, grid), device=dev, generator=g) * 8).to(torch.uint8)
    u8b = torch.empty_like(u8)
    i32 = torch.empty((grid, grid), dtype=torch.int32, device=dev)
    f32 = torch.rand((grid, grid), device=dev, generator=g)
    f32b = torch.empty_like(f32)
    f64 = torch.empty((grid, grid), dtype=torch.float64, device=dev)
    d64 = torch.empty((grid, grid), dtype=torch.float64, device=dev)
    for _ in range(2):
        ops.layer_fill(u8b, 3)
        ops.layer_affine_clamp(u8, 1.0, 1.0, 0.0, 4.0, out=u8b)
        ops.layer_binary(u8, u8b, "min", out=u8b)
        ops.layer_reduce(u8, "count_nonzero")
        ops.layer_reduce(u8, "max")
        ops.neighborhood_sum(u8, 1, True, False, True, out=i32)
        ops.neighborhood_sum(u8, 2, False, False, False, out=i32)
        ops.layer_affine_clamp(f32, 0.9, 0.01, 0.0, 1.0, out=f32b)
        ops.layer_reduce(f32, "sum")
        ops.layer_diffuse(f32, 0.3, True, out=f32b)
        ops.neighborhood_sum(f32, 1, True, False, True, out=f64)
        ops.layer_diffuse(f64, 0.3, True, out=d64)
        torch.cuda.synchronize()
    print("ok")
