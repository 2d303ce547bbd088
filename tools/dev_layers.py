"""Time the property-layer kernels (bench.layer_workloads) on one GPU; one JSON object per line."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else bench.GRID
    for row in bench.layer_workloads(torch.device("cuda:0"), grid=grid):
        print(json.dumps(row), flush=True)
