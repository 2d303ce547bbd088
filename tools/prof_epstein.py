"""A few activations of the Epstein model on a side x side grid -- the target of ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dev_epstein import make
side = int(sys.argv[1]) if len(sys.argv) > 1 else 40
m = make(side)
for _ in range(6):
    m.agents.shuffle_do("step")
torch.cuda.synchronize()
