#!/usr/bin/env python3
"""time the forward on an input where every tile needs the max-8 pass (tuning tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools.tf_check import front, timeit
B = 4096
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1000)
x = torch.empty(B, 480000, device=dev).normal_(0.0, 1e-5, generator=g)
tt = torch.arange(400, device=dev) / 16000.0
pos = torch.randint(0, 480000 - 400, (B,), generator=g, device=dev)
x[torch.arange(B, device=dev)[:, None], pos[:, None] + torch.arange(400, device=dev)[None, :]] += (0.9 * torch.sin(2 * np.pi * 440.0 * tt))[None, :]
out = torch.empty(B, 128, 3000, device=dev)
f = front(128, 0)
print("clamp-everywhere 4096x128: median %.3f ms best %.3f ms" % timeit(f, x, out, 5))
