#!/usr/bin/env python3
"""Per-phase clock stamps of CTA 0 (debug build with -DLM_TIMELINE).  Usage on the GPU box:
   python tools/timeline.py <path to liblogmel_timeline.so> [variant]"""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlx8_ws_audio_transformer_b200 import _native as N
N.LIB_PATH = sys.argv[1]
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
from mlx8_ws_audio_transformer_b200 import LogMelFrontend
from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank
fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, 128), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True, variant=variant)
B = 592
x = torch.randn(B, 480000, device="cuda") * 0.1
stamps = torch.zeros(48 * 16 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    stamps.zero_()
    fe.forward(x, clip_max=stamps.view(torch.float32))
torch.cuda.synchronize()
s = stamps.cpu().numpy().reshape(48, 16, 8)
nw = int((s[1, :, 1] != 0).sum())
base = s[0, :nw, 0].min()
names = ["top", "tma_ok", "s1_done", "bar1", "s2_done", "bar2", "mel_done"]
print("warps", nw)
for t in range(2, 10):
    print(f"tile {t}: tile period {s[t, 0, 0] - s[t - 1, 0, 0]} cycles")
    for w in range(nw):
        r = s[t, w, :7] - s[t, :nw, 0].min()
        print(f"  w{w:2d} " + " ".join(f"{n}={int(v):6d}" for n, v in zip(names, r)))
