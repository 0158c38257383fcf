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
stamps = torch.zeros(48 * 16 * 8 + 16 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    stamps.zero_()
    fe.forward(x, clip_max=stamps.view(torch.float32))
torch.cuda.synchronize()
allst = stamps.cpu().numpy()
s = allst[:48 * 16 * 8].reshape(48, 16, 8)
c = allst[48 * 16 * 8:].reshape(16, 8)
print("clip stamps (start, loop_end, sync_done, fixup_done) relative to first clip start; deltas:")
for i in range(1, 6):
    print(f"  clip {i}: tiles {c[i,1]-c[i,0]:7d}  group-sync {c[i,2]-c[i,1]:6d}  fixup {c[i,3]-c[i,2]:6d}  gap-to-next {c[i+1,0]-c[i,3]:6d}  total {c[i+1,0]-c[i,0]:7d}")
nw = int((s[1, :, 2] != 0).sum())
names = ["top", "tma_ok", "s1_done", "bar1", "s2_done", "bar2", "mel_done"]
print("warps", nw)
for t in range(2, 10):
    print(f"tile {t}: tile period {s[t, 0, 2] - s[t - 1, 0, 2]} cycles (stamps: 1 wave ready, 2 S1 done, 4 S2 done, 6 mel done)")
    for w in range(nw):
        r = s[t, w, :7] - s[t, :nw, 2].min()
        print(f"  w{w:2d} " + " ".join(f"{n}={int(v):6d}" for n, v in zip(names, r)))
