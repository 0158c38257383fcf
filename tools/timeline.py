#!/usr/bin/env python3
"""Per-phase clock stamps of CTA 0 (debug build with -DLM_TIMELINE).  Usage on the GPU box:
   python tools/timeline.py <path to liblogmel_timeline.so> [variant] [clips]

Stamp slots (LM_STAMP in the kernels): 1 waveform tile ready (before stage 1), 2 top of the tile
loop, 3 TMA copies issued, 4 stage 2 done, 5 P_FULL reached (mel warps), 6 mel done.
A stamp taken right after a barrier shows when the warp ARRIVED (BAR.SYNC defers blocking)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlx8_ws_audio_transformer_b200 import _native as N
N.LIB_PATH = sys.argv[1]
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 592
from mlx8_ws_audio_transformer_b200 import LogMelFrontend
from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank
fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, 128), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True, variant=variant)
x = torch.randn(B, 480000, device="cuda") * 0.1
stamps = torch.zeros(48 * 16 * 8 + 16 * 8 + 160 + 160 * 24, dtype=torch.int64, device="cuda")
for _ in range(2):
    stamps.zero_()
    fe.forward(x, clip_max=stamps.view(torch.float32))
torch.cuda.synchronize()
allst = stamps.cpu().numpy()
s = allst[:48 * 16 * 8].reshape(48, 16, 8)
per_cta = allst[48 * 16 * 8 + 16 * 8:48 * 16 * 8 + 16 * 8 + 160]
per_clip = allst[48 * 16 * 8 + 16 * 8 + 160:].reshape(160, 24)
if per_cta.any():
    cyc, sm = per_cta >> 12, per_cta & 4095
    live = cyc > 0
    print("whole-kernel cycles per CTA (mel leader): min %d  median %d  max %d" % (cyc[live].min(), np.median(cyc[live]), cyc[live].max()))
    order = np.argsort(cyc)
    print("  fastest CTAs (cta:smid:cycles):", " ".join(f"{i}:{sm[i]}:{cyc[i]}" for i in order if live[i])[:400])
    print("  slowest CTAs:", " ".join(f"{i}:{sm[i]}:{cyc[i]}" for i in order[::-1][:12]))
    for r in range(4):
        sel = live & (np.arange(len(cyc)) % 4 == r)
        print(f"  rank {r}: mean {cyc[sel].mean():.0f}")
if per_clip.any():
    cyc = per_cta >> 12
    order = [i for i in np.argsort(cyc) if cyc[i] > 0]
    for name, i in (("fastest", order[0]), ("median", order[len(order) // 2]), ("slowest", order[-1])):
        for c in (i - i % 4 + r for r in range(4)):
            d = np.diff(np.concatenate([[0], per_clip[c][per_clip[c] > 0]]))
            print(f"  {name} group, CTA {c} (rank {c % 4}): cycles per clip " + " ".join(str(int(v)) for v in d))
nw = int((s[1, :, 2] != 0).sum())
print("warps", nw)
names = {1: "wave_ok", 2: "top", 3: "tma_issued", 4: "s2_done", 5: "pfull", 6: "mel_done"}
top = np.where(s[:, :nw, 2] != 0, s[:, :nw, 2], np.iinfo(np.int64).max).min(axis=1)
print("tile: period | per slot: min..max over the warps that stamped it, relative to the earliest loop top of the tile")
for t in range(1, 47):
    if top[t] == np.iinfo(np.int64).max or top[t + 1] == np.iinfo(np.int64).max:
        break
    line = f"tile {t:2d}: period {top[t + 1] - top[t]:6d} |"
    for slot in (2, 3, 4, 1, 5, 6):
        v = s[t, :nw, slot]
        v = v[v != 0] - top[t]
        if len(v):
            line += f" {names[slot]} {v.min():5d}..{v.max():5d}"
    print(line)
if len(sys.argv) > 4:
    for t in (2, 20):
        print(f"tile {t} per warp:")
        for w in range(nw):
            print(f"  w{w:2d} " + " ".join(f"{names[k]}={int(s[t, w, k] - top[t]) if s[t, w, k] else -1:6d}" for k in (2, 3, 4, 1, 5, 6)))
