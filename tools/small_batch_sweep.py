#!/usr/bin/env python3
"""Small batches: thread-per-frame kernel with a clip spread over several CTAs against the CTA-tiled kernel.

    python tools/small_batch_sweep.py [mels]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mlx8_ws_audio_transformer_b200 import LogMelFrontend
from mlx8_ws_audio_transformer_b200 import _native as N
from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank


def front(nm, slices):
    os.environ["LM_TF_SLICES"] = "1" if slices else "0"
    fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True)
    os.environ.pop("LM_TF_SLICES")
    return fe


def timeit(fe, x, out, iters=20):
    for _ in range(5):
        fe.forward(x, out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fe.forward(x, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


nm = int(sys.argv[1]) if len(sys.argv) > 1 else 80
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(160, 480000, generator=g, device="cuda") * 0.1
x[3] *= 1e-4
x[3, 5000:5400] = 0.8                      # a clip whose max-8 clamp bites in every tile
out = torch.empty(160, nm, 3000, device="cuda")
f1, f0 = front(nm, True), front(nm, False)
for B in (1, 2, 4, 8, 16, 32, 48, 64, 74, 75, 90, 98, 120):
    a = f1.forward(x[:B]).clone()
    b = f0.forward(x[:B])
    t1, t0 = timeit(f1, x[:B], out[:B]), timeit(f0, x[:B], out[:B])
    print(f"batch {B:4d}: {f1.kernel_name(B, 480000)[4:30]:26s} {t1 * 1e3:7.1f} us   slices off: {f0.kernel_name(B, 480000)[4:30]:26s} {t0 * 1e3:7.1f} us"
          f"   max|diff| {float((a - b).abs().max()):.2e}")
