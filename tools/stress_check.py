#!/usr/bin/env python3
"""Race hunt: the same launch repeated many times must give bit-identical features every time (float32 and 16-bit PCM
input, thread-per-frame and CTA-tiled kernels, with ragged lengths).  Tuning / bring-up tool, not a test.

    python tools/stress_check.py [--iters 100]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from mlx8_ws_audio_transformer_b200 import LogMelFrontend
from mlx8_ws_audio_transformer_b200 import _native as N
from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(11)
    bad = 0
    for nm, B, T in ((128, 1500, 480000), (80, 700, 480000), (80, 40, 480000), (128, 900, 100000)):
        fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True)
        x = torch.randn(B, T, generator=g, device="cuda") * 0.1
        x[::7] *= 1e-4                                   # quiet clips: the max-8 pass has work to do
        x[::7, 1000:1400] = 0.9
        lengths = torch.randint(0, T + 1, (B,), generator=g, device="cuda", dtype=torch.int32)
        xi = (x * 32768).clamp_(-32768, 32767).to(torch.int16)
        for name, inp, ln in (("float32", x, None), ("float32+lengths", x, lengths), ("pcm16", xi, None), ("pcm16+lengths", xi, lengths)):
            first = fe.forward(inp, lengths=ln).clone()
            out = torch.empty_like(first)
            diff = 0
            for _ in range(a.iters):
                fe.forward(inp, lengths=ln, out=out)
                diff += int((out != first).sum())
            torch.cuda.synchronize()
            bad += diff
            print(f"{fe.kernel_name(B, T):50s} {nm:3d} mels B={B:5d} T={T} {name:16s}: {a.iters} launches, mismatching values {diff}")
    print("STRESS OK" if bad == 0 else "STRESS FAILED")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
