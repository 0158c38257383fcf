#!/usr/bin/env python3
"""Timing of the fused 16-bit PCM ingest against the float32 path on the same clips (tuning tool).

    LM_LIB_PATH=tools/_dbg/liblogmel_X.so python tools/pcm_check.py [--clips 4096] [--mels 128] [--stereo]
"""
from __future__ import annotations

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
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--mels", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--stereo", action="store_true")
    a = ap.parse_args()
    fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, a.mels), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True)
    g = torch.Generator(device="cuda").manual_seed(3)
    shape = (a.clips, 480000, 2) if a.stereo else (a.clips, 480000)
    xi = (torch.randn(shape, device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
    out = torch.empty(a.clips, a.mels, 3000, device="cuda")

    def timed(x):
        for _ in range(3):
            fe.forward(x, out=out)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        ev[0].record()
        for i in range(a.iters):
            fe.forward(x, out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        return min(ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters))

    t_pcm = timed(xi)
    got = out[:640].clone()
    n = min(a.clips, 1184)
    xf = (xi[:n].float().sum(-1) / 65536.0) if a.stereo else (xi[:n].float() / 32768.0)
    ref = fe.forward(xf)
    same = float((ref[:640] - got[:min(640, n)]).abs().max())
    del xi
    print(f"kernel {fe.kernel_name(a.clips, 480000)}  pcm16{' stereo' if a.stereo else ''}: {t_pcm:.3f} ms   max|pcm - float path| = {same:.3g}")
    x = torch.randn(a.clips, 480000, device="cuda", generator=g) * 0.1
    print(f"float32: {timed(x):.3f} ms")


if __name__ == "__main__":
    main()
