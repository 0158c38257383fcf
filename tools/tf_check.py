#!/usr/bin/env python3
"""GPU check of the thread-per-frame kernel (variant 3) against the oracle and the CTA-tiled kernels.

    python tools/tf_check.py [--clips 4096] [--mels 128] [--no-parity] [--iters 10]

Prints parity of variant 3 on the edge-case batch (oracle = checker) and CUDA-event times of
variants 3 and 0-tiled on the benchmark input.  Tuning / bring-up tool, not a test.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from mlx8_ws_audio_transformer_b200 import LogMelFrontend, synth
from mlx8_ws_audio_transformer_b200 import _native as N
from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank


def front(nm, variant):
    return LogMelFrontend(400, 160, slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True,
                          variant=variant)


def parity():
    from oracle import logmel_oracle as O
    n = 480000
    clips = [synth.gaussian_clips(1, n, seed=21)[0], synth.sine_clip(7000.0), synth.impulse_clip(0),
             synth.impulse_clip(n - 1), synth.int16_uniform_clip(), np.zeros(n, np.float32),
             synth.midi_piano_clips(1, seed=5)[0][0], (synth.gaussian_clips(1, n, seed=22)[0] * 30).astype(np.float32),
             synth.chirp_clip(), (synth.gaussian_clips(1, n, seed=23)[0] * 1e-4).astype(np.float32),
             np.full(n, 0.3, np.float32)]
    x = np.stack(clips)
    xt = torch.from_numpy(x).cuda()
    lengths = torch.tensor([480000, 100000, 5, 480000, 333333, 0, 123456, 480000, 7777, 160 * 32 * 3, 1], dtype=torch.int32)
    ok = True
    for nm in (80, 128):
        ref = O.whisper_logmel(x, n_mels=nm)
        f3, f2 = front(nm, 3), front(nm, 2)
        cm = torch.empty(len(clips), device="cuda")
        got = f3.forward(xt, clip_max=cm)
        torch.cuda.synchronize()
        g2 = f2.forward(xt)
        for i in range(len(clips)):
            mx, mean = O.parity(got[i].cpu().numpy(), ref[i])
            d = float((got[i] - g2[i]).abs().max())
            flag = "" if (mx < 1e-3 and mean < 1e-5 and np.isfinite(got[i].cpu().numpy()).all()) else "  <-- FAIL"
            ok = ok and not flag
            print(f"mels {nm} clip {i:2d}: vs oracle max {mx:.2e} mean {mean:.2e}; vs variant 2 max {d:.2e}{flag}")
        cmr = (ref.reshape(len(clips), -1).max(axis=1) * 4 - 4)
        print(f"   clip_max err {np.abs(cm.cpu().numpy() - cmr).max():.2e}")
        # lengths: dirty tails must be ignored
        dirty = xt.clone()
        xz = x.copy()
        for i, L in enumerate(lengths.tolist()):
            dirty[i, L:] = 7.0
            xz[i, L:] = 0.0
        refl = O.whisper_logmel(xz, n_mels=nm)
        gl = f3.forward(dirty, lengths=lengths.cuda())
        mx, mean = O.parity(gl.cpu().numpy(), refl)
        flag = "" if (mx < 1e-3 and mean < 1e-5) else "  <-- FAIL"
        ok = ok and not flag
        print(f"mels {nm} lengths batch: max {mx:.2e} mean {mean:.2e}{flag}")
    print("PARITY", "OK" if ok else "FAILED")
    return ok


def timeit(f, x, out, iters):
    for _ in range(3):
        f.forward(x, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f.forward(x, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--mels", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-tiled", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="tf vs tiled kernel over batch sizes (dispatch threshold)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ok = True
    if not args.no_parity:
        ok = parity()
    if args.sweep:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(2400, 480000, generator=g, device="cuda") * 0.1
        out = torch.empty(2400, args.mels, 3000, device="cuda")
        f3 = front(args.mels, 3)
        os.environ["LM_TF_MIN_BATCH"] = "1000000000"
        f0 = front(args.mels, 0)
        os.environ.pop("LM_TF_MIN_BATCH")
        for b in (1, 8, 32, 74, 148, 222, 296, 444, 592, 888, 1184, 1776, 2368):
            t3 = timeit(f3, x[:b], out[:b], args.iters)[0]
            t0 = timeit(f0, x[:b], out[:b], args.iters)[0]
            print(f"batch {b:5d}: thread-per-frame {t3 * 1e3:8.1f} us   CTA-tiled {t0 * 1e3:8.1f} us   ratio {t0 / t3:.2f}")
        sys.exit(0)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(args.clips, 480000, generator=g, device="cuda") * 0.1
    out = torch.empty(args.clips, args.mels, 3000, device="cuda")
    bytes_ = args.clips * (480000 * 4 + args.mels * 3000 * 4)
    for name, variant, env in (("thread-per-frame (variant 3)", 3, None), ("CTA-tiled (variant 0, tf disabled)", 0, "1000000000")):
        if variant == 0 and args.no_tiled:
            continue
        if env:
            os.environ["LM_TF_MIN_BATCH"] = env
        f = front(args.mels, variant)
        med, best = timeit(f, x, out, args.iters)
        print(f"{name}: {args.clips} clips x {args.mels} mels: median {med:.3f} ms best {best:.3f} ms -> "
              f"{args.clips / med * 1e3:.0f} clips/s, {bytes_ / med / 1e6:.0f} GB/s = {bytes_ / med / 1e6 / 6553.6:.3f} of HBM roofline")
        print(f"   checksum {float(out[::97].double().sum()):.6f}")
        os.environ.pop("LM_TF_MIN_BATCH", None)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
