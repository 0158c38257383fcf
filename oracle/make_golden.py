#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from the LIVE reference libraries.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where ``transformers`` and
``torchaudio`` (the libraries that hold the reference's arithmetic for this path, see
oracle/logmel_oracle.py) are importable.  The reference has no golden vectors of its own
(SURVEY.md §4), so these outputs of its actual call --
``WhisperFeatureExtractor.__call__`` as at /root/reference/AB/fineTune.py:88 and
``MelSpectrogram`` + ``torch.log(mel + 1e-6)`` as at /root/reference/.charles/spectrogram.py:161-162
-- are what pins the oracle and the CUDA path.  The GPU box has no /root/reference and may
have other library versions, so the tests read these files, never the libraries.

    python oracle/make_golden.py          # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mlx8_ws_audio_transformer_b200 import synth  # noqa: E402  (input generators only)

OUT = os.path.join(ROOT, "tests", "golden")
SHORT = 16000       # 1 s containers keep the fixtures small: 100 frames per clip
SLICE = 37          # full 30 s clips are stored as every 37th frame


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def short_inputs():
    """one-second clips covering the edge cases of SURVEY.md §8d"""
    n = SHORT
    g = synth.gaussian_clips(2, n, seed=0)
    clips = {
        "gauss0": g[0], "gauss1": g[1],
        "sine440": synth.sine_clip(440.0, n), "sine7k": synth.sine_clip(7000.0, n),
        "chirp": synth.chirp_clip(20.0, 8000.0, n),
        "impulse_first": synth.impulse_clip(0, n), "impulse_last": synth.impulse_clip(n - 1, n),
        "zeros": np.zeros(n, np.float32),
        "int16": synth.int16_uniform_clip(n, seed=1),
        "len1": np.array([0.25], np.float32),                       # zero padded by the extractor
        "len399": synth.gaussian_clips(1, 399, seed=3)[0],
        "len16001": synth.gaussian_clips(1, n + 1, seed=4)[0],      # truncated by the extractor
        "piano": synth.midi_piano_clips(1, seed=0, n_samples=n)[0][0],
        "quiet": (synth.gaussian_clips(1, n, seed=5)[0] * 1e-3).astype(np.float32),
    }
    return clips


def main():
    import torch
    import torchaudio
    import transformers
    from transformers import WhisperFeatureExtractor

    os.makedirs(OUT, exist_ok=True)
    versions = f"transformers {transformers.__version__}; torch {torch.__version__}; torchaudio {torchaudio.__version__}"
    print(versions)

    # ---- Whisper, short containers --------------------------------------------------------
    clips = short_inputs()
    names = sorted(clips)
    save = {"versions": np.array(versions), "names": np.array(names)}
    for nm in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=nm)
        out = fe([clips[k] for k in names], sampling_rate=16000, max_length=SHORT, return_tensors="np")
        save[f"feat{nm}"] = out["input_features"].astype(np.float32)
        save[f"fbank{nm}"] = fe.mel_filters.astype(np.float64)
    for k in names:
        save[f"in_{k}"] = clips[k]
    np.savez_compressed(os.path.join(OUT, "whisper_short.npz"), **save)

    # ---- Whisper, full 30 s containers (config 1 seeds), stored as strided frame slices ------
    x = synth.gaussian_clips(3, synth.WHISPER_SAMPLES, seed=0)
    piano, plen = synth.midi_piano_clips(2, seed=0)
    full = np.concatenate([x, piano, synth.sine_clip(440.0)[None], synth.chirp_clip()[None]])
    save = {"versions": np.array(versions), "slice": np.array(SLICE), "input_digest": np.array(digest(full)),
            "desc": np.array("gaussian seed0 clips 0-2; midi_piano seed0 clips 0-1; sine 440; chirp 20-8000")}
    for nm in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=nm)
        out = fe(list(full), sampling_rate=16000, return_tensors="np")["input_features"]
        save[f"feat{nm}"] = np.ascontiguousarray(out[:, :, ::SLICE]).astype(np.float32)
        save[f"max{nm}"] = out.reshape(len(full), -1).max(axis=1)
        save[f"mean{nm}"] = out.reshape(len(full), -1).mean(axis=1, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "whisper_30s.npz"), **save)

    # ---- torchaudio frontend (.charles/spectrogram.py), 4 s clips --------------------------
    w, lengths = synth.urbansound_clips(6, seed=0)
    w[5] = 0.0
    save = {"versions": np.array(versions), "lengths": lengths, "input_digest": np.array(digest(w))}
    for hop, nm in ((512, 128), (128, 128), (512, 64)):
        ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, hop_length=hop, n_mels=nm,
                                                  f_min=0, f_max=8000, power=2.0)
        mel = ms(torch.from_numpy(w))
        save[f"mel_{hop}_{nm}"] = mel.numpy()[:3]                         # raw mel power, 3 clips
        save[f"logmel_{hop}_{nm}"] = torch.log(mel + 1e-6).numpy()
        save[f"fb_{hop}_{nm}"] = ms.mel_scale.fb.numpy()
    np.savez_compressed(os.path.join(OUT, "torchaudio_4s.npz"), **save)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
