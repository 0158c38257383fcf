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
    make_refwav(versions)
    make_configs(versions)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


REFWAV_FRAMES = 32000      # first 2 s of every reference WAV (keeps the fixture at a few MB)
REFWAV_STRIDE = 3


def make_refwav(versions: str):
    """The 18 real-audio fixtures the reference ships (/root/reference/.charles/samples/**/*.wav: 16 kHz,
    stereo, s16), through the reference's own two calls:
      * every channel as a clip and the mono mix through WhisperFeatureExtractor (the [2, N] array of
        /root/reference/AB/wavToWhisper.py:52-55 is a batch of two clips, feature_extraction_whisper.py:274-279);
      * the mono mix (waveform.mean(dim=0), /root/reference/.charles/spectrogram.py:147-148), zero padded to
        4 s (:152-157), through torchaudio MelSpectrogram + torch.log(mel + 1e-6) (:161-162).
    The decoded PCM is stored as int16 so that the fused int16 / stereo ingest is tested on it too."""
    import glob
    import wave

    import torch
    import torchaudio
    from transformers import WhisperFeatureExtractor

    files = sorted(glob.glob("/root/reference/.charles/samples/**/*.wav", recursive=True))
    if not files:
        print("no /root/reference: refwav.npz left as it is")
        return
    pcm = np.zeros((len(files), REFWAV_FRAMES, 2), np.int16)
    names = []
    for i, f in enumerate(files):
        with wave.open(f) as w:
            assert w.getnchannels() == 2 and w.getsampwidth() == 2 and w.getframerate() == 16000
            raw = np.frombuffer(w.readframes(REFWAV_FRAMES), dtype="<i2").reshape(-1, 2)
        pcm[i, :len(raw)] = raw
        names.append(os.path.relpath(f, "/root/reference/.charles/samples"))
    chan = (pcm.astype(np.float32) / 32768.0)                              # torchaudio.load's normalisation
    mono = chan.mean(axis=2).astype(np.float32)                            # == (l + r) / 65536 exactly
    per_channel = np.ascontiguousarray(chan.transpose(0, 2, 1)).reshape(-1, REFWAV_FRAMES)   # [36, N]
    save = {"versions": np.array(versions), "names": np.array(names), "pcm": pcm, "stride": np.array(REFWAV_STRIDE)}
    fe80, fe128 = WhisperFeatureExtractor(feature_size=80), WhisperFeatureExtractor(feature_size=128)
    kw = dict(sampling_rate=16000, max_length=REFWAV_FRAMES, return_tensors="np")
    save["feat80_channels"] = fe80(list(per_channel), **kw)["input_features"][:, :, ::REFWAV_STRIDE].astype(np.float32)
    save["feat80_mono"] = fe80(list(mono), **kw)["input_features"][:, :, ::REFWAV_STRIDE].astype(np.float32)
    save["feat128_mono"] = fe128(list(mono), **kw)["input_features"][:, :, ::REFWAV_STRIDE].astype(np.float32)
    padded = np.zeros((len(files), 64000), np.float32)
    padded[:, :REFWAV_FRAMES] = mono
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, hop_length=512, n_mels=128, f_min=0, f_max=8000, power=2.0)
    save["ta_logmel_512_128"] = torch.log(ms(torch.from_numpy(padded)) + 1e-6).numpy()[:, :, ::REFWAV_STRIDE]
    np.savez_compressed(os.path.join(OUT, "refwav.npz"), **save)


def make_configs(versions: str):
    """BASELINE configs 1 and 4 at their exact sizes, plus the __call__ branches the drop-in forwards to HF."""
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor(feature_size=80)
    save = {"versions": np.array(versions)}
    # config 1: default_rng(0), B = 32, 80 mels, 30 s
    x = synth.gaussian_clips(32, seed=0)
    out = fe(list(x), sampling_rate=16000, return_tensors="np")["input_features"]
    save["cfg1_stride"] = np.array(97)
    save["cfg1_feat"] = np.ascontiguousarray(out[:, :, ::97]).astype(np.float32)
    save["cfg1_max"] = out.reshape(32, -1).max(axis=1)
    save["cfg1_mean"] = out.reshape(32, -1).mean(axis=1, dtype=np.float64)
    # config 4: 1000 piano clips (one per row of /root/reference/AB/mididataset.csv), mostly zero padding
    w, n = synth.midi_piano_clips(1000, seed=0)
    mx, mean, keep = [], [], {}
    for s0 in range(0, 1000, 50):
        o = fe(list(w[s0:s0 + 50]), sampling_rate=16000, return_tensors="np")["input_features"]
        mx.append(o.reshape(50, -1).max(axis=1))
        mean.append(o.reshape(50, -1).mean(axis=1, dtype=np.float64))
        for i in range(s0, s0 + 50):
            if i % 125 == 0:
                keep[i] = np.ascontiguousarray(o[i - s0][:, ::29]).astype(np.float32)
    save["cfg4_lengths"] = n
    save["cfg4_max"] = np.concatenate(mx)
    save["cfg4_mean"] = np.concatenate(mean)
    save["cfg4_keep_idx"] = np.array(sorted(keep))
    save["cfg4_keep_feat"] = np.stack([keep[i] for i in sorted(keep)])
    save["cfg4_stride"] = np.array(29)
    # the other __call__ branches: attention mask (feature_extraction_whisper.py:328-337) and
    # do_normalize (:306-312) on ragged one-second-ish clips
    ragged = [synth.gaussian_clips(1, L, seed=50 + i)[0] * (1 + i) for i, L in enumerate((16000, 9000, 399, 12345))]
    for i, c in enumerate(ragged):
        save[f"ragged_in{i}"] = c
    o = fe(ragged, sampling_rate=16000, max_length=16000, return_attention_mask=True, return_tensors="np")
    save["ragged_mask"] = o["attention_mask"].astype(np.int32)
    save["ragged_feat"] = o["input_features"].astype(np.float32)
    o = fe(ragged, sampling_rate=16000, max_length=16000, do_normalize=True, return_attention_mask=True, return_tensors="np")
    save["ragged_feat_normalized"] = o["input_features"].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "whisper_cfg.npz"), **save)


if __name__ == "__main__":
    main()
