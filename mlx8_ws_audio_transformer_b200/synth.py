"""Deterministic synthetic clips for the five BASELINE.json configs.

There is no network and no dataset on the GPU box, so every parity test and every bench
line runs on clips made here.  Shapes follow the reference's own data:

* 30 s @ 16 kHz Whisper containers (``/root/reference/AB/fineTune.py:81-88``),
* 4 s UrbanSound8K excerpts, zero padded (``/root/reference/.charles/spectrogram.py:152-157``),
* 5-note piano clips from the MIDI generator (``/root/reference/AB/midiDatasetGen.py:7-40``),
  rendered with an additive synth because fluidsynth is not installed.

Only numpy is used, so the oracle, the tests and bench.py can all share these inputs.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
WHISPER_SAMPLES = 480000
URBAN_SAMPLES = 64000


def gaussian_clips(n_clips: int, n_samples: int = WHISPER_SAMPLES, seed: int = 0,
                   sigma: float = 0.1) -> np.ndarray:
    """Config 1/2/5: ``default_rng(seed).standard_normal((B, L)) * sigma`` as float32."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n_clips, n_samples)) * sigma).astype(np.float32)


def sine_clip(freq_hz: float, n_samples: int = WHISPER_SAMPLES, amp: float = 0.5) -> np.ndarray:
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    return (amp * np.sin(2.0 * np.pi * freq_hz * t)).astype(np.float32)


def chirp_clip(f0: float = 20.0, f1: float = 8000.0, n_samples: int = WHISPER_SAMPLES,
               amp: float = 0.5) -> np.ndarray:
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    dur = n_samples / SAMPLE_RATE
    phase = 2.0 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    return (amp * np.sin(phase)).astype(np.float32)


def impulse_clip(position: int, n_samples: int = WHISPER_SAMPLES, amp: float = 1.0) -> np.ndarray:
    x = np.zeros(n_samples, dtype=np.float32)
    x[position] = amp
    return x


def int16_uniform_clip(n_samples: int = WHISPER_SAMPLES, seed: int = 1) -> np.ndarray:
    """Uniform noise quantised to the s16 grid, as a decoded 16-bit WAV would be."""
    rng = np.random.default_rng(seed)
    q = rng.integers(-32768, 32768, size=n_samples, dtype=np.int64)
    return (q.astype(np.float64) / 32768.0).astype(np.float32)


def urbansound_clips(n_clips: int, seed: int = 0, n_samples: int = URBAN_SAMPLES,
                     sigma: float = 0.1):
    """Config 3: Gaussian excerpts with active length U{8000..n_samples}, zero padded.

    Returns ``(wave[B, n_samples] float32, lengths[B] int32)``; samples past ``lengths[i]``
    are already zero, so the array can be used with or without the lengths.
    """
    rng = np.random.default_rng(seed)
    wave = (rng.standard_normal((n_clips, n_samples)) * sigma).astype(np.float32)
    lengths = rng.integers(8000, n_samples + 1, size=n_clips).astype(np.int32)
    mask = np.arange(n_samples)[None, :] < lengths[:, None]
    wave *= mask
    return wave, lengths


_NOTE_OFFSETS = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}


def note_to_midi(name: str) -> int:
    """'G#6' -> 92 (pretty_midi.note_name_to_number convention: C4 = 60)."""
    name = name.strip()
    pitch = _NOTE_OFFSETS[name[0].upper()]
    rest = name[1:]
    while rest and rest[0] in "#b":
        pitch += 1 if rest[0] == "#" else -1
        rest = rest[1:]
    return 12 * (int(rest) + 1) + pitch


def midi_piano_clips(n_clips: int, seed: int = 0, n_samples: int = WHISPER_SAMPLES,
                     notes_per_clip: int = 5):
    """Config 4: five decaying piano-like notes, then an exact-zero tail up to 30 s.

    Mirrors the clip shape of ``/root/reference/AB/midiDatasetGen.py:8,27-39`` (pitches drawn
    from a fixed set, gaps 1.1-1.5 s) with an additive synth:
    ``sum_h (0.5/h) sin(2 pi h f0 t) exp(-6 t)``, harmonics above 8 kHz dropped, peak 0.5.
    Returns ``(wave[B, n_samples] float32, lengths[B] int32)``.
    """
    rng = np.random.default_rng(seed)
    wave = np.zeros((n_clips, n_samples), dtype=np.float32)
    lengths = np.zeros(n_clips, dtype=np.int32)
    for i in range(n_clips):
        pitches = rng.integers(21, 109, size=notes_per_clip)
        onset = 0.0
        clip = np.zeros(n_samples, dtype=np.float64)
        end = 0
        for p in pitches:
            f0 = 440.0 * 2.0 ** ((int(p) - 69) / 12.0)
            dur = float(rng.uniform(0.4, 1.0))
            n0 = int(round(onset * SAMPLE_RATE))
            n1 = min(n_samples, n0 + int(round(dur * SAMPLE_RATE)))
            t = np.arange(n1 - n0, dtype=np.float64) / SAMPLE_RATE
            note = np.zeros_like(t)
            for h in range(1, 7):
                if h * f0 >= 8000.0:
                    break
                note += (0.5 / h) * np.sin(2.0 * np.pi * h * f0 * t)
            clip[n0:n1] += note * np.exp(-6.0 * t)
            end = max(end, n1)
            onset += float(rng.uniform(1.1, 1.5))
        peak = np.max(np.abs(clip))
        if peak > 0:
            clip *= 0.5 / peak
        wave[i] = clip.astype(np.float32)
        lengths[i] = end
    return wave, lengths
