"""CPU oracle for the log-mel hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.

This is a float64 NumPy restatement of the arithmetic the reference reaches through its
feature-extraction calls (``/root/reference/AB/fineTune.py:88``,
``/root/reference/AB/wavToWhisper.py:55``, ``/root/reference/.charles/music2midi/model.py:100-104``,
``/root/reference/.charles/spectrogram.py:79-87,161-162``).  The arithmetic itself lives in
third-party libraries that are NOT under /root/reference:

* transformers (pinned ==4.35.2 in ``/root/reference/AB/pyproject.toml:25`` and 4.53.1 in
  ``/root/reference/.charles/uv.lock:1590-1591``; 5.5.0 installed here), files
  ``transformers/models/whisper/feature_extraction_whisper.py`` and ``transformers/audio_utils.py``;
* torchaudio 2.7.1 (``/root/reference/.charles/uv.lock:1502-1538``; 2.11.0 installed here), files
  ``torchaudio/transforms/_transforms.py`` and ``torchaudio/functional/functional.py``.

Pinning: the reference holds no golden vectors or known-answer tests for this path
(SURVEY.md §4, §8c).  The oracle is pinned instead against the live libraries in the build
container: ``oracle/make_golden.py`` runs ``WhisperFeatureExtractor.__call__`` and
``torchaudio.transforms.MelSpectrogram`` on seeded inputs and commits their outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against those vectors
(and, when the libraries are importable, against the live calls).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------
# constants: window and filter banks
# ----------------------------------------------------------------------------------------


def hann_periodic(n: int) -> np.ndarray:
    """``torch.hann_window(n)`` / ``window_function(n, "hann")`` (periodic), float64.

    follows transformers/audio_utils.py:560-620 and feature_extraction_whisper.py:141.
    """
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def _hz_to_mel_slaney(f):
    # transformers/audio_utils.py:263-297 (mel_scale="slaney")
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    log_region = f >= 1000.0
    with np.errstate(divide="ignore", invalid="ignore"):
        mels = np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, mels)
    return mels


def _mel_to_hz_slaney(m):
    # transformers/audio_utils.py:300-332 (mel_scale="slaney")
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    log_region = m >= 15.0
    return np.where(log_region, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def slaney_mel_filter_bank(n_freq: int = 201, n_mels: int = 80, f_min: float = 0.0,
                           f_max: float = 8000.0, sample_rate: int = 16000) -> np.ndarray:
    """float64 ``[n_freq, n_mels]`` bank of ``WhisperFeatureExtractor.__init__``.

    follows transformers/models/whisper/feature_extraction_whisper.py:94-102 and
    transformers/audio_utils.py:453-544, 356-375 (norm="slaney", mel_scale="slaney").
    """
    mel_pts = np.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2)
    filter_freqs = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sample_rate // 2, n_freq)
    filter_diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    return fb * enorm[None, :]


def htk_mel_filter_bank_f32(n_freq: int, n_mels: int, f_min: float, f_max: float,
                            sample_rate: int) -> np.ndarray:
    """float32 ``[n_freq, n_mels]`` bank of ``torchaudio.functional.melscale_fbanks``.

    follows torchaudio/functional/functional.py:518-587, 425-515 (mel_scale="htk", norm=None).
    torchaudio builds this bank in float32 with torch kernels (``torch.linspace``, ``**``), and
    neither a float64 nor a NumPy-float32 re-derivation reproduces its roundings (they differ
    by up to 1.4e-5 per weight), so the formulas are restated here with the same torch float32
    primitives in the same order.  ``tests/test_oracle.py`` checks it against the library
    object bit for bit.
    """
    import math
    import torch
    all_freqs = torch.linspace(0, sample_rate // 2, n_freq)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    return fb.numpy().astype(np.float32)


# ----------------------------------------------------------------------------------------
# the operator
# ----------------------------------------------------------------------------------------


def _frames_f64(x: np.ndarray, n_fft: int, hop: int, n_frames: int) -> np.ndarray:
    """reflect-pad by n_fft//2 and slice ``n_frames`` overlapping frames (float64 view)."""
    pad = n_fft // 2
    xp = np.pad(x.astype(np.float64), (pad, pad), mode="reflect")
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    return xp[idx]


def power_spectrogram(x: np.ndarray, n_fft: int, hop: int, n_frames: int) -> np.ndarray:
    """``|rfft(hann * frame)|^2`` as float64 ``[n_frames, n_fft//2+1]``.

    follows transformers/audio_utils.py:769-808 (center=True, reflect, periodic Hann, float64)
    and torchaudio/functional/functional.py:123-144 (same operator in float32).
    """
    frames = _frames_f64(x, n_fft, hop, n_frames) * hann_periodic(n_fft)[None, :]
    spec = np.fft.rfft(frames, axis=1)
    return spec.real ** 2 + spec.imag ** 2


def whisper_logmel(wave: np.ndarray, mel_filters: np.ndarray | None = None, n_mels: int = 80,
                   n_fft: int = 400, hop: int = 160, n_samples: int | None = 480000) -> np.ndarray:
    """``WhisperFeatureExtractor`` features, float32 ``[B, n_mels, n_samples // hop]``.

    ``wave`` is ``[B, T]`` (or a list of 1-D arrays of any length); each clip is right
    zero-padded / truncated to ``n_samples`` first
    (feature_extraction_whisper.py:296-303; feature_extraction_sequence_utils.py:276-277).
    Then, per clip (feature_extraction_whisper.py:105-133,135-164; audio_utils.py:769-830):
    reflect pad 200, periodic Hann, 400-point rfft every 160 samples, ``|X|^2``, drop the last
    frame, ``mel_filters.T @ P`` with the float64 bank cast to float32, clamp at 1e-10,
    ``log10``, ``max(S, max(S) - 8)`` with a per-clip max, ``(S + 4) / 4``.
    """
    clips = [np.asarray(w, dtype=np.float32).reshape(-1) for w in wave]
    if n_samples is None:
        n_samples = max(len(c) for c in clips)
    if mel_filters is None:
        mel_filters = slaney_mel_filter_bank(n_fft // 2 + 1, n_mels)
    fb = np.asarray(mel_filters).astype(np.float32).astype(np.float64)   # :152 casts to f32
    n_frames = n_samples // hop            # 1 + L//hop frames, last one dropped (:150 / :128)
    out = np.empty((len(clips), fb.shape[1], n_frames), dtype=np.float32)
    for i, c in enumerate(clips):
        x = np.zeros(n_samples, dtype=np.float32)
        m = min(len(c), n_samples)
        x[:m] = c[:m]
        p = power_spectrogram(x, n_fft, hop, n_frames)          # [frames, bins]
        mel = fb.T @ p.T                                        # [n_mels, frames]
        s = np.log10(np.maximum(mel, 1e-10))
        s = np.maximum(s, s.max() - 8.0)
        out[i] = ((s + 4.0) / 4.0).astype(np.float32)
    return out


def torchaudio_mel(wave: np.ndarray, fb: np.ndarray, n_fft: int = 1024, hop: int = 512,
                   log_offset: float | None = 1e-6, lengths=None) -> np.ndarray:
    """``MelSpectrogram(...)(w)`` and optionally ``torch.log(mel + log_offset)``.

    float32 ``[B, n_mels, 1 + T // hop]``; follows torchaudio/transforms/_transforms.py:621-631,
    407-419, torchaudio/functional/functional.py:123-144 and
    /root/reference/.charles/spectrogram.py:161-162.  No frame is dropped and nothing is
    normalised.  ``lengths`` zeroes the tail of each clip first (spectrogram.py:152-157).
    """
    wave = np.asarray(wave, dtype=np.float32)
    if wave.ndim == 1:
        wave = wave[None, :]
    n_frames = 1 + wave.shape[1] // hop
    fb64 = np.asarray(fb, dtype=np.float32).astype(np.float64)
    out = np.empty((wave.shape[0], fb64.shape[1], n_frames), dtype=np.float32)
    for i in range(wave.shape[0]):
        x = wave[i].copy()
        if lengths is not None:
            x[int(lengths[i]):] = 0.0
        p = power_spectrogram(x, n_fft, hop, n_frames)
        mel = fb64.T @ p.T
        if log_offset is not None:
            mel = np.log(mel + log_offset)
        out[i] = mel.astype(np.float32)
    return out


def parity(a: np.ndarray, b: np.ndarray):
    """(max-abs, mean-abs) difference — the two figures BASELINE.json's north_star bounds."""
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return float(d.max()) if d.size else 0.0, float(d.mean()) if d.size else 0.0
