"""Filter banks the frontend is configured with (host-side constants, computed once).

The kernel takes the bank as data (``lm_config.fbank``), so these functions only have to
reproduce the two banks the reference's libraries build:

* :func:`slaney_mel_filter_bank` -- ``WhisperFeatureExtractor.mel_filters``
  (transformers/models/whisper/feature_extraction_whisper.py:94-102;
  transformers/audio_utils.py:453-544), float64, Slaney scale and area normalisation;
* :func:`torchaudio_mel_filter_bank` -- ``MelScale.fb``
  (torchaudio/functional/functional.py:518-587), float32 arithmetic with torch primitives so the
  roundings match the library bit for bit (a float64 derivation differs by 1e-5 per weight).
"""
from __future__ import annotations

import math

import numpy as np


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    with np.errstate(divide="ignore"):
        log = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)


def slaney_mel_filter_bank(n_freq: int = 201, n_mels: int = 80, f_min: float = 0.0,
                           f_max: float = 8000.0, sample_rate: int = 16000) -> np.ndarray:
    """float64 ``[n_freq, n_mels]``; equal to ``WhisperFeatureExtractor(feature_size=n_mels).mel_filters``."""
    pts = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2))
    freqs = np.linspace(0, sample_rate // 2, n_freq)
    diff = np.diff(pts)
    slopes = pts[None, :] - freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]))
    return fb * (2.0 / (pts[2:n_mels + 2] - pts[:n_mels]))[None, :]


def torchaudio_mel_filter_bank(n_freq: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                               norm: str | None = None, mel_scale: str = "htk"):
    """float32 torch tensor ``[n_freq, n_mels]``; equal to ``torchaudio.functional.melscale_fbanks``."""
    import torch

    if norm is not None and norm != "slaney":
        raise ValueError('norm must be one of None or "slaney"')
    if mel_scale not in ("htk", "slaney"):
        raise ValueError('mel_scale should be one of "htk" or "slaney".')

    def hz_to_mel(freq: float) -> float:
        if mel_scale == "htk":
            return 2595.0 * math.log10(1.0 + (freq / 700.0))
        f_sp = 200.0 / 3
        mels = freq / f_sp
        if freq >= 1000.0:
            mels = 1000.0 / f_sp + math.log(freq / 1000.0) / (math.log(6.4) / 27.0)
        return mels

    def mel_to_hz(mels):
        if mel_scale == "htk":
            return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
        f_sp = 200.0 / 3
        freqs = f_sp * mels
        min_log_mel = 1000.0 / f_sp
        log_t = mels >= min_log_mel
        freqs[log_t] = 1000.0 * torch.exp((math.log(6.4) / 27.0) * (mels[log_t] - min_log_mel))
        return freqs

    all_freqs = torch.linspace(0, sample_rate // 2, n_freq)
    m_pts = torch.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2)
    f_pts = mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb
