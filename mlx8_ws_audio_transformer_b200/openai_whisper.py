"""``whisper.log_mel_spectrogram`` / ``whisper.pad_or_trim`` semantics on the B200 kernel.

``/root/reference/AB/wavToWhisper.py:10-13`` and ``AB/UI/Asmo.py:45,69`` transcribe whole files with
``openai-whisper``'s ``model.transcribe``, which computes its own log-mel: the same N_FFT = 400 / HOP = 160
Hann STFT and Slaney bank as the HF extractor, but over the WHOLE file (plus ``padding`` zeros --
``transcribe`` passes 30 s) with ONE maximum for the whole file, and only then cuts 3000-frame segments
(``whisper/audio.py: log_mel_spectrogram``, published algorithm restated below; openai-whisper itself is
not installed here, so parity of this entry point is pinned to that restatement through the oracle, not to
the library: "parity unpinned" for 8f-4).

    stft = torch.stft(audio (+ padding zeros), 400, 160, hann_window(400), return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    log_spec = clamp(filters @ magnitudes, min=1e-10).log10()
    log_spec = maximum(log_spec, log_spec.max() - 8.0);  log_spec = (log_spec + 4.0) / 4.0

That is the fused kernel with the file as the one "clip": ``n_samples = len(audio) + padding``,
``drop_last`` and the per-clip maximum.  Files longer than ~41 s take the CTA-tiled kernel (a clip spread
over many CTAs); shorter ones, batched, the thread-per-frame kernel.  The same 128-filter extractor is what
Qwen2-Audio's processor builds (``.charles/music2midi/test/qwen2_audio_tests.py:34,51-52``): there the drop-in
``LogMelWhisperFeatureExtractor(feature_size=128)`` with ``return_attention_mask=True`` is the frontend.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .filters import slaney_mel_filter_bank
from .frontend import LogMelFrontend

SAMPLE_RATE, N_FFT, HOP_LENGTH, CHUNK_LENGTH = 16000, 400, 160, 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE          # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH              # 3000

_FRONTENDS: dict = {}


def _frontend(n_mels: int, device: int) -> LogMelFrontend:
    key = (n_mels, device)
    fe = _FRONTENDS.get(key)
    if fe is None:
        fe = LogMelFrontend(N_FFT, HOP_LENGTH, slaney_mel_filter_bank(N_FFT // 2 + 1, n_mels), N.LOG10_CLAMP_WHISPER_NORM,
                            log_param=1e-10, drop_last=True, device=device)
        _FRONTENDS[key] = fe
    return fe


def pad_or_trim(array, length: int = N_SAMPLES, axis: int = -1):
    """``whisper.pad_or_trim``: cut or right-zero-pad ``axis`` to ``length`` (torch tensor or ndarray)."""
    import torch

    if isinstance(array, torch.Tensor):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - array.shape[axis])
            array = torch.nn.functional.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array
    array = np.asarray(array)
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, device=None):
    """``whisper.log_mel_spectrogram(audio, n_mels, padding, device)`` -> float32 ``[..., n_mels, n_frames]`` CUDA tensor.

    ``audio``: ndarray / tensor ``[..., n]`` of 16 kHz samples (file paths are not decoded here: ffmpeg is the
    reference's L0, out of scope).  Every row is normalised with its OWN maximum, as the library does for a
    batch of one; ``padding`` zeros are appended on the device by the kernel (never materialised).
    """
    import torch

    if isinstance(audio, str):
        raise TypeError("log_mel_spectrogram(b200) takes samples, not a path: decode and resample on the host first")
    if not isinstance(audio, torch.Tensor):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    if device is None:
        device = audio.device if audio.is_cuda else torch.device("cuda", torch.cuda.current_device())
    audio = audio.to(device=device, dtype=torch.float32)
    lead = audio.shape[:-1]
    flat = audio.reshape(-1, audio.shape[-1])
    fe = _frontend(int(n_mels), torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device())
    out = fe.forward(flat, n_samples=flat.shape[-1] + int(padding))
    return out.reshape(lead + out.shape[-2:])
