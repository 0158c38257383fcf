"""Drop-in for the ``torchaudio.transforms.MelSpectrogram`` frontend of ``.charles/spectrogram.py``.

Reference call sites: ``/root/reference/.charles/spectrogram.py:79-87`` (construction),
``:161-162``, ``:299-300``, ``:306-307`` (``mel_spectrogram(waveform)`` then
``torch.log(mel + 1e-6)``).  Same constructor keywords, ``.to(device)``, and
``forward(waveform[..., T]) -> [..., n_mels, 1 + T // hop]`` float32 on the input's device.
The arithmetic of ``Spectrogram`` + ``MelScale``
(torchaudio/functional/functional.py:123-144; torchaudio/transforms/_transforms.py:407-419)
runs in the sm_100a kernel; ``log_offset=1e-6`` fuses the reference's following ``torch.log``.
"""
from __future__ import annotations

import torch

from . import _native as N
from .filters import torchaudio_mel_filter_bank
from .frontend import LogMelFrontend


class MelSpectrogram(torch.nn.Module):
    """``torchaudio.transforms.MelSpectrogram`` computed by ``liblogmel_b200.so``.

    Supported: the configuration family the reference uses -- ``power=2.0``, ``center=True``,
    ``pad_mode="reflect"``, ``normalized=False``, ``win_length == n_fft``, ``pad=0``, periodic
    Hann window (or any explicit window through ``window_fn``), ``(n_fft, hop)`` in
    {(1024, 512), (1024, 128), (400, 160)}.  Anything else raises ``NotImplementedError`` rather
    than silently computing something different.

    ``log_offset``: if not ``None`` the module returns ``log(mel + log_offset)`` (one pass).
    """

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, win_length=None, hop_length=None,
                 f_min: float = 0.0, f_max=None, pad: int = 0, n_mels: int = 128,
                 window_fn=torch.hann_window, power: float = 2.0, normalized: bool = False, wkwargs=None,
                 center: bool = True, pad_mode: str = "reflect", onesided=None, norm=None,
                 mel_scale: str = "htk", log_offset=None, lm_variant: int = 0):
        super().__init__()
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.win_length = win_length if win_length is not None else n_fft
        self.hop_length = hop_length if hop_length is not None else self.win_length // 2
        self.pad = pad
        self.power = power
        self.normalized = normalized
        self.n_mels = n_mels
        self.f_max = float(f_max) if f_max is not None else float(sample_rate // 2)
        self.f_min = float(f_min)
        self.log_offset = log_offset
        self.lm_variant = lm_variant
        if power != 2.0 or normalized or not center or pad_mode != "reflect" or pad != 0 \
                or self.win_length != n_fft or onesided is False:
            raise NotImplementedError(
                "MelSpectrogram(b200): only power=2.0, normalized=False, center=True, pad_mode='reflect', "
                "pad=0, win_length=n_fft, onesided spectra are implemented")
        if (n_fft, self.hop_length) not in ((1024, 512), (1024, 128), (400, 160)):
            raise NotImplementedError(
                f"MelSpectrogram(b200): no kernel for n_fft={n_fft}, hop_length={self.hop_length}; "
                "implemented: (1024, 512), (1024, 128), (400, 160)")
        window = window_fn(self.win_length) if wkwargs is None else window_fn(self.win_length, **wkwargs)
        self.register_buffer("window", window.to(torch.float32), persistent=False)
        fb = torchaudio_mel_filter_bank(n_fft // 2 + 1, self.f_min, self.f_max, n_mels, sample_rate, norm, mel_scale)
        self.register_buffer("fb", fb, persistent=False)
        self._frontends = {}

    def __getstate__(self):       # device handles are not copied / pickled with the module
        d = self.__dict__.copy()
        d["_frontends"] = {}
        return d

    def _frontend(self, idx: int) -> LogMelFrontend:
        fe = self._frontends.get(idx)
        if fe is None:
            mode = N.LOG_NONE if self.log_offset is None else N.LN_PLUS_EPS
            fe = LogMelFrontend(self.n_fft, self.hop_length, self.fb, mode,
                                log_param=0.0 if self.log_offset is None else float(self.log_offset),
                                drop_last=False, device=idx, window=self.window, variant=self.lm_variant)
            self._frontends[idx] = fe
        return fe

    def forward(self, waveform: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        """``waveform[..., T]`` -> ``[..., n_mels, 1 + T // hop_length]``.

        ``lengths`` (optional, ``[...]`` int) marks how many leading samples of each row are
        audio; the rest is treated as the zero padding of spectrogram.py:152-157 without the
        caller having to materialise it.
        """
        shape = waveform.shape
        flat = waveform.reshape(-1, shape[-1])
        flat_len = None if lengths is None else lengths.reshape(-1)
        if flat.is_cuda:
            out = self._frontend(flat.device.index).forward(flat, lengths=flat_len)
        else:
            idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
            out = self._frontend(idx).forward_host(flat, lengths=flat_len)
        return out.reshape(shape[:-1] + out.shape[-2:])


class LogMelSpectrogram(MelSpectrogram):
    """``torch.log(MelSpectrogram(...)(w) + 1e-6)`` in one kernel (spectrogram.py:161-162)."""

    def __init__(self, *args, log_offset: float = 1e-6, **kwargs):
        super().__init__(*args, log_offset=log_offset, **kwargs)
