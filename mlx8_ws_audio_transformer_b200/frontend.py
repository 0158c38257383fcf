"""``LogMelFrontend`` -- a thin Python owner of one ``lm_handle`` (one geometry, one device).

PyTorch is used here for what it is good at -- device memory, streams, pinned host memory --
and nothing else: every number is produced by ``liblogmel_b200.so``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as N


class LogMelFrontend:
    """One configured log-mel operator on one GPU.

    Parameters mirror ``lm_config`` (include/logmel.h).  ``fbank`` is ``[n_fft//2+1, n_mels]``
    (numpy or torch, any float dtype; rounded to float32 once, exactly as
    feature_extraction_whisper.py:152 and MelScale do).
    """

    def __init__(self, n_fft: int, hop: int, fbank, log_mode: int, log_param: float = 1e-10,
                 drop_last: bool = True, device: int | None = None, window=None, variant: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("LogMelFrontend needs a CUDA device: the log-mel path has no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        fb = np.ascontiguousarray(_to_numpy(fbank), dtype=np.float32)
        if fb.ndim != 2 or fb.shape[0] != n_fft // 2 + 1:
            raise ValueError(f"fbank must be [{n_fft // 2 + 1}, n_mels], got {fb.shape}")
        win = None if window is None else np.ascontiguousarray(_to_numpy(window), dtype=np.float32)
        if win is not None and win.shape != (n_fft,):
            raise ValueError(f"window must have {n_fft} samples")
        self.n_fft, self.hop, self.n_mels = int(n_fft), int(hop), int(fb.shape[1])
        self.log_mode, self.log_param, self.drop_last = int(log_mode), float(log_param), bool(drop_last)
        cfg = N.LmConfig(self.n_fft, self.hop, self.n_mels, self.log_mode, self.log_param,
                         1 if drop_last else 0, self.device_index, int(variant),
                         fb.ctypes.data_as(ctypes.c_void_p),
                         None if win is None else win.ctypes.data_as(ctypes.c_void_p))
        self._lib = N.lib()
        h = ctypes.c_void_p()
        N.check(self._lib.lm_create(ctypes.byref(h), ctypes.byref(cfg)), "lm_create")
        self._h = h

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lm_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    def num_frames(self, n_samples: int) -> int:
        return int(self._lib.lm_num_frames(self._h, int(n_samples)))

    def kernel_info(self) -> dict:
        v = [ctypes.c_int32() for _ in range(5)]
        N.check(self._lib.lm_kernel_info(self._h, *[ctypes.byref(x) for x in v]), "lm_kernel_info")
        keys = ("n_sm", "ctas_per_sm", "smem_bytes", "threads", "frames_per_tile")
        return {k: int(x.value) for k, x in zip(keys, v)}

    def kernel_name(self, batch: int, n_samples: int) -> str:
        """the kernel ``forward`` launches for this batch (as ncu prints it)"""
        return self._lib.lm_kernel_name(self._h, int(batch), int(n_samples)).decode()

    def _scratch_for(self, batch: int):
        """Per-CALL scratch from torch's caching allocator: the block belongs to the current stream, so
        two forward() calls of one (shared, cached) frontend on different streams never see each
        other's clip counters, and a block is only reused after the launch that used it."""
        import torch

        need = int(self._lib.lm_scratch_bytes(self._h, batch))
        return torch.empty(max(need, 16), dtype=torch.uint8, device=f"cuda:{self.device_index}")

    def _check_out(self, t, shape, name: str, device_type: str):
        import torch

        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != torch.float32 or tuple(t.shape) != tuple(shape) or not t.is_contiguous() or t.device.type != device_type:
            raise ValueError(f"{name} must be a contiguous float32 {device_type} tensor of shape {tuple(shape)}, "
                             f"got {t.dtype} {tuple(t.shape)} on {t.device} (contiguous={t.is_contiguous()})")
        if device_type == "cuda" and t.device.index != self.device_index:
            raise ValueError(f"{name} is on cuda:{t.device.index}, handle on cuda:{self.device_index}")

    def forward(self, wave, lengths=None, n_samples: int | None = None, out=None, clip_max=None):
        """Device-resident path: ``wave`` float32 CUDA ``[B, T]`` -> float32 CUDA ``[B, n_mels, frames]``.

        ``n_samples`` (default ``T``) is the length every clip is zero-padded / truncated to;
        ``lengths`` (int32 CUDA ``[B]``) marks how much of each row is real audio.  Runs
        asynchronously on the current torch stream.
        """
        import torch

        if not (isinstance(wave, torch.Tensor) and wave.is_cuda):
            raise TypeError("forward() takes a CUDA tensor; use forward_host() for host buffers")
        channels = 0
        if wave.dtype == torch.int16:
            # 16-bit PCM as a WAV file holds it: [B, T] mono or [B, T, C] interleaved frames (C = 1, 2);
            # converted (x / 32768) and down-mixed (mean over channels) inside the kernel's tile loader
            if wave.dim() == 1:
                wave = wave[None, :]
            channels = 1 if wave.dim() == 2 else int(wave.shape[2]) if wave.dim() == 3 else -1
            if channels not in (1, 2):
                raise ValueError("int16 PCM must be [B, T] (mono) or [B, T, C] with C = 1 or 2 interleaved channels")
            wave = wave.contiguous()
            pcm = wave
            wave = wave.reshape(wave.shape[0], -1)[:, ::channels]      # [B, T] view: shape bookkeeping only
        elif wave.dtype != torch.float32:
            wave = wave.float()
        if wave.dim() == 1:
            wave = wave[None, :]
        if wave.dim() != 2:
            raise ValueError("wave must be [B, T]")
        if not channels and wave.stride(1) != 1:
            wave = wave.contiguous()
        if wave.device.index != self.device_index:
            raise ValueError(f"wave is on cuda:{wave.device.index}, handle on cuda:{self.device_index}")
        B, T = wave.shape
        L = T if n_samples is None else int(n_samples)
        if lengths is None and L > T:
            lengths = torch.full((B,), T, dtype=torch.int32, device=wave.device)
        if lengths is not None:
            # never read past a row: the kernel clamps to n_samples, the row may be shorter than that
            lengths = lengths.to(device=wave.device, dtype=torch.int32).clamp(0, T).contiguous()
            if lengths.shape != (B,):
                raise ValueError(f"lengths must have shape ({B},), got {tuple(lengths.shape)}")
        frames = self.num_frames(L)
        if out is None:
            out = torch.empty((B, self.n_mels, max(frames, 0)), dtype=torch.float32, device=wave.device)
        else:
            self._check_out(out, (B, self.n_mels, max(frames, 0)), "out", "cuda")
        if clip_max is not None:
            self._check_out(clip_max, (B,), "clip_max", "cuda")
        if B == 0:
            return out
        scratch = self._scratch_for(B)
        stream = torch.cuda.current_stream(wave.device).cuda_stream
        if channels:
            rc = self._lib.lm_forward_pcm16(
                self._h, pcm.data_ptr(), channels, B, max(T, 1), L,
                None if lengths is None else lengths.data_ptr(), out.data_ptr(),
                None if clip_max is None else clip_max.data_ptr(), scratch.data_ptr(), scratch.numel(),
                ctypes.c_void_p(stream))
        else:
            rc = self._lib.lm_forward(
                self._h, wave.data_ptr(), B, wave.stride(0) if B > 1 else max(T, 1), L,
                None if lengths is None else lengths.data_ptr(), out.data_ptr(),
                None if clip_max is None else clip_max.data_ptr(), scratch.data_ptr(), scratch.numel(),
                ctypes.c_void_p(stream))
        N.check(rc, "lm_forward")
        return out

    def forward_host(self, wave, lengths=None, n_samples: int | None = None, out=None):
        """Host-buffer path through ``lm_forward_host``: numpy / CPU tensor in, same kind out.

        H2D copy, kernel and D2H copy of successive chunks overlap on three streams inside the
        library.  Pinned inputs / outputs (``torch.Tensor.pin_memory``) get full PCIe speed.
        """
        import torch

        is_torch = isinstance(wave, torch.Tensor)
        w = wave if is_torch else torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32))
        if not is_torch and np.asarray(wave).dtype == np.int16:
            w = torch.from_numpy(np.ascontiguousarray(wave))
        if w.is_cuda:
            raise TypeError("forward_host() takes host buffers")
        channels = 0
        if w.dtype == torch.int16:                     # 16-bit PCM: [B, T] or [B, T, C] (see forward())
            if w.dim() == 1:
                w = w[None, :]
            channels = 1 if w.dim() == 2 else int(w.shape[2]) if w.dim() == 3 else -1
            if channels not in (1, 2):
                raise ValueError("int16 PCM must be [B, T] (mono) or [B, T, C] with C = 1 or 2 interleaved channels")
        elif w.dtype != torch.float32:
            w = w.float()
        if w.dim() == 1:
            w = w[None, :]
        if not channels and w.dim() != 2:
            raise ValueError("wave must be [B, T]")
        w = w.contiguous()
        B, T = w.shape[0], w.shape[1]
        L = T if n_samples is None else int(n_samples)
        len_t = None
        if lengths is not None:
            len_t = torch.as_tensor(lengths, dtype=torch.int32).clamp(0, T).contiguous()
            if len_t.shape != (B,):
                raise ValueError(f"lengths must have shape ({B},), got {tuple(len_t.shape)}")
        elif L > T:
            len_t = torch.full((B,), T, dtype=torch.int32)
        frames = self.num_frames(L)
        if out is None:
            out = torch.empty((B, self.n_mels, max(frames, 0)), dtype=torch.float32)
        else:
            if not isinstance(out, torch.Tensor):
                out = torch.from_numpy(out)
            self._check_out(out, (B, self.n_mels, max(frames, 0)), "out", "cpu")
        if B and channels:
            rc = self._lib.lm_forward_host_pcm16(self._h, w.data_ptr(), channels, B, T, L,
                                                 None if len_t is None else len_t.data_ptr(), out.data_ptr())
            N.check(rc, "lm_forward_host_pcm16")
        elif B:
            rc = self._lib.lm_forward_host(self._h, w.data_ptr(), B, T, L,
                                           None if len_t is None else len_t.data_ptr(), out.data_ptr())
            N.check(rc, "lm_forward_host")
        return out if is_torch else out.numpy()


def launch_count() -> int:
    """Kernels launched by liblogmel_b200.so in this process (bench.py reports the delta)."""
    return int(N.lib().lm_launch_count())


def _to_numpy(x):
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(x)
