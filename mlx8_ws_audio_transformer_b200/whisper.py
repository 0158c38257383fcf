"""Drop-in for the ``WhisperFeatureExtractor`` the reference calls through ``WhisperProcessor``.

Call sites replaced: ``/root/reference/AB/fineTune.py:88,107``,
``AB/fineTuneMidi.py:88,107``, ``AB/wavToWhisper.py:55``, ``AB/fineTuneMidiTester.py:33``,
``.charles/music2midi/model.py:100-104``.  Same signature, same ``BatchFeature`` with
``input_features`` float32 ``[B, feature_size, 3000]`` (numpy by default, torch with
``return_tensors="pt"``), same errors -- but the arithmetic of
``_torch_extract_fbank_features`` / ``_np_extract_fbank_features``
(transformers/models/whisper/feature_extraction_whisper.py:105-164) runs in the sm_100a kernel.

Beyond the HF contract, a CUDA tensor in gives a CUDA tensor out with no host round trip
(the stock extractor always ends in ``.cpu()``, feature_extraction_whisper.py:162-163).
"""
from __future__ import annotations

import weakref

import numpy as np
from transformers import WhisperFeatureExtractor
from transformers.feature_extraction_utils import BatchFeature
from transformers.utils import logging as hf_logging

from . import _native as N
from .frontend import LogMelFrontend

logger = hf_logging.get_logger(__name__)

_FRONTENDS: dict = {}
# float32 image of an extractor's filter bank and its hash, computed once per instance (not per call:
# the music2midi loop of /root/reference/.charles/music2midi/model.py:96-110 calls the extractor per
# clip).  Kept outside the instance so that the extractor stays a plain JSON-serialisable HF object.
_BANKS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


class LogMelWhisperFeatureExtractor(WhisperFeatureExtractor):
    """``WhisperFeatureExtractor`` whose log-mel is computed by ``liblogmel_b200.so``.

    Extra keyword: ``lm_variant`` (kernel tuning knob, see ``lm_config.variant``).
    """

    def __init__(self, *args, lm_variant: int = 0, **kwargs):
        super().__init__(*args, **kwargs)
        self.lm_variant = int(lm_variant)     # plain int: survives to_dict()/save_pretrained()

    @classmethod
    def from_hf(cls, fe: WhisperFeatureExtractor, **kw):
        """Same configuration as an existing HF extractor (e.g. ``processor.feature_extractor``)."""
        return cls(feature_size=fe.feature_size, sampling_rate=fe.sampling_rate, hop_length=fe.hop_length,
                   chunk_length=fe.chunk_length, n_fft=fe.n_fft, padding_value=fe.padding_value,
                   dither=getattr(fe, "dither", 0.0), return_attention_mask=fe.return_attention_mask, **kw)

    # ------------------------------------------------------------------------------------
    def _frontend(self, device=None) -> LogMelFrontend:
        import torch

        idx = _device_index(device)
        if idx is None:
            idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
        # handles are cached per (geometry, bank, device) outside the instance, so the extractor
        # itself stays a plain, copyable, JSON-serialisable HF object
        bank = _BANKS.get(self)
        if bank is None or bank[0] is not self.mel_filters:
            fb32 = np.ascontiguousarray(self.mel_filters, dtype=np.float32)
            bank = (self.mel_filters, fb32, hash(fb32.tobytes()))
            _BANKS[self] = bank
        fb32 = bank[1]
        key = (self.n_fft, self.hop_length, fb32.shape, bank[2], idx, self.lm_variant)
        fe = _FRONTENDS.get(key)
        if fe is None:
            fe = LogMelFrontend(self.n_fft, self.hop_length, fb32, N.LOG10_CLAMP_WHISPER_NORM,
                                log_param=1e-10, drop_last=True, device=idx, variant=self.lm_variant)
            _FRONTENDS[key] = fe
        return fe

    # the two extractor hooks HF's __call__ picks from (feature_extraction_whisper.py:317-320)
    def _torch_extract_fbank_features(self, waveform: np.ndarray, device: str = "cpu") -> np.ndarray:
        waveform = np.asarray(waveform, dtype=np.float32)
        squeeze = waveform.ndim == 1
        if squeeze:
            waveform = waveform[None, :]
        dither = getattr(self, "dither", 0.0)
        if dither != 0.0:
            waveform = waveform + dither * np.random.randn(*waveform.shape).astype(np.float32)
        out = self._frontend(device).forward_host(waveform)
        return out[0] if squeeze else out

    def _np_extract_fbank_features(self, waveform_batch: np.ndarray, device: str = "cpu") -> np.ndarray:
        return self._torch_extract_fbank_features(np.asarray(waveform_batch, dtype=np.float32), device)

    # ------------------------------------------------------------------------------------
    def __call__(self, raw_speech, truncation: bool = True, pad_to_multiple_of=None, return_tensors=None,
                 return_attention_mask=None, padding="max_length", max_length=None, sampling_rate=None,
                 do_normalize=None, device="cpu", **kwargs) -> BatchFeature:
        import torch

        if sampling_rate is not None:
            if sampling_rate != self.sampling_rate:
                raise ValueError(
                    f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a"
                    f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                    f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        else:
            logger.warning(
                f"It is strongly recommended to pass the `sampling_rate` argument to `{self.__class__.__name__}()`. "
                "Failing to do so can result in silent errors that might be hard to debug.")

        # the kernel's own padding is zero padding, so the fast path needs padding_value == 0 (the default)
        fast = (padding == "max_length" and truncation and pad_to_multiple_of is None and not do_normalize
                and getattr(self, "dither", 0.0) == 0.0 and self.padding_value == 0.0)
        cuda_list = (isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0
                     and all(isinstance(c, torch.Tensor) and c.is_cuda for c in raw_speech))
        if not fast:
            if isinstance(raw_speech, torch.Tensor):
                raw_speech = raw_speech.detach().cpu().numpy()
            elif cuda_list:
                raw_speech = [c.detach().cpu().numpy() for c in raw_speech]
            # HF's own host glue (pad / normalise / mask), with the kernel behind the extractor hooks
            return super().__call__(raw_speech, truncation=truncation, pad_to_multiple_of=pad_to_multiple_of,
                                    return_tensors=return_tensors, return_attention_mask=return_attention_mask,
                                    padding=padding, max_length=max_length, sampling_rate=sampling_rate,
                                    do_normalize=do_normalize, device=device, **kwargs)

        n_samples = max_length if max_length else self.n_samples
        want_mask = return_attention_mask if return_attention_mask is not None else self.return_attention_mask

        # ---- device-resident path: CUDA tensor(s) in, CUDA tensor out -----------------------
        # A [B, T] tensor, a 1-D tensor, or a LIST of ragged 1-D CUDA tensors -- the batch
        # WhisperAudioEncoder.forward (/root/reference/.charles/music2midi/model.py:94-123) walks
        # clip by clip through the processor: here it is ONE launch with per-clip lengths.
        if cuda_list or (isinstance(raw_speech, torch.Tensor) and raw_speech.is_cuda):
            if cuda_list:
                if any(c.dim() != 1 for c in raw_speech):
                    raise ValueError(f"Only mono-channel audio is supported for input to {self}")
                lens = [min(int(c.shape[0]), n_samples) for c in raw_speech]
                width = max(max(lens), 4)
                width += (-width) % 4                      # 16-byte aligned rows
                w = torch.zeros((len(raw_speech), width), dtype=torch.float32, device=raw_speech[0].device)
                for i, c in enumerate(raw_speech):
                    w[i, :lens[i]] = c[:lens[i]]
                lengths = torch.tensor(lens, dtype=torch.int32)
                feats = self._frontend(w.device).forward(w, lengths=lengths.to(w.device), n_samples=n_samples)
            else:
                if raw_speech.dim() > 2:
                    raise ValueError(f"Only mono-channel audio is supported for input to {self}")
                w = raw_speech if raw_speech.dim() == 2 else raw_speech[None, :]
                lengths = torch.full((w.shape[0],), min(w.shape[1], n_samples), dtype=torch.int32)
                feats = self._frontend(w.device).forward(w, n_samples=n_samples)
            data = {"input_features": feats}
            if want_mask:
                data["attention_mask"] = torch.from_numpy(self._frame_mask(lengths.numpy(), n_samples)).to(w.device)
            if return_tensors in (None, "np"):
                data = {k: v.cpu().numpy() for k, v in data.items()}
            elif str(return_tensors) not in ("pt", "TensorType.PYTORCH"):
                raise ValueError(f"return_tensors={return_tensors!r} is not supported for CUDA input")
            return BatchFeature(data)

        # ---- host path ---------------------------------------------------------------------
        if isinstance(raw_speech, torch.Tensor):
            raw_speech = raw_speech.detach().numpy()
        is_batched_numpy = isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1
        if is_batched_numpy and raw_speech.ndim > 2:
            raise ValueError(f"Only mono-channel audio is supported for input to {self}")
        is_batched = is_batched_numpy or (
            isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0
            and isinstance(raw_speech[0], (np.ndarray, tuple, list)))
        clips = list(raw_speech) if is_batched else [raw_speech]
        if is_batched_numpy and raw_speech.dtype == np.float32:
            width = min(raw_speech.shape[1], n_samples)
            stage = np.ascontiguousarray(raw_speech[:, :width])
            lengths = np.full(len(clips), width, dtype=np.int32)
        else:
            clips = [np.asarray(c, dtype=np.float32).reshape(-1) for c in clips]
            lengths = np.array([min(len(c), n_samples) for c in clips], dtype=np.int32)
            width = max(int(lengths.max()) if len(clips) else 0, 1)
            stage = np.zeros((len(clips), width), dtype=np.float32)
            for i, c in enumerate(clips):
                stage[i, :lengths[i]] = c[:lengths[i]]
        feats = self._frontend(device).forward_host(stage, lengths=lengths, n_samples=n_samples)
        data = {"input_features": feats}
        if want_mask:
            data["attention_mask"] = self._frame_mask(lengths, n_samples)
        out = BatchFeature(data)
        if return_tensors is not None:
            out = out.convert_to_tensors(return_tensors)
        return out

    def _frame_mask(self, lengths: np.ndarray, n_samples: int) -> np.ndarray:
        """attention mask rescaled from samples to frames (feature_extraction_whisper.py:328-337)."""
        pos = np.arange(0, n_samples, self.hop_length)[None, :]
        mask = (pos < lengths[:, None]).astype(np.int32)
        if n_samples % self.hop_length != 0:
            mask = mask[:, :-1]
        return mask


def _device_index(device):
    """'cuda:1' / torch.device / int -> ordinal; 'cpu' / None -> None (use the current GPU)."""
    if device is None:
        return None
    if isinstance(device, int):
        return device
    s = str(device)
    if s.startswith("cuda"):
        return int(s.split(":")[1]) if ":" in s else None
    return None
