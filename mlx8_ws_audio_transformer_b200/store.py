"""Feature store / collator side of the hot path.

The reference materialises every example's features as Python / Arrow lists and re-stacks them each step:

* ``/root/reference/.charles/spectrogram.py:165-181`` -- per file: ``.cpu().numpy().astype(float32).flatten()`` into a
  list of dicts, a DataFrame, ``to_parquet``; read back and reshaped per item at ``:204-212``;
* ``/root/reference/AB/fineTune.py:89,95`` -- ``input_features[0]`` stored in the ``datasets`` Arrow table per example,
  ``:104-118`` -- the collator calls ``feature_extractor.pad`` on the list of dicts to re-stack them.

Here the kernel's output goes straight into ONE pinned host block (asynchronous D2H on a side stream while
the next batch is computed), that block IS the Arrow column (zero-copy ``FixedSizeList<float32>``) the Parquet
writer consumes, and the collator stacks device tensors without a host round trip.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np


class PinnedFeatureWriter:
    """``[capacity, n_mels, frames]`` float32 host block filled batch by batch from CUDA feature tensors."""

    def __init__(self, capacity: int, n_mels: int, frames: int, pin: bool = True):
        import torch

        self.shape = (int(capacity), int(n_mels), int(frames))
        use_pin = bool(pin) and torch.cuda.is_available()
        self.buffer = torch.empty(self.shape, dtype=torch.float32, pin_memory=use_pin)
        self.count = 0
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._events = []

    def write(self, feats, start: Optional[int] = None) -> slice:
        """copy ``feats [b, n_mels, frames]`` (CUDA or host tensor / ndarray) into rows ``start : start + b``
        (default: append).  CUDA sources are copied asynchronously behind the producing stream; call
        :meth:`synchronize` (or any of the ``to_*`` methods) before reading the block."""
        import torch

        if not isinstance(feats, torch.Tensor):
            feats = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32))
        b = int(feats.shape[0])
        lo = self.count if start is None else int(start)
        if tuple(feats.shape[1:]) != self.shape[1:] or lo < 0 or lo + b > self.shape[0]:
            raise ValueError(f"features {tuple(feats.shape)} do not fit rows {lo}:{lo + b} of a {self.shape} store")
        dst = self.buffer[lo:lo + b]
        if feats.is_cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(feats.device))
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ready)
                dst.copy_(feats, non_blocking=True)
                feats.record_stream(self._stream)
                done = torch.cuda.Event()
                done.record(self._stream)
            self._events.append(done)
        else:
            dst.copy_(feats.to(torch.float32))
        self.count = max(self.count, lo + b)
        return slice(lo, lo + b)

    def synchronize(self):
        for e in self._events:
            e.synchronize()
        self._events.clear()

    def numpy(self) -> np.ndarray:
        self.synchronize()
        return self.buffer[:self.count].numpy()

    def to_arrow(self, columns: Optional[Dict[str, Sequence[Any]]] = None, flat_name: str = "log_mel_flat",
                 shape_name: str = "log_mel_shape"):
        """pyarrow Table with the reference's two columns (spectrogram.py:171-172) plus ``columns``; the flat
        feature column is a zero-copy view of the pinned block."""
        import pyarrow as pa

        feats = self.numpy()
        n, per = feats.shape[0], feats.shape[1] * feats.shape[2]
        flat = pa.FixedSizeListArray.from_arrays(pa.array(feats.reshape(-1), type=pa.float32()), per)
        shapes = pa.array([[feats.shape[1], feats.shape[2]]] * n, type=pa.list_(pa.int64()))
        names, arrays = [], []
        for k, v in (columns or {}).items():
            if len(v) != n:
                raise ValueError(f"column {k!r} has {len(v)} rows, the store {n}")
            names.append(k)
            arrays.append(pa.array(list(v)))
        return pa.table(arrays + [flat, shapes], names=names + [flat_name, shape_name])

    def to_parquet(self, path: str, columns: Optional[Dict[str, Sequence[Any]]] = None, **kw):
        """what ``preprocess_to_parquet`` ends with (spectrogram.py:177-181); readable by ``UrbanSoundDataSet``
        (``np.array(row["log_mel_flat"]).reshape(tuple(row["log_mel_shape"]))``, :204-212)."""
        import pyarrow.parquet as pq

        pq.write_table(self.to_arrow(columns), path, **kw)


class DeviceCollator:
    """``DataCollatorSpeechSeq2SeqWithPadding`` (/root/reference/AB/fineTune.py:99-118) for features that are
    already (CUDA) tensors: ``torch.stack`` instead of ``feature_extractor.pad`` over lists of lists, labels
    padded with -100, the leading ``decoder_start_token_id`` cut exactly as the reference does."""

    def __init__(self, decoder_start_token_id: int, pad_label: int = -100):
        self.decoder_start_token_id = int(decoder_start_token_id)
        self.pad_label = int(pad_label)

    def __call__(self, features: List[Dict[str, Any]]) -> Dict[str, Any]:
        import torch

        feats = [f["input_features"] if isinstance(f["input_features"], torch.Tensor) else torch.as_tensor(f["input_features"])
                 for f in features]
        batch = {"input_features": torch.stack(feats)}
        if "labels" in features[0]:
            dev = batch["input_features"].device
            width = max(len(f["labels"]) for f in features)
            labels = torch.full((len(features), width), self.pad_label, dtype=torch.long)
            for i, f in enumerate(features):
                labels[i, :len(f["labels"])] = torch.as_tensor(f["labels"], dtype=torch.long)
            if bool((labels[:, 0] == self.decoder_start_token_id).all()):
                labels = labels[:, 1:]
            batch["labels"] = labels.to(dev)
        return batch
