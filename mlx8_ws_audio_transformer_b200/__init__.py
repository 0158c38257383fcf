"""B200-native log-mel feature frontend (drop-in for the reference's feature-extraction call).

Public surface:

* :class:`LogMelWhisperFeatureExtractor` -- ``WhisperFeatureExtractor`` replacement
  (``AB/fineTune.py:88``, ``AB/wavToWhisper.py:55``, ``.charles/music2midi/model.py:100-104``);
* :class:`MelSpectrogram` / :class:`LogMelSpectrogram` -- ``torchaudio.transforms.MelSpectrogram``
  replacement (``.charles/spectrogram.py:79-87,161-162``);
* :class:`LogMelFrontend` -- the raw operator over device or host buffers;
* :class:`ShardedFrontend`, :func:`shard_bounds` -- clip sharding across ranks;
* :class:`PinnedFeatureWriter`, :class:`DeviceCollator` -- feature store / collator side
  (``.charles/spectrogram.py:165-181``, ``AB/fineTune.py:99-118``);
* :mod:`openai_whisper` -- ``whisper.log_mel_spectrogram`` / ``pad_or_trim`` semantics (``AB/wavToWhisper.py:10-13``).

Heavy imports (torch, transformers) happen on first attribute access, not at package import.
"""
from __future__ import annotations

__all__ = [
    "LogMelFrontend", "LogMelWhisperFeatureExtractor", "MelSpectrogram", "LogMelSpectrogram",
    "ShardedFrontend", "shard_bounds", "shard_sizes", "launch_count", "build",
    "PinnedFeatureWriter", "DeviceCollator", "openai_whisper",
]

_LAZY = {
    "LogMelFrontend": ("frontend", "LogMelFrontend"),
    "launch_count": ("frontend", "launch_count"),
    "LogMelWhisperFeatureExtractor": ("whisper", "LogMelWhisperFeatureExtractor"),
    "MelSpectrogram": ("melspec", "MelSpectrogram"),
    "LogMelSpectrogram": ("melspec", "LogMelSpectrogram"),
    "ShardedFrontend": ("sharding", "ShardedFrontend"),
    "shard_bounds": ("sharding", "shard_bounds"),
    "shard_sizes": ("sharding", "shard_sizes"),
    "build": ("_native", "build"),
    "PinnedFeatureWriter": ("store", "PinnedFeatureWriter"),
    "DeviceCollator": ("store", "DeviceCollator"),
}


def __getattr__(name):
    if name == "openai_whisper":
        import importlib

        return importlib.import_module(f"{__name__}.openai_whisper")
    if name in _LAZY:
        import importlib

        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
