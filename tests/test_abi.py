"""The C-ABI library loads and exports every symbol include/logmel.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from mlx8_ws_audio_transformer_b200 import _native as N
from mlx8_ws_audio_transformer_b200 import shard_bounds, shard_sizes


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "logmel.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    N.build()
    lib = N.lib()
    declared = _header_symbols()
    assert len(declared) >= 10
    assert sorted(N.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert lib.lm_version() == 1


def test_argument_errors_before_any_launch():
    lib = N.lib()
    h = ctypes.c_void_p()
    assert lib.lm_create(ctypes.byref(h), None) == -1                      # LM_ERR_NULL
    assert b"non-NULL" in lib.lm_last_error()
    fb = np.zeros((201, 80), np.float32)
    cfg = N.LmConfig(400, 160, 80, 99, 1e-10, 1, 0, 0, fb.ctypes.data_as(ctypes.c_void_p), None)
    assert lib.lm_create(ctypes.byref(h), ctypes.byref(cfg)) == -7         # LM_ERR_MODE
    cfg = N.LmConfig(400, 160, 500, 1, 1e-10, 1, 0, 0, fb.ctypes.data_as(ctypes.c_void_p), None)
    assert lib.lm_create(ctypes.byref(h), ctypes.byref(cfg)) == -3         # LM_ERR_FBANK
    assert lib.lm_forward(None, None, 1, 1, 1, None, None, None, None, 0, None) == -1
    assert lib.lm_scratch_bytes(None, 4) == 0


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = N.lib()
    fb = np.zeros((201, 80), np.float32)
    cfg = N.LmConfig(400, 160, 80, 1, 1e-10, 1, 0, 0, fb.ctypes.data_as(ctypes.c_void_p), None)
    h = ctypes.c_void_p()
    assert lib.lm_create(ctypes.byref(h), ctypes.byref(cfg)) == -6         # LM_ERR_NO_DEVICE
    from mlx8_ws_audio_transformer_b200 import LogMelWhisperFeatureExtractor
    fe = LogMelWhisperFeatureExtractor()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fe(np.zeros(16000, np.float32), sampling_rate=16000)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mlx8_ws_audio_transformer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text.replace("oracle/logmel_oracle.py)", ""), f


def test_shard_bounds():
    assert [shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_sizes(8192, 8) == [1024] * 8
    assert sum(shard_sizes(7, 8)) == 7 and shard_bounds(7, 8, 7) == (7, 7)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_whisper_dropin_validation_matches_hf():
    """Errors raised before any compute mirror feature_extraction_whisper.py:261-276."""
    from transformers import WhisperFeatureExtractor
    from mlx8_ws_audio_transformer_b200 import LogMelWhisperFeatureExtractor
    ours, hf = LogMelWhisperFeatureExtractor(feature_size=128), WhisperFeatureExtractor(feature_size=128)
    for attr in ("feature_size", "sampling_rate", "hop_length", "chunk_length", "n_fft", "n_samples",
                 "nb_max_frames", "padding_value", "model_input_names"):
        assert getattr(ours, attr) == getattr(hf, attr), attr
    assert np.array_equal(ours.mel_filters, hf.mel_filters)
    assert isinstance(ours, WhisperFeatureExtractor)
    x = np.zeros(100, np.float32)
    for fe in (ours, hf):
        with pytest.raises(ValueError, match="sampling rate"):
            fe(x, sampling_rate=8000)
        with pytest.raises(ValueError, match="mono-channel"):
            fe(np.zeros((2, 2, 10), np.float32), sampling_rate=16000)
    # the collator path (AB/fineTune.py:107) is pure host glue and works without a GPU
    feats = [{"input_features": np.full((128, 3000), i, np.float32)} for i in range(3)]
    a = ours.pad(feats, return_tensors="pt")["input_features"]
    b = hf.pad(feats, return_tensors="pt")["input_features"]
    assert a.shape == (3, 128, 3000) and (a == b).all()
    assert "lm_variant" in ours.to_dict()


def test_filter_banks_equal_the_libraries(golden_whisper_short, golden_torchaudio):
    from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank, torchaudio_mel_filter_bank
    for nm in (80, 128):
        assert np.abs(slaney_mel_filter_bank(201, nm) - golden_whisper_short[f"fbank{nm}"]).max() < 1e-15
    for hop, nm in ((512, 128), (512, 64)):
        fb = torchaudio_mel_filter_bank(513, 0.0, 8000.0, nm, 16000).numpy()
        assert np.array_equal(fb, golden_torchaudio[f"fb_{hop}_{nm}"])
    torchaudio = pytest.importorskip("torchaudio")
    ref = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 40, 16000, norm="slaney", mel_scale="slaney")
    assert np.array_equal(torchaudio_mel_filter_bank(513, 0.0, 8000.0, 40, 16000, "slaney", "slaney").numpy(), ref.numpy())


def test_generated_mel_code_is_current_and_bakes_the_hf_bank():
    """csrc/tf_mel_gen.cuh holds the Whisper banks' weights as literals: it must be what tools/gen_tf_mel.py
    emits today, and the literals must be the float32 image of WhisperFeatureExtractor.mel_filters."""
    import re
    import subprocess
    import sys

    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tf_mel.py")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(ROOT, "mlx8_ws_audio_transformer_b200", "csrc", "tf_mel_gen.cuh")) as f:
        assert f.read() == gen
    from transformers import WhisperFeatureExtractor
    for nm in (80, 128):
        blk = gen[gen.index(f"struct TfMelPattern<{nm}>"):]
        arr = lambda name: [int(v.rstrip("u"), 0) for v in re.search(name + r"\[NNZ\] = \{([^}]*)\}", blk).group(1).split(",")]
        bins, mels, bits = arr("bin"), arr("mel"), arr("bits")
        fb = WhisperFeatureExtractor(feature_size=nm).mel_filters.astype(np.float32)
        assert len(bits) == int((fb != 0).sum())
        assert np.array_equal(fb[bins, mels].view(np.uint32), np.array(bits, np.uint32))
