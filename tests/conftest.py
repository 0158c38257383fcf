"""Shared fixtures.  `-m "not gpu"` runs on the CPU build box, `-m gpu` on a B200."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# the north_star's parity bar on the normalised log-mel
TOL_MAX, TOL_MEAN = 1e-3, 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_whisper_short():
    return np.load(os.path.join(GOLDEN, "whisper_short.npz"))


@pytest.fixture(scope="session")
def golden_whisper_30s():
    return np.load(os.path.join(GOLDEN, "whisper_30s.npz"))


@pytest.fixture(scope="session")
def golden_torchaudio():
    return np.load(os.path.join(GOLDEN, "torchaudio_4s.npz"))


def padded(clip: np.ndarray, n: int) -> np.ndarray:
    out = np.zeros(n, np.float32)
    m = min(len(clip), n)
    out[:m] = clip[:m]
    return out


@pytest.fixture(scope="session")
def emul():
    """Host emulation of the kernel (tests/emul/host_emul.cpp), built with g++ on first use."""
    src = os.path.join(ROOT, "tests", "emul", "host_emul.cpp")
    so = os.path.join(ROOT, "tests", "emul", "libhost_emul.so")
    deps = [src] + [os.path.join(ROOT, "mlx8_ws_audio_transformer_b200", "csrc", f)
                    for f in ("codelets_gen.cuh", "vec_ops.cuh", "logmel_core.cuh", "logmel_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    lib = ctypes.CDLL(so)

    def run(n_fft, hop, pk, wave, fbank, log_mode, log_param, drop_last, lengths=None):
        wave = np.ascontiguousarray(wave, dtype=np.float32)
        B, L = wave.shape
        frames = 1 + L // hop - (1 if drop_last else 0)
        fb = np.ascontiguousarray(fbank, dtype=np.float32)
        out = np.zeros((B, fb.shape[1], frames), np.float32)
        lp = None
        if lengths is not None:
            larr = np.ascontiguousarray(lengths, dtype=np.int32)
            lp = larr.ctypes.data_as(ctypes.c_void_p)
        rc = lib.emul_logmel(n_fft, hop, pk, wave.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(B),
                             ctypes.c_long(L), lp, L, frames, fb.ctypes.data_as(ctypes.c_void_p),
                             fb.shape[1], log_mode, ctypes.c_float(log_param),
                             out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, rc
        return out

    return run


@pytest.fixture(scope="session")
def golden_refwav():
    return np.load(os.path.join(GOLDEN, "refwav.npz"))


@pytest.fixture(scope="session")
def golden_cfg():
    return np.load(os.path.join(GOLDEN, "whisper_cfg.npz"))
