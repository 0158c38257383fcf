"""ctypes binding of ``liblogmel_b200.so`` (C ABI declared in ``include/logmel.h``).

The shared library is built in-tree by :func:`build` (``nvcc`` for sm_100a) and loaded from
``csrc/``.  There is deliberately no fallback: if the library is missing, or no CUDA device
is present when a handle is created, the call raises -- a log-mel computed anywhere else
would not be this product.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("LM_LIB_PATH") or os.path.join(CSRC, "liblogmel_b200.so")   # override: tuning builds only
HEADER = os.path.join(os.path.dirname(_HERE), "include", "logmel.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]

# every symbol include/logmel.h declares
SYMBOLS = [
    "lm_version", "lm_last_error", "lm_create", "lm_destroy", "lm_num_frames", "lm_scratch_bytes",
    "lm_forward", "lm_forward_host", "lm_host_register", "lm_host_unregister", "lm_launch_count",
    "lm_kernel_info", "lm_kernel_name", "lm_forward_pcm16", "lm_forward_host_pcm16",
]

LOG_NONE, LOG10_CLAMP_WHISPER_NORM, LN_PLUS_EPS, LOG10_CLAMP = 0, 1, 2, 3


class LmConfig(ctypes.Structure):
    _fields_ = [
        ("n_fft", ctypes.c_int32), ("hop", ctypes.c_int32), ("n_mels", ctypes.c_int32),
        ("log_mode", ctypes.c_int32), ("log_param", ctypes.c_float), ("drop_last", ctypes.c_int32),
        ("device", ctypes.c_int32), ("variant", ctypes.c_int32),
        ("fbank", ctypes.c_void_p), ("window", ctypes.c_void_p),
    ]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + [HEADER]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/logmel_api.cu`` for sm_100a into ``csrc/liblogmel_b200.so``."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build liblogmel_b200.so")
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, os.path.join(CSRC, "logmel_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the log-mel path.")
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, i32p = ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)
        L.lm_version.restype = ctypes.c_int
        L.lm_last_error.restype = ctypes.c_char_p
        L.lm_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(LmConfig)]
        L.lm_create.restype = ctypes.c_int
        L.lm_destroy.argtypes = [vp]
        L.lm_destroy.restype = None
        L.lm_num_frames.argtypes = [vp, i64]
        L.lm_num_frames.restype = i64
        L.lm_scratch_bytes.argtypes = [vp, i64]
        L.lm_scratch_bytes.restype = ctypes.c_size_t
        L.lm_forward.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, ctypes.c_size_t, vp]
        L.lm_forward.restype = ctypes.c_int
        L.lm_forward_host.argtypes = [vp, vp, i64, i64, i64, vp, vp]
        L.lm_forward_host.restype = ctypes.c_int
        L.lm_forward_pcm16.argtypes = [vp, vp, ctypes.c_int32, i64, i64, i64, vp, vp, vp, vp, ctypes.c_size_t, vp]
        L.lm_forward_pcm16.restype = ctypes.c_int
        L.lm_forward_host_pcm16.argtypes = [vp, vp, ctypes.c_int32, i64, i64, i64, vp, vp]
        L.lm_forward_host_pcm16.restype = ctypes.c_int
        L.lm_host_register.argtypes = [vp, ctypes.c_size_t]
        L.lm_host_register.restype = ctypes.c_int
        L.lm_host_unregister.argtypes = [vp]
        L.lm_host_unregister.restype = ctypes.c_int
        L.lm_launch_count.restype = i64
        L.lm_kernel_info.argtypes = [vp, i32p, i32p, i32p, i32p, i32p]
        L.lm_kernel_info.restype = ctypes.c_int
        L.lm_kernel_name.argtypes = [vp, i64, i64]
        L.lm_kernel_name.restype = ctypes.c_char_p
        _lib = L
    return _lib


def last_error() -> str:
    return lib().lm_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    """Map the C status convention onto Python exceptions (negative: argument, positive: CUDA)."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (status {rc})"
    if rc < 0:
        raise ValueError(msg)
    raise RuntimeError(msg)
