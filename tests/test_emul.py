"""Host emulation of the CUDA kernel (same headers: codelets, index maps, tables) vs golden."""
import numpy as np
import pytest

from conftest import TOL_MAX, TOL_MEAN, padded
from mlx8_ws_audio_transformer_b200 import synth
from oracle import logmel_oracle as O

WHISPER, LN = 1, 2


@pytest.mark.parametrize("nm", [80, 128])
@pytest.mark.parametrize("pk", [1, 2, 3])      # 3 = warp-specialised geometry
def test_emul_whisper_short(emul, golden_whisper_short, nm, pk):
    g = golden_whisper_short
    names = [str(n) for n in g["names"]]
    x = np.stack([padded(g[f"in_{k}"], 16000) for k in names])
    got = emul(400, 160, pk, x, g[f"fbank{nm}"], WHISPER, 1e-10, True)
    for i, k in enumerate(names):
        mx, mean = O.parity(got[i], g[f"feat{nm}"][i])
        assert mx < TOL_MAX and mean < TOL_MEAN, (k, mx, mean)
        assert mx < 3e-4 and mean < 3e-6, (k, mx, mean)            # where the fp32 kernel actually sits


def test_emul_lengths_equal_explicit_padding(emul, golden_whisper_short):
    g = golden_whisper_short
    rng = np.random.default_rng(0)
    x = synth.gaussian_clips(3, 16000, seed=11)
    lengths = np.array([16000, 5000, 399], np.int32)
    xz = x.copy()
    for i, n in enumerate(lengths):
        xz[i, n:] = 0
    x_dirty = x.copy()
    for i, n in enumerate(lengths):
        x_dirty[i, n:] = rng.standard_normal(16000 - n)            # must be ignored
    a = emul(400, 160, 2, xz, g["fbank80"], WHISPER, 1e-10, True)
    b = emul(400, 160, 2, x_dirty, g["fbank80"], WHISPER, 1e-10, True, lengths)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("hop,nm", [(512, 128), (128, 128), (512, 64)])
def test_emul_torchaudio(emul, golden_torchaudio, hop, nm):
    g = golden_torchaudio
    w, lengths = synth.urbansound_clips(6, seed=0)
    w[5] = 0.0
    got = emul(1024, hop, 1, w, g[f"fb_{hop}_{nm}"], LN, 1e-6, False, lengths)
    mx, mean = O.parity(got, g[f"logmel_{hop}_{nm}"])
    assert mx < TOL_MAX and mean < TOL_MEAN, (mx, mean)
    raw = emul(1024, hop, 1, w[:3], g[f"fb_{hop}_{nm}"], 0, 0.0, False)
    ref = g[f"mel_{hop}_{nm}"]
    assert np.abs(raw - ref).max() <= 2e-5 * max(1.0, float(np.abs(ref).max()))
