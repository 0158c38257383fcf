"""The oracle against the golden vectors produced by the live reference libraries."""
import os
import wave as wavmod

import numpy as np
import pytest

from conftest import padded
from mlx8_ws_audio_transformer_b200 import synth
from oracle import logmel_oracle as O


def test_slaney_bank_equals_hf(golden_whisper_short):
    for nm in (80, 128):
        fb = O.slaney_mel_filter_bank(201, nm)
        ref = golden_whisper_short[f"fbank{nm}"]
        assert fb.shape == ref.shape and np.abs(fb - ref).max() < 1e-15
        assert (fb[0] == 0).all() and (fb[200] == 0).all()
        assert ((fb != 0).sum(axis=1) <= 2).all()


def test_htk_bank_bit_exact(golden_torchaudio):
    for hop, nm in ((512, 128), (512, 64)):
        fb = O.htk_mel_filter_bank_f32(513, nm, 0.0, 8000.0, 16000)
        assert np.array_equal(fb, golden_torchaudio[f"fb_{hop}_{nm}"])


@pytest.mark.parametrize("nm", [80, 128])
def test_whisper_short_golden(golden_whisper_short, nm):
    g = golden_whisper_short
    names = [str(n) for n in g["names"]]
    x = np.stack([padded(g[f"in_{k}"], 16000) for k in names])
    got = O.whisper_logmel(x, g[f"fbank{nm}"], n_samples=16000)
    ref = g[f"feat{nm}"]
    assert got.shape == ref.shape == (len(names), nm, 100)
    for i, k in enumerate(names):
        mx, mean = O.parity(got[i], ref[i])
        assert mx < 2e-4 and mean < 2e-6, (k, mx, mean)
    z = names.index("zeros")
    assert np.all(got[z] == -1.5) and np.all(ref[z] == -1.5)         # silence is exactly -1.5


@pytest.mark.parametrize("nm", [80, 128])
def test_whisper_30s_golden(golden_whisper_30s, nm):
    g = golden_whisper_30s
    x = np.concatenate([synth.gaussian_clips(3, seed=0), synth.midi_piano_clips(2, seed=0)[0],
                        synth.sine_clip(440.0)[None], synth.chirp_clip()[None]])
    got = O.whisper_logmel(x, n_mels=nm)
    assert got.shape == (7, nm, 3000)
    mx, mean = O.parity(got[:, :, ::int(g["slice"])], g[f"feat{nm}"])
    assert mx < 2e-4 and mean < 2e-6, (mx, mean)
    assert np.abs(got.reshape(7, -1).max(axis=1) - g[f"max{nm}"]).max() < 1e-5


@pytest.mark.parametrize("hop,nm", [(512, 128), (128, 128), (512, 64)])
def test_torchaudio_golden(golden_torchaudio, hop, nm):
    g = golden_torchaudio
    w, lengths = synth.urbansound_clips(6, seed=0)
    w[5] = 0.0
    fb = g[f"fb_{hop}_{nm}"]
    got = O.torchaudio_mel(w, fb, 1024, hop, log_offset=1e-6)
    assert got.shape == g[f"logmel_{hop}_{nm}"].shape == (6, nm, 1 + 64000 // hop)
    mx, mean = O.parity(got, g[f"logmel_{hop}_{nm}"])
    assert mx < 5e-4 and mean < 5e-6, (mx, mean)
    raw = O.torchaudio_mel(w[:3], fb, 1024, hop, log_offset=None)
    ref = g[f"mel_{hop}_{nm}"]
    assert np.abs(raw - ref).max() <= 2e-5 * max(1.0, float(np.abs(ref).max()))


def test_live_libraries_when_present():
    """In the build container the libraries are importable: check the oracle against them live."""
    transformers = pytest.importorskip("transformers")
    x = np.stack([synth.gaussian_clips(1, 48000, seed=7)[0], synth.sine_clip(1000.0, 48000)])
    fe = transformers.WhisperFeatureExtractor()
    ref = fe(list(x), sampling_rate=16000, max_length=48000, return_tensors="np")["input_features"]
    got = O.whisper_logmel(x, fe.mel_filters, n_samples=48000)
    mx, mean = O.parity(got, ref)
    assert mx < 2e-4 and mean < 2e-6


REF_WAVS = "/root/reference/.charles/samples"


@pytest.mark.skipif(not os.path.isdir(REF_WAVS), reason="reference checkout not present (GPU box)")
def test_reference_wavs_against_hf():
    """The 18 WAVs shipped with the reference (16 kHz stereo s16), each channel as a clip."""
    transformers = pytest.importorskip("transformers")
    paths = []
    for root, _, files in os.walk(REF_WAVS):
        paths += [os.path.join(root, f) for f in files if f.endswith(".wav")]
    assert len(paths) >= 10
    clips = []
    for p in sorted(paths)[:6]:
        with wavmod.open(p, "rb") as wf:
            assert wf.getframerate() == 16000 and wf.getsampwidth() == 2
            pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16).reshape(-1, wf.getnchannels())
        for ch in range(pcm.shape[1]):
            clips.append((pcm[:, ch].astype(np.float32) / 32768.0))
    fe = transformers.WhisperFeatureExtractor()
    ref = fe(clips, sampling_rate=16000, return_tensors="np")["input_features"]
    got = O.whisper_logmel(clips, fe.mel_filters)
    mx, mean = O.parity(got, ref)
    assert mx < 5e-4 and mean < 2e-6, (mx, mean)
