"""GPU parity at the reference's own call sites and on its own audio fixtures.

* the 18 WAVs of /root/reference/.charles/samples (first 2 s, committed as int16 in tests/golden/refwav.npz with
  the outputs of the live HF / torchaudio calls) through the float path AND the fused int16 / stereo ingest;
* BASELINE config 1 (exact seeded batch) and config 4 (1000 piano clips with per-clip lengths);
* the WhisperFeatureExtractor.__call__ branches (attention mask, do_normalize), a WhisperProcessor composed
  with the drop-in (/root/reference/AB/fineTune.py:62,88), the batched encoder feed that replaces the per-clip
  loop of /root/reference/.charles/music2midi/model.py:94-123, and two streams sharing one cached frontend.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import TOL_MAX, TOL_MEAN
from mlx8_ws_audio_transformer_b200 import (LogMelFrontend, LogMelSpectrogram, LogMelWhisperFeatureExtractor, launch_count,
                                            synth)
from mlx8_ws_audio_transformer_b200 import _native as N
from oracle import logmel_oracle as O

pytestmark = pytest.mark.gpu


def _parity(got, ref, what):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), what
    mx, mean = O.parity(got, ref)
    assert mx < TOL_MAX and mean < TOL_MEAN, (what, mx, mean)


def _front(nm, variant=0):
    return LogMelFrontend(400, 160, O.slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True, variant=variant)


# ---------------------------------------------------------------------------------------------
# the reference's 18 WAV fixtures
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 3])
def test_reference_wavs_whisper(golden_refwav, variant):
    g = golden_refwav
    pcm = g["pcm"]                                            # int16 [18, 32000, 2]
    st = int(g["stride"])
    n = pcm.shape[1]
    chan = pcm.astype(np.float32) / 32768.0                   # torchaudio.load (wavToWhisper.py:52)
    per_channel = np.ascontiguousarray(chan.transpose(0, 2, 1)).reshape(-1, n)     # the [2, N] array is a batch of 2
    mono = chan.mean(axis=2).astype(np.float32)               # spectrogram.py:147-148
    fe80, fe128 = _front(80, variant), _front(128, variant)
    # float path, as the reference's loaders hand the audio over
    ch = fe80.forward(torch.from_numpy(per_channel).cuda())
    _parity(ch[:, :, ::st], g["feat80_channels"], "channels")
    m80 = fe80.forward(torch.from_numpy(mono).cuda())
    _parity(m80[:, :, ::st], g["feat80_mono"], "mono 80")
    m128 = fe128.forward(torch.from_numpy(mono).cuda())
    _parity(m128[:, :, ::st], g["feat128_mono"], "mono 128")
    # fused ingest: the WAV's own int16 stereo frames in, conversion and down-mix in the tile loader
    n0 = launch_count()
    p80 = fe80.forward(torch.from_numpy(pcm).cuda())
    assert launch_count() == n0 + 1
    assert torch.equal(p80, m80)                              # exact: sums of int16 and 2^-16 are exact in float32
    assert torch.equal(fe128.forward(torch.from_numpy(pcm).cuda()), m128)
    # int16 mono (memoToWav.py:19 writes s16 mono): every channel as its own mono clip
    pc = np.ascontiguousarray(pcm.transpose(0, 2, 1)).reshape(-1, n)
    assert torch.equal(fe80.forward(torch.from_numpy(pc).cuda()), ch)
    # host entry point with int16 buffers, lengths shorter than the row, odd (unaligned) row width
    odd = np.ascontiguousarray(pcm[:, :31999])
    lengths = np.full(len(odd), 20001, np.int32)
    got = fe80.forward_host(odd, lengths=lengths, n_samples=n)
    mz = mono.copy()
    mz[:, 20001:] = 0.0
    assert np.array_equal(got, fe80.forward(torch.from_numpy(mz).cuda()).cpu().numpy())
    with pytest.raises(ValueError, match="int16 PCM"):
        fe80.forward(torch.zeros(2, 1000, 3, dtype=torch.int16, device="cuda"))


def test_reference_wavs_torchaudio(golden_refwav):
    """mono mix, zero padded to 4 s, MelSpectrogram + log (spectrogram.py:144-162) -- float and fused int16 stereo"""
    g = golden_refwav
    pcm, st = g["pcm"], int(g["stride"])
    mono = (pcm.astype(np.float32) / 32768.0).mean(axis=2).astype(np.float32)
    logm = LogMelSpectrogram(sample_rate=16000, n_fft=1024, hop_length=512, n_mels=128, f_min=0, f_max=8000, power=2.0).to("cuda")
    padded = np.zeros((len(mono), 64000), np.float32)
    padded[:, :mono.shape[1]] = mono
    a = logm(torch.from_numpy(padded).cuda())
    _parity(a[:, :, ::st], g["ta_logmel_512_128"], "padded")
    fe = logm._frontend(torch.cuda.current_device())
    b = fe.forward(torch.from_numpy(pcm).cuda(), n_samples=64000)          # 2 s of stereo PCM, container of 4 s
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------
# BASELINE configs 1 and 4 at their exact sizes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 3])
def test_config1_exact_seeded_batch(golden_cfg, variant):
    g = golden_cfg
    x = synth.gaussian_clips(32, seed=0)                      # default_rng(0), B = 32
    y = _front(80, variant).forward(torch.from_numpy(x).cuda()).cpu().numpy()
    _parity(y[:, :, ::int(g["cfg1_stride"])], g["cfg1_feat"], "cfg1")
    assert np.abs(y.reshape(32, -1).max(axis=1) - g["cfg1_max"]).max() < 1e-5
    assert np.abs(y.reshape(32, -1).mean(axis=1, dtype=np.float64) - g["cfg1_mean"]).max() < 1e-5


@pytest.mark.parametrize("variant", [0, 3])
def test_config4_piano_1000_with_lengths(golden_cfg, variant):
    g = golden_cfg
    w, n = synth.midi_piano_clips(1000, seed=0)
    assert np.array_equal(n, g["cfg4_lengths"])
    dirty = torch.from_numpy(w).cuda()
    lengths = torch.from_numpy(n).cuda()
    for i in range(0, 1000, 97):
        dirty[i, int(n[i]):] = 3.0                            # the tail is padding: never read
    y = _front(80, variant).forward(dirty, lengths=lengths)
    flat = y.reshape(1000, -1)
    assert np.abs(flat.max(dim=1).values.cpu().numpy() - g["cfg4_max"]).max() < 1e-4
    assert np.abs(flat.double().mean(dim=1).cpu().numpy() - g["cfg4_mean"]).max() < 1e-5
    keep = g["cfg4_keep_idx"]
    _parity(y[torch.from_numpy(keep).cuda()][:, :, ::int(g["cfg4_stride"])], g["cfg4_keep_feat"], "cfg4 kept clips")


# ---------------------------------------------------------------------------------------------
# the extractor's other branches, the processor, the batched encoder feed
# ---------------------------------------------------------------------------------------------
def test_attention_mask_and_do_normalize_match_hf(golden_cfg):
    g = golden_cfg
    ragged = [g[f"ragged_in{i}"] for i in range(4)]
    fe = LogMelWhisperFeatureExtractor(feature_size=80)
    o = fe(ragged, sampling_rate=16000, max_length=16000, return_attention_mask=True, return_tensors="np")
    assert np.array_equal(o["attention_mask"], g["ragged_mask"])              # _frame_mask == attention_mask[:, ::160]
    _parity(o["input_features"], g["ragged_feat"], "ragged")
    o = fe(ragged, sampling_rate=16000, max_length=16000, do_normalize=True, return_attention_mask=True, return_tensors="np")
    assert np.array_equal(o["attention_mask"], g["ragged_mask"])
    _parity(o["input_features"], g["ragged_feat_normalized"], "do_normalize")
    # CUDA in -> CUDA out, with the mask
    w = torch.zeros(4, 16000, device="cuda")
    for i, c in enumerate(ragged):
        w[i, :len(c)] = torch.from_numpy(c)
    o = fe([torch.from_numpy(c).cuda() for c in ragged], sampling_rate=16000, max_length=16000, return_attention_mask=True,
           return_tensors="pt")
    assert o["input_features"].is_cuda and np.array_equal(o["attention_mask"].cpu().numpy(), g["ragged_mask"])
    _parity(o["input_features"], g["ragged_feat"], "ragged cuda list")


def test_whisper_processor_composition(tmp_path, golden_whisper_short):
    """processor = WhisperProcessor(feature_extractor, tokenizer); processor(audio, sampling_rate=..., text=...)
    as at /root/reference/AB/fineTune.py:62,88 -- with the drop-in extractor and an offline toy tokenizer."""
    from transformers import WhisperProcessor, WhisperTokenizer

    vocab = {"<|endoftext|>": 0, "<|startoftranscript|>": 1, "<|notimestamps|>": 2, "h": 3, "i": 4, "hi": 5}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text("#version: 0.2\nh i\n")
    tok = WhisperTokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    proc = WhisperProcessor(feature_extractor=LogMelWhisperFeatureExtractor(feature_size=80), tokenizer=tok)
    g = golden_whisper_short
    names = [str(n) for n in g["names"]]
    out = proc(g["in_sine440"], sampling_rate=16000, text="hi", max_length=16000)       # prepare_dataset(), fineTune.py:88
    assert list(out["labels"]) == [1, 2, 5, 0]
    feats = np.asarray(out["input_features"])
    assert feats.shape == (1, 80, 100) and feats.dtype == np.float32
    _parity(feats[0], g["feat80"][names.index("sine440")], "processor")
    pt = proc(g["in_gauss0"], sampling_rate=16000, return_tensors="pt", max_length=16000)   # wavToWhisper.py:55
    assert isinstance(pt["input_features"], torch.Tensor)
    _parity(pt["input_features"][0], g["feat80"][names.index("gauss0")], "processor pt")
    # the collator's pad() on already extracted features (fineTune.py:107)
    batch = proc.feature_extractor.pad([{"input_features": feats[0]}, {"input_features": feats[0]}], return_tensors="pt")
    assert batch["input_features"].shape == (2, 80, 100)


def test_batched_encoder_feed_equals_the_per_clip_loop():
    """WhisperAudioEncoder.forward walks the batch clip by clip through the processor
    (/root/reference/.charles/music2midi/model.py:94-123); a list of ragged CUDA waveforms is one launch here."""
    fe = LogMelWhisperFeatureExtractor(feature_size=80)
    rng = np.random.default_rng(3)
    clips = [(rng.standard_normal(L) * 0.1).astype(np.float32) for L in (160000, 479999, 1, 480000, 250003, 777)]
    dev = [torch.from_numpy(c).cuda() for c in clips]
    fe(dev[:1], sampling_rate=16000, return_tensors="pt")                     # handle creation outside the launch count
    n0 = launch_count()
    batched = fe(dev, sampling_rate=16000, return_tensors="pt")["input_features"]
    assert launch_count() == n0 + 1 and batched.is_cuda and batched.shape == (6, 80, 3000)
    loop = torch.cat([fe(c, sampling_rate=16000, return_tensors="pt")["input_features"] for c in dev])   # the reference's loop
    assert torch.equal(batched, loop)
    _parity(batched, O.whisper_logmel(clips, n_mels=80), "batched feed")


def test_two_streams_share_one_cached_frontend():
    """the extractor caches its frontend process-wide; calls on different streams must not share the
    per-clip counters the small-batch (clip group > 1) kernel spins on."""
    fe = _front(128)
    x = torch.from_numpy(synth.gaussian_clips(12, seed=77)).cuda()
    ref = fe.forward(x)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for it in range(20):
        with torch.cuda.stream(s1):
            a = fe.forward(x[:5])
        with torch.cuda.stream(s2):
            b = fe.forward(x[5:])
        outs.append((a, b))
    s1.synchronize()
    s2.synchronize()
    for a, b in outs:
        assert torch.equal(a, ref[:5]) and torch.equal(b, ref[5:])


def test_out_and_length_arguments_are_validated():
    fe = _front(80)
    x = torch.zeros(2, 16000, device="cuda")
    with pytest.raises(ValueError, match="out must be"):
        fe.forward(x, out=torch.empty(2, 80, 99, device="cuda"))
    with pytest.raises(ValueError, match="out must be"):
        fe.forward(x, out=torch.empty(2, 80, 100, device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError, match="clip_max must be"):
        fe.forward(x, clip_max=torch.empty(3, device="cuda"))
    with pytest.raises(ValueError, match="out must be"):
        fe.forward_host(np.zeros((2, 16000), np.float32), out=np.zeros((2, 80, 101), np.float32))
    # lengths beyond the row are clamped to the row, not read past it
    big = torch.full((2,), 10 ** 6, dtype=torch.int32, device="cuda")
    assert torch.equal(fe.forward(x, lengths=big, n_samples=32000), fe.forward(x, n_samples=32000))
    dev0 = torch.cuda.current_device()
    LogMelFrontend(400, 160, O.slaney_mel_filter_bank(201, 80), N.LOG10_CLAMP_WHISPER_NORM, device=dev0).forward(x)
    assert torch.cuda.current_device() == dev0                                # the library restores the caller's device


def test_openai_whisper_log_mel_semantics():
    """whisper.log_mel_spectrogram (wavToWhisper.py:10-13 -> model.transcribe): whole file + padding, ONE maximum."""
    from mlx8_ws_audio_transformer_b200 import openai_whisper as W
    rng = np.random.default_rng(11)
    audio = (rng.standard_normal(16000 * 47 + 123) * 0.05).astype(np.float32)        # a 47 s "file": longer than one container
    audio[: 16000 * 3] *= 20.0                                                        # its loud start sets the file's maximum
    mel = W.log_mel_spectrogram(audio, n_mels=80, padding=W.N_SAMPLES)               # as transcribe() calls it
    n = len(audio) + W.N_SAMPLES
    assert mel.is_cuda and mel.shape == (80, n // 160)
    ref = O.whisper_logmel(np.concatenate([audio, np.zeros(W.N_SAMPLES, np.float32)])[None], n_mels=80, n_samples=n)[0]
    _parity(mel, ref, "whole file")
    seg = W.pad_or_trim(mel, W.N_FRAMES)                                              # the first 30 s segment fed to the encoder
    assert seg.shape == (80, 3000)
    # the HF container of the same first 30 s has its own (different) maximum: the two frontends are not interchangeable
    short = W.log_mel_spectrogram(audio[:16000 * 5], n_mels=128)
    _parity(short, O.whisper_logmel(audio[None, :16000 * 5], n_mels=128, n_samples=16000 * 5)[0], "5 s, 128 mels")
    assert np.array_equal(W.pad_or_trim(np.arange(5.0), 8), np.array([0, 1, 2, 3, 4, 0, 0, 0.0]))
    assert W.pad_or_trim(torch.arange(10.0), 4).tolist() == [0, 1, 2, 3]


def test_qwen2_audio_extractor_is_the_128_mel_dropin():
    """Qwen2-Audio's processor (qwen2_audio_tests.py:34,51-52) calls WhisperFeatureExtractor(feature_size=128) with
    return_attention_mask=True: the drop-in serves it (oracle for the features, HF's formula for the mask)."""
    fe = LogMelWhisperFeatureExtractor(feature_size=128)
    rng = np.random.default_rng(5)
    clips = [(rng.standard_normal(L) * 0.1).astype(np.float32) for L in (480000, 100001, 31)]
    o = fe(clips, sampling_rate=16000, return_attention_mask=True, padding="max_length", return_tensors="pt")
    assert o["input_features"].shape == (3, 128, 3000) and o["attention_mask"].shape == (3, 3000)
    _parity(o["input_features"], O.whisper_logmel(clips, n_mels=128), "qwen2-audio extractor")
    for i, L in enumerate((480000, 100001, 31)):
        assert int(o["attention_mask"][i].sum()) == (L + 159) // 160
