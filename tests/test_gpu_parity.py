"""Parity of the CUDA path, called through the C ABI, with the golden vectors and the oracle.

Bar (BASELINE.json north_star): max-abs 1e-3 and mean-abs 1e-5 on the normalised log-mel.
The oracle is only the checker here; every number under test comes from liblogmel_b200.so.
"""
import numpy as np
import pytest
import torch

from conftest import TOL_MAX, TOL_MEAN, padded
from mlx8_ws_audio_transformer_b200 import (LogMelFrontend, LogMelSpectrogram, LogMelWhisperFeatureExtractor,
                                            MelSpectrogram, launch_count, synth)
from mlx8_ws_audio_transformer_b200 import _native as N
from oracle import logmel_oracle as O

pytestmark = pytest.mark.gpu


def _assert_parity(got, ref, what, tmax=TOL_MAX, tmean=TOL_MEAN):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), what
    mx, mean = O.parity(got, ref)
    assert mx < tmax and mean < tmean, (what, mx, mean)
    return mx, mean


@pytest.fixture(scope="module")
def fronts():
    cache = {}

    def get(nm, variant):
        key = (nm, variant)
        if key not in cache:
            cache[key] = LogMelFrontend(400, 160, O.slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM,
                                        1e-10, True, variant=variant)
        return cache[key]

    return get


# ---------------------------------------------------------------------------------------------
# golden vectors (outputs of the live HF / torchaudio calls, committed under tests/golden)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nm", [80, 128])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_whisper_short_golden(golden_whisper_short, fronts, nm, variant):
    g = golden_whisper_short
    names = [str(n) for n in g["names"]]
    x = np.stack([padded(g[f"in_{k}"], 16000) for k in names])
    n0 = launch_count()
    got = fronts(nm, variant).forward(torch.from_numpy(x).cuda()).cpu().numpy()
    assert launch_count() == n0 + 1                       # one fused kernel, nothing else
    for i, k in enumerate(names):
        _assert_parity(got[i], g[f"feat{nm}"][i], k)
    z = names.index("zeros")
    assert np.abs(got[z] + 1.5).max() < 1e-6              # silence: -1.5 everywhere


@pytest.mark.parametrize("nm", [80, 128])
def test_whisper_dropin_call_golden(golden_whisper_short, nm):
    """the reference's own call form: extractor(list_of_arrays, sampling_rate=16000, ...)"""
    g = golden_whisper_short
    names = [str(n) for n in g["names"]]
    fe = LogMelWhisperFeatureExtractor(feature_size=nm)
    out = fe([g[f"in_{k}"] for k in names], sampling_rate=16000, max_length=16000)
    feats = out["input_features"]
    assert isinstance(feats, np.ndarray) and feats.dtype == np.float32
    _assert_parity(feats, g[f"feat{nm}"], "dropin")
    one = fe(g["in_gauss0"], sampling_rate=16000, max_length=16000, return_tensors="pt")["input_features"]
    assert isinstance(one, torch.Tensor) and one.shape == (1, nm, 100) and one.device.type == "cpu"
    _assert_parity(one[0], g[f"feat{nm}"][names.index("gauss0")], "single")
    # HF glue path (padding='longest' is not the fast path) must agree with HF semantics too
    slow = fe([g["in_gauss0"], g["in_len399"]], sampling_rate=16000, padding="longest", return_tensors="np")
    assert slow["input_features"].shape == (2, nm, 100)
    _assert_parity(slow["input_features"][0], g[f"feat{nm}"][names.index("gauss0")], "longest")


@pytest.mark.parametrize("nm", [80, 128])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_whisper_30s_golden(golden_whisper_30s, fronts, nm, variant):
    g = golden_whisper_30s
    x = np.concatenate([synth.gaussian_clips(3, seed=0), synth.midi_piano_clips(2, seed=0)[0],
                        synth.sine_clip(440.0)[None], synth.chirp_clip()[None]])
    cmax = torch.empty(7, device="cuda")
    got = fronts(nm, variant).forward(torch.from_numpy(x).cuda(), clip_max=cmax).cpu().numpy()
    assert got.shape == (7, nm, 3000)
    _assert_parity(got[:, :, ::int(g["slice"])], g[f"feat{nm}"], "30s")
    assert np.abs(got.reshape(7, -1).max(axis=1) - g[f"max{nm}"]).max() < 1e-5
    assert np.abs(got.reshape(7, -1).mean(axis=1, dtype=np.float64) - g[f"mean{nm}"]).max() < 1e-5
    assert np.abs((cmax.cpu().numpy() + 4) / 4 - g[f"max{nm}"]).max() < 1e-5


@pytest.mark.parametrize("hop,nm", [(512, 128), (128, 128), (512, 64)])
def test_torchaudio_golden(golden_torchaudio, hop, nm):
    g = golden_torchaudio
    w, lengths = synth.urbansound_clips(6, seed=0)
    w[5] = 0.0
    kw = dict(sample_rate=16000, n_fft=1024, hop_length=hop, n_mels=nm, f_min=0, f_max=8000, power=2.0)
    logm = LogMelSpectrogram(**kw).to("cuda")
    assert np.array_equal(logm.fb.cpu().numpy(), g[f"fb_{hop}_{nm}"])
    wt = torch.from_numpy(w).cuda()
    _assert_parity(logm(wt), g[f"logmel_{hop}_{nm}"], "logmel")
    # per-file call shape of spectrogram.py:160-162: [1, 64000] -> [1, n_mels, frames]
    one = logm(wt[2:3])
    assert one.shape == (1, nm, 1 + 64000 // hop)
    _assert_parity(one, g[f"logmel_{hop}_{nm}"][2:3], "single")
    # lengths instead of materialised zero padding
    dirty = wt.clone()
    for i, n in enumerate(lengths):
        dirty[i, int(n):] = 7.0
    _assert_parity(logm(dirty, lengths=torch.from_numpy(lengths)), g[f"logmel_{hop}_{nm}"], "lengths")
    # raw mel power (MelSpectrogram.forward) then the caller's own torch.log, as the reference writes it
    mel = MelSpectrogram(**kw).to("cuda")(wt[:3])
    ref = g[f"mel_{hop}_{nm}"]
    assert np.abs(mel.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, float(np.abs(ref).max()))
    _assert_parity(torch.log(mel + 1e-6), g[f"logmel_{hop}_{nm}"][:3], "log(mel)")
    # CPU tensor in -> CPU tensor out (same device as the input, like torchaudio)
    host = logm(torch.from_numpy(w[:2]))
    assert host.device.type == "cpu"
    _assert_parity(host, g[f"logmel_{hop}_{nm}"][:2], "host")


# ---------------------------------------------------------------------------------------------
# oracle on seeded inputs, edge cases
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nm", [80, 128])
def test_whisper_vs_oracle_mixed_batch(fronts, nm):
    n = 480000
    clips = [synth.gaussian_clips(1, n, seed=21)[0], synth.sine_clip(7000.0), synth.impulse_clip(0),
             synth.impulse_clip(n - 1), synth.int16_uniform_clip(), np.zeros(n, np.float32),
             synth.midi_piano_clips(1, seed=5)[0][0], (synth.gaussian_clips(1, n, seed=22)[0] * 30).astype(np.float32),
             synth.chirp_clip(), (synth.gaussian_clips(1, n, seed=23)[0] * 1e-4).astype(np.float32), np.full(n, 0.3, np.float32)]
    x = np.stack(clips)                                             # odd batch of 11
    ref = O.whisper_logmel(x, n_mels=nm)
    # every kernel: 0 = warp-specialised CTA (what a batch of 11 takes by default), 1 / 2 = phase-synchronous
    # CTA, 3 = thread-per-frame kernel (what batches >= 4 clips per SM take by default)
    for variant in (0, 1, 2, 3):
        got = fronts(nm, variant).forward(torch.from_numpy(x).cuda())
        for i in range(len(clips)):
            _assert_parity(got[i], ref[i], f"clip{i} v{variant}")
        assert np.abs(got[5].cpu().numpy() + 1.5).max() < 1e-6      # the all-zero clip: exactly -1.5


@pytest.mark.parametrize("variant", [2, 3])
def test_lengths_truncation_and_short_clips(fronts, variant):
    fe = fronts(80, variant)
    n = 480000
    x = synth.gaussian_clips(4, n + 1000, seed=31)                  # longer than the container
    lengths = np.array([n + 1000, 1, 399, 250001], np.int32)
    ref = O.whisper_logmel([x[i, :lengths[i]] for i in range(4)], n_mels=80)
    got = fe.forward(torch.from_numpy(x).cuda(), lengths=torch.from_numpy(lengths).cuda(), n_samples=n)
    assert got.shape == (4, 80, 3000)
    _assert_parity(got, ref, "lengths")
    # a [B, T<L] tensor is zero padded on the device, never on the host
    short = synth.gaussian_clips(2, 70000, seed=32)
    _assert_parity(fe.forward(torch.from_numpy(short).cuda(), n_samples=n), O.whisper_logmel(short, n_mels=80), "short")
    # other container lengths (HF max_length=...), including one that is not a multiple of hop or 4
    for L in (16000, 16100, 201, 48000 + 7):
        w = synth.gaussian_clips(3, L, seed=L)
        _assert_parity(fe.forward(torch.from_numpy(w).cuda()), O.whisper_logmel(w, n_mels=80, n_samples=L), f"L={L}")


@pytest.mark.parametrize("variant", [0, 2, 3])
def test_batch_position_group_size_and_paths_are_bit_identical(fronts, variant):
    """A clip's features do not depend on batch size, batch position, CTA grouping or entry point
    (within one kernel: variant 0 keeps batches this small on the CTA-tiled kernel, variant 3 forces
    the thread-per-frame kernel for all of them)."""
    fe = fronts(128, variant)
    x = torch.from_numpy(synth.gaussian_clips(40, seed=41)).cuda()
    full = fe.forward(x)
    torch.cuda.synchronize()
    assert torch.equal(fe.forward(x[7:8])[0], full[7])             # B=1: the clip is spread over 47 CTAs
    assert torch.equal(fe.forward(x[3:9]), full[3:9])              # small batch: larger groups
    perm = torch.randperm(40, generator=torch.Generator().manual_seed(0)).cuda()
    assert torch.equal(fe.forward(x[perm]), full[perm])
    host = fe.forward_host(x[:5].cpu().numpy())
    assert np.array_equal(host, full[:5].cpu().numpy())
    pinned = x[:70].cpu() if x.shape[0] >= 70 else torch.cat([x, x])[:70].cpu()
    out_h = fe.forward_host(pinned.pin_memory())                    # > 1 chunk through the 3-stream pipeline
    ref = fe.forward(pinned.cuda())
    assert torch.equal(out_h, ref.cpu())
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):                                      # honours the caller's stream
        y = fe.forward(x[:4])
    s.synchronize()
    assert torch.equal(y, full[:4])


def test_full_size_properties():
    """Config-2 shape at a size the GPU test can afford: 1184 x 30 s clips, 128 mels -- the default
    dispatch sends this batch (and its 592-clip shard below) to the thread-per-frame kernel."""
    fe = LogMelFrontend(400, 160, O.slaney_mel_filter_bank(201, 128), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True)
    B = 1184
    assert "logmel_tf_kernel<128, 3000, 0>" in fe.kernel_name(B, 480000)         # a clip per warp pair
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, 480000, generator=g, device="cuda") * 0.1
    x[17] = 0.0
    x[100] *= 50.0
    y = fe.forward(x)
    torch.cuda.synchronize()
    assert y.shape == (B, 128, 3000) and torch.isfinite(y).all()
    flat = y.reshape(B, -1)
    mx, mn = flat.max(dim=1).values, flat.min(dim=1).values
    assert (mx - mn <= 2.0 + 1e-6).all()                           # the max-8 clamp: range <= 8/4
    assert torch.allclose(y[17], torch.full_like(y[17], -1.5), atol=1e-6)
    # scaling a clip by c shifts the un-clamped log-mel by 2 log10(c) / 4
    z = fe.forward((x[100] / 50.0)[None])
    d = (y[100] - z[0])
    keep = (z[0] > z[0].min() + 1e-3) & (y[100] > y[100].min() + 1e-3)
    assert (d[keep] - 2 * np.log10(50.0) / 4).abs().max() < 1e-4
    sub = [0, 17, 100, 311, 591, 1183]
    ref = O.whisper_logmel(x[sub].cpu().numpy(), n_mels=128)
    _assert_parity(y[sub], ref, "subset")
    assert torch.equal(fe.forward(x[592:]), y[592:])               # contiguous shard == slice of the whole
    # the dispatch threshold (2/3 clip per SM) and the clip-per-CTA work distribution mid-size batches take:
    # 1184 clips went a clip per warp pair, these go a clip per CTA -- same kernel, same bits
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    # work distributions of the kernel: 1184 clips went a clip per warp pair; mid-size batches go a clip per CTA,
    # small ones a clip over several CTAs (cooperative launch) -- same kernel, same bits
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    for nb in (1, 3, 32, n_sm // 2, n_sm // 2 + 1, n_sm + 3, 4 * n_sm + 8):
        assert "logmel_tf_kernel" in fe.kernel_name(nb, 480000)
        assert torch.equal(fe.forward(x[5:5 + nb]), y[5:5 + nb]), nb
    assert "logmel_ws_kernel" in fe.kernel_name(8, 16000)           # a handful of 1 s clips: the CTA-tiled kernel
    # the CTA-tiled kernel (here: forced by making the clips unaligned) agrees to float32 rounding, not bit for bit
    xu = torch.empty(40 * 480001 + 1, device="cuda")[1:].view(40, 480001)[:, :480000]
    xu.copy_(x[:40])
    assert (fe.forward(xu) - y[:40]).abs().max() < 2e-6


@pytest.mark.parametrize("kind", ["float32", "pcm16", "pcm16_stereo"])
def test_guard_bands_thread_per_frame_kernel(fronts, kind):
    """(compute-sanitizer is not available on the GPU pool.)  Output and clip-max buffers sit inside larger
    allocations filled with a sentinel, the waveform rows are followed by NaN / extreme padding that `lengths`
    hides: nothing outside the buffers may change, nothing from the padding may reach the features.  700 ragged
    clips of 3.3 s -> the thread-per-frame kernel, run-time frame count, clip-edge and all-padding tiles."""
    fe = fronts(80, 3)
    B, T, stride = 700, 52800, 52800 + 64
    g = torch.Generator(device="cuda").manual_seed(5)
    lengths = torch.randint(0, T + 1, (B,), generator=g, device="cuda", dtype=torch.int32)
    lengths[:4] = torch.tensor([0, 1, T, 399], dtype=torch.int32)
    clean = torch.randn(B, T, generator=g, device="cuda") * 0.2
    mask = torch.arange(T, device="cuda")[None, :] < lengths[:, None]
    if kind == "float32":
        buf = torch.full((B, stride), float("nan"), device="cuda")
        buf[:, :T] = torch.where(mask, clean, torch.full_like(clean, float("nan")))
        x = buf[:, :T]                                             # row stride 52864 floats, 16-byte aligned rows
        want = fe.forward(torch.where(mask, clean, torch.zeros_like(clean)))
    else:
        ch = 2 if kind == "pcm16_stereo" else 1
        pcm = (torch.randn(B, T, ch, generator=g, device="cuda") * 6000).clamp_(-32768, 32767).to(torch.int16)
        m3 = mask[:, :, None].expand(B, T, ch)
        buf = torch.full((B, stride, ch), 32767, device="cuda", dtype=torch.int16)
        buf[:, :T] = torch.where(m3, pcm, torch.full_like(pcm, -32768))
        x = buf[:, :T] if ch == 2 else buf[:, :T, 0]
        mono = torch.where(m3, pcm, torch.zeros_like(pcm)).float().sum(-1) / (32768.0 * ch)
        want = fe.forward(mono)
    n_frames = T // 160
    SENT = 1234.5
    big = torch.full((B + 2, 80, n_frames), SENT, device="cuda")
    cm = torch.full((B + 2,), SENT, device="cuda")
    assert "logmel_tf_kernel<80, 0, 1>" in fe.kernel_name(B, T)                    # run-time frame count, a clip per CTA
    got = fe.forward(x, lengths=lengths, out=big[1:B + 1], clip_max=cm[1:B + 1])
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert torch.equal(got, want)
    assert (big[0] == SENT).all() and (big[B + 1] == SENT).all()
    assert cm[0] == SENT and cm[B + 1] == SENT and torch.isfinite(cm[1:B + 1]).all()


@pytest.mark.parametrize("nm", [80, 128])
@pytest.mark.parametrize("how", ["auto", "cta_tiled", "clip_per_pair", "clip_per_cta"])
def test_clamp_everywhere_one_cta_per_clip(nm, how, monkeypatch):
    """> 148 clips (every CTA / warp pair owns whole clips) whose quiet parts lie > 80 dB under one loud
    burst: every tile is revisited by the max-8 pass after the clip maximum is known.  Run on the CTA-tiled
    kernel and on the thread-per-frame kernel with both of its work distributions (the knob is read at lm_create)."""
    variant = 0 if how in ("auto", "cta_tiled") else 3
    if how == "cta_tiled":
        monkeypatch.setenv("LM_TF_MIN_BATCH", "1000000000")          # the warp-specialised CTA-tiled kernel
    if how.startswith("clip_per"):
        monkeypatch.setenv("LM_TF_PAIRS_PER_CLIP", "1" if how == "clip_per_pair" else "4")
    fe = LogMelFrontend(400, 160, O.slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True,
                        variant=variant)
    assert ("logmel_tf_kernel" in fe.kernel_name(300, 480000)) == (how != "cta_tiled")
    B = 300
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, 480000, generator=g, device="cuda") * 1e-5
    t = torch.arange(400, device="cuda") / 16000.0
    pos = torch.randint(0, 479000, (B,), generator=torch.Generator().manual_seed(6))
    pos[7] = pos[9] = 1000                                           # keep these two bursts inside what survives below
    for i in range(B):
        x[i, pos[i]:pos[i] + 400] += 0.9 * torch.sin(2 * np.pi * (300.0 + 20 * i) * t)
    x[7, 200000:] = 0.0                                              # silent tiles inside such a clip
    lengths = torch.full((B,), 480000, dtype=torch.int32, device="cuda")
    lengths[9] = 123457
    y = fe.forward(x, lengths=lengths)
    torch.cuda.synchronize()
    flat = y.reshape(B, -1)
    assert torch.isfinite(y).all()
    assert ((flat.max(dim=1).values - flat.min(dim=1).values) - 2.0).abs().max() < 1e-5   # clamp active in every clip
    sub = [0, 7, 9, 151, 299]
    xs = x[sub].cpu().numpy()
    xs[2, 123457:] = 0.0
    _assert_parity(y[sub], O.whisper_logmel(xs, n_mels=nm), "clamped subset")
    assert torch.equal(fe.forward(x[150:], lengths=lengths[150:]), y[150:])


def test_cls_transformer_consumer_logits():
    """Config 3: features feed the CLS-token encoder of spectrogram.py:944-1057 ([B, n_mels, T] input)."""
    w, lengths = synth.urbansound_clips(8, seed=3)
    kw = dict(sample_rate=16000, n_fft=1024, hop_length=512, n_mels=128, f_min=0, f_max=8000, power=2.0)
    ours = LogMelSpectrogram(**kw).to("cuda")(torch.from_numpy(w).cuda())
    fb = O.htk_mel_filter_bank_f32(513, 128, 0.0, 8000.0, 16000)
    ref = torch.from_numpy(O.torchaudio_mel(w, fb, 1024, 512, 1e-6)).cuda()
    torch.manual_seed(0)
    proj = torch.nn.Linear(128, 128).cuda()
    enc = torch.nn.TransformerEncoder(torch.nn.TransformerEncoderLayer(128, 4, 256, 0.0, batch_first=True), 2).cuda().eval()
    head = torch.nn.Linear(128, 10).cuda()
    cls = torch.zeros(1, 1, 128, device="cuda")

    def logits(f):
        h = proj(f.transpose(1, 2))                                 # [B, T, 128] (spectrogram.py:998-1006)
        h = torch.cat([cls.expand(h.shape[0], -1, -1), h], dim=1)
        return head(enc(h)[:, 0])

    with torch.no_grad():
        a, b = logits(ours), logits(ref)
    assert (a - b).abs().max() < 1e-3


def test_errors_mirror_the_reference():
    fe = LogMelFrontend(400, 160, O.slaney_mel_filter_bank(201, 80), N.LOG10_CLAMP_WHISPER_NORM)
    with pytest.raises(ValueError, match="reflect"):                # torch.stft refuses pad >= length too
        fe.forward(torch.zeros(2, 200, device="cuda"))
    with pytest.raises(ValueError, match="no kernel"):
        LogMelFrontend(512, 128, np.zeros((257, 40), np.float32), N.LOG_NONE)
    with pytest.raises(ValueError, match="banded"):
        LogMelFrontend(1024, 512, np.ones((513, 128), np.float32), N.LOG_NONE)
    with pytest.raises(TypeError):
        fe.forward(np.zeros((1, 16000), np.float32))
    assert fe.forward(torch.zeros(0, 16000, device="cuda")).shape == (0, 80, 100)
    lib = N.lib()
    x = torch.zeros(2, 16000, device="cuda")
    out = torch.empty(2, 80, 100, device="cuda")
    small = torch.empty(8, dtype=torch.uint8, device="cuda")
    rc = lib.lm_forward(fe._h, x.data_ptr(), 2, 16000, 16000, None, out.data_ptr(), None, small.data_ptr(), 8, None)
    assert rc == -5 and b"scratch" in lib.lm_last_error()
    info = fe.kernel_info()
    assert info["n_sm"] >= 100 and info["ctas_per_sm"] >= 1 and info["frames_per_tile"] in (32, 64)
