"""Feature store / collator: host logic on the CPU, device path under -m gpu."""
import os

import numpy as np
import pytest
import torch

from mlx8_ws_audio_transformer_b200 import DeviceCollator, PinnedFeatureWriter


def test_writer_parquet_round_trip_reads_like_the_reference(tmp_path):
    """the reference's reader: np.array(row["log_mel_flat"]).reshape(tuple(row["log_mel_shape"])) (spectrogram.py:204-212)"""
    import pandas as pd

    rng = np.random.default_rng(0)
    feats = rng.standard_normal((7, 16, 11)).astype(np.float32)
    w = PinnedFeatureWriter(8, 16, 11, pin=False)
    w.write(feats[:3])
    w.write(torch.from_numpy(feats[3:7]))
    with pytest.raises(ValueError):
        w.write(feats[:2])                                    # would overflow the 8-row store
    with pytest.raises(ValueError):
        w.write(np.zeros((1, 16, 12), np.float32))
    path = os.path.join(tmp_path, "processed.parquet")
    w.to_parquet(path, columns={"class_id": list(range(7)), "fold": [1] * 7})
    df = pd.read_parquet(path)
    assert list(df.columns) == ["class_id", "fold", "log_mel_flat", "log_mel_shape"]
    for i in range(7):
        row = df.iloc[i]
        got = np.array(row["log_mel_flat"], dtype=np.float32).reshape(tuple(row["log_mel_shape"]))
        assert np.array_equal(got, feats[i]) and int(row["class_id"]) == i
    table = w.to_arrow()
    assert table.column("log_mel_flat").chunk(0).values.to_numpy(zero_copy_only=True).ctypes.data == w.numpy().ctypes.data


def test_device_collator_matches_the_reference_collator_semantics():
    f = [{"input_features": torch.full((4, 6), float(i)), "labels": [50258, 7, 8, 9][: 2 + i]} for i in range(3)]
    out = DeviceCollator(decoder_start_token_id=50258)(f)
    assert out["input_features"].shape == (3, 4, 6)
    # labels: right-padded with -100, then the common leading start token is cut (fineTune.py:111-115)
    assert out["labels"].tolist() == [[7, -100, -100], [7, 8, -100], [7, 8, 9]]
    g = [{"input_features": np.zeros((4, 6), np.float32), "labels": [1, 2]}, {"input_features": np.zeros((4, 6), np.float32), "labels": [3]}]
    out = DeviceCollator(decoder_start_token_id=50258)(g)
    assert out["labels"].tolist() == [[1, 2], [3, -100]]


@pytest.mark.gpu
def test_store_and_collator_on_the_device(tmp_path):
    from mlx8_ws_audio_transformer_b200 import LogMelSpectrogram, LogMelWhisperFeatureExtractor, synth
    w, lengths = synth.urbansound_clips(40, seed=9)
    logm = LogMelSpectrogram(sample_rate=16000, n_fft=1024, hop_length=512, n_mels=128, f_min=0, f_max=8000, power=2.0).to("cuda")
    store = PinnedFeatureWriter(40, 128, 126)
    ref = []
    for s in range(0, 40, 16):                                 # batches computed and drained concurrently
        y = logm(torch.from_numpy(w[s:s + 16]).cuda())
        store.write(y)
        ref.append(y)
    got = store.numpy()
    assert store.buffer.is_pinned() and np.array_equal(got, torch.cat(ref).cpu().numpy())
    fe = LogMelWhisperFeatureExtractor(feature_size=80)
    clips = [torch.from_numpy(synth.gaussian_clips(1, 16000, seed=i)[0]).cuda() for i in range(3)]
    feats = fe(clips, sampling_rate=16000, max_length=16000, return_tensors="pt")["input_features"]
    batch = DeviceCollator(50258)([{"input_features": feats[i], "labels": [50258, 5, 6]} for i in range(3)])
    assert batch["input_features"].is_cuda and torch.equal(batch["input_features"], feats) and batch["labels"].is_cuda
