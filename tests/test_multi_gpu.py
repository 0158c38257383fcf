"""Clip sharding on real GPUs: one process per GPU over NCCL (needs >= 2 GPUs, otherwise skipped).

Every rank computes its contiguous slice with the CUDA kernel; the optional NCCL all-gather returns the whole
batch on every rank.  Checked bit for bit against the same rank computing the whole batch alone: sharding changes
nothing (SURVEY.md 8e -- the operator has no inter-clip dependency)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from mlx8_ws_audio_transformer_b200 import LogMelFrontend, ShardedFrontend
        from mlx8_ws_audio_transformer_b200 import _native as N
        from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank

        fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, 80), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True, device=rank)
        sf = ShardedFrontend(fe.forward)
        res = {}
        # 1300 clips: 650 per rank, a clip per warp pair / per CTA; 9 clips: 5 + 4, uneven slices, each clip spread over
        # several CTAs (cooperative launch) -- one kernel, so the shards equal the single-GPU result bit for bit
        for n_clips in (1300, 9):
            wave = torch.empty(n_clips, 160000, device=dev)
            wave.normal_(0.0, 0.1, generator=torch.Generator(device=dev).manual_seed(7))     # same values on every rank
            local = sf.forward_local(wave)
            full = sf.all_gather(local, n_clips)                                                 # NCCL over NVLink
            alone = fe.forward(wave)
            sl = sf.local_slice(n_clips)
            res[n_clips] = (bool(torch.equal(local, alone[sl])), bool(torch.equal(full, alone)), tuple(full.shape),
                            fe.kernel_name(local.shape[0], 160000))
            del wave, local, full, alone
        q.put((rank, res, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, None, repr(e)))


def test_sharded_equals_single_gpu_bit_for_bit_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, res, err in out:
        assert err is None, (rank, err)
        for n_clips, (local_ok, full_ok, shape, kname) in res.items():
            assert local_ok and full_ok, (rank, n_clips)
            assert shape == (n_clips, 80, 1000)
        assert "logmel_tf_kernel" in res[1300][3] and "logmel_tf_kernel" in res[9][3]
