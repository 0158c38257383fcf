"""Clip sharding with world_size 2 on the gloo backend (no GPU): slices and the optional gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mlx8_ws_audio_transformer_b200 import ShardedFrontend

    def op(w):   # stand-in operator with the real output contract [b, n_mels, frames]
        return w[:, None, :16].repeat(1, 3, 1) * 2.0 + 1.0

    g = torch.Generator().manual_seed(0)
    wave = torch.randn(n_clips, 64, generator=g)
    sf = ShardedFrontend(op)
    local = sf.forward_local(wave)
    full = sf.all_gather(local, n_clips)
    ok = torch.equal(full, op(wave)) and local.shape[0] == sf.local_slice(n_clips).stop - sf.local_slice(n_clips).start
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7, 1])
def test_sharded_forward_and_gather_gloo(n_clips):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, rank
        assert shape == (n_clips, 3, 16)
