#!/usr/bin/env python3
"""Copy-only probe: the host <-> device ceiling that bounds the end-to-end (host-buffer) figure.

    python tools/copy_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/copy_probe.py

Every rank moves what one e2e step of bench.py moves -- 512 clips x (1.92 MB waveform in, 1.536 MB features out,
128 mels) -- between pinned host memory and its GPU, H2D and D2H on two streams at the same time, with NO kernel in
between, all ranks at once.  The aggregate clips/s this allows is the ceiling of `e2e` at that GPU count; bench.py's
e2e divided by it says how much of the ceiling lm_forward_host's three-stream pipeline reaches.
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    clips, n_in, n_out = 512, 480000, 128 * 3000
    hx = torch.empty((clips, n_in), dtype=torch.float32).pin_memory()
    hy = torch.empty((clips, n_out), dtype=torch.float32).pin_memory()
    dx = torch.empty((clips, n_in), dtype=torch.float32, device=dev)
    dy = torch.empty((clips, n_out), dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(h2d=True, d2h=True):
        if h2d:
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)

    res = {}
    for name, kw in (("both", {}), ("h2d_only", {"d2h": False}), ("d2h_only", {"h2d": False})):
        for _ in range(2):
            step(**kw)
        barrier()
        t0 = time.perf_counter()
        n = 8
        for _ in range(n):
            step(**kw)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        t = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
        res[name] = {"clips_per_s": world * clips * n / el,
                     "h2d_GBps_total": (world * clips * n_in * 4 * n / el / 1e9) if kw.get("h2d", True) else 0.0,
                     "d2h_GBps_total": (world * clips * n_out * 4 * n / el / 1e9) if kw.get("d2h", True) else 0.0}
        barrier()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "clips_per_step_per_gpu": clips, "pinned": True, "host_cpus": os.cpu_count(), **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
