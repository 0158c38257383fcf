"""Clip sharding over the GPUs of one node.

The operator has no inter-clip dependency (the Whisper maximum is per clip,
feature_extraction_whisper.py:156-158), so the batch is cut into contiguous slices, one per
rank, each computed on that rank's GPU with no data-path collective.  One process per GPU
(``torch.distributed`` over NCCL for the plumbing); the only collective is the OPTIONAL
all-gather of the finished features for a consumer that needs all of them on every rank.
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_clips: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice ``[lo, hi)`` of rank ``rank``; the remainder goes to the low ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(int(n_clips), world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_clips: int, world_size: int):
    return [shard_bounds(n_clips, world_size, r)[1] - shard_bounds(n_clips, world_size, r)[0]
            for r in range(world_size)]


class ShardedFrontend:
    """Run a per-rank operator on this rank's slice; optionally gather the features.

    ``op`` is any callable ``wave[b, T] -> feats[b, n_mels, frames]`` (a ``LogMelFrontend.forward``,
    a ``MelSpectrogram`` module, or -- in the CPU tests -- a stand-in), so the slicing and the
    collective are testable with the ``gloo`` backend and no GPU.
    """

    def __init__(self, op, group=None):
        import torch.distributed as dist

        self.op = op
        self.group = group
        self.dist = dist
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world_size = 0, 1

    def local_slice(self, n_clips: int) -> slice:
        lo, hi = shard_bounds(n_clips, self.world_size, self.rank)
        return slice(lo, hi)

    def forward_local(self, wave_global):
        """``wave_global[B, T]`` is addressable on every rank (or a view of it); compute our slice."""
        return self.op(wave_global[self.local_slice(wave_global.shape[0])])

    def all_gather(self, feats_local, n_clips: int):
        """Optional: ``[b_r, n_mels, frames]`` on every rank -> ``[B, n_mels, frames]`` on every rank.

        Uneven slices are padded to the largest one for ``all_gather_into_tensor`` and trimmed
        afterwards.  Kept out of every throughput figure (at N=8 it moves >10x
        the bytes the kernel does).
        """
        import torch

        if self.world_size == 1:
            return feats_local
        sizes = shard_sizes(n_clips, self.world_size)
        bmax = max(sizes)
        pad = feats_local
        if feats_local.shape[0] < bmax:
            pad = torch.zeros((bmax,) + tuple(feats_local.shape[1:]), dtype=feats_local.dtype,
                              device=feats_local.device)
            pad[: feats_local.shape[0]] = feats_local
        gathered = torch.empty((self.world_size * bmax,) + tuple(feats_local.shape[1:]),
                               dtype=feats_local.dtype, device=feats_local.device)
        self.dist.all_gather_into_tensor(gathered, pad.contiguous(), group=self.group)
        if all(s == bmax for s in sizes):
            return gathered
        parts = [gathered[r * bmax: r * bmax + sizes[r]] for r in range(self.world_size)]
        return torch.cat(parts, dim=0)
