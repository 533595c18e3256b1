"""Batch sharding of the set-abstraction path across the GPUs of one box: one process per GPU, contiguous
slices of the batch dimension, no data-path collective (every op of pointnet2_utils.py is per cloud in eval
mode; BatchNorm uses running statistics).

The only coupling between clouds in the reference is the ONE torch.randint draw of FPS start indices per
set-abstraction layer (pointnet2_utils.py:59), made for the whole batch.  `GlobalStarts` reproduces that draw
for the global batch on every rank (same seed -> same CPU generator stream) and hands each rank its slice, so
the concatenation of the per-rank outputs is bit-identical to a single-process run of the whole batch.
"""
from __future__ import annotations

import contextlib

import torch

import pointnet2_utils as _pu


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` clouds owned by `rank`; the first total % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GlobalStarts:
    """fps_start_source callable: draws for the GLOBAL batch, returns this rank's slice."""

    def __init__(self, global_batch: int, world: int, rank: int, seed: int | None = None, device=None):
        self.global_batch, self.world, self.rank = global_batch, world, rank
        self.lo, self.hi = shard_bounds(global_batch, world, rank)
        self.gen = None if seed is None else torch.Generator().manual_seed(seed)
        self.device = device
        self.draws = 0

    def __call__(self, B: int, N: int) -> torch.Tensor:
        if B != self.hi - self.lo:
            raise RuntimeError(f"GlobalStarts: local batch {B} != shard size {self.hi - self.lo}")
        full = torch.randint(0, N, (self.global_batch,), dtype=torch.long, generator=self.gen)
        self.draws += 1
        out = full[self.lo:self.hi]
        return out.to(self.device) if self.device is not None else out


@contextlib.contextmanager
def sharded_starts(global_batch: int, world: int, rank: int, seed: int | None = None, device=None):
    src = GlobalStarts(global_batch, world, rank, seed, device)
    with _pu.fps_start_source(src):
        yield src


def forward_shard(model, inputs, world: int, rank: int, seed: int | None = None):
    """Run `model` on this rank's slice of every tensor in `inputs` (each [B_global, ...]); returns the local
    outputs.  No communication: gather on the host (or leave sharded) as the application needs."""
    B = inputs[0].shape[0]
    lo, hi = shard_bounds(B, world, rank)
    local = [t[lo:hi] for t in inputs]
    with torch.no_grad(), sharded_starts(B, world, rank, seed):
        return model(*local)
