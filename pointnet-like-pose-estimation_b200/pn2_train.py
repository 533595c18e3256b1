"""Data-parallel training step for the PointNet++ pose heads (BASELINE.json configs[4]: rotation_msg, B=256, N=2048).

One step of the reference's loop (train_rotation.py:99-133): ``optimizer.zero_grad`` (:107) -> forward in training
mode (:121; BatchNorm on batch statistics, pointnet2_utils.py:170,217) -> ``MSELoss/L1Loss`` (:122, built by
``get_loss(type, reduction)``, rotation_msg.py:40-50) -> ``backward`` (:124) -> ``optimizer.step`` (:128, Adam with
lr 1e-4, betas (0.9, 0.999), eps 1e-8, weight decay 1e-4, :50-57).  The reference is single-GPU; the data-parallel
version keeps its semantics per replica (local BatchNorm statistics -- the reference has no SyncBN) and adds ONE
collective per step:

    all gradients live in one flat fp32 buffer (``GradBucket``: every ``param.grad`` is a view into it, so backward
    writes straight into the bucket), and the bucket is all-reduced (sum, then / world) over NCCL / NVLink in a single
    call on the compute stream right after backward -- 1.74 M elements = 6.97 MB for rotation_msg, i.e. launch-latency
    sized for NVSwitch; splitting it into per-layer buckets would only add launches.

Sampling and grouping (FPS, ball query, neighbour gather / scatter-add) run on the hand-written kernels in every mode; in
training mode the shared MLP runs layer by layer on the grouped rows: the 1x1 convolutions, their input and weight gradients
on our tcgen05 row GEMMs (TF32, the policy of the reference's cuDNN convolutions; csrc/pn2_gemm.cu), batch-statistics
BatchNorm + ReLU + max-pool forward and backward on the fused kernels of csrc/pn2_bn_train.cu, wired into autograd by
``torch.autograd.Function``s in pointnet2_utils.py.  The FPS start indices are drawn on the CPU generator once per layer like
the reference (:59); ``pn2_shard.GlobalStarts`` makes every rank slice one global draw.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

__all__ = ["GradBucket", "get_loss", "make_optimizer", "train_step"]


class get_loss(nn.Module):
    """rotation_msg.py:40-50 (identical in every pose head): MSELoss for 'L2_loss', L1Loss otherwise."""

    def __init__(self, type="L2_loss", reduction="mean"):
        super().__init__()
        self.loss = nn.MSELoss(reduction=reduction) if type == "L2_loss" else nn.L1Loss(reduction=reduction)

    def forward(self, pred, target):
        return self.loss(pred, target)


def make_optimizer(model: nn.Module, learning_rate: float = 1e-4, decay_rate: float = 1e-4) -> torch.optim.Optimizer:
    """train_rotation.py:50-57 with the script's defaults."""
    return torch.optim.Adam(model.parameters(), lr=learning_rate, betas=(0.9, 0.999), eps=1e-08, weight_decay=decay_rate)


class GradBucket:
    """One flat fp32 gradient buffer for a module; ``param.grad`` of every trainable parameter is a view into it.

    ``zero()`` replaces ``optimizer.zero_grad()`` (one memset, the views stay attached); ``all_reduce_mean()`` is the
    step's only collective.  Works on any device / backend (NCCL on the GPUs, gloo in the CPU tests)."""

    def __init__(self, params: Iterable[nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("GradBucket: parameters must share one device and dtype")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        """True while every parameter's .grad still aliases the bucket (``zero_grad(set_to_none=True)`` detaches them)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def all_reduce_mean(self, group: Optional[dist.ProcessGroup] = None) -> None:
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)


def train_step(model: nn.Module, criterion: nn.Module, optimizer: torch.optim.Optimizer, bucket: GradBucket,
               points: torch.Tensor, target: torch.Tensor, mean: Optional[torch.Tensor] = None,
               group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """One optimisation step on this rank's shard; returns the (local, detached) loss.

    ``points`` [B, C, N] and ``target`` are this rank's slice of the global batch; ``mean`` is the extra input of the
    translation heads (translation_ssg.py:31-33)."""
    if not bucket.attached():
        raise RuntimeError("train_step: parameter gradients no longer alias the bucket (use bucket.zero(), not zero_grad())")
    model.train()
    bucket.zero()                                              # optimizer.zero_grad() (:107)
    pred = model(points) if mean is None else model(points, mean)
    if isinstance(pred, (tuple, list)):
        pred = pred[0]
    loss = criterion(pred.float(), target.float())            # (:122)
    loss.backward()                                            # (:124) gradients land in the bucket
    bucket.all_reduce_mean(group)                              # the step's one collective
    optimizer.step()                                           # (:128)
    return loss.detach()
