"""The reference's PointNet-v1-like pose heads on the fused set-abstraction kernel.

``model/width.py`` :7-44, ``model/rotation.py`` :7-50, ``model/translation.py`` :7-51 and ``model/sign.py`` :7-43 share one
body: ``Conv1d(k=1) + BatchNorm1d + ReLU`` x 5 (10|13 -> 64 -> 64 -> 64 -> 128 -> 1024) over all N points, a global max
over the points, then ``relu(dropout(bn(fc)))`` x 2 and a last ``fc``.  The point-wise stack followed by the max is
exactly the fused kernel's ``group_all`` mode without any grouping: the [B, C, N] input is the "coordinate" operand
(C <= 16), there are no features, one group per cloud holds all N rows, and the [B, N, 1024] activations never reach HBM.
In eval mode ``PointNetV1Head`` runs that (CUDA tensors only: there is no CPU implementation); in training mode
(batch-statistics BatchNorm, autograd) it runs the reference's op sequence on ATen.

Attribute names (``conv, bn_conv, fc, bn_fc, dropout``, rotation's ``ftnet``, translation's ``fc1, fc2, bn1``) and the
construction order follow the reference files, so seeded construction and ``state_dict``s are interchangeable
(tests/test_reference_live.py).  rotation.py builds a 64x64 T-Net (:27) whose output it computes and discards (:37); it
is kept for checkpoint compatibility and, like in the reference, only evaluated in training mode where its BatchNorm
statistics are a side effect.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

import pointnet2_utils as _pu
from _pn2_native import ORDER_ALL

__all__ = ["pointwise_mlp_max", "TNetkd", "PointNetV1Head", "get_v1_model", "V1_ARCH"]


def pointwise_mlp_max(x: torch.Tensor, convs, bns, cache: "_pu._FoldedMLP | None" = None) -> torch.Tensor:
    """max over points of relu(bn(conv(.))) x len(convs): x [B, C, N] (C <= 16, CUDA, fp32) -> [B, Cout].

    The fused kernel in group_all mode (one group per cloud = all N rows, raw channels, no centre)."""
    be = _pu._backend()
    B, C, N = x.shape
    pts = x.permute(0, 2, 1)                                   # [B, N, C] view; the kernel reads it through strides
    cache = cache if cache is not None else _pu._FoldedMLP()
    packed, dims, prec = cache.get(convs, bns, C, ORDER_ALL)
    alloc = torch.zeros if be.sa_needs_zero(1, N, dims, prec) else torch.empty
    pooled = alloc(B, 1, dims[-1], dtype=torch.float32, device=x.device)
    be.sa_fused(pts, None, None, None, ORDER_ALL, dims, packed, prec, pooled, 0)
    return pooled.view(B, dims[-1])


class TNetkd(nn.Module):
    """model/pointnet_utils.py:45-82, restated for rotation.py's unused feature transform."""

    def __init__(self, channel):
        super().__init__()
        self.conv1 = nn.Conv1d(channel, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, channel * channel)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)
        self.channel = channel

    def forward(self, x):
        B = x.size(0)
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = F.relu(self.bn3(self.conv3(x)))
        x = torch.max(x, 2, keepdim=True)[0].view(-1, 1024)
        x = F.relu(self.bn4(self.fc1(x)))
        x = F.relu(self.bn5(self.fc2(x)))
        x = self.fc3(x)
        iden = torch.eye(self.channel, dtype=x.dtype, device=x.device).view(1, -1).repeat(B, 1)
        return (x + iden).view(-1, self.channel, self.channel)


# name -> (input channels beyond the one-hot category, output width, kind)
V1_ARCH = {
    "width": (6, 1, "reg"),           # width.py: normal_channel=True -> 6 + num_category input channels
    "rotation": (3, 3, "rot"),
    "translation": (3, 3, "trans"),
    "sign": (3, 1, "sign"),
}


class PointNetV1Head(nn.Module):
    def __init__(self, name: str, num_category: int = 7, mlp_list=(64, 64, 64, 128, 1024), linear_list=None, mean_mlp: str = "True"):
        super().__init__()
        if name not in V1_ARCH:
            raise ValueError(f"unknown v1 head {name!r}; one of {sorted(V1_ARCH)}")
        base, out_w, kind = V1_ARCH[name]
        self.name, self.kind, self.mean_mlp = name, kind, mean_mlp
        channel = base + num_category
        linear_list = list(linear_list) if linear_list is not None else [512, 256, out_w]
        self.conv, self.bn_conv = nn.ModuleList(), nn.ModuleList()
        width = channel
        for out_channel in mlp_list:
            self.conv.append(nn.Conv1d(width, out_channel, 1))
            self.bn_conv.append(nn.BatchNorm1d(out_channel))
            width = out_channel
        self.fc, self.bn_fc = nn.ModuleList(), nn.ModuleList()
        for out_channel in linear_list:
            self.fc.append(nn.Linear(width, out_channel))
            self.bn_fc.append(nn.BatchNorm1d(out_channel))
            width = out_channel
        if kind == "rot":
            self.ftnet = TNetkd(channel=64)                    # rotation.py:27
        if kind == "trans" and mean_mlp == "True":             # translation.py:28-31
            self.fc1 = nn.Linear(3, 6)
            self.fc2 = nn.Linear(6, 3)
            self.bn1 = nn.BatchNorm1d(6)
        self.dropout = nn.Dropout(p=0.4)
        self._folded = _pu._FoldedMLP()
        self._folded_head = _pu._FoldedHead()
        self.pooled_width = mlp_list[-1]

    def _encode(self, x):
        if not self.training:                                  # eval: the fused kernel; a CPU tensor raises (no CPU path)
            return pointwise_mlp_max(x, self.conv, self.bn_conv, self._folded)
        for i in range(len(self.conv)):                        # training mode: the reference's op sequence under autograd
            x = F.relu(self.bn_conv[i](self.conv[i](x)))
            if self.kind == "rot" and i == 1 and self.training:
                self.ftnet(x)                                  # computed and discarded (rotation.py:37)
        return torch.max(x, 2, keepdim=True)[0].view(-1, self.pooled_width)

    def forward(self, x, mean=None):
        fused = not self.training and len(self.fc) == 3 and _pu._FoldedHead.supported(tuple(self.fc))
        if self.kind == "trans" and self.mean_mlp == "True" and not fused:
            mean = self.fc2(F.relu(self.bn1(self.fc1(mean))))
        x = self._encode(x)
        if fused:      # eval: fc/bn/relu x 2 -> fc -> epilogue in one launch (dropout is the identity)
            from _pn2_native import HEAD_PLAIN, HEAD_SIGN, HEAD_TRANS
            kind = {"trans": HEAD_TRANS, "sign": HEAD_SIGN}.get(self.kind, HEAD_PLAIN)
            mm = (self.fc1, self.bn1, self.fc2) if (self.kind == "trans" and self.mean_mlp == "True") else None
            out, aux = _pu.fc_head(x, tuple(self.fc), (self.bn_fc[0], self.bn_fc[1]), kind, self._folded_head,
                                   mean=mean if self.kind == "trans" else None, mean_mlp=mm)
            return (out, aux) if self.kind == "sign" else out
        for i in range(len(self.fc)):
            if i < len(self.fc) - 1:
                x = F.relu(self.dropout(self.bn_fc[i](self.fc[i](x))))       # dropout BEFORE relu, as in the reference
            else:
                x = self.fc[i](x)
        if self.kind == "trans":
            return x + mean
        if self.kind == "sign":
            x = torch.sigmoid(x)
            return x, torch.sign(x - 0.5)
        return x


def get_v1_model(name: str, num_category: int = 7, **kw) -> PointNetV1Head:
    return PointNetV1Head(name, num_category, **kw)
