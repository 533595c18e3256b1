"""CPU timing baseline -- TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

The reference's set-abstraction path is PyTorch-on-CPU code (model/pointnet2_utils.py): its cost is the cost
of the ATen / MKL / oneDNN kernels it calls.  The reference checkout does not travel to the GPU box, so this
module restates the same sequence of tensor operations (same algorithm, same intermediate tensors: the
[B,S,N] distance matrix, the full sort over N, the materialised [B,S,K,Cin] grouped tensor, conv2d + batch_norm
+ relu + max) for bench.py's `cpu_baseline` and `--impl reference` legs to time on the host cores.  The CPU
test-suite checks it against the golden vectors recorded from the live reference (bit-exact indices), so the
thing timed is the thing the reference computes.  Bit-exact parity checking itself uses pn2_oracle.c.

Reference call sites restated: square_distance :5-26, index_points :28-45, farthest_point_sample :47-68,
query_ball_point :70-90, sample_and_group :92-120, sample_and_group_all :122-141, PointNetSetAbstraction
:143-174, PointNetSetAbstractionMsg :176-223.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

_start_source = None


def set_start_source(fn):
    """fn(B, N) -> int64 [B]: replaces the torch.randint draw of pointnet2_utils.py:59 (None = randint)."""
    global _start_source
    _start_source = fn


def pairwise_sqdist(a, b):
    """[B,S,C] x [B,N,C] -> [B,S,N], expanded form -2ab + |a|^2 + |b|^2 (:23-25)."""
    out = torch.matmul(a, b.transpose(1, 2)) * -2
    out += (a ** 2).sum(-1).unsqueeze(2)
    out += (b ** 2).sum(-1).unsqueeze(1)
    return out


def take_rows(table, idx):
    """table [B,N,C], idx [B,...] -> table[b, idx[b,...]] (:28-45)."""
    B = table.shape[0]
    lead = torch.arange(B, device=table.device).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return table[lead, idx, :]


def fps(points, number, start=None):
    """Iterative farthest point sampling (:47-68)."""
    B, N, _ = points.shape
    if start is None:
        start = _start_source(B, N) if _start_source is not None else torch.randint(0, N, (B,), dtype=torch.long)
    chosen = torch.zeros(B, number, dtype=torch.long)
    nearest = torch.full((B, N), 1e10)
    cur = start.clone()
    rows = torch.arange(B)
    for i in range(number):
        chosen[:, i] = cur
        centre = points[rows, cur, :].unsqueeze(1)
        d = ((points - centre) ** 2).sum(-1)
        closer = d < nearest
        nearest[closer] = d[closer]
        cur = nearest.max(-1)[1]
    return chosen


def ball_query(radius, number, points, centres):
    """First `number` in-radius indices in ascending order, padded with the first (:70-90)."""
    B, N, _ = points.shape
    S = centres.shape[1]
    cand = torch.arange(N, dtype=torch.long).view(1, 1, N).repeat(B, S, 1)
    cand[pairwise_sqdist(centres, points) > radius ** 2] = N
    cand = cand.sort(dim=-1)[0][:, :, :number]
    first = cand[:, :, :1].expand(-1, -1, number)
    pad = cand == N
    cand[pad] = first[pad]
    return cand


def group_ssg(points, feature, S, K, radius):
    picked = fps(points, S)
    centres = take_rows(points, picked)
    nbr = ball_query(radius, K, points, centres)
    local = take_rows(points, nbr) - centres.unsqueeze(2)
    if feature is not None:
        local = torch.cat([local, take_rows(feature, nbr)], dim=-1)          # [xyz_norm, feat]  (:113-114)
    return centres, local


def group_all(points, feature):
    B, N, C = points.shape
    rows = points.reshape(B, 1, N, C)
    if feature is not None:
        rows = torch.cat([rows, feature.reshape(B, 1, N, -1)], dim=-1)       # [xyz_raw, feat]   (:138)
    return torch.zeros(B, 1, C), rows


def _mlp_pool(rows_bskc, convs, bns):
    x = rows_bskc.permute(0, 3, 2, 1)                                         # [B,Cin,K,S]        (:167)
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return x.max(dim=2)[0]


class PointNetSetAbstraction(nn.Module):
    """Same constructor, parameter names and forward contract as the reference module (:143-174)."""

    def __init__(self, point_number, sample_number, radius, in_channel, mlp, group_all=False):
        super().__init__()
        self.point_number, self.sample_number, self.radius, self.group_all = point_number, sample_number, radius, group_all
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        c = in_channel
        for w in mlp:
            self.mlp_convs.append(nn.Conv2d(c, w, 1))
            self.mlp_bns.append(nn.BatchNorm2d(w))
            c = w

    def forward(self, points, feature):
        pts = points.permute(0, 2, 1)
        feat = None if feature is None else feature.permute(0, 2, 1)
        if self.group_all:
            centres, rows = group_all(pts, feat)
        else:
            centres, rows = group_ssg(pts, feat, self.point_number, self.sample_number, self.radius)
        return centres.permute(0, 2, 1), _mlp_pool(rows, self.mlp_convs, self.mlp_bns)


class PointNetSetAbstractionMsg(nn.Module):
    """Multi-radius variant (:176-223): one FPS, per radius ball query + [feat, xyz_norm] rows + MLP + max."""

    def __init__(self, point_number, sample_number_list, radius_list, in_channel, mlp_list, num_category=0):
        super().__init__()
        self.point_number, self.sample_number_list, self.radius_list = point_number, sample_number_list, radius_list
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for widths in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            c = in_channel + 3 + num_category
            for w in widths:
                convs.append(nn.Conv2d(c, w, 1))
                bns.append(nn.BatchNorm2d(w))
                c = w
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, points, feature):
        pts = points.permute(0, 2, 1)
        feat = None if feature is None else feature.permute(0, 2, 1)
        centres = take_rows(pts, fps(pts, self.point_number))
        outs = []
        for r, K, convs, bns in zip(self.radius_list, self.sample_number_list, self.conv_blocks, self.bn_blocks):
            nbr = ball_query(r, K, pts, centres)
            rows = take_rows(pts, nbr) - centres.unsqueeze(2)
            if feat is not None:
                rows = torch.cat([take_rows(feat, nbr), rows], dim=-1)        # [feat, xyz_norm]  (:208-209)
            outs.append(_mlp_pool(rows, convs, bns))
        return centres.permute(0, 2, 1), torch.cat(outs, dim=1)


def v1_forward(model, x, mean=None):
    """The PointNet-v1-like pose heads (model/width.py:32-44, rotation.py:34-50, translation.py:37-51, sign.py:31-43) as the
    reference runs them on the CPU: Conv1d(k=1) + BatchNorm1d + ReLU x 5 over all points, global max, FC stack.  `model`
    only supplies the parameters (a pn2_v1.PointNetV1Head: same attribute names as the reference files); timing baseline
    of configs[3]'s width head."""
    if model.kind == "trans" and model.mean_mlp == "True":
        mean = model.fc2(F.relu(model.bn1(model.fc1(mean))))
    for conv, bn in zip(model.conv, model.bn_conv):
        x = F.relu(bn(conv(x)))
    x = torch.max(x, 2, keepdim=True)[0].view(-1, model.pooled_width)
    for i in range(len(model.fc)):
        x = F.relu(model.dropout(model.bn_fc[i](model.fc[i](x)))) if i < len(model.fc) - 1 else model.fc[i](x)
    if model.kind == "trans":
        return x + mean
    if model.kind == "sign":
        x = torch.sigmoid(x)
        return x, torch.sign(x - 0.5)
    return x
