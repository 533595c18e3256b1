"""Device-side input preparation of the pose heads: the step immediately before the path.

The reference prepares every batch on the host in numpy (train_rotation.py:108-116): ``provider.normalization`` (per cloud:
subtract the centroid, divide by the largest norm; provider.py:5-21), ``provider.splice_torch`` (append the one-hot category
vector, a Python loop over the batch; provider.py:166-180) and a transpose to channels-first.  ``prepare_pose_input`` does
the three in one kernel launch (one CTA per cloud) and returns the [B, 3 + num_category, N] tensor the models consume.
``normalization`` / ``splice_torch`` keep the reference's names, argument order and output shapes for CUDA tensors.
fp32 like the reference's numpy path; the sums are tree reductions, so results agree to rounding (<= 1e-5 on coordinates in [-1, 1]), not bitwise.
"""
from __future__ import annotations

import torch

import pointnet2_utils as _pu

__all__ = ["normalization", "splice_torch", "prepare_pose_input"]


def prepare_pose_input(xyz: torch.Tensor, category: torch.Tensor | None = None, num_category: int = 7) -> torch.Tensor:
    """xyz [B,N,3] (CUDA fp32, any strides), category [B] int64 -> [B, 3 + num_category, N] normalised + one-hot, channels first."""
    if category is None:
        num_category = 0
    else:
        category = category.to(device=xyz.device, dtype=torch.int64).contiguous()
        if category.numel() and (int(category.min()) < 0 or int(category.max()) >= num_category):
            raise IndexError("category out of range")        # the reference's class_vector[i, category[i]] would raise too
    return _pu._backend().normalize_splice(xyz, category, num_category)


def normalization(point_cloud: torch.Tensor) -> torch.Tensor:
    """provider.normalization for a CUDA tensor [B,N,3] -> [B,N,3]."""
    return _pu._backend().normalize_splice(point_cloud, None, 0).permute(0, 2, 1)


def splice_torch(point_cloud: torch.Tensor, category: torch.Tensor, num_category: int = 7) -> torch.Tensor:
    """provider.splice_torch: [B,N,C], category [B] -> [B,N,C+num_category] (one-hot appended), without the loop over B."""
    B, N, _ = point_cloud.shape
    onehot = torch.nn.functional.one_hot(category.to(point_cloud.device).long().view(B), num_category).to(point_cloud.dtype)
    return torch.cat([point_cloud, onehot[:, None, :].expand(B, N, num_category)], dim=-1)
