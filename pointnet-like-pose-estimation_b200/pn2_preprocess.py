"""Device-side input preparation of the pose heads: the step immediately before the path.

The reference prepares every batch on the host in numpy (train_rotation.py:108-116): ``provider.normalization`` (per cloud:
subtract the centroid, divide by the largest norm; provider.py:5-21), ``provider.splice_torch`` (append the one-hot category
vector, a Python loop over the batch; provider.py:166-180) and a transpose to channels-first.  ``prepare_pose_input`` does
the three in one kernel launch (one CTA per cloud) and returns the [B, 3 + num_category, N] tensor the models consume.
``normalization`` / ``splice_torch`` keep the reference's names, argument order and output shapes for CUDA tensors.
fp32 like the reference's numpy path; the sums are tree reductions, so results agree to rounding (<= 1e-5 on coordinates in [-1, 1]), not bitwise.

``farthest_point_sample`` / ``random_sample`` are the numpy-array resamplers the data generators, loaders and the DBSCAN
cluster step call once per cloud (provider.py:183-204, data_utils/ModelDataLoader.py:10-46, point_collect/transformer.py:
107-141, point_collect/collect.py:64-68): same names, arguments, host-RNG draws and return values, with the FPS loop on
the GPU in the array's own dtype (float64 or float32) -- bit-identical selections.
"""
from __future__ import annotations

import numpy as np
import torch

import pointnet2_utils as _pu

__all__ = ["normalization", "splice_torch", "prepare_pose_input", "farthest_point_sample", "random_sample"]


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


def farthest_point_sample(point_cloud, number=1024, device="cuda"):
    """provider.farthest_point_sample: [N, C] numpy array (float64 | float32) -> the `number` sampled rows [number, C].

    The start index is drawn with ``np.random.randint(0, N)`` like :193; the selection runs on the first three columns
    in the array's dtype.  A list / [B, N, C] stack of equal-sized clouds is sampled in one launch (one draw per cloud,
    in order) and returns [B, number, C]."""
    arr = np.asarray(point_cloud)
    batched = arr.ndim == 3
    if arr.ndim not in (2, 3) or arr.shape[-1] < 3:
        raise ValueError("farthest_point_sample expects an [N, C] (or [B, N, C]) array with C >= 3")
    if arr.dtype not in (np.float64, np.float32):
        arr = arr.astype(np.float64)                       # integer input: numpy's arithmetic would promote the same way
    clouds = arr if batched else arr[None]
    B, N, _ = clouds.shape
    start = torch.tensor([np.random.randint(0, N) for _ in range(B)], dtype=torch.int64)
    xyz = torch.from_numpy(np.ascontiguousarray(clouds[:, :, :3])).to(device)
    idx = _pu._backend().fps_xyz(xyz, start.to(device), int(number)).cpu().numpy()
    out = np.stack([clouds[b][idx[b].astype(np.int32)] for b in range(B)])
    return out if batched else out[0]


def random_sample(point_cloud, number=1024):
    """data_utils/ModelDataLoader.py:33-46: `number` rows without replacement (one ``np.random.choice`` draw), or the
    array itself when it has no more than that.  numpy arrays and torch tensors (any device: the rows are gathered
    where the tensor lives; the draw stays on the host generator)."""
    N = point_cloud.shape[0]
    if N <= number:
        return point_cloud
    sample_idx = np.random.choice(N, number, replace=False)
    if isinstance(point_cloud, torch.Tensor):
        return point_cloud.index_select(0, torch.from_numpy(sample_idx).to(point_cloud.device))
    return point_cloud[sample_idx]
