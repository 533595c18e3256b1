"""ctypes front-end of the CPU oracle (oracle/pn2_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (pointnet-like-pose-estimation_b200/) never does.

All functions take torch CPU tensors (fp32 / int64) and keep the argument names of the reference's
model/pointnet2_utils.py (this fork calls the coordinates ``points`` and the per-point features
``feature``).  The one deliberate difference: the FPS start indices that the reference draws with
``torch.randint`` inside the function (pointnet2_utils.py:59) are an explicit ``start`` argument here.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpn2_oracle.so")
_lib = None

_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_ptr = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile the C restatement in place (gcc via oracle/Makefile)."""
    src = os.path.join(_HERE, "pn2_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "-B", "libpn2_oracle.so"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.pn2o_fps.argtypes = [_ptr, _i64, _i64, _i64, _int, _int, _int, _ptr, _int, _ptr]
        L.pn2o_square_distance.argtypes = [_ptr, _ptr, _i64, _i64, _i64, _int, _int, _int, _int, _ptr]
        L.pn2o_ball_query.argtypes = [_ptr, _ptr, _i64, _i64, _i64, _int, _int, _int, _int, _f32, _int, _ptr]
        L.pn2o_group.argtypes = [_ptr, _i64, _i64, _i64, _ptr, _i64, _i64, _i64, _ptr, _ptr,
                                 _int, _int, _int, _int, _int, _int, _int, _ptr]
        L.pn2o_mlp_max.argtypes = [_ptr, _i64, _int, _int, _ptr, _ptr, _f32, _ptr]
        L.pn2o_sum_order.argtypes = [_ptr, _int, _int]
        L.pn2o_sum_order.restype = _f32
        for f in (L.pn2o_fps_xyz_f64, L.pn2o_fps_xyz_f32):
            f.argtypes = [_ptr, _i64, _i64, _int, _i64, _int, _ptr]
            f.restype = _int
        for f in (L.pn2o_fps, L.pn2o_square_distance, L.pn2o_ball_query, L.pn2o_group, L.pn2o_mlp_max):
            f.restype = _int
        _lib = L
    return _lib


def _chk(t: torch.Tensor, dtype) -> torch.Tensor:
    assert t.device.type == "cpu" and t.dtype == dtype, (t.device, t.dtype)
    return t


def radius_sq_f32(radius: float) -> float:
    """``radius ** 2`` in double, rounded once to fp32 (pointnet2_utils.py:85 compares in fp32)."""
    return float(np.float32(float(radius) ** 2))


def farthest_point_sample(points: torch.Tensor, number: int, start: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:47-68 with the :59 randint passed in as ``start`` [B] int64."""
    _chk(points, torch.float32); _chk(start, torch.int64)
    B, N, C = points.shape
    out = torch.empty(B, number, dtype=torch.int64)
    sb, sn, sc = points.stride()
    st = start.contiguous()
    rc = lib().pn2o_fps(points.data_ptr(), sb, sn, sc, B, N, C, st.data_ptr(), number, out.data_ptr())
    if rc:
        raise RuntimeError(f"pn2o_fps rc={rc}")
    return out


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:5-26 (src contiguous [B,S,C]; dst any strides [B,N,C])."""
    _chk(src, torch.float32); _chk(dst, torch.float32)
    src = src.contiguous()
    B, S, C = src.shape
    N = dst.shape[1]
    out = torch.empty(B, S, N, dtype=torch.float32)
    sb, sn, sc = dst.stride()
    rc = lib().pn2o_square_distance(src.data_ptr(), dst.data_ptr(), sb, sn, sc, B, S, N, C, out.data_ptr())
    if rc:
        raise RuntimeError(f"pn2o_square_distance rc={rc}")
    return out


def query_ball_point(radius: float, number: int, points: torch.Tensor, new_points: torch.Tensor,
                     allow_empty: bool = False) -> torch.Tensor:
    """pointnet2_utils.py:70-90."""
    _chk(points, torch.float32); _chk(new_points, torch.float32)
    new_points = new_points.contiguous()
    B, N, C = points.shape
    S = new_points.shape[1]
    out = torch.empty(B, S, number, dtype=torch.int64)
    sb, sn, sc = points.stride()
    rc = lib().pn2o_ball_query(new_points.data_ptr(), points.data_ptr(), sb, sn, sc, B, S, N, C,
                               radius_sq_f32(radius), number, out.data_ptr())
    if rc < 0 or (rc == 1 and not allow_empty):
        raise RuntimeError(f"pn2o_ball_query rc={rc}")
    return out


def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:28-45 -- batched row gather."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.stack([points[b][flat[b]] for b in range(B)], 0)
    return out.reshape(*idx.shape, points.shape[-1])


ORDER_SSG, ORDER_MSG, ORDER_ALL = 0, 1, 2


def group(points, feature, idx, new_points, order: int) -> torch.Tensor:
    """Grouped rows [B,S,K,C+D] in SSG (:109-116), MSG (:205-211) or group_all (:134-140) order."""
    _chk(points, torch.float32)
    B, N, C = points.shape
    if order == ORDER_ALL:
        S, K = 1, N
        idx_t = torch.zeros(1, dtype=torch.int64)
        cen = torch.zeros(B, 1, C)
    else:
        _, S, K = idx.shape
        idx_t = _chk(idx, torch.int64).contiguous()
        cen = _chk(new_points, torch.float32).contiguous()
    D = 0 if feature is None else feature.shape[-1]
    out = torch.empty(B, S, K, C + D, dtype=torch.float32)
    sb, sn, sc = points.stride()
    fb = fn = fc = 0
    fptr = None
    if feature is not None:
        _chk(feature, torch.float32)
        fb, fn, fc = feature.stride()
        fptr = feature.data_ptr()
    rc = lib().pn2o_group(points.data_ptr(), sb, sn, sc, fptr, fb, fn, fc, idx_t.data_ptr(), cen.data_ptr(),
                          B, N, S, K, C, D, order, out.data_ptr())
    if rc:
        raise RuntimeError(f"pn2o_group rc={rc}")
    return out


def pack_layers(layers) -> tuple[np.ndarray, np.ndarray]:
    """layers: list of dicts/tuples (W[Cout,Cin], b, gamma, beta, mean, var) as torch tensors."""
    dims = [int(layers[0][0].shape[1])]
    chunks = []
    for (W, b, g, be, mu, var) in layers:
        dims.append(int(W.shape[0]))
        for t in (W.reshape(W.shape[0], -1), b, g, be, mu, var):
            chunks.append(t.detach().to(torch.float32).contiguous().reshape(-1).numpy())
    return np.asarray(dims, dtype=np.int32), np.concatenate(chunks).astype(np.float32)


def mlp_max(rows: torch.Tensor, layers, eps: float = 1e-5) -> torch.Tensor:
    """(Conv1x1 -> eval BatchNorm -> ReLU) x L then max over K (pointnet2_utils.py:167-172).

    rows [G,K,Cin] -> [G,CL].  layers as for :func:`pack_layers`.
    """
    _chk(rows, torch.float32)
    rows = rows.contiguous()
    G, K, Cin = rows.shape
    dims, params = pack_layers(layers)
    assert dims[0] == Cin, (dims, Cin)
    out = torch.empty(G, int(dims[-1]), dtype=torch.float32)
    rc = lib().pn2o_mlp_max(rows.data_ptr(), G, K, len(layers), dims.ctypes.data, params.ctypes.data,
                            eps, out.data_ptr())
    if rc:
        raise RuntimeError(f"pn2o_mlp_max rc={rc}")
    return out


def layers_of(convs, bns):
    """(W,b,gamma,beta,mean,var) tuples from parallel lists of nn.Conv2d / nn.BatchNorm2d modules."""
    return [(c.weight, c.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var) for c, bn in zip(convs, bns)]


def set_abstraction(points_bcn, feature_bdn, point_number, sample_number, radius, layers, group_all, start):
    """PointNetSetAbstraction.forward (pointnet2_utils.py:158-174) in eval mode.

    Returns (new_points [B,C,S] view of contiguous [B,S,C], new_feature [B,Cout,S], fps_idx, group_idx).
    """
    points = points_bcn.permute(0, 2, 1)
    feature = None if feature_bdn is None else feature_bdn.permute(0, 2, 1)
    B, N, C = points.shape
    if group_all:
        rows = group(points, feature, None, None, ORDER_ALL)
        new_points = torch.zeros(B, 1, C)
        fps_idx = gidx = None
    else:
        fps_idx = farthest_point_sample(points, point_number, start)
        new_points = index_points(points, fps_idx)
        gidx = query_ball_point(radius, sample_number, points, new_points)
        rows = group(points, feature, gidx, new_points, ORDER_SSG)
    _, S, K, W = rows.shape
    pooled = mlp_max(rows.reshape(B * S, K, W), layers).reshape(B, S, -1)
    return new_points.permute(0, 2, 1), pooled.permute(0, 2, 1), fps_idx, gidx


def set_abstraction_msg(points_bcn, feature_bdn, point_number, sample_number_list, radius_list, layers_list, start):
    """PointNetSetAbstractionMsg.forward (pointnet2_utils.py:195-223) in eval mode."""
    points = points_bcn.permute(0, 2, 1)
    feature = None if feature_bdn is None else feature_bdn.permute(0, 2, 1)
    B, N, C = points.shape
    fps_idx = farthest_point_sample(points, point_number, start)
    new_points = index_points(points, fps_idx)
    outs, gidxs = [], []
    for radius, K, layers in zip(radius_list, sample_number_list, layers_list):
        gidx = query_ball_point(radius, K, points, new_points)
        rows = group(points, feature, gidx, new_points, ORDER_MSG)
        pooled = mlp_max(rows.reshape(B * point_number, K, rows.shape[-1]), layers).reshape(B, point_number, -1)
        outs.append(pooled.permute(0, 2, 1))
        gidxs.append(gidx)
    return new_points.permute(0, 2, 1), torch.cat(outs, 1), fps_idx, gidxs


# ---------------------------------------------------------------------------------------------------------------------
# input preparation (host-side numpy in the reference; provider.py cannot be imported here -- it imports open3d at :3)
def normalization(point_cloud):
    """provider.py:5-21, the same numpy operations in the same order (float32 in -> float64 container, like the reference)."""
    import numpy as np
    B, N, C = point_cloud.shape
    normalize = np.zeros((B, N, C))
    for i in range(B):
        pc = point_cloud[i]
        centroid = np.mean(pc, axis=0)
        pc = pc - centroid
        m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
        pc = pc / m
        normalize[i] = pc
    return normalize


def farthest_point_sample_np(point_cloud, number, start):
    """provider.py:183-204 / data_utils/ModelDataLoader.py:10-31 / point_collect/transformer.py:120-141 with the :193
    np.random.randint draw passed in: [N, C] float64 | float32 array -> (sampled rows [number, C], indices [number])."""
    import numpy as np
    assert point_cloud.ndim == 2 and point_cloud.shape[1] >= 3 and point_cloud.dtype in (np.float64, np.float32)
    N = point_cloud.shape[0]
    es = point_cloud.itemsize
    assert point_cloud.strides[0] % es == 0 and point_cloud.strides[1] % es == 0
    idx = np.empty(number, dtype=np.int64)
    fn = lib().pn2o_fps_xyz_f64 if point_cloud.dtype == np.float64 else lib().pn2o_fps_xyz_f32
    rc = fn(point_cloud.ctypes.data, point_cloud.strides[0] // es, point_cloud.strides[1] // es, N, int(start), int(number),
            idx.ctypes.data)
    if rc:
        raise RuntimeError(f"pn2o_fps_xyz rc={rc}")
    return point_cloud[idx.astype(np.int32)], idx


def splice_torch(point_cloud, category, num_category=7):
    """provider.py:166-180."""
    point_cloud = point_cloud.transpose(2, 1)
    B, C, N = point_cloud.size()
    class_vector = torch.zeros(B, num_category, N)
    for i in range(B):
        class_vector[i, category[i].item(), :] = 1
    return torch.cat([point_cloud, class_vector], 1).transpose(2, 1)
