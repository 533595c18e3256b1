"""TEST-ONLY stand-in for _pn2_native.Native backed by the CPU oracle.

Lets the `not gpu` suite exercise the host logic of pointnet2_utils.py (RNG call order, argument plumbing,
channel orders, branch concatenation, views / strides, weight-cache invalidation) and run the reference's
unchanged model files on top of it.  Nothing under pointnet-like-pose-estimation_b200/ imports this file.
"""
import numpy as np
import torch

import pn2_oracle as orc
from _pn2_native import ORDER_ALL, ORDER_MSG, ORDER_SSG


class OracleNative:
    name = "oracle-fake"

    def __init__(self):
        self.calls = []

    def fps(self, points, start, number, want_idx64=True, want_idx32=False, want_new_points=False):
        self.calls.append("fps")
        idx = orc.farthest_point_sample(points, number, start)
        return (idx if want_idx64 else None, idx.int() if want_idx32 else None,
                orc.index_points(points, idx).contiguous() if want_new_points else None)

    def ball_query_multi(self, centres, points, r2_list, numbers, want_idx64=False, want_idx32=True, empty_flag=None):
        self.calls.append("ball_query_multi")
        n = len(self.calls)
        outs = [self.ball_query(centres, points, r2, K, want_idx64, want_idx32, empty_flag) for r2, K in zip(r2_list, numbers)]
        del self.calls[n:]
        return outs

    def fps_xyz(self, points, start, S):
        self.calls.append("fps_xyz")
        rows = [orc.farthest_point_sample_np(points[b].numpy(), S, int(start[b]))[1] for b in range(points.shape[0])]
        return torch.from_numpy(np.stack(rows))

    def bn_relu_train_fwd(self, y, gamma, beta, eps, K=0, running=None):
        self.calls.append("bn_relu_train_fwd")
        M, C = y.shape
        if M < 2:
            raise ValueError("Expected more than 1 value per channel when training")
        yd = y.double()
        mean, var = yd.mean(0), yd.var(0, unbiased=False)
        invstd = (1.0 / torch.sqrt(var + eps)).float()
        a = gamma * invstd
        b = torch.addcmul(beta, -mean.float(), a)
        stats = torch.stack([mean.float(), var.float(), invstd, a, b])
        if running is not None:
            rm, rv, cb, f = running
            rm.mul_(1.0 - f).add_(mean.float() + (cb if cb is not None else 0.0), alpha=f)
            rv.mul_(1.0 - f).add_(var.float(), alpha=f * M / (M - 1))
        z = torch.relu(torch.addcmul(b, y, a))
        if not K:
            return z, stats, None
        zm, arg = z.view(M // K, K, C).max(1)
        return zm.contiguous(), stats, arg.int()

    def bn_relu_train_bwd(self, dz, y, arg, stats, K=0):
        self.calls.append("bn_relu_train_bwd")
        M, C = y.shape
        mean, invstd, a, b = stats[0], stats[2], stats[3], stats[4]
        if K:
            full = torch.zeros(M // K, K, C).scatter_(1, arg.long().unsqueeze(1), dz.view(M // K, 1, C)).view(M, C)
        else:
            full = dz
        g = torch.where(torch.addcmul(b, y, a) > 0, full, torch.zeros_like(full))
        xh = (y - mean) * invstd
        s1, s2 = g.double().sum(0), (g.double() * xh.double()).sum(0)
        dy = a * (g - (s1 / M).float() - xh * (s2 / M).float())
        return dy, torch.stack([s1.float(), s2.float()])

    def index_points(self, src, idx, check=True):
        self.calls.append("index_points")
        N = src.shape[1]
        if check and ((idx < -N) | (idx >= N)).any():
            raise IndexError("index out of range in index_points")
        return orc.index_points(src, idx % N)

    def square_distance(self, src, dst):
        self.calls.append("square_distance")
        return orc.square_distance(src, dst)

    def ball_query(self, centres, points, r2, number, want_idx64=True, want_idx32=False, empty_flag=None):
        self.calls.append("ball_query")
        B, N, C = points.shape
        S = centres.shape[1]
        out = torch.empty(B, S, number, dtype=torch.int64)
        sb, sn, sc = points.stride()
        cen = centres.contiguous()
        rc = orc.lib().pn2o_ball_query(cen.data_ptr(), points.data_ptr(), sb, sn, sc, B, S, N, C, r2, number, out.data_ptr())
        assert rc >= 0
        if empty_flag is not None and bool((out == N).any()):
            empty_flag.fill_(1)
        return (out if want_idx64 else None, out.int() if want_idx32 else None)

    def group(self, points, feature, idx32, centres, order):
        self.calls.append("group")
        return orc.group(points, feature, None if idx32 is None else idx32.long(), centres, order)

    def pack(self, layers, eps, prec, C, order):
        self.calls.append("pack")
        dims = [int(layers[0][0].shape[1])] + [int(l[0].shape[0]) for l in layers]
        return ([tuple(t.detach().clone() for t in tup) for tup in layers], eps), dims

    def fc_head(self, x, w1p, b1, w2p, b2, w3, b3, kind, mean=None, mean_mlp=None):
        self.calls.append("fc_head")
        unpack = lambda wp: wp.permute(1, 0, 2).reshape(wp.shape[1], -1)          # [K/4][H][4] -> [H, K]
        h = torch.relu(x @ unpack(w1p).t() + b1)
        h = torch.relu(h @ unpack(w2p).t() + b2)
        y = h @ w3.t() + b3
        if kind == 1:
            y = torch.log_softmax(y, -1)
            return y, y.max(1)[1]
        if kind == 2:
            if mean_mlp is not None:
                q = mean_mlp
                mean = torch.relu(mean @ q[:18].view(6, 3).t() + q[18:24]) @ q[24:42].view(3, 6).t() + q[42:45]
            return y + mean, None
        if kind == 3:
            y = torch.sigmoid(y)
            return y, torch.sign(y - 0.5)
        return y, None

    def rows_gemm(self, a, w, out_dtype=None, out=None):
        self.calls.append("rows_gemm")
        y = a.float() @ w.float().t()
        return y.to(out_dtype or (torch.bfloat16 if w.dtype == torch.bfloat16 else torch.float32))

    def gather_rows(self, feat, idx32):
        self.calls.append("gather_rows")
        B, N, D = feat.shape
        flat = idx32.reshape(B, -1, 1).long().clamp(0, N - 1).expand(-1, -1, D)
        out = torch.gather(feat, 1, flat).view(*idx32.shape, D)
        return out * ((idx32 >= 0) & (idx32 < N)).unsqueeze(-1)

    def scatter_add_rows(self, grad, idx32, N):
        self.calls.append("scatter_add_rows")
        B, D = grad.shape[0], grad.shape[-1]
        ok = ((idx32 >= 0) & (idx32 < N)).reshape(B, -1, 1)
        flat = idx32.reshape(B, -1, 1).long().clamp(0, N - 1).expand(-1, -1, D)
        return torch.zeros(B, N, D).scatter_add_(1, flat, grad.reshape(B, -1, D) * ok)

    def rows_wgrad(self, dy, x):
        self.calls.append("rows_wgrad")
        return dy.t() @ x

    def sa_needs_zero(self, S, K, dims, prec):
        return False

    def sa_bf16_side_io(self, K, C, order, dims, prec):
        return False

    def sa_fused(self, points, feature, centres, idx32, order, dims, packed, prec, out, co_off=0,
                 feature_bf16=None, out_bf16=None, algo_dims=None):
        self.calls.append("sa_fused")
        layers, eps = packed
        rows = orc.group(points, feature, None if idx32 is None else idx32.long(), centres, order)
        B, S, K, W = rows.shape
        pooled = orc.mlp_max(rows.reshape(B * S, K, W), layers, eps).reshape(B, S, -1)
        out[:, :, co_off:co_off + pooled.shape[-1]] = pooled
        return out
