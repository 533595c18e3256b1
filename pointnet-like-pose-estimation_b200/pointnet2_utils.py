"""Drop-in replacement for the reference's ``model/pointnet2_utils.py`` on B200 (sm_100a).

Put this directory first on ``sys.path`` (the reference's scripts do ``sys.path.append(<root>/model)`` and
``importlib.import_module(args.model)``, train_rotation.py:16-18,32) and the unchanged model files
``pointnet2_cls_{ssg,msg}.py``, ``{rotation,translation,sign}_{ssg,msg}.py`` pick up these definitions through
``from pointnet2_utils import PointNetSetAbstraction, PointNetSetAbstractionMsg``.

Same public names, argument names (this fork: ``points`` = coordinates, ``feature`` = per-point features),
shapes and dtypes as the reference; the arithmetic runs in hand-written CUDA kernels behind the C ABI of
``include/pn2.h`` (see ``_pn2_native.py``).  There is no CPU implementation here.

What is deliberately different from the reference (all documented in DESIGN.md):
  * pooled features are stored channels-last ([B,S,Cout]) and returned as the [B,Cout,S] *view* of that
    buffer (same shape, dtype and values; strides (S*Cout, 1, Cout) instead of contiguous) so that the next
    layer gathers whole rows; ``set_reference_strides(True)`` returns contiguous tensors instead.
  * ``set_precision('bf16')`` runs the shared MLP on the tensor cores (bf16 in, fp32 accumulate);
    default is 'fp32'.
"""
from __future__ import annotations

import contextlib
import math
import os
import threading

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import _pn2_native as _nat
from _pn2_native import ORDER_ALL, ORDER_MSG, ORDER_SSG, PREC_BF16, PREC_F32

__all__ = ["square_distance", "index_points", "farthest_point_sample", "query_ball_point", "sample_and_group",
           "sample_and_group_all", "PointNetSetAbstraction", "PointNetSetAbstractionMsg",
           "set_precision", "get_precision", "set_reference_strides", "fps_start_source",
           "invalidate_caches", "empty_ball_seen"]

_native = None              # created on first use; tests may monkeypatch this attribute
_precision = os.environ.get("PN2_PRECISION", "fp32")
_reference_strides = False
_factorise_first_layer = os.environ.get("PN2_FACTORISE", "1") != "0"     # see _FactorisedFirstLayer
_train_fused = os.environ.get("PN2_TRAIN_FUSED", "1") != "0"            # see _shared_mlp_train_rows
_train_own_gemm = os.environ.get("PN2_TRAIN_OWN_GEMM", "1") != "0"       # training convolutions on pn2_rows_gemm (else cuBLAS)
_start_source = None        # optional callable (B, N) -> int64 CPU tensor [B]; default = torch.randint
_cache_generation = 0       # part of every derived-weight cache key, see invalidate_caches()
_empty_flags = {}           # device -> int32 [1]: set by the ball queries of the modules when a centre has no neighbour


def _backend():
    global _native
    if _native is None:
        _native = _nat.Native()
    return _native


def set_precision(mode: str) -> None:
    """'fp32' (FFMA, <= 1e-5 of the reference) or 'bf16' (tcgen05 tensor cores, <= 2e-2)."""
    global _precision
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _precision = mode


def get_precision() -> str:
    return _precision


def set_reference_strides(flag: bool) -> None:
    """When True the modules return contiguous [B,Cout,S] features exactly like the reference."""
    global _reference_strides
    _reference_strides = bool(flag)


@contextlib.contextmanager
def fps_start_source(fn):
    """Route the FPS start-index draw (pointnet2_utils.py:59) through ``fn(B, N) -> int64 CPU tensor [B]``.

    Used by the batch-sharded runner so that every rank slices ONE global draw (bit-identical to a
    single-GPU run of the whole batch)."""
    global _start_source
    prev, _start_source = _start_source, fn
    try:
        yield
    finally:
        _start_source = prev


class _PinnedDraws:
    """A small ring of pinned [B] int64 buffers per batch size for the FPS start draw: `torch.randint(..., out=pinned)` makes
    exactly the draw of the reference (same generator, same values); a kernel then copies the mapped pinned buffer to the
    device on the compute stream (pn2_upload_i64).  The pageable `.to(device)` of the plain form stalls the enqueueing thread
    once per sampling layer (-1.7 % on the B = 2048 step), and a pinned cudaMemcpyAsync queues on the host-to-device copy
    engine behind the next batch's input upload (-4 % end to end).  A buffer is reused only after the stream has passed the
    launch that read it (one event per buffer)."""
    SLOTS = 8
    KERNEL_UPLOAD_MIN_B = 512

    def __init__(self):
        self.rings = {}

    def draw(self, B, N, device):
        key = (B, device.index if device.index is not None else torch.cuda.current_device(), threading.get_ident())
        ring = self.rings.get(key)
        if ring is None:
            ring = self.rings[key] = {"bufs": [torch.empty(B, dtype=torch.long).pin_memory() for _ in range(self.SLOTS)],
                                      "evs": [torch.cuda.Event() for _ in range(self.SLOTS)], "i": 0, "used": [False] * self.SLOTS}
        i = ring["i"]
        ring["i"] = (i + 1) % self.SLOTS
        if ring["used"][i]:
            ring["evs"][i].synchronize()
        buf = ring["bufs"][i]
        torch.randint(0, N, (B,), dtype=torch.long, out=buf)
        # small batches are launch-bound and their input uploads short: the plain asynchronous copy costs the enqueueing thread
        # less (B = 24 eager forward: 61 k clouds/s against 50 k through the kernel's ctypes call)
        dev_t = _backend().upload_i64(buf, device) if B >= self.KERNEL_UPLOAD_MIN_B else buf.to(device, non_blocking=True)
        ring["evs"][i].record(torch.cuda.current_stream(device))
        ring["used"][i] = True
        return dev_t


_pinned_draws = _PinnedDraws()


def _draw_start(B: int, N: int, device) -> torch.Tensor:
    # The reference draws on the CPU default generator even when the model is on the GPU (:59); keep that,
    # one draw per farthest_point_sample call, so seeded runs choose the same centroids.
    if _start_source is not None:
        start = _start_source(B, N)
        return start.to(device, non_blocking=start.is_pinned())     # pinned sources (pn2_graph.py) must not synchronise
    device = torch.device(device)
    if device.type == "cuda" and not torch.cuda.is_current_stream_capturing() and hasattr(_backend(), "upload_i64"):
        return _pinned_draws.draw(B, N, device)
    return torch.randint(0, N, (B,), dtype=torch.long).to(device)


def invalidate_caches() -> None:
    """Drop every derived weight cache (BN-folded / bf16-packed weights) so that the next forward repacks.

    The caches follow parameter updates through ``tensor._version`` (optimizer steps, ``load_state_dict``, in-place ops on
    the parameter).  Writes through ``param.data`` (``p.data.copy_()``, ``m.weight.data.normal_()``, EMA code) bump a
    different version counter and are NOT seen: call this after such an update.  CUDA graphs captured by
    ``pn2_graph.GraphedForward`` notice the new generation and re-capture."""
    global _cache_generation
    _cache_generation += 1


def _empty_flag(device) -> torch.Tensor:
    key = (device.type, device.index if (device.index is not None or device.type != "cuda") else torch.cuda.current_device())
    t = _empty_flags.get(key)
    if t is None:
        t = _empty_flags[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return t


def empty_ball_seen(device=None, reset: bool = True) -> bool:
    """True if, since the last reset, a set-abstraction module on `device` met a centre with NO point inside its radius.

    The reference writes index N for such a centre and crashes in the next ``index_points`` (:85-89, :44); that cannot
    happen when the centres are FPS picks of the same cloud unless the expanded-form distance d(s,s) rounds above r**2
    (un-normalised, large-magnitude clouds).  The modules here neither crash nor read out of bounds: the group's rows
    are zeros (training) / its pooled feature is 0 (eval), and this flag records it without a device synchronisation on
    the forward path.  Reading it synchronises.  The stand-alone ``sample_and_group`` raises IndexError like the reference."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    flag = _empty_flag(torch.device(device))
    seen = bool(flag.item())
    if seen and reset:
        flag.zero_()
    return seen


def _radius_sq(radius) -> float:
    # Python squares in double (0.2**2 = 0.04000000000000001), ATen then compares in fp32 (:85)
    return float(np.float32(float(radius) ** 2))


# ------------------------------------------------------------------------------------------------------
# functions (reference: pointnet2_utils.py:5-141)
# ------------------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """[B,N,C], [B,M,C] -> [B,N,M] squared distances in the reference's expanded form (:5-26)."""
    return _backend().square_distance(src, dst)


def index_points(points, idx):
    """points [B,N,C], idx [B,S] or [B,S,K] (int64) -> points[b, idx[b,...], :] (:28-45)."""
    return _backend().index_points(points, idx)


def farthest_point_sample(points, number):
    """points [B,N,C] -> sampled indices [B,number] int64 (:47-68)."""
    B, N, _ = points.shape
    start = _draw_start(B, N, points.device)
    return _backend().fps(points, start, number, want_idx64=True)[0]


def query_ball_point(radius, number, points, new_points):
    """First `number` in-radius indices in ascending order, padded with the first one (:70-90)."""
    if number > points.shape[1]:
        raise RuntimeError(f"query_ball_point: number={number} > N={points.shape[1]} (the reference fails at :89)")
    return _backend().ball_query(new_points, points, _radius_sq(radius), number, want_idx64=True)[0]


def sample_and_group(points, feature, point_number, sample_number, radius, returnfps=False):
    """(:92-120) -> new_points [B,S,C], new_feature [B,S,K,C+D] = cat([xyz - centre, feat])."""
    be = _backend()
    B, N, C = points.shape
    start = _draw_start(B, N, points.device)
    fps_idx, _, new_points = be.fps(points, start, point_number, want_idx64=True, want_new_points=True)
    flag = torch.zeros(1, dtype=torch.int32, device=points.device)
    _, idx32 = be.ball_query(new_points, points, _radius_sq(radius), sample_number, want_idx64=False, want_idx32=True, empty_flag=flag)
    if int(flag.item()):        # the reference indexes with N here and fails in index_points (:109 -> :44)
        raise IndexError(f"sample_and_group: a centre has no point within radius {radius} (index {N} is out of bounds for dimension 1 with size {N})")
    new_feature = be.group(points, feature, idx32, new_points, ORDER_SSG)
    if returnfps:
        grouped = be.index_points(feature if feature is not None else points, idx32.long(), check=False)
        return new_points, new_feature, grouped, fps_idx
    return new_points, new_feature


def sample_and_group_all(points, feature):
    """(:122-141) -> zeros [B,1,C], cat([xyz, feat]) viewed [B,1,N,C+D]."""
    B, N, C = points.shape
    new_points = torch.zeros(B, 1, C, device=points.device)
    return new_points, _backend().group(points, feature, None, None, ORDER_ALL)


# ------------------------------------------------------------------------------------------------------
# modules (reference: pointnet2_utils.py:143-223)
# ------------------------------------------------------------------------------------------------------
def _prec_code() -> int:
    return PREC_BF16 if _precision == "bf16" else PREC_F32


def _stack_tensors(convs, bns):
    """The tensors a folded (conv, bn) stack depends on, read straight from the modules' registries (this runs on every forward
    to validate the caches: nn.Module.__getattr__ and parameters() cost several microseconds per tensor)."""
    out = []
    for conv, bn in zip(convs, bns):
        cp, bp, bb = conv._parameters, bn._parameters, bn._buffers
        out += [cp["weight"], cp["bias"], bp["weight"], bp["bias"], bb["running_mean"], bb["running_var"]]
    return out


def _leaf_tensors(m):
    """parameters() + buffers() of a leaf module, without the generator machinery."""
    return [t for t in m._parameters.values() if t is not None] + [t for t in m._buffers.values() if t is not None]


class _FoldedMLP:
    """Derived cache: BN-folded, kernel-layout weights of one (Conv2d, BatchNorm2d) stack.

    Never part of state_dict; rebuilt whenever a parameter / buffer changes (tensor ._version), moves, or
    the precision mode changes."""

    def __init__(self):
        self.key = None
        self.packed = None
        self.dims = None

    def get(self, convs, bns, C, order):
        tensors = _stack_tensors(convs, bns)
        prec = _prec_code()
        key = (_cache_generation, prec, C, order, tuple((t.data_ptr(), t._version) for t in tensors))
        if key != self.key:
            layers = [(conv.weight.detach().reshape(conv.out_channels, conv.in_channels), conv.bias.detach(),
                       bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var)
                      for conv, bn in zip(convs, bns)]
            eps = bns[0].eps
            assert all(bn.eps == eps for bn in bns)
            self.packed, self.dims = _backend().pack(layers, eps, prec, C, order)
            self.key = key
        return self.packed, self.dims, prec


class _FactorisedFirstLayer:
    """Derived cache for the first-layer factorisation of an MSG branch (bf16 mode).

    The first layer is linear in the concatenated row [feat[n], xyz[n] - centre] (:208-217), so its feature part does not
    depend on the group a point was gathered into:
        y0 = relu(s (Wf feat[n] + Wx (xyz[n] - c) + b - mu) + beta) = relu(P[n] + (s Wx)(xyz[n] - c) + b'),   P = feat (s Wf)^T
    P is computed ONCE PER POINT (a plain library GEMM, bf16 inputs / fp32 accumulation like the fused kernel) instead of
    once per grouped row, and the fused kernel gathers the N0-wide rows of P through an identity block in place of the
    D-wide feature rows: for cls_msg's sa2 (D = 320, N0 = 64 / 128, K = 32 / 64 / 128 neighbours) the first layer's
    work drops 16-64x, the stack's weights fit shared memory again (the pipelined kernel instead of the streaming one)
    and the gather moves 2.5-5x fewer bytes.  Exactly the reference's arithmetic up to the extra bf16 rounding of P."""

    def __init__(self):
        self.key = None
        self.packed = self.dims = self.wf = None

    def get(self, convs, bns, C, D):
        tensors = _stack_tensors(convs, bns)
        key = (_cache_generation, C, D, tuple((t.data_ptr(), t._version) for t in tensors))
        if key != self.key:
            conv0, bn0 = convs[0], bns[0]
            n0 = conv0.out_channels
            w0 = conv0.weight.detach().reshape(n0, D + C)                       # MSG channel order: [feat(D), xyz(C)]
            scale = bn0.weight.detach() / torch.sqrt(bn0.running_var + bn0.eps)
            bias = (conv0.bias.detach() - bn0.running_mean) * scale + bn0.bias.detach()
            self.wf = (w0[:, :D] * scale[:, None]).to(torch.bfloat16).contiguous()    # [N0, D]
            w_syn = torch.cat([torch.eye(n0, dtype=w0.dtype, device=w0.device), w0[:, D:] * scale[:, None]], dim=1)
            one, zero = torch.ones_like(bias), torch.zeros_like(bias)
            eps = bn0.eps
            layers = [(w_syn.contiguous(), bias.contiguous(), one, zero, zero, one - eps)]     # folds to scale 1, bias `bias`
            layers += [(conv.weight.detach().reshape(conv.out_channels, conv.in_channels), conv.bias.detach(),
                        bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var)
                       for conv, bn in zip(list(convs)[1:], list(bns)[1:])]
            assert all(bn.eps == eps for bn in bns)
            self.packed, self.dims = _backend().pack(layers, eps, PREC_BF16, C, ORDER_MSG)
            self.key = key
        return self.packed, self.dims, self.wf


class _FoldedHead:
    """Derived cache of the FC head behind the last set abstraction (fc1/bn1 -> fc2/bn2 -> fc3, pointnet2_cls_ssg.py:31-38 and
    the same tail in every pose head): eval-mode BatchNorm1d folded into the Linear weights, packed [K/4][H][4] for
    pn2_fc_head_f32 (include/pn2.h).  Same invalidation rule as _FoldedMLP."""

    def __init__(self):
        self.key = None
        self.tensors = None          # (w1p, b1, w2p, b2, w3, b3, mean_mlp | None)

    @staticmethod
    def supported(fcs) -> bool:
        def ok(K, H):
            if H % 8 or H // 8 > 256 or 256 % (H // 8):
                return False
            return K % (16 * (256 // (H // 8))) == 0
        k0, h1, h2, out = fcs[0].in_features, fcs[0].out_features, fcs[1].out_features, fcs[2].out_features
        return len(fcs) == 3 and 256 <= k0 <= 2048 and h1 <= 1024 and h2 <= 1024 and out <= 16 and ok(k0, h1) and ok(h1, h2)

    @staticmethod
    def _fold(fc, bn):
        s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        return fc.weight.detach() * s[:, None], (fc.bias.detach() - bn.running_mean) * s + bn.bias.detach()

    @staticmethod
    def _pack(w):                    # [H, K] -> [K/4][H][4]
        H, K = w.shape
        return w.reshape(H, K // 4, 4).permute(1, 0, 2).contiguous()

    def get(self, fcs, bns, mean_mlp=None):
        """fcs = (fc1, fc2, fc3), bns = (bn1, bn2); mean_mlp = (mean_fc1, mean_bn1, mean_fc2) of the translation heads or None."""
        mods = list(fcs) + list(bns) + (list(mean_mlp) if mean_mlp else [])
        ts = [t for m in mods for t in _leaf_tensors(m)]
        key = (_cache_generation, tuple((t.data_ptr(), t._version) for t in ts))
        if key != self.key:
            with torch.no_grad():
                w1, b1 = self._fold(fcs[0], bns[0])
                w2, b2 = self._fold(fcs[1], bns[1])
                mm = None
                if mean_mlp:
                    f1, bn, f2 = mean_mlp
                    mw1, mb1 = self._fold(f1, bn)
                    mm = torch.cat([mw1.reshape(-1), mb1, f2.weight.detach().reshape(-1), f2.bias.detach()]).float().contiguous()
                self.tensors = (self._pack(w1.float()), b1.float().contiguous(), self._pack(w2.float()), b2.float().contiguous(),
                                fcs[2].weight.detach().float().contiguous(), fcs[2].bias.detach().float().contiguous(), mm)
            self.key = key
        return self.tensors


def fc_head(x, fcs, bns, kind, cache, mean=None, mean_mlp=None):
    """Eval-mode FC head in ONE launch: x [B,K0] (CUDA fp32) -> (out [B,OUT], aux): see _pn2_native.Native.fc_head."""
    w1p, b1, w2p, b2, w3, b3, mm = cache.get(fcs, bns, mean_mlp)
    return _backend().fc_head(x, w1p, b1, w2p, b2, w3, b3, kind, mean=mean, mean_mlp=mm)


def _finish(new_points_bsc, pooled_bsc, pooled_bf16=None):
    new_points = new_points_bsc.permute(0, 2, 1)       # view of contiguous [B,S,C], like the reference (:173)
    new_feature = pooled_bsc.permute(0, 2, 1)
    if _reference_strides:
        new_feature = new_feature.contiguous()
    if pooled_bf16 is not None:
        # bf16 mode: the same values rounded to bf16 ride along for the next set-abstraction layer, which then
        # gathers them with cp.async instead of converting fp32 rows (derived data, dropped if the tensor changes)
        new_feature._pn2_bf16 = (pooled_bf16, new_feature._version)
    return new_points, new_feature


def _bf16_twin(feature_bdn, B, N, D):
    """The bf16 [B,N,D] copy attached by _finish to a feature tensor this module returned, if still current."""
    twin = getattr(feature_bdn, "_pn2_bf16", None)
    if twin is None or _precision != "bf16":
        return None
    t, version = twin
    if version != feature_bdn._version or t.shape != (B, N, D) or t.device != feature_bdn.device:
        return None
    return t


def _shared_mlp_train(x, convs, bns):
    """The reference's op sequence on ATen (:168-172 / :214-218); x [B, Cin, K, S].  Kept as the comparator of the fused
    training path (PN2_TRAIN_FUSED=0 selects it)."""
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return torch.max(x, 2)[0]


class _gemm_precision_of_the_reference_convs:
    """The reference's 1x1 convolutions run through cuDNN, whose TF32 switch is ``torch.backends.cudnn.allow_tf32``
    (default True); the GEMMs that stand in for them follow that switch instead of the matmul one (default False)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(torch.backends.cudnn.allow_tf32)

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def _own_gemm(rows, w) -> bool:
    """The training GEMMs run on pn2_rows_gemm (TF32 tensor-core product) whenever the reference's own convolutions would run
    in TF32 (torch.backends.cudnn.allow_tf32, default True); with TF32 switched off they stay full-fp32 library GEMMs."""
    return (_train_own_gemm and rows.is_cuda and bool(torch.backends.cudnn.allow_tf32) and rows.dtype == torch.float32
            and rows.is_contiguous() and rows.shape[1] % 4 == 0 and w.shape[0] % 4 == 0 and w.shape[1] % 4 == 0)


class _RowsLinear(torch.autograd.Function):
    """y [M, Cout] = x [M, Cin] @ w[Cout, Cin]^T: a 1x1 convolution on channels-last rows, bias left to the BatchNorm."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        if _own_gemm(x, w):                                        # TF32 tcgen05 product, fp32 rows out
            return _backend().rows_gemm(x, w if (w.stride(1) == 1 and w.stride(0) % 4 == 0) else w.contiguous())
        with _gemm_precision_of_the_reference_convs():
            return torch.mm(x, w.t())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        dy = dy.contiguous()
        with _gemm_precision_of_the_reference_convs():
            if ctx.needs_input_grad[0]:
                dx = _backend().rows_gemm(dy, w.t().contiguous()) if _own_gemm(dy, w) else torch.mm(dy, w)
            if ctx.needs_input_grad[1]:
                own = _own_gemm(dy, w) and x.is_contiguous() and 0 < x.shape[0] < 2 ** 31
                dw = _backend().rows_wgrad(dy, x) if own else _rows_outer(dy, x)
        return dx, dw


def _rows_linear(x, w):
    """x [M, Cin] @ w[Cout, Cin]^T.  Rows that are not a multiple of 16 bytes (the 10 / 13-wide grouped coordinates of the
    pose heads, 330-wide MSG inputs) are padded with zero columns: cuBLAS only has its slow element-aligned kernels for them."""
    pad = (-x.shape[1]) % 4
    if pad:
        x, w = F.pad(x, (0, pad)), F.pad(w, (0, pad))
    return _RowsLinear.apply(x, w)


def _rows_outer(dy, x, chunk=8192):
    """dy^T x for tall operands [M, Cout], [M, Cin] (the weight gradient of a 1x1 convolution: M = B*S*K in the millions,
    a few thousand output elements).  A single GEMM call leaves cuBLAS with a split-K problem it runs at a third of the
    HBM rate; as a batch of M/chunk independent GEMMs plus a small sum it streams both operands once at full rate
    (B200, TF32: 386 -> 211 us for 2.1 M x 96 x 64, 691 -> 233 us for Cin = 10)."""
    M = dy.shape[0]
    c = math.gcd(M, chunk)
    if c < 1024 or M // c < 2:
        return torch.mm(dy.t(), x)
    P = M // c
    return torch.bmm(dy.reshape(P, c, -1).transpose(1, 2), x.reshape(P, c, -1)).sum(0)


class _BnReluRows(torch.autograd.Function):
    """relu(batch_norm(y + bias)) on rows y [M, C] with batch statistics, optionally followed by the max over groups of K
    rows -- the product kernels of csrc/pn2_bn_train.cu.  The bias of the convolution cancels against the batch mean, so
    it is not added (its gradient is exactly zero; the running mean is shifted by it outside)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, bias, eps, K, running):
        z, stats, arg = _backend().bn_relu_train_fwd(y, gamma, beta, eps, K, running)
        ctx.save_for_backward(y, stats, arg)
        ctx.K = K
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(stats)
        return z, stats

    @staticmethod
    def backward(ctx, dz, _dstats):
        y, stats, arg = ctx.saved_tensors
        dy, sums = _backend().bn_relu_train_bwd(dz, y, arg, stats, ctx.K)
        dbias = torch.zeros_like(sums[0]) if ctx.has_bias else None
        return dy, sums[1], sums[0], dbias, None, None, None


def _bn_relu_rows(y, conv_bias, bn, K=0):
    """One ``relu(bn(.))`` of the shared MLP in training mode on rows [M, C] (+ fused max over K rows): product kernels
    for the data, nn.BatchNorm2d's bookkeeping (running statistics, momentum, num_batches_tracked) here."""
    running = None
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1
        f = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)     # cumulative average: one sync
        running = (bn.running_mean, bn.running_var, conv_bias.detach() if conv_bias is not None else None, float(f))
    z, _ = _BnReluRows.apply(y, bn.weight, bn.bias, conv_bias, bn.eps, K, running)            # running statistics in the same launch
    return z


def _shared_mlp_train_rows(rows, convs, bns, first=None):
    """Training-mode shared MLP + max over K on the grouped rows [B, S, K, Cin] as they leave the grouping kernel:
    per layer one GEMM and the fused BatchNorm + ReLU kernels; the last layer's max-pool is fused into its BatchNorm, so
    its activations never reach HBM.  ``first`` = pre-activations of layer 0 [B, S, K, C0] when the caller has already
    evaluated the first convolution (factorised MSG branch).  Returns [B, Cout, S] like torch.max(x, 2)[0] of (:172)."""
    B, S, K = rows.shape[:3]
    convs, bns = list(convs), list(bns)
    x = rows.reshape(B * S * K, -1)
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        if i == 0 and first is not None:
            y = first.reshape(B * S * K, -1)
        else:
            y = _rows_linear(x, conv.weight.view(conv.out_channels, -1))
        x = _bn_relu_rows(y, conv.bias, bn, K if i == len(convs) - 1 else 0)
    out = x.view(B, S, -1).permute(0, 2, 1)
    return out.contiguous() if _reference_strides else out


class PointNetSetAbstraction(nn.Module):
    def __init__(self, point_number, sample_number, radius, in_channel, mlp, group_all=False):
        super().__init__()
        self.point_number = point_number
        self.radius = radius
        self.sample_number = sample_number
        self.group_all = group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        width = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(width, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            width = out_channel
        self._folded = _FoldedMLP()

    def forward(self, points, feature):
        """points [B,C,N], feature [B,D,N] or None -> new_points [B,C,S], new_feature [B,Cout,S]."""
        be = _backend()
        pts = points.permute(0, 2, 1)
        feat = feature.permute(0, 2, 1) if feature is not None else None
        B, N, C = pts.shape
        if self.training:
            return self._forward_train(pts, feat)
        order = ORDER_ALL if self.group_all else ORDER_SSG
        packed, dims, prec = self._folded.get(self.mlp_convs, self.mlp_bns, C, order)
        if self.group_all:
            new_points = torch.zeros(B, 1, C, device=pts.device)
            S, K, idx32 = 1, N, None
        else:
            S, K = self.point_number, self.sample_number
            start = _draw_start(B, N, pts.device)
            _, _, new_points = be.fps(pts, start, S, want_idx64=False, want_new_points=True)
            _, idx32 = be.ball_query(new_points, pts, _radius_sq(self.radius), K, want_idx64=False, want_idx32=True,
                                     empty_flag=_empty_flag(pts.device))
        alloc = torch.zeros if be.sa_needs_zero(S, K, dims, prec) else torch.empty
        pooled = alloc(B, S, dims[-1], dtype=torch.float32, device=pts.device)
        side = be.sa_bf16_side_io(K, C, order, dims, prec)
        # the previous layer's bf16 rows, if it left any: every bf16 kernel gathers them instead of converting the fp32 rows
        twin_in = _bf16_twin(feature, B, N, feat.shape[-1]) if feat is not None else None
        twin_out = torch.empty(B, S, dims[-1], dtype=torch.bfloat16, device=pts.device) if side else None
        be.sa_fused(pts, feat, new_points if not self.group_all else None, idx32, order, dims, packed, prec, pooled, 0,
                    feature_bf16=twin_in, out_bf16=twin_out)
        return _finish(new_points, pooled, twin_out)

    def _forward_train(self, pts, feat):
        # Training mode (batch-statistics BatchNorm + autograd): sampling and grouping run on our kernels, the shared MLP
        # layer by layer on the grouped rows (_shared_mlp_train_rows: GEMM + the fused BatchNorm / ReLU / max-pool kernels
        # of csrc/pn2_bn_train.cu, forward and backward).
        be = _backend()
        B, N, C = pts.shape
        if self.group_all:
            new_points = torch.zeros(B, 1, C, device=pts.device)
            rows = pts.reshape(B, 1, N, C) if pts.is_contiguous() else pts.contiguous().view(B, 1, N, C)
            if feat is not None:
                rows = torch.cat([rows, feat.reshape(B, 1, N, -1)], dim=-1)
        else:
            S, K = self.point_number, self.sample_number
            start = _draw_start(B, N, pts.device)
            _, _, new_points = be.fps(pts.detach(), start, S, want_idx64=False, want_new_points=True)
            _, idx32 = be.ball_query(new_points, pts.detach(), _radius_sq(self.radius), K, want_idx64=False, want_idx32=True,
                                     empty_flag=_empty_flag(pts.device))
            rows = be.group(pts.detach(), None, idx32, new_points, ORDER_SSG)
            if feat is not None:
                rows = torch.cat([rows, _gather_rows(feat, idx32)], dim=-1)
        if _train_fused:
            pooled = _shared_mlp_train_rows(rows, self.mlp_convs, self.mlp_bns)
        else:
            pooled = _shared_mlp_train(rows.permute(0, 3, 2, 1), self.mlp_convs, self.mlp_bns)
        return new_points.permute(0, 2, 1), pooled


class _GatherRows(torch.autograd.Function):
    """index_points(feature, idx) of training mode (:112 / :208) on the product kernels: gather forward, scatter-add backward."""

    @staticmethod
    def forward(ctx, feat, idx32):
        ctx.save_for_backward(idx32)
        ctx.N = feat.shape[1]
        return _backend().gather_rows(feat, idx32)

    @staticmethod
    def backward(ctx, grad):
        (idx32,) = ctx.saved_tensors
        return _backend().scatter_add_rows(grad, idx32, ctx.N), None


def _gather_rows(feat, idx32):
    """Differentiable feature gather [B,N,D], [B,S,K] -> [B,S,K,D] (scatter-add in backward)."""
    return _GatherRows.apply(feat, idx32)


class PointNetSetAbstractionMsg(nn.Module):
    def __init__(self, point_number, sample_number_list, radius_list, in_channel, mlp_list, num_category=0):
        super().__init__()
        self.point_number = point_number
        self.radius_list = radius_list
        self.sample_number_list = sample_number_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for widths in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            width = in_channel + 3 + num_category
            for out_channel in widths:
                convs.append(nn.Conv2d(width, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                width = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self._folded = [_FoldedMLP() for _ in mlp_list]
        self._factorised = [_FactorisedFirstLayer() for _ in mlp_list]

    def _use_factorised(self, i, K, C, D):
        """bf16 mode, features wider than the branch's first layer, and the pipelined kernel covers the factorised stack."""
        if _precision != "bf16" or not _factorise_first_layer or D == 0:
            return False
        n0 = self.conv_blocks[i][0].out_channels
        if D <= n0 or n0 % 64 != 0:
            return False
        dims = [n0 + C] + [c.out_channels for c in self.conv_blocks[i]]
        return bool(_backend().sa_bf16_side_io(K, C, ORDER_MSG, dims, PREC_BF16))

    def forward(self, points, feature):
        be = _backend()
        pts = points.permute(0, 2, 1)
        feat = feature.permute(0, 2, 1) if feature is not None else None
        B, N, C = pts.shape
        S = self.point_number
        start = _draw_start(B, N, pts.device)            # ONE draw per layer, shared by all radii (:201)
        if self.training:
            return self._forward_train(pts, feat, start)
        _, _, new_points = be.fps(pts, start, S, want_idx64=False, want_new_points=True)
        plans, fact = [], []
        D = feat.shape[-1] if feat is not None else 0
        for i in range(len(self.radius_list)):
            if self._use_factorised(i, self.sample_number_list[i], C, D):
                packed, dims, wf = self._factorised[i].get(self.conv_blocks[i], self.bn_blocks[i], C, D)
                plans.append((packed, dims, PREC_BF16))
                fact.append(wf)
            else:
                packed, dims, prec = self._folded[i].get(self.conv_blocks[i], self.bn_blocks[i], C, ORDER_MSG)
                plans.append((packed, dims, prec))
                fact.append(None)
        total = sum(d[-1] for _, d, _ in plans)
        zero = any(be.sa_needs_zero(S, K, d, prec) for K, (_, d, prec) in zip(self.sample_number_list, plans))
        pooled = (torch.zeros if zero else torch.empty)(B, S, total, dtype=torch.float32, device=pts.device)
        sides = [be.sa_bf16_side_io(K, C, ORDER_MSG, d, prec) for K, (_, d, prec) in zip(self.sample_number_list, plans)]
        twin_in = _bf16_twin(feature, B, N, feat.shape[-1]) if feat is not None else None
        # the bf16 twin of the concatenated output is only complete if every branch's kernel writes its slice
        twin_out = torch.empty(B, S, total, dtype=torch.bfloat16, device=pts.device) if all(sides) else None
        co = 0
        feat_bf16 = None
        # one scan of the cloud serves every radius of the layer (the distances do not depend on it)
        groups = be.ball_query_multi(new_points, pts, [_radius_sq(r) for r in self.radius_list], self.sample_number_list,
                                     empty_flag=_empty_flag(pts.device))
        for i, ((_, idx32), K, (packed, dims, prec), wf) in enumerate(zip(groups, self.sample_number_list, plans, fact)):
            if wf is not None:
                # first layer factorised: P = feat (s Wf)^T once per point, gathered in place of the feature rows
                if feat_bf16 is None:       # bf16 twin of the previous layer if present, else the fp32 rows (converted on load)
                    feat_bf16 = twin_in if twin_in is not None else (feat if feat.is_contiguous() else feat.contiguous())
                # [B,N,N0] bf16 rows, fp32 accumulation: pn2_rows_gemm (tcgen05), one launch per branch
                p_rows = be.rows_gemm(feat_bf16.view(B * N, D), wf, torch.bfloat16).view(B, N, -1)
                be.sa_fused(pts, None, new_points, idx32, ORDER_MSG, dims, packed, prec, pooled, co,
                            feature_bf16=p_rows, out_bf16=twin_out,
                            algo_dims=[D + C] + [c.out_channels for c in self.conv_blocks[i]])
            else:
                be.sa_fused(pts, feat, new_points, idx32, ORDER_MSG, dims, packed, prec, pooled, co,   # branch concat (:222) is the channel offset
                            feature_bf16=twin_in, out_bf16=twin_out)
            co += dims[-1]
        return _finish(new_points, pooled, twin_out)

    def _forward_train(self, pts, feat, start):
        be = _backend()
        S = self.point_number
        _, _, new_points = be.fps(pts.detach(), start, S, want_idx64=False, want_new_points=True)
        outs = []
        groups = be.ball_query_multi(new_points, pts.detach(), [_radius_sq(r) for r in self.radius_list], self.sample_number_list,
                                     empty_flag=_empty_flag(pts.device))
        for i, (_, idx32) in enumerate(groups):
            rows = be.group(pts.detach(), None, idx32, new_points, ORDER_MSG)            # [B,S,K,C] xyz - centre
            convs, bns = self.conv_blocks[i], self.bn_blocks[i]
            if feat is not None and _factorise_first_layer and feat.shape[-1] > convs[0].out_channels:
                outs.append(_msg_branch_train_factorised(rows, feat, idx32, convs, bns))
                continue
            if feat is not None:
                rows = torch.cat([_gather_rows(feat, idx32), rows], dim=-1)
            if _train_fused:
                outs.append(_shared_mlp_train_rows(rows, convs, bns))
            else:
                outs.append(_shared_mlp_train(rows.permute(0, 3, 2, 1), convs, bns))
        return new_points.permute(0, 2, 1), torch.cat(outs, dim=1)


def _msg_branch_train_factorised(rows_xyz, feat, idx32, convs, bns):
    """Training-mode MSG branch with the first convolution factorised (see _FactorisedFirstLayer): the 1x1 convolution is
    linear in [feat[n], xyz[n] - c], so its feature part is evaluated once per POINT and gathered, instead of gathering the
    D-wide rows and convolving every grouped copy.  Same function as `_shared_mlp_train(cat([gather(feat), xyz]))` up to
    fp32 reassociation (the sum is split in two); BatchNorm sees the same pre-activations, autograd the same graph
    leaves.  The [B, D+C, K, S] tensor of the reference (:209) is never built."""
    conv0, bn0 = convs[0], bns[0]
    D = feat.shape[-1]
    w = conv0.weight.reshape(conv0.out_channels, -1)                  # [N0, D + C], MSG order [feat, xyz] (:209)
    if _train_fused:       # bias left to the BatchNorm (it cancels against the batch mean)
        B, N = feat.shape[:2]
        p_rows = _rows_linear(feat.reshape(B * N, D), w[:, :D]).view(B, N, -1)
        y = _gather_rows(p_rows, idx32)
        y += _rows_linear(rows_xyz.reshape(-1, rows_xyz.shape[-1]), w[:, D:]).view(y.shape)
        return _shared_mlp_train_rows(rows_xyz, convs, bns, first=y)
    p_rows = torch.matmul(feat, w[:, :D].t())                         # [B, N, N0] once per point
    y = _gather_rows(p_rows, idx32) + torch.matmul(rows_xyz, w[:, D:].t()) + conv0.bias      # [B, S, K, N0]
    x = F.relu(bn0(y.permute(0, 3, 2, 1)))
    for conv, bn in zip(list(convs)[1:], list(bns)[1:]):
        x = F.relu(bn(conv(x)))
    return torch.max(x, 2)[0]
