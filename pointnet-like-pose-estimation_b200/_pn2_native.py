"""ctypes binding of libpn2_sm100.so (include/pn2.h) at tensor level.

This is the only module that touches the C ABI.  Every method takes CUDA fp32 / int64 / int32 tensors,
allocates the outputs with torch (the library never allocates), launches on the calling thread's current
CUDA stream and maps a non-zero return code to RuntimeError(pn2_last_error_string()).  There is no CPU
path: a CPU tensor, or a missing shared library, is an error.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libpn2_sm100.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

ORDER_SSG, ORDER_MSG, ORDER_ALL = 0, 1, 2
PREC_F32, PREC_BF16 = 0, 1
MAX_COORD = 16
ABI_VERSION = 15
HEAD_PLAIN, HEAD_CLS, HEAD_TRANS, HEAD_SIGN = 0, 1, 2, 3

_i64, _i32, _f32, _vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

_SIGNATURES = {
    "pn2_version": ([], _i32),
    "pn2_last_error_string": ([], ctypes.c_char_p),
    "pn2_fps_f32": ([_vp, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp], _i32),
    "pn2_index_points_f32": ([_vp, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp], _i32),
    "pn2_square_distance_f32": ([_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "pn2_ball_query_f32": ([_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _i32,
                            _vp, _vp, _vp, _vp], _i32),
    "pn2_ball_query_multi_f32": ([_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32,
                                  _vp, _vp, _vp, _vp, _vp, _vp], _i32),
    "pn2_group_f32": ([_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _vp,
                       _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "pn2_packed_bytes": ([_i32, _vp, _i32], _i64),
    "pn2_pack_weights": ([_i32, _vp, _vp, _f32, _i32, _i32, _i32, _vp, _vp], _i32),
    "pn2_sa_needs_zero_init": ([_i32, _i32, _i32, _vp, _i32], _i32),
    "pn2_sa_bf16_side_io": ([_i32, _i32, _i32, _i32, _vp, _i32], _i32),
    "pn2_sa_fused": ([_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _vp,
                      _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32,
                      _vp, _i64, _i64, _i32, _vp, _vp, _vp], _i32),
    "pn2_normalize_splice_f32": ([_vp, _i64, _i64, _i64, _i32, _i32, _vp, _i32, _vp, _vp], _i32),
    "pn2_fps_xyz_scratch_bytes": ([_i32, _i32, _i32], _i64),
    "pn2_fps_xyz_f64": ([_vp, _i64, _i64, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp], _i32),
    "pn2_fps_xyz_f32": ([_vp, _i64, _i64, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp], _i32),
    "pn2_bn_train_ws_bytes": ([_i32], _i64),
    "pn2_bn_relu_train_fwd_f32": ([_vp, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp], _i32),
    "pn2_bn_relu_train_bwd_f32": ([_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp], _i32),
    "pn2_fc_head_f32": ([_vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp], _i32),
    "pn2_rows_gemm": ([_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _i64, _i32, _i32, _vp], _i32),
    "pn2_rows_wgrad_ws_bytes": ([_i64, _i32, _i32], _i64),
    "pn2_rows_wgrad_tf32": ([_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp], _i32),
    "pn2_gather_rows_f32": ([_vp, _i64, _i64, _i64, _vp, _i32, _i32, _i64, _i32, _vp, _vp], _i32),
    "pn2_scatter_add_rows_f32": ([_vp, _vp, _i32, _i32, _i64, _i32, _vp, _vp], _i32),
    "pn2_upload_i64": ([_vp, _vp, _i64, _vp], _i32),
    "pn2_selftest_umma": ([_i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "pn2_debug_trace": ([_vp], _i32),
    "pn2_selftest_tma_gather": ([_vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp], _i32),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_lock = threading.Lock()


def load_library(path: str = LIB_PATH):
    """dlopen the shared library and attach signatures.  Raises if it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(path):
                raise ImportError(
                    f"{path} is missing: build it with `python __graft_entry__.py build` "
                    f"(nvcc -gencode arch=compute_100a,code=sm_100a); there is no fallback implementation")
            lib = ctypes.CDLL(path)
            for name, (argtypes, restype) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = restype
            if lib.pn2_version() != ABI_VERSION:
                raise ImportError(f"{path}: ABI version {lib.pn2_version()} != {ABI_VERSION}; rebuild")
            _lib = lib
    return _lib


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t: torch.Tensor) -> int:
    """cudaStream_t of torch's current stream on t's device (the raw-handle query: the Stream object costs ~8 us per call)."""
    if _raw_stream is not None:
        idx = t.device.index
        return _raw_stream(idx if idx is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(t.device).cuda_stream


class _NoCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_CTX = _NoCtx()


def _on(device):
    """Context that makes `device` current for a launch; nothing when it already is (the common case: ~6 us saved per call)."""
    idx = device.index if isinstance(device, torch.device) else torch.device(device).index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_CTX
    return torch.cuda.device(idx)


def _need_cuda(name: str, t: torch.Tensor, dtype) -> None:
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        dev = getattr(t, "device", type(t))
        raise RuntimeError(f"pn2: `{name}` must be a CUDA tensor (got {dev}); this build has no CPU path")
    if t.dtype != dtype:
        raise RuntimeError(f"pn2: `{name}` must be {dtype} (got {t.dtype})")


def _check(rc: int) -> None:
    if rc != 0:
        msg = load_library().pn2_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"pn2 (rc={rc}): {msg}")


def _dims_array(dims) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(dims, dtype=np.int32))


class Native:
    """Tensor-level façade over the C ABI; stateless apart from the loaded library."""

    name = "sm_100a"

    def __init__(self):
        self.lib = load_library()
        self.launches = 0      # kernels launched through the C ABI (bench.py reports it as gpu_launches)
        self.trace = None      # set to a list to record (name, meta, start_event, end_event) per launch
        self._memo = {}        # results of the pure plan queries (pn2_sa_needs_zero_init, pn2_sa_bf16_side_io)

    @contextlib.contextmanager
    def _span(self, name, **meta):
        """Count the launch; when tracing, bracket it with CUDA events on the launching (current) stream."""
        self.launches += 1
        tr = self.trace
        if tr is None:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        tr.append((name, meta, s, e))

    # -- sampling ------------------------------------------------------------------------------------------
    def fps(self, points, start, number, want_idx64=True, want_idx32=False, want_new_points=False):
        """points [B,N,C] (any strides), start [B] int64 -> (idx64 [B,S] | None, idx32 | None, new_points [B,S,C] | None)."""
        _need_cuda("points", points, torch.float32)
        _need_cuda("start", start, torch.int64)
        B, N, C = points.shape
        S = int(number)
        dev = points.device
        i64 = torch.empty(B, S, dtype=torch.int64, device=dev) if want_idx64 else None
        i32 = torch.empty(B, S, dtype=torch.int32, device=dev) if want_idx32 else None
        npts = torch.empty(B, S, C, dtype=torch.float32, device=dev) if want_new_points else None
        start = start.contiguous()
        with _on(dev), self._span("fps", B=B, N=N, C=C, S=S):
            _check(self.lib.pn2_fps_f32(points.data_ptr(), *points.stride(), B, N, C, start.data_ptr(), S,
                                        i64.data_ptr() if want_idx64 else None,
                                        i32.data_ptr() if want_idx32 else None,
                                        npts.data_ptr() if want_new_points else None, _stream(points)))
        return i64, i32, npts

    def upload_i64(self, pinned, device):
        """pinned int64 CPU tensor (page-locked) -> device tensor, copied by a kernel on the current stream (pn2_upload_i64)."""
        assert pinned.dtype == torch.int64 and pinned.is_pinned() and pinned.is_contiguous()
        out = torch.empty(pinned.shape, dtype=torch.int64, device=device)
        with _on(device):
            _check(self.lib.pn2_upload_i64(pinned.data_ptr(), out.data_ptr(), pinned.numel(), _stream(out)))
        return out

    def index_points(self, src, idx, check=True):
        """src [B,N,C] strided, idx [B,...] int64 -> [B,...,C] contiguous."""
        _need_cuda("points", src, torch.float32)
        _need_cuda("idx", idx, torch.int64)
        B, N, C = src.shape
        idx_c = idx.contiguous()
        M = idx_c.numel() // max(B, 1)
        out = torch.empty(*idx.shape, C, dtype=torch.float32, device=src.device)
        flag = torch.zeros(1, dtype=torch.int32, device=src.device) if check else None
        with _on(src.device), self._span("index_points", B=B, N=N, C=C, M=M):
            _check(self.lib.pn2_index_points_f32(src.data_ptr(), *src.stride(), B, N, C, idx_c.data_ptr(), M,
                                                 out.data_ptr(), flag.data_ptr() if check else None, _stream(src)))
        if check and int(flag.item()) != 0:
            raise IndexError(f"index out of range in index_points (valid range [-{N}, {N}))")
        return out

    def square_distance(self, src, dst):
        _need_cuda("src", src, torch.float32)
        _need_cuda("dst", dst, torch.float32)
        B, S, C = src.shape
        N = dst.shape[1]
        out = torch.empty(B, S, N, dtype=torch.float32, device=src.device)
        with _on(src.device), self._span("square_distance", B=B, S=S, N=N, C=C):
            _check(self.lib.pn2_square_distance_f32(src.data_ptr(), *src.stride(), dst.data_ptr(), *dst.stride(),
                                                    B, S, N, C, out.data_ptr(), _stream(src)))
        return out

    def ball_query(self, centres, points, r2, number, want_idx64=True, want_idx32=False, empty_flag=None):
        """centres [B,S,C], points [B,N,C] (any strides), r2 = fp32(radius**2) -> (idx64 | None, idx32 | None) [B,S,K].
        empty_flag: optional int32 [1] device tensor, OR-ed with 1 when some centre has no point in radius (its row is N)."""
        _need_cuda("new_points", centres, torch.float32)
        _need_cuda("points", points, torch.float32)
        B, N, C = points.shape
        S = centres.shape[1]
        K = int(number)
        dev = points.device
        i64 = torch.empty(B, S, K, dtype=torch.int64, device=dev) if want_idx64 else None
        i32 = torch.empty(B, S, K, dtype=torch.int32, device=dev) if want_idx32 else None
        with _on(dev), self._span("ball_query", B=B, N=N, C=C, S=S, K=K, i64=bool(want_idx64)):
            _check(self.lib.pn2_ball_query_f32(centres.data_ptr(), *centres.stride(), points.data_ptr(), *points.stride(),
                                               B, S, N, C, float(r2), K,
                                               i64.data_ptr() if want_idx64 else None,
                                               i32.data_ptr() if want_idx32 else None,
                                               empty_flag.data_ptr() if empty_flag is not None else None, _stream(points)))
        return i64, i32

    def ball_query_multi(self, centres, points, r2_list, numbers, want_idx64=False, want_idx32=True, empty_flag=None):
        """All radii of one MSG layer in one scan: -> list of (idx64 | None, idx32 | None) [B,S,K_r] per radius; every list
        is bit-identical to ball_query(centres, points, r2_list[r], numbers[r])."""
        _need_cuda("new_points", centres, torch.float32)
        _need_cuda("points", points, torch.float32)
        B, N, C = points.shape
        S = centres.shape[1]
        R = len(r2_list)
        dev = points.device
        if R > 3:
            return [self.ball_query(centres, points, r2, K, want_idx64, want_idx32, empty_flag) for r2, K in zip(r2_list, numbers)]
        outs = [(torch.empty(B, S, int(K), dtype=torch.int64, device=dev) if want_idx64 else None,
                 torch.empty(B, S, int(K), dtype=torch.int32, device=dev) if want_idx32 else None) for K in numbers]
        r2 = (ctypes.c_float * R)(*[float(v) for v in r2_list])
        ks = (ctypes.c_int * R)(*[int(K) for K in numbers])
        p64 = (ctypes.c_void_p * R)(*[o[0].data_ptr() if o[0] is not None else None for o in outs])
        p32 = (ctypes.c_void_p * R)(*[o[1].data_ptr() if o[1] is not None else None for o in outs])
        with _on(dev), self._span("ball_query_multi", B=B, N=N, C=C, S=S, K=[int(K) for K in numbers]):
            _check(self.lib.pn2_ball_query_multi_f32(centres.data_ptr(), *centres.stride(), points.data_ptr(), *points.stride(),
                                                     B, S, N, C, R, r2, ks, p64 if want_idx64 else None,
                                                     p32 if want_idx32 else None,
                                                     empty_flag.data_ptr() if empty_flag is not None else None, _stream(points)))
        return outs

    # -- grouping ------------------------------------------------------------------------------------------
    def group(self, points, feature, idx32, centres, order):
        """Materialised grouped rows [B,S,K,C+D] (SSG / MSG / group_all channel order)."""
        _need_cuda("points", points, torch.float32)
        B, N, C = points.shape
        D = 0
        fptr, fstr = None, (0, 0, 0)
        if feature is not None:
            _need_cuda("feature", feature, torch.float32)
            D = feature.shape[-1]
            fptr, fstr = feature.data_ptr(), feature.stride()
        if order == ORDER_ALL:
            S, K = 1, N
            iptr = cptr = None
        else:
            _need_cuda("idx", idx32, torch.int32)
            _need_cuda("new_points", centres, torch.float32)
            idx32 = idx32.contiguous()
            centres = centres.contiguous()
            _, S, K = idx32.shape
            iptr, cptr = idx32.data_ptr(), centres.data_ptr()
        out = torch.empty(B, S, K, C + D, dtype=torch.float32, device=points.device)
        with _on(points.device), self._span("group", B=B, N=N, S=S, K=K, C=C, D=D):
            _check(self.lib.pn2_group_f32(points.data_ptr(), *points.stride(), fptr, *fstr, iptr, cptr,
                                          B, N, S, K, C, D, order, out.data_ptr(), _stream(points)))
        return out

    # -- shared MLP ------------------------------------------------------------------------------------------
    def pack(self, layers, eps, prec, C, order):
        """layers: [(W [Cout,Cin(,1,1)], b, gamma, beta, mean, var)] CUDA fp32 -> (packed uint8 tensor, dims)."""
        dims = [int(layers[0][0].shape[1])] + [int(l[0].shape[0]) for l in layers]
        dev = layers[0][0].device
        keep, ptrs = [], []
        for tup in layers:
            for t in tup:
                _need_cuda("layer parameter", t, torch.float32)
                t = t.detach().contiguous()
                keep.append(t)
                ptrs.append(t.data_ptr())
        d = _dims_array(dims)
        nbytes = self.lib.pn2_packed_bytes(len(layers), d.ctypes.data, prec)
        if nbytes < 0:
            _check(-1)
        packed = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        parr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        with _on(dev), self._span("pack", dims=tuple(dims)):
            _check(self.lib.pn2_pack_weights(len(layers), d.ctypes.data, ctypes.cast(parr, _vp), float(eps), prec,
                                             int(C), int(order), packed.data_ptr(), _stream(packed)))
        return packed, dims

    def sa_needs_zero(self, S, K, dims, prec) -> bool:
        key = ("z", int(S), int(K), tuple(dims), prec)
        hit = self._memo.get(key)
        if hit is not None:
            return hit
        d = _dims_array(dims)
        rc = self.lib.pn2_sa_needs_zero_init(int(S), int(K), len(dims) - 1, d.ctypes.data, prec)
        if rc < 0:
            _check(rc)
        self._memo[key] = bool(rc)
        return bool(rc)

    def sa_bf16_side_io(self, K, C, order, dims, prec) -> bool:
        """True when the kernel serving this configuration writes `out_bf16` (`feature_bf16` is read by every bf16 kernel)."""
        key = ("s", int(K), int(C), int(order), tuple(dims), prec)
        hit = self._memo.get(key)
        if hit is not None:
            return hit
        d = _dims_array(dims)
        rc = self.lib.pn2_sa_bf16_side_io(int(K), int(C), int(order), len(dims) - 1, d.ctypes.data, prec)
        if rc < 0:
            _check(rc)
        self._memo[key] = bool(rc)
        return bool(rc)

    def sa_fused(self, points, feature, centres, idx32, order, dims, packed, prec, out, co_off=0,
                 feature_bf16=None, out_bf16=None, algo_dims=None):
        """Fused gather -> MLP -> max.  out [B,S,Ctot] (channels-last) receives channels [co_off, co_off+dims[-1]).
        feature_bf16 [B,N,D] / out_bf16 [B,S,Ctot]: optional bf16 side tensors (see include/pn2.h).
        algo_dims: trace metadata only -- the layer widths of the module when `dims` describes a factorised first layer."""
        _need_cuda("points", points, torch.float32)
        _need_cuda("out", out, torch.float32)
        B, N, C = points.shape
        D = 0
        fptr, fstr = None, (0, 0, 0)
        if feature is not None:
            _need_cuda("feature", feature, torch.float32)
            D = feature.shape[-1]
            fptr, fstr = feature.data_ptr(), feature.stride()
        elif feature_bf16 is not None:          # bf16 rows only (bf16 mode): strides of the contiguous [B,N,D] tensor
            D = feature_bf16.shape[-1]
            fstr = (N * D, D, 1)
        if order == ORDER_ALL:
            S, K = 1, N
            iptr = cptr = None
        else:
            _need_cuda("idx", idx32, torch.int32)
            _need_cuda("new_points", centres, torch.float32)
            assert idx32.is_contiguous() and centres.is_contiguous()
            _, S, K = idx32.shape
            iptr, cptr = idx32.data_ptr(), centres.data_ptr()
        assert out.shape[0] == B and out.shape[1] == S and out.stride(2) == 1
        if feature_bf16 is not None:
            _need_cuda("feature_bf16", feature_bf16, torch.bfloat16)
            assert feature_bf16.shape == (B, N, D) and feature_bf16.is_contiguous()
        if out_bf16 is not None:
            _need_cuda("out_bf16", out_bf16, torch.bfloat16)
            assert out_bf16.shape == out.shape and out_bf16.stride() == out.stride()
        d = _dims_array(dims)
        with _on(points.device), self._span("sa_fused", B=B, N=N, S=S, K=K, C=C, D=D, dims=tuple(dims), prec=prec,
                                                             algo_dims=tuple(algo_dims) if algo_dims else None):
            _check(self.lib.pn2_sa_fused(points.data_ptr(), *points.stride(), fptr, *fstr, cptr, iptr,
                                         B, N, S, K, C, D, order, len(dims) - 1, d.ctypes.data, packed.data_ptr(), prec,
                                         out.data_ptr(), out.stride(0), out.stride(1), int(co_off),
                                         feature_bf16.data_ptr() if feature_bf16 is not None else None,
                                         out_bf16.data_ptr() if out_bf16 is not None else None, _stream(points)))
        return out

    def normalize_splice(self, xyz, category=None, num_category=0):
        """provider.normalization + provider.splice_torch + transpose, fused: xyz [B,N,3] -> [B, 3+num_category, N]."""
        _need_cuda("xyz", xyz, torch.float32)
        B, N, C = xyz.shape
        if C != 3:
            raise ValueError("normalize_splice expects xyz [B,N,3]")
        cptr = None
        if num_category:
            _need_cuda("category", category, torch.int64)
            assert category.shape == (B,) and category.is_contiguous()
            cptr = category.data_ptr()
        out = torch.empty(B, 3 + num_category, N, dtype=torch.float32, device=xyz.device)
        with _on(xyz.device), self._span("normalize_splice", B=B, N=N):
            _check(self.lib.pn2_normalize_splice_f32(xyz.data_ptr(), *xyz.stride(), B, N, cptr, int(num_category),
                                                     out.data_ptr(), _stream(xyz)))
        return out

    def fps_xyz(self, points, start, S):
        """The numpy FPS copies (provider.py:183-204 ...): points [B,N,>=3] float64 | float32 (first three columns are
        sampled), start [B] int64 -> idx [B,S] int64.  Arithmetic in the tensor's dtype, bit-exact with numpy."""
        if points.dtype not in (torch.float64, torch.float32):
            raise RuntimeError(f"pn2: `points` must be float64 or float32 (got {points.dtype})")
        _need_cuda("points", points, points.dtype)
        _need_cuda("start", start, torch.int64)
        B, N, C = points.shape
        if C < 3:
            raise ValueError("fps_xyz samples on the first three columns; the array has fewer")
        assert start.shape == (B,) and start.is_contiguous()
        idx = torch.empty(B, S, dtype=torch.int64, device=points.device)
        nbytes = self.lib.pn2_fps_xyz_scratch_bytes(B, N, points.element_size())
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=points.device) if nbytes > 0 else None
        fn = self.lib.pn2_fps_xyz_f64 if points.dtype == torch.float64 else self.lib.pn2_fps_xyz_f32
        with _on(points.device), self._span("fps_xyz", B=B, N=N, S=S):
            _check(fn(points.data_ptr(), *points.stride(), B, N, start.data_ptr(), int(S), idx.data_ptr(),
                      scratch.data_ptr() if scratch is not None else None, _stream(points)))
        return idx

    def bn_relu_train_fwd(self, y, gamma, beta, eps, K=0, running=None):
        """Training-mode relu(bn(y)) on rows y [M, C] (batch statistics): -> (z, stats [5, C], arg).
        K = 0: z [M, C], arg None.  K > 0: z [M/K, C] = max over each group of K rows, arg [M/K, C] int32.
        running = (running_mean, running_var, conv_bias | None, factor): updated in place in the same launch."""
        _need_cuda("y", y, torch.float32)
        _need_cuda("gamma", gamma, torch.float32)
        _need_cuda("beta", beta, torch.float32)
        M, C = y.shape
        assert y.is_contiguous() and gamma.is_contiguous() and beta.is_contiguous() and gamma.numel() == C == beta.numel()
        if M < 2:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(y.shape)}")
        ws = torch.empty(self.lib.pn2_bn_train_ws_bytes(C), dtype=torch.uint8, device=y.device)
        stats = torch.empty(5, C, dtype=torch.float32, device=y.device)
        z = torch.empty(M // K if K else M, C, dtype=torch.float32, device=y.device)
        arg = torch.empty(M // K, C, dtype=torch.int32, device=y.device) if K else None
        rm = rv = cb = None
        factor = 0.0
        if running is not None:
            rm_t, rv_t, cb_t, factor = running
            for name, t in (("running_mean", rm_t), ("running_var", rv_t)) + ((("conv_bias", cb_t),) if cb_t is not None else ()):
                _need_cuda(name, t, torch.float32)
                assert t.is_contiguous() and t.numel() == C
            rm, rv, cb = rm_t.data_ptr(), rv_t.data_ptr(), (cb_t.data_ptr() if cb_t is not None else None)
        with _on(y.device), self._span("bn_relu_train_fwd", M=M, C=C, K=K):
            _check(self.lib.pn2_bn_relu_train_fwd_f32(y.data_ptr(), M, C, int(K), gamma.data_ptr(), beta.data_ptr(), float(eps),
                                                      ws.data_ptr(), stats.data_ptr(), z.data_ptr(),
                                                      arg.data_ptr() if K else None, rm, rv, cb, float(factor), _stream(y)))
        return z, stats, arg

    def bn_relu_train_bwd(self, dz, y, arg, stats, K=0):
        """Backward of bn_relu_train_fwd: dz (gradient of z) -> (dy [M, C], sums [2, C] = dbeta, dgamma)."""
        _need_cuda("dz", dz, torch.float32)
        M, C = y.shape
        dz = dz.contiguous()
        assert dz.shape == ((M // K, C) if K else (M, C))
        ws = torch.empty(self.lib.pn2_bn_train_ws_bytes(C), dtype=torch.uint8, device=y.device)
        sums = torch.empty(2, C, dtype=torch.float32, device=y.device)
        dy = torch.empty_like(y)
        with _on(y.device), self._span("bn_relu_train_bwd", M=M, C=C, K=K):
            _check(self.lib.pn2_bn_relu_train_bwd_f32(dz.data_ptr(), y.data_ptr(), arg.data_ptr() if K else None, M, C, int(K),
                                                      stats.data_ptr(), ws.data_ptr(), sums.data_ptr(), dy.data_ptr(), _stream(y)))
        return dy, sums

    def fc_head(self, x, w1p, b1, w2p, b2, w3, b3, kind, mean=None, mean_mlp=None):
        """Fused FC head (include/pn2.h: pn2_fc_head_f32): x [B,K0] -> (out [B,OUT], aux | None); folded / packed weights
        from pointnet2_utils._FoldedHead."""
        _need_cuda("x", x, torch.float32)
        B, K0 = x.shape
        H1, H2, OUT = b1.numel(), b2.numel(), b3.numel()
        assert x.stride(1) == 1
        out = torch.empty(B, OUT, dtype=torch.float32, device=x.device)
        aux = None
        if kind == HEAD_CLS:
            aux = torch.empty(B, dtype=torch.int64, device=x.device)
        elif kind == HEAD_SIGN:
            aux = torch.empty(B, OUT, dtype=torch.float32, device=x.device)
        mptr, mstr = None, 0
        if kind == HEAD_TRANS:
            _need_cuda("mean", mean, torch.float32)
            assert mean.shape == (B, 3) and mean.stride(1) == 1
            mptr, mstr = mean.data_ptr(), mean.stride(0)
        with _on(x.device), self._span("fc_head", B=B, K0=K0, H1=H1, H2=H2, OUT=OUT):
            _check(self.lib.pn2_fc_head_f32(x.data_ptr(), x.stride(0), B, K0, H1, H2, OUT, w1p.data_ptr(), b1.data_ptr(),
                                            w2p.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), int(kind), mptr, mstr,
                                            mean_mlp.data_ptr() if mean_mlp is not None else None, out.data_ptr(),
                                            aux.data_ptr() if aux is not None else None, _stream(x)))
        return out, aux

    def rows_gemm(self, a, w, out_dtype=None, out=None):
        """C [M,N] = a [M,K] @ w [N,K]^T on the tensor cores (include/pn2.h: pn2_rows_gemm).  w bf16: bf16 product (fp32 rows of
        `a` converted on load); w fp32: TF32 product.  Rows of a / w may be strided views (unit stride along K)."""
        dt = {torch.float32: 0, torch.bfloat16: 1}
        if a.dtype not in dt or w.dtype not in dt:
            raise RuntimeError(f"pn2: rows_gemm takes fp32 / bf16 operands (got {a.dtype}, {w.dtype})")
        _need_cuda("a", a, a.dtype)
        _need_cuda("w", w, w.dtype)
        M, K = a.shape
        N = w.shape[0]
        assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
        out_dtype = out_dtype or (torch.bfloat16 if w.dtype == torch.bfloat16 else torch.float32)
        if out is None:
            out = torch.empty(M, N, dtype=out_dtype, device=a.device)
        assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in dt
        with _on(a.device), self._span("rows_gemm", M=M, N=N, K=K, tf32=w.dtype == torch.float32):
            _check(self.lib.pn2_rows_gemm(a.data_ptr(), a.stride(0), dt[a.dtype], w.data_ptr(), w.stride(0), dt[w.dtype],
                                          out.data_ptr(), out.stride(0), dt[out.dtype], M, N, K, _stream(a)))
        return out

    def gather_rows(self, feat, idx32):
        """feat [B,N,D] (unit channel stride), idx32 [B,S,K] -> [B,S,K,D] (training mode's feature gather)."""
        _need_cuda("feature", feat, torch.float32)
        _need_cuda("idx", idx32, torch.int32)
        B, N, D = feat.shape
        if feat.stride(2) != 1:
            feat = feat.contiguous()
        idx32 = idx32.contiguous()
        M = idx32.numel() // max(B, 1)
        out = torch.empty(*idx32.shape, D, dtype=torch.float32, device=feat.device)
        with _on(feat.device), self._span("gather_rows", B=B, N=N, M=M, D=D):
            _check(self.lib.pn2_gather_rows_f32(feat.data_ptr(), feat.stride(0), feat.stride(1), 1, idx32.data_ptr(), B, N, M, D,
                                                out.data_ptr(), _stream(feat)))
        return out

    def scatter_add_rows(self, grad, idx32, N):
        """Backward of gather_rows: grad [B,S,K,D] -> grad_feat [B,N,D]."""
        _need_cuda("grad", grad, torch.float32)
        grad = grad.contiguous()
        idx32 = idx32.contiguous()
        B, D = grad.shape[0], grad.shape[-1]
        M = idx32.numel() // max(B, 1)
        gfeat = torch.zeros(B, N, D, dtype=torch.float32, device=grad.device)
        with _on(grad.device), self._span("scatter_add_rows", B=B, N=N, M=M, D=D):
            _check(self.lib.pn2_scatter_add_rows_f32(grad.data_ptr(), idx32.data_ptr(), B, N, M, D, gfeat.data_ptr(), _stream(grad)))
        return gfeat

    def rows_wgrad(self, dy, x):
        """dW [Cout,Cin] = dy [M,Cout]^T @ x [M,Cin] (TF32 tensor-core product, deterministic): include/pn2.h pn2_rows_wgrad_tf32."""
        _need_cuda("dy", dy, torch.float32)
        _need_cuda("x", x, torch.float32)
        M, Cout = dy.shape
        Cin = x.shape[1]
        assert x.shape[0] == M and dy.stride(1) == 1 and x.stride(1) == 1
        ws = torch.empty(max(1, self.lib.pn2_rows_wgrad_ws_bytes(M, Cout, Cin)), dtype=torch.uint8, device=dy.device)
        dw = torch.empty(Cout, Cin, dtype=torch.float32, device=dy.device)
        with _on(dy.device), self._span("rows_wgrad", M=M, Cout=Cout, Cin=Cin):
            _check(self.lib.pn2_rows_wgrad_tf32(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, Cout, Cin,
                                                ws.data_ptr(), dw.data_ptr(), _stream(dy)))
        return dw

    def selftest_umma(self, n_cols, k_dim, variant, device="cuda:0", dump=False):
        err = torch.full((1,), -1.0, dtype=torch.float32, device=device)
        d = torch.zeros(256 if variant == 4 else 128, n_cols, dtype=torch.float32, device=device) if dump else None
        with _on(err.device):
            _check(self.lib.pn2_selftest_umma(n_cols, k_dim, variant, err.data_ptr(),
                                              d.data_ptr() if dump else None, _stream(err)))
        return err, d
