"""Readers for the fixtures under tests/golden/ (written by oracle/make_golden.py from the live reference)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out = {}
    for key in z.files:
        case, field = key.split("/", 1)
        out.setdefault(case, {})[field] = z[key]
    return out


def index_cases():
    """case -> dict(points [B,N,C] f32, start, S, r, K, fps, ball, sqdist_sample, layout)."""
    cases = _load("indices.npz")
    for name, d in cases.items():
        d["layout"] = name.rsplit("_", 1)[1]
        for f in ("S", "K"):
            d[f] = int(d[f])
        d["r"] = float(d["r"])
    return cases


def module_cases():
    cases = _load("modules.npz")
    for d in cases.values():
        d["sd"] = {k[3:]: torch.from_numpy(v) for k, v in d.items() if k.startswith("sd/")}
    return cases


def model_cases():
    return _load("models.npz")


def v1_cases():
    """head name -> dict(points [B,N,C], mean [B,3], pooled [B,1024], out0[, out1], sd_checksum, sd_keys)."""
    return _load("v1.npz")


def npfps_cases():
    """case -> dict(points (stored or regenerated from the seed), start, idx, rows_sha, number) for the numpy FPS copies."""
    import hashlib
    cases = _load("npfps.npz")
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "make_golden.py")).read()
    ns = {"np": np}
    start = src.index("def np_cloud(")
    exec(src[start:src.index("def gen_npfps(")], ns)          # np_cloud + NPFPS_SPECS only (the rest needs the reference)
    for name, N, C, dt, number, seed, dup, store in ns["NPFPS_SPECS"]:
        d = cases[name]
        if "points" not in d:
            d["points"] = ns["np_cloud"](N, C, np.dtype(dt), seed, dup)
        d["input_ok"] = hashlib.sha256(d["points"].tobytes()).hexdigest() == str(d["sha"][0])
        d["number"], d["start"] = number, int(d["start"][0])
    return cases


def as_layout(points_np, layout, device="cpu"):
    """[B,N,C] array -> tensor in the layout the reference saw: permuted view of [B,C,N], or contiguous."""
    x = torch.from_numpy(np.ascontiguousarray(points_np)).to(device)
    if layout == "strided":
        return x.permute(0, 2, 1).contiguous().permute(0, 2, 1)
    return x.contiguous()
