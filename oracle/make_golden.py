"""Generate tests/golden/*.npz from the LIVE reference -- TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference, read-only).  It imports the reference's own
model/pointnet2_utils.py and model files unmodified, feeds them the seeded synthetic clouds of
tests/cloudgen.py and stores inputs + reference outputs as small fixtures, after asserting that the C
restatement (oracle/pn2_oracle.c) reproduces them (bit-exact indices, 1e-5 features).  The fixtures are what
pins the oracle -- and, through the `-m gpu` tests, the CUDA kernels -- on machines without the reference.

    python oracle/make_golden.py            # writes tests/golden/
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
REF = "/root/reference/model"
OUT = os.path.join(ROOT, "tests", "golden")

import cloudgen  # noqa: E402
import pn2_oracle as orc  # noqa: E402


def ref_module(name):
    """Import a reference module by bare name from /root/reference/model (its own import convention)."""
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k == "pointnet2_utils"]:
            if not getattr(sys.modules[k], "__file__", "").startswith(REF):
                del sys.modules[k]
        m = importlib.import_module(name)
    finally:
        sys.path.remove(REF)
    assert m.__file__.startswith(REF), m.__file__
    if not hasattr(m, "torch"):
        m.torch = torch          # sign_ssg.py / sign_msg.py forget `import torch` (sign_ssg.py:34)
    return m


def layouts(x_bnc):
    """The two memory layouts that occur on the model path: permuted [B,C,N] view, and contiguous."""
    return {"strided": cloudgen.channels_first(x_bnc).permute(0, 2, 1), "contig": x_bnc.contiguous()}


def gen_indices(ref):
    """FPS + ball query call sites: C in {3,10}, both layouts, uniform / surface / duplicate-heavy, S>N."""
    cases = {}
    specs = [
        ("uni3", cloudgen.uniform_cloud(3, 256, 11), 64, 0.25, 16),
        ("surf3", cloudgen.surface_cloud(3, 300, 12), 100, 0.2, 32),
        ("dup3", cloudgen.duplicate_heavy(cloudgen.uniform_cloud(2, 200, 13), 0.6, 14), 64, 0.3, 24),
        ("uni10", cloudgen.with_category(cloudgen.uniform_cloud(3, 256, 15)), 64, 0.25, 16),
        ("dup10", cloudgen.with_category(cloudgen.duplicate_heavy(cloudgen.surface_cloud(2, 160, 16), 0.5, 17)), 48, 0.4, 32),
        ("over3", cloudgen.uniform_cloud(2, 40, 18), 64, 0.5, 8),       # S > N (pointnet2_utils.py:240-245)
        # N % 32 != 0 with dense 10-D / 6-D coordinates: exercises the leftover-point summation order of the strided layout
        ("rand10", torch.rand(2, 300, 10, generator=torch.Generator().manual_seed(19)), 64, 0.9, 16),
        ("rand6", torch.rand(3, 77, 6, generator=torch.Generator().manual_seed(20)), 40, 0.6, 12),
    ]
    for name, x, S, r, K in specs:
        for lay, pts in layouts(x).items():
            B, N, C = pts.shape
            torch.manual_seed(42)
            start = torch.randint(0, N, (B,), dtype=torch.long)
            torch.manual_seed(42)
            fps = ref.farthest_point_sample(pts, S)
            mine = orc.farthest_point_sample(pts, S, start)
            assert torch.equal(fps, mine), ("fps", name, lay)
            new_pts = ref.index_points(pts, fps)
            bq = ref.query_ball_point(r, K, pts, new_pts)
            mine = orc.query_ball_point(r, K, pts, new_pts)
            assert torch.equal(bq, mine), ("ball", name, lay, (bq != mine).sum())
            sd = ref.square_distance(new_pts, pts)
            mine = orc.square_distance(new_pts, pts)
            assert torch.equal(sd, mine), ("sqdist", name, lay, (sd != mine).sum())
            cases[f"{name}_{lay}"] = dict(points=x.numpy(), start=start.numpy(), S=S, r=r, K=K, fps=fps.numpy().astype(np.int32),
                                          ball=bq.numpy().astype(np.int32), sqdist_sample=sd[:, :4].numpy())
    flat = {}
    for k, d in cases.items():
        for f, v in d.items():
            flat[f"{k}/{f}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "indices.npz"), **flat)
    print("indices.npz:", len(cases), "cases")


def rand_bn_(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in mod.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                n = m.num_features
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.5)
                m.running_var.copy_(torch.rand(n, generator=g) * 1.5 + 0.5)
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.2)


def gen_modules(ref):
    """Small PointNetSetAbstraction / ...Msg layers: weights + inputs + reference outputs."""
    flat = {}
    x3 = cloudgen.uniform_cloud(2, 128, 21)
    f3 = torch.randn(2, 12, 128, generator=torch.Generator().manual_seed(22))
    x10 = cloudgen.with_category(cloudgen.surface_cloud(2, 128, 23))
    specs = {
        "ssg3": (lambda: ref.PointNetSetAbstraction(32, 16, 0.4, 3, [16, 24, 40], False), x3, None),
        "ssg3f": (lambda: ref.PointNetSetAbstraction(32, 8, 0.3, 12 + 3, [32, 20], False), x3, f3),
        "ssg10": (lambda: ref.PointNetSetAbstraction(24, 32, 0.5, 10, [16, 16, 32], False), x10, None),
        "all3f": (lambda: ref.PointNetSetAbstraction(None, None, None, 12 + 3, [32, 64], True), x3, f3),
        "msg3": (lambda: ref.PointNetSetAbstractionMsg(32, [8, 16, 64], [0.2, 0.4, 0.8], 0, [[8, 16], [16, 24], [16, 32]]), x3, None),
        "msg3f": (lambda: ref.PointNetSetAbstractionMsg(16, [16, 32], [0.3, 0.6], 12, [[16, 16, 32], [24, 40]]), x3, f3),
        "msg10": (lambda: ref.PointNetSetAbstractionMsg(32, [16, 32], [0.3, 0.6], 0, [[16, 32], [32, 48]], num_category=7), x10, None),
    }
    for name, (ctor, x, f) in specs.items():
        torch.manual_seed(5)
        m = ctor().eval()
        rand_bn_(m, 6)
        pts = cloudgen.channels_first(x)
        with torch.no_grad():
            torch.manual_seed(42)
            new_pts, new_feat = m(pts, f)
        flat[f"{name}/points"] = x.numpy()
        if f is not None:
            flat[f"{name}/feature"] = f.numpy()
        flat[f"{name}/new_points"] = new_pts.contiguous().numpy()
        flat[f"{name}/new_feature"] = new_feat.contiguous().numpy()
        for k, v in m.state_dict().items():
            flat[f"{name}/sd/{k}"] = v.numpy()
        # oracle check
        B, N = x.shape[:2]
        torch.manual_seed(42)
        start = torch.randint(0, N, (B,), dtype=torch.long)
        if isinstance(m, ref.PointNetSetAbstraction):
            o_pts, o_feat, _, _ = orc.set_abstraction(pts, f, m.point_number, m.sample_number, m.radius,
                                                      orc.layers_of(m.mlp_convs, m.mlp_bns), m.group_all, start)
        else:
            o_pts, o_feat, _, _ = orc.set_abstraction_msg(pts, f, m.point_number, m.sample_number_list, m.radius_list,
                                                          [orc.layers_of(c, b) for c, b in zip(m.conv_blocks, m.bn_blocks)], start)
        assert torch.equal(o_pts, new_pts), name
        err = (o_feat - new_feat).abs().max() / new_feat.abs().max()
        assert err < 1e-5, (name, float(err))
        print(f"  {name}: oracle vs reference features {float(err):.2e}")
    np.savez_compressed(os.path.join(OUT, "modules.npz"), **flat)
    print("modules.npz:", len(specs), "modules")


def gen_models():
    """Whole reference models (seeded construction, randomised BN) at B=2: inputs, state-dict checksum, outputs."""
    flat = {}
    names = ["pointnet2_cls_ssg", "pointnet2_cls_msg", "rotation_ssg", "rotation_msg",
             "translation_ssg", "translation_msg", "sign_ssg", "sign_msg"]
    for i, name in enumerate(names):
        mod = ref_module(name)
        torch.manual_seed(0)
        model = mod.get_model(7).eval()
        rand_bn_(model, 7)
        xyz = cloudgen.uniform_cloud(2, 1024, 100 + i)
        pose = not name.startswith("pointnet2_cls")
        x = cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz)
        args = (x,)
        if name.startswith("translation"):
            mean = xyz[:, :3, :].mean(1)             # train_translation.py:113 (first three points)
            args = (x, mean)
            flat[f"{name}/mean"] = mean.numpy()
        with torch.no_grad():
            torch.manual_seed(42)
            out = model(*args)
        outs = out if isinstance(out, tuple) else (out,)
        flat[f"{name}/xyz"] = xyz.numpy()
        for j, o in enumerate(outs):
            flat[f"{name}/out{j}"] = o.contiguous().numpy()
        sd = model.state_dict()
        flat[f"{name}/sd_checksum"] = np.array([float(sum(v.double().abs().sum() for v in sd.values()))])
        flat[f"{name}/sd_keys"] = np.array(list(sd.keys()))
        print(f"  {name}: {[tuple(o.shape) for o in outs]}")
    np.savez_compressed(os.path.join(OUT, "models.npz"), **flat)
    print("models.npz:", len(names), "models")


def gen_v1():
    """The four PointNet-v1-like pose heads (model/width.py, rotation.py, translation.py, sign.py): seeded construction,
    randomised BN, eval forward at B=4, N=256.  Weights are not stored (they are reproduced by seeded construction and
    checked through a checksum); inputs, the pooled 1024-wide feature and the head outputs are."""
    flat = {}
    for i, name in enumerate(["width", "rotation", "translation", "sign"]):
        sys.modules.pop("pointnet_utils", None)
        mod = ref_module(name)
        torch.manual_seed(0)
        model = mod.get_model(num_category=7).eval()
        rand_bn_(model, 7)
        xyz = cloudgen.uniform_cloud(4, 256, 300 + i)
        pts = cloudgen.with_category(xyz)
        if name == "width":                          # width.py:12 normal_channel=True -> 6 + num_category channels
            nrm = torch.randn(4, 256, 3, generator=torch.Generator().manual_seed(400))
            pts = torch.cat([xyz, nrm, pts[..., 3:]], -1)
        x = cloudgen.channels_first(pts)
        mean = xyz[:, :3, :].mean(1)
        args = (x, mean) if name == "translation" else (x,)
        with torch.no_grad():
            out = model(*args)
            h = x
            for conv, bn in zip(model.conv, model.bn_conv):      # the point-wise stack + max the fused kernel replaces
                h = torch.relu(bn(conv(h)))
            pooled = h.max(2)[0]
        outs = out if isinstance(out, tuple) else (out,)
        flat[f"{name}/points"] = pts.numpy()
        flat[f"{name}/mean"] = mean.numpy()
        flat[f"{name}/pooled"] = pooled.numpy()
        for j, o in enumerate(outs):
            flat[f"{name}/out{j}"] = o.contiguous().numpy()
        sd = model.state_dict()
        flat[f"{name}/sd_checksum"] = np.array([float(sum(v.double().abs().sum() for v in sd.values()))])
        flat[f"{name}/sd_keys"] = np.array(list(sd.keys()))
        print(f"  v1/{name}: {[tuple(o.shape) for o in outs]}")
    np.savez_compressed(os.path.join(OUT, "v1.npz"), **flat)
    print("v1.npz: 4 heads")


def np_cloud(N, C, dtype, seed, dup=False):
    """Seeded numpy cloud for the numpy-FPS fixtures (regenerated by the tests for the large cases; sha-guarded)."""
    rng = np.random.default_rng(seed)
    pc = rng.standard_normal((N, C)).astype(dtype)
    if dup:
        pc[rng.random(N) < 0.5] = pc[0]                # like provider.random_point_dropout (:157-164)
    return pc


NPFPS_SPECS = [  # name, N, C, dtype, number, seed, dup, store input
    ("f64_small", 300, 3, "float64", 64, 1, False, True),
    ("f64_cols6", 500, 6, "float64", 128, 2, False, True),
    ("f64_dup", 400, 3, "float64", 450, 3, True, True),          # number > N and duplicates: ties, zero distances
    ("f32_small", 300, 3, "float32", 64, 4, False, True),
    ("f32_modelnet", 10000, 6, "float32", 1024, 5, False, False),  # ModelNetDataLoader.py:102-105 sizes
    ("f64_scratch", 7000, 3, "float64", 256, 6, False, False),     # beyond the on-chip limit of the CUDA kernel
    ("f32_scratch", 13000, 3, "float32", 128, 7, False, False),
]


def gen_npfps():
    """The numpy copies of FPS, run live: data_utils/ModelDataLoader.py:10-31 and point_collect/transformer.py:120-141
    (provider.py:183-204 is the same text but cannot be imported: open3d)."""
    import hashlib
    sys.path.insert(0, "/root/reference/data_utils"); sys.path.insert(0, "/root/reference/point_collect")
    try:
        import ModelDataLoader as mdl
        import transformer as tr
    finally:
        sys.path.remove("/root/reference/data_utils"); sys.path.remove("/root/reference/point_collect")
    assert mdl.__file__.startswith("/root/reference") and tr.__file__.startswith("/root/reference")
    flat = {}
    for name, N, C, dt, number, seed, dup, store in NPFPS_SPECS:
        pc = np_cloud(N, C, np.dtype(dt), seed, dup)
        np.random.seed(1000 + seed)
        ref = mdl.farthest_point_sample(pc.copy(), number)
        np.random.seed(1000 + seed)
        ref2 = tr.farthest_point_sample(pc.copy(), number)
        np.random.seed(1000 + seed)
        start = np.random.randint(0, N)
        assert np.array_equal(ref, ref2)
        got, idx = orc.farthest_point_sample_np(pc, number, start)
        assert np.array_equal(got, ref), name                    # the C restatement reproduces the live reference
        if store:
            flat[f"{name}/points"] = pc
        flat[f"{name}/sha"] = np.array([hashlib.sha256(pc.tobytes()).hexdigest()])
        flat[f"{name}/start"] = np.array([start])
        flat[f"{name}/rows_sha"] = np.array([hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest()])
        flat[f"{name}/idx"] = idx.astype(np.int32)
        print(f"  npfps/{name}: N={N} C={C} {dt} -> {ref.shape}")
    np.savez_compressed(os.path.join(OUT, "npfps.npz"), **flat)
    print("npfps.npz:", len(NPFPS_SPECS), "cases")


def main():
    os.makedirs(OUT, exist_ok=True)
    orc.build()
    ref = ref_module("pointnet2_utils")
    gen_indices(ref)
    gen_modules(ref)
    gen_models()
    gen_v1()
    gen_npfps()


if __name__ == "__main__":
    main()
