"""-m gpu: the CUDA kernels, called through the C ABI (include/pn2.h via _pn2_native), against the CPU oracle
and the golden vectors recorded from the live reference.  Indices are compared bit-exactly; features within
the stated tolerance (fp32 mode 1e-5, bf16 mode 2e-2, both as max|a-b| / max|b|)."""
import numpy as np
import pytest
import torch

import cloudgen
import golden_io

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 2e-2


@pytest.fixture(scope="module")
def nat():
    import _pn2_native
    return _pn2_native.Native()


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def dev(x):
    return x.to("cuda:0")


# ---- tcgen05 building blocks -------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 2, 3])
@pytest.mark.parametrize("n_cols,k_dim", [(32, 16), (64, 64), (128, 144), (256, 256), (128, 32), (128, 96)])
def test_umma_selftest(nat, n_cols, k_dim, variant):
    """Hand-built tcgen05 descriptors (no-swizzle operands; 128-byte-swizzled A; 128-byte-swizzled B) against a
    scalar loop on the device, integer-valued operands: exact."""
    if variant == 3 and n_cols != 128:
        pytest.skip("the swizzled-B form is used with N = 128 only")
    err, _ = nat.selftest_umma(n_cols, k_dim, variant)
    torch.cuda.synchronize()
    assert float(err.item()) == 0.0


@pytest.mark.parametrize("n_cols,k_dim", [(64, 16), (128, 64), (256, 128), (128, 256)])
def test_umma_pair_selftest(nat, n_cols, k_dim):
    """tcgen05.mma.cta_group::2 (M = 256 over a CTA pair): each CTA holds its 128 rows of A and half of B's columns at the same
    shared-memory offsets, CTA 0 issues and commits to both; every element of the 256 x N product against a scalar loop."""
    err, d = nat.selftest_umma(n_cols, k_dim, 4, dump=True)
    torch.cuda.synchronize()
    assert float(err.item()) == 0.0, (float(err.item()), d[:2, :8], d[128:130, :8])
    assert bool(d[128:].abs().sum() > 0)          # the second CTA wrote its half


# ---- golden vectors (reference outputs) ----------------------------------------------------------------------
@pytest.mark.parametrize("case", sorted(golden_io.index_cases()))
def test_golden_fps_ball_sqdist(nat, case):
    d = golden_io.index_cases()[case]
    pts = golden_io.as_layout(d["points"], d["layout"], "cuda:0")
    start = dev(torch.from_numpy(d["start"]))
    i64, i32, new_pts = nat.fps(pts, start, d["S"], want_idx64=True, want_idx32=True, want_new_points=True)
    ref = torch.from_numpy(d["fps"]).long()
    assert torch.equal(i64.cpu(), ref)
    assert torch.equal(i32.cpu().long(), ref)
    cpu_pts = golden_io.as_layout(d["points"], d["layout"])
    exp_new = torch.stack([cpu_pts[b][ref[b]] for b in range(ref.shape[0])])
    assert torch.equal(new_pts.cpu(), exp_new)
    import pointnet2_utils as pu
    b64, b32 = nat.ball_query(new_pts, pts, pu._radius_sq(d["r"]), d["K"], want_idx64=True, want_idx32=True)
    assert torch.equal(b64.cpu(), torch.from_numpy(d["ball"]).long())
    assert torch.equal(b32.cpu().long(), b64.cpu())
    sd = nat.square_distance(new_pts, pts)
    assert torch.equal(sd[:, :4].cpu(), torch.from_numpy(d["sqdist_sample"]))


# ---- oracle comparisons on seeded inputs ----------------------------------------------------------------------
FPS_CASES = [  # (B, N, C, S, layout, kind)
    (4, 1024, 3, 512, "strided", "uni"), (4, 512, 3, 128, "contig", "uni"), (3, 1024, 10, 512, "strided", "uni"),
    (3, 512, 10, 128, "contig", "surf"), (2, 2048, 10, 512, "strided", "surf"), (2, 2048, 3, 512, "contig", "dup"),
    (2, 4096, 3, 1024, "contig", "uni"), (2, 8192, 3, 2048, "contig", "uni"), (2, 8192, 3, 64, "strided", "surf"),
    (16, 8192, 3, 2048, "contig", "uni"), (3, 6000, 3, 700, "contig", "dup"), (5, 4128, 3, 300, "strided", "uni"),   # one cloud per cluster
    (48, 4096, 3, 96, "contig", "uni"), (64, 5000, 3, 64, "contig", "dup"),      # 4-CTA clusters, two CTAs per SM
    (2, 100, 3, 512, "strided", "uni"), (2, 1000, 3, 100, "contig", "dup"), (2, 777, 10, 99, "contig", "dup"),
    (2, 3000, 10, 64, "strided", "uni"), (2, 300, 5, 64, "contig", "uni"), (2, 300, 16, 64, "strided", "uni"),
    (1, 2, 3, 4, "contig", "uni"), (2, 33, 8, 40, "contig", "uni"),
    # many clouds: a one-ulp error in a distance only moves the argmax when the two farthest candidates are that close,
    # about once in ten clouds (this is what catches a contracted multiply-add, see tools/probes/f32x2_contract.cu)
    (256, 1024, 3, 512, "strided", "uni"), (256, 1024, 3, 512, "contig", "uni"), (96, 1024, 10, 512, "strided", "uni"),
    (96, 512, 10, 128, "contig", "uni"),
    # C = 10 without / with a broken constant tail: the general ten-channel loop, not the three-channel shortcut
    (8, 1024, 10, 512, "strided", "mix"), (8, 512, 10, 128, "contig", "mix"), (8, 1024, 10, 512, "strided", "odd1"),
    (8, 512, 10, 128, "contig", "odd1"),
    # >= 592 clouds: the issue-optimised launch shape (8 points per thread)
    (600, 1024, 3, 512, "strided", "uni"), (600, 512, 3, 128, "contig", "uni"), (600, 1000, 3, 100, "contig", "dup"),
]


def make_cloud(B, N, C, kind, seed):
    base = {"uni": cloudgen.uniform_cloud, "surf": cloudgen.surface_cloud}.get(kind, cloudgen.uniform_cloud)(B, N, seed)
    if kind == "dup":
        base = cloudgen.duplicate_heavy(base, 0.6, seed + 1)
    if C == 3:
        return base
    if C == 10 and kind == "mix":          # ten genuinely varying coordinates (no one-hot tail)
        return torch.cat([base, torch.rand(B, N, 7, generator=torch.Generator().manual_seed(seed + 3))], -1)
    if C == 10 and kind == "odd1":         # one-hot tail except for ONE point of every cloud: the uniform-tail test must fail
        x = cloudgen.with_category(base).clone()
        x[:, N // 2, 5] += 0.25
        return x
    if C == 10:
        return cloudgen.with_category(base)
    g = torch.Generator().manual_seed(seed + 2)
    extra = torch.rand(B, N, C - 3, generator=g) if C > 3 else None
    return torch.cat([base, extra], -1) if C > 3 else base[..., :C].contiguous()


def lay(x, layout):
    return x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "strided" else x.contiguous()


@pytest.mark.parametrize("B,N,C,S,layout,kind", FPS_CASES)
def test_fps_vs_oracle(nat, oracle, B, N, C, S, layout, kind):
    x = lay(make_cloud(B, N, C, kind, 1000 + N + C), layout)
    g = torch.Generator().manual_seed(7)
    start = torch.randint(0, N, (B,), generator=g, dtype=torch.long)
    exp = oracle.farthest_point_sample(x, S, start)
    i64, _, new_pts = nat.fps(dev(x), dev(start), S, want_idx64=True, want_new_points=True)
    assert torch.equal(i64.cpu(), exp)
    assert torch.equal(new_pts.cpu(), oracle.index_points(x, exp))


BQ_CASES = [  # (B, N, C, S, K, r, layout, kind)
    (4, 1024, 3, 512, 32, 0.2, "strided", "uni"), (4, 512, 3, 128, 64, 0.4, "contig", "uni"),
    (3, 1024, 10, 512, 16, 0.1, "strided", "surf"), (3, 1024, 10, 512, 128, 0.4, "strided", "uni"),
    (3, 512, 10, 128, 128, 0.8, "contig", "surf"), (2, 2048, 10, 512, 32, 0.2, "strided", "dup"),
    (2, 8192, 3, 2048, 32, 0.2, "contig", "uni"), (2, 8192, 3, 256, 32, 0.2, "strided", "surf"),
    (1, 20000, 3, 64, 48, 0.1, "contig", "uni"), (1, 6000, 10, 70, 20, 0.3, "contig", "uni"),
    (2, 300, 5, 64, 7, 0.5, "contig", "uni"), (2, 300, 16, 64, 300, 5.0, "strided", "uni"),
    (2, 64, 3, 64, 64, 10.0, "contig", "uni"), (2, 500, 3, 37, 1, 0.05, "contig", "dup"),
    # C = 10 whose tail is NOT one constant one-hot vector: the full ten-channel chain, not the three-channel shortcut
    (4, 1024, 10, 512, 32, 0.9, "strided", "mix"), (4, 512, 10, 128, 64, 1.2, "contig", "mix"),
    (4, 1024, 10, 512, 32, 0.2, "strided", "odd1"), (4, 512, 10, 128, 64, 0.4, "contig", "odd1"),
    (24, 1024, 10, 512, 128, 0.4, "strided", "uni"), (24, 512, 10, 128, 32, 0.2, "contig", "uni"),
]


@pytest.mark.parametrize("B,N,C,S,K,r,layout,kind", BQ_CASES)
def test_ball_query_vs_oracle(nat, oracle, B, N, C, S, K, r, layout, kind):
    x = lay(make_cloud(B, N, C, kind, 2000 + N + C + K), layout)
    g = torch.Generator().manual_seed(8)
    start = torch.randint(0, N, (B,), generator=g, dtype=torch.long)
    fps = oracle.farthest_point_sample(x, S, start)
    cen = oracle.index_points(x, fps).contiguous()
    exp = oracle.query_ball_point(r, K, x, cen)
    i64, i32 = nat.ball_query(dev(cen), dev(x), oracle.radius_sq_f32(r), K, want_idx64=True, want_idx32=True)
    assert torch.equal(i64.cpu(), exp)
    assert torch.equal(i32.cpu().long(), exp)
    sd = nat.square_distance(dev(cen[:, :8]), dev(x))
    assert torch.equal(sd.cpu(), oracle.square_distance(cen[:, :8], x))


BQ_MULTI_CASES = [  # (B, N, C, S, Ks, radii, layout, kind) -- the MSG layers (pointnet2_cls_msg.py:9-10, translation_msg.py:11) + edge cases
    (4, 1024, 3, 512, [16, 32, 128], [0.1, 0.2, 0.4], "strided", "uni"), (4, 512, 3, 128, [32, 64, 128], [0.2, 0.4, 0.8], "contig", "uni"),
    (4, 1024, 10, 512, [16, 32, 128], [0.1, 0.2, 0.4], "strided", "surf"), (4, 512, 10, 128, [32, 64, 128], [0.2, 0.4, 0.8], "contig", "uni"),
    (3, 1024, 10, 512, [16, 32, 64], [0.1, 0.2, 0.4], "strided", "dup"), (3, 2048, 10, 512, [16, 32, 128], [0.1, 0.2, 0.4], "strided", "uni"),
    (3, 1000, 10, 300, [8, 200], [0.3, 0.3], "contig", "mix"), (3, 777, 10, 99, [64, 5, 32], [0.2, 0.2, 0.9], "strided", "odd1"),
    (2, 333, 3, 70, [333, 1], [0.05, 50.0], "contig", "dup"),
    (2, 512, 3, 128, [16, 32, 64], [0.4, 0.2, 0.1], "contig", "uni"),        # descending radii: served by one query per radius
    (2, 300, 5, 64, [7, 9], [0.5, 0.6], "contig", "uni"),                    # other widths: generic kernel per radius
    (1, 6000, 3, 64, [16, 32], [0.05, 0.1], "contig", "uni"),                # large cloud
    (40, 1024, 10, 512, [16, 32, 128], [0.1, 0.2, 0.4], "strided", "uni"),
]


@pytest.mark.parametrize("B,N,C,S,Ks,radii,layout,kind", BQ_MULTI_CASES)
def test_multi_radius_ball_query_equals_one_query_per_radius(nat, oracle, B, N, C, S, Ks, radii, layout, kind):
    x = lay(make_cloud(B, N, C, kind, 3000 + N + C + sum(Ks)), layout)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(9), dtype=torch.long)
    cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
    r2 = [oracle.radius_sq_f32(r) for r in radii]
    outs = nat.ball_query_multi(dev(cen), dev(x), r2, Ks, want_idx64=True, want_idx32=True)
    assert len(outs) == len(Ks)
    for (i64, i32), r, K in zip(outs, radii, Ks):
        exp = oracle.query_ball_point(r, K, x, cen)
        assert torch.equal(i64.cpu(), exp), (r, K)
        assert torch.equal(i32.cpu().long(), exp)
    only32 = nat.ball_query_multi(dev(cen), dev(x), r2, Ks)
    assert all(a is None and torch.equal(b, o[1]) for (a, b), o in zip(only32, outs))


def test_ball_query_empty_rows_are_flagged_not_oob(nat, oracle):
    """No point in radius: the reference writes N and fails in index_points; we fill N and never read OOB."""
    x = cloudgen.uniform_cloud(1, 256, 5)
    cen = torch.full((1, 4, 3), 50.0)
    i64, _ = nat.ball_query(dev(cen), dev(x), oracle.radius_sq_f32(0.1), 8)
    assert (i64.cpu() == 256).all()


def test_index_points_and_group_vs_oracle(nat, oracle):
    x = make_cloud(3, 400, 10, "surf", 31)
    f = torch.randn(3, 400, 20, generator=torch.Generator().manual_seed(3))
    for layout in ("strided", "contig"):
        xs, fs = lay(x, layout), lay(f, layout)
        g = torch.Generator().manual_seed(9)
        idx = torch.randint(0, 400, (3, 50, 12), generator=g)
        assert torch.equal(nat.index_points(dev(xs), dev(idx)).cpu(), oracle.index_points(xs, idx))
        assert torch.equal(nat.index_points(dev(fs), dev(idx[:, :, 0])).cpu(), oracle.index_points(fs, idx[:, :, 0]))
        cen = oracle.index_points(xs, idx[:, :, 0]).contiguous()
        for order in (0, 1):
            for feat in (None, fs):
                exp = oracle.group(xs, feat, idx, cen, order)
                got = nat.group(dev(xs), None if feat is None else dev(feat), dev(idx.int()), dev(cen), order)
                assert torch.equal(got.cpu(), exp)
        exp = oracle.group(xs, fs, None, None, 2)
        assert torch.equal(nat.group(dev(xs), dev(fs), None, None, 2).cpu(), exp)
        # coordinates only (training mode's grouped rows): the 32-rows-per-warp kernel, C = 10 and C = 3, ragged row counts
        assert torch.equal(nat.group(dev(xs), None, None, None, 2).cpu(), oracle.group(xs, None, None, None, 2))
        x3 = lay(x[..., :3].contiguous(), layout)
        for order in (0, 1):
            for cut in (12, 7):
                exp = oracle.group(x3, None, idx[:, :, :cut].contiguous(), cen[..., :3].contiguous(), order)
                got = nat.group(dev(x3), None, dev(idx[:, :, :cut].int().contiguous()), dev(cen[..., :3].contiguous()), order)
                assert torch.equal(got.cpu(), exp)
    with pytest.raises(IndexError):
        nat.index_points(dev(x), dev(torch.full((3, 4), 400, dtype=torch.long)))


# ---- fused set abstraction --------------------------------------------------------------------------------------
def rand_layers(dims, seed):
    g = torch.Generator().manual_seed(seed)
    layers = []
    for ci, co in zip(dims[:-1], dims[1:]):
        W = (torch.rand(co, ci, generator=g) * 2 - 1) / ci ** 0.5
        b = (torch.rand(co, generator=g) * 2 - 1) / ci ** 0.5
        gam = torch.rand(co, generator=g) + 0.5
        bet = torch.randn(co, generator=g) * 0.2
        mu = torch.randn(co, generator=g) * 0.5
        var = torch.rand(co, generator=g) * 1.5 + 0.5
        layers.append((W, b, gam, bet, mu, var))
    return layers


SA_CASES = [  # (name, B, N, C, D, S, K, r, order, mlp)
    ("ssg_sa1", 3, 1024, 3, 0, 512, 32, 0.2, 0, [64, 64, 128]),
    ("ssg_sa2", 3, 512, 3, 128, 128, 64, 0.4, 0, [128, 128, 256]),
    ("ssg_sa3", 5, 128, 3, 256, 1, 128, None, 2, [256, 512, 1024]),
    ("pose_sa1", 2, 1024, 10, 0, 512, 32, 0.2, 0, [64, 64, 128]),
    ("pose_sa2", 2, 512, 10, 128, 128, 64, 0.4, 0, [128, 128, 256]),
    ("trans_sa2", 3, 512, 10, 128, 1, 512, None, 2, [256, 512, 1024]),
    ("msg_sa1_k16", 2, 1024, 3, 0, 512, 16, 0.1, 1, [32, 32, 64]),
    ("msg_sa1_k128", 2, 1024, 3, 0, 512, 128, 0.4, 1, [64, 96, 128]),
    ("msg_sa2_k32", 2, 512, 3, 320, 128, 32, 0.2, 1, [64, 64, 128]),
    ("msg_sa2_k128", 2, 512, 10, 320, 128, 128, 0.8, 1, [128, 128, 256]),
    ("msg_all", 2, 128, 3, 640, 1, 128, None, 2, [256, 512, 1024]),
    ("odd", 2, 300, 5, 7, 50, 12, 0.5, 0, [20, 33]),
    ("odd_msg", 2, 300, 3, 9, 41, 24, 0.5, 1, [17]),
    ("odd_all", 3, 77, 3, 5, 1, 77, None, 2, [24, 40]),
    ("bigk", 1, 700, 3, 8, 9, 300, 2.0, 0, [32, 64]),
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name,B,N,C,D,S,K,r,order,mlp", SA_CASES, ids=[c[0] for c in SA_CASES])
def test_sa_fused_vs_oracle(nat, oracle, prec, name, B, N, C, D, S, K, r, order, mlp):
    x = make_cloud(B, N, C, "uni", 3000 + N)
    f = torch.randn(B, N, D, generator=torch.Generator().manual_seed(5)) if D else None
    dims = [C + D] + mlp
    layers = rand_layers(dims, 11)
    if order == 2:
        idx = cen = None
        rows = oracle.group(x, f, None, None, 2)
    else:
        g = torch.Generator().manual_seed(8)
        start = torch.randint(0, N, (B,), generator=g, dtype=torch.long)
        fps = oracle.farthest_point_sample(x, S, start)
        cen = oracle.index_points(x, fps).contiguous()
        idx = oracle.query_ball_point(r, K, x, cen)
        rows = oracle.group(x, f, idx, cen, order)
    exp = oracle.mlp_max(rows.reshape(B * S, K, C + D), layers).reshape(B, S, -1)
    code = 0 if prec == "fp32" else 1
    packed, pdims = nat.pack([tuple(dev(t) for t in l) for l in layers], 1e-5, code, C, order)
    assert pdims == dims
    co_off, total = 3, mlp[-1] + 8
    alloc = torch.zeros if nat.sa_needs_zero(S, K, dims, code) else (lambda *a, **k: torch.full(*a, fill_value=-7.0, **k))
    out = alloc((B, S, total), dtype=torch.float32, device="cuda:0")
    before = out.clone()
    nat.sa_fused(dev(x), None if f is None else dev(f), None if cen is None else dev(cen),
                 None if idx is None else dev(idx.int()), order, dims, packed, code, out, co_off)
    torch.cuda.synchronize()
    got = out[:, :, co_off:co_off + mlp[-1]].cpu()
    err = rel_err(got, exp)
    assert err <= (TOL_F32 if prec == "fp32" else TOL_BF16), err
    # channels outside the branch's slice are untouched (MSG concat by offset, pointnet2_utils.py:222)
    assert torch.equal(out[:, :, :co_off], before[:, :, :co_off])
    assert torch.equal(out[:, :, co_off + mlp[-1]:], before[:, :, co_off + mlp[-1]:])


def test_cpu_tensor_is_an_error(nat):
    with pytest.raises(RuntimeError):
        nat.fps(torch.zeros(1, 8, 3), torch.zeros(1, dtype=torch.long), 2)


@pytest.mark.parametrize("K,order", [(64, 0), (32, 1), (16, 1), (128, 1), (8, 0)])
def test_sa_fused_bf16_side_tensors(nat, oracle, K, order):
    """bf16 twin of the features in (cp.async gather) and of the pooled output: same values as the fp32 route."""
    B, N, C, D, S, r, mlp = 2, 512, 3, 64, 96, 0.6, [64, 32, 128]
    x = make_cloud(B, N, C, "uni", 77)
    f = torch.randn(B, N, D, generator=torch.Generator().manual_seed(6))
    dims = [C + D] + mlp
    layers = rand_layers(dims, 12)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
    cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
    idx = oracle.query_ball_point(r, K, x, cen)
    assert nat.sa_bf16_side_io(K, C, order, dims, 1)
    packed, _ = nat.pack([tuple(dev(t) for t in l) for l in layers], 1e-5, 1, C, order)
    outs = []
    for twin in (False, True):
        out = torch.full((B, S, mlp[-1]), -7.0, device="cuda:0")
        ob = torch.zeros(B, S, mlp[-1], dtype=torch.bfloat16, device="cuda:0")
        nat.sa_fused(dev(x), dev(f), dev(cen), dev(idx.int()), order, dims, packed, 1, out, 0,
                     feature_bf16=dev(f).bfloat16().contiguous() if twin else None, out_bf16=ob)
        outs.append((out.cpu(), ob.cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[0][0].bfloat16())
    rows = oracle.group(x, f, idx, cen, order)
    exp = oracle.mlp_max(rows.reshape(B * S, K, C + D), layers).reshape(B, S, -1)
    assert rel_err(outs[0][0], exp) <= TOL_BF16


@pytest.mark.parametrize("D,K,order,mlp", [(128, 64, 0, [128, 128, 256]), (192, 32, 1, [64, 64, 128]), (256, 64, 0, [64, 64, 128]),
                                          (256, 16, 1, [32, 64]), (64, 128, 1, [128, 128, 256])])
def test_sa_fused_bf16_rows_only(nat, oracle, D, K, order, mlp):
    """Features given as the bf16 side tensor ALONE (no fp32 rows: the factorised MSG branches do this), several
    64-channel blocks per row (the producer warps' batches), one and several hidden layers, two- and four-unit stacks."""
    B, N, C, S, r = 3, 512, 3, 128, 0.7
    x = make_cloud(B, N, C, "uni", 91)
    f = torch.randn(B, N, D, generator=torch.Generator().manual_seed(16)).bfloat16()
    dims = [C + D] + mlp
    layers = rand_layers(dims, 17)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
    cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
    idx = oracle.query_ball_point(r, K, x, cen)
    assert nat.sa_bf16_side_io(K, C, order, dims, 1)
    packed, _ = nat.pack([tuple(dev(t) for t in l) for l in layers], 1e-5, 1, C, order)
    out = torch.full((B, S, mlp[-1]), -7.0, device="cuda:0")
    nat.sa_fused(dev(x), None, dev(cen), dev(idx.int()), order, dims, packed, 1, out, 0, feature_bf16=f.cuda().contiguous())
    rows = oracle.group(x, f.float(), idx, cen, order)
    exp = oracle.mlp_max(rows.reshape(B * S, K, C + D), layers).reshape(B, S, -1)
    assert rel_err(out.cpu(), exp) <= TOL_BF16


@pytest.mark.parametrize("name,B,N,C,D,S,K,r,order,mlp", [
    ("ssg_sa3", 5, 128, 3, 256, 1, 128, None, 2, [256, 512, 1024]), ("trans_sa2", 3, 512, 10, 128, 1, 512, None, 2, [256, 512, 1024]),
    ("msg_all", 2, 128, 3, 640, 1, 128, None, 2, [256, 512, 1024]), ("odd_pairs", 2, 300, 3, 24, 9, 300, 2.0, 0, [32, 64]),
    ("one_group", 2, 200, 3, 8, 20, 150, 1.0, 0, [48, 300])])
def test_sa_fused_streaming_kernel_takes_the_bf16_rows(nat, oracle, name, B, N, C, D, S, K, r, order, mlp):
    """The streaming (weight-ring) bf16 kernel -- wide group_all stacks, groups beyond 128 rows -- gathers the previous layer's
    bf16 rows with cp.async when they are given: bit-identical to converting the fp32 rows itself, also as the only feature
    tensor, also with an odd number of 16-byte segments per row (D = 24, 8) and with absent rows (K > ball size)."""
    x = make_cloud(B, N, C, "uni", 4100 + N)
    f = torch.randn(B, N, D, generator=torch.Generator().manual_seed(5))
    dims = [C + D] + mlp
    layers = rand_layers(dims, 13)
    if order == 2:
        idx = cen = None
        rows = oracle.group(x, f.bfloat16().float(), None, None, 2)
    else:
        start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
        cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
        idx = oracle.query_ball_point(r, K, x, cen)
        rows = oracle.group(x, f.bfloat16().float(), idx, cen, order)
    assert not nat.sa_bf16_side_io(K, C, order, dims, 1)           # not a stack of the pipelined kernels
    packed, _ = nat.pack([tuple(dev(t) for t in l) for l in layers], 1e-5, 1, C, order)
    twin = dev(f).bfloat16().contiguous()
    outs = []
    for feat, bf in ((dev(f), None), (dev(f), twin), (None, twin)):
        alloc = torch.zeros if nat.sa_needs_zero(S, K, dims, 1) else (lambda *a, **k: torch.full(*a, fill_value=-7.0, **k))
        out = alloc((B, S, mlp[-1]), dtype=torch.float32, device="cuda:0")
        nat.sa_fused(dev(x), feat, None if cen is None else dev(cen), None if idx is None else dev(idx.int()), order, dims, packed, 1,
                     out, 0, feature_bf16=bf)
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    exp = oracle.mlp_max(rows.reshape(B * S, K, C + D), layers).reshape(B, S, -1)
    assert rel_err(outs[0], exp) <= TOL_BF16


def test_tma_gather4_matches_the_swizzled_operand_layout(nat):
    """cp.async.bulk.tensor tile::gather4 + 128-byte swizzle: 4 arbitrary rows -> 4 consecutive swizzled smem rows
    (chunk j of row r at r*128 + ((j ^ (r & 7)) << 4)); rows outside the tensor read as zeros."""
    R, D, n_inst = 300, 128, 8
    g = torch.Generator().manual_seed(4)
    src = torch.randn(R, D, generator=g).bfloat16().cuda()
    rows = torch.randint(0, R, (n_inst * 4,), generator=g, dtype=torch.int32)
    rows[5] = R + 7                      # out of range -> zeros
    out = torch.zeros(n_inst * 512, dtype=torch.uint8, device="cuda:0")
    for col0 in (0, 64):
        rc = nat.lib.pn2_selftest_tma_gather(src.data_ptr(), R, D, rows.cuda().data_ptr(), n_inst, col0, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        assert rc == 0, nat.lib.pn2_last_error_string()
        torch.cuda.synchronize()
        img = out.cpu().view(n_inst * 4, 8, 16)                      # [smem row][chunk position][16 B]
        srcb = src.cpu().view(torch.uint8).view(R, D * 2)
        for r in range(n_inst * 4):
            want_row = srcb[rows[r], col0 * 2: col0 * 2 + 128].view(8, 16) if rows[r] < R else torch.zeros(8, 16, dtype=torch.uint8)
            for j in range(8):
                assert torch.equal(img[r, j ^ (r & 7)], want_row[j]), (col0, r, j)


@pytest.mark.parametrize("env", [{"PN2_TC4": "0"}, {"PN2_TC4": "1", "PN2_TC4_STATIC": "0"}, {"PN2_TC4": "0", "PN2_TC_COOP": "0"},
                                 {"PN2_FACTORISE": "0"}])
def test_alternative_schedules_keep_parity(env):
    """The kernel-selection switches are read once per process: run the bf16 parity cases in a child process with the
    cooperative TMA kernel (PN2_TC4=0), the dynamically polled warp-specialised schedule, the unit kernel everywhere, and
    the unfactorised MSG branches."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_kernels.py"), os.path.join(root, "tests", "test_gpu_models.py"),
                        "-m", "gpu", "-q", "-x", "-k", "bf16 and not alternative"], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


@pytest.mark.parametrize("variant", ["0", "1", "2", "4"])
def test_fps_cluster_variants_keep_parity(variant):
    """PN2_FPS_CLUSTER (read once per process) selects how few large clouds are sampled: 0 = one CTA per cloud, 1 = 8-CTA
    cluster with a block reduction and per-CTA slots, 2 / 4 = 8- / 4-CTA cluster with per-warp slots (the default picks between
    the last two by the number of clouds).  Every variant must return the oracle's indices."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, PN2_FPS_CLUSTER=variant)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_kernels.py"), "-m", "gpu", "-q", "-x",
                        "-k", "test_fps_vs_oracle and (4096 or 8192 or 6000 or 4128 or 5000)"], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


# ---- fused FC head ---------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 24, 70, 2048])
@pytest.mark.parametrize("kind,out_w,mean_mlp", [("cls", 7, False), ("reg", 3, False), ("trans", 3, True), ("trans", 3, False), ("sign", 1, False)])
def test_fc_head_vs_torch_fp32(nat, B, kind, out_w, mean_mlp):
    """pn2_fc_head_f32 (one launch: fc1/bn1/relu -> fc2/bn2/relu -> fc3 -> log_softmax+argmax | +mean(MLP) | sigmoid+sign)
    against the reference's op sequence in plain PyTorch fp32 on the CPU (pointnet2_cls_ssg.py:31-38, translation_ssg.py:31-33,
    39-44, sign_ssg.py:29-36), eval mode, non-trivial BatchNorm statistics."""
    import torch.nn as nn
    import torch.nn.functional as F
    import pointnet2_utils as pu
    import pn2_models
    from _pn2_native import HEAD_CLS, HEAD_PLAIN, HEAD_SIGN, HEAD_TRANS
    torch.manual_seed(3)
    fc1, fc2, fc3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, out_w)
    bn1, bn2 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)
    mfc1, mbn, mfc2 = nn.Linear(3, 6), nn.BatchNorm1d(6), nn.Linear(6, 3)
    holder = nn.ModuleList([fc1, fc2, fc3, bn1, bn2, mfc1, mbn, mfc2]).eval()
    pn2_models.randomize_bn_(holder, 5)
    g = torch.Generator().manual_seed(B)
    x = torch.relu(torch.randn(B, 1024, generator=g))                       # pooled features are post-ReLU maxima
    mean = torch.randn(B, 3, generator=g)
    with torch.no_grad():
        h = F.relu(bn1(fc1(x)))
        h = F.relu(bn2(fc2(h)))
        y = fc3(h)
        if kind == "cls":
            exp, exp_aux = F.log_softmax(y, -1), None
            exp_aux = exp.max(1)[1]
        elif kind == "trans":
            m = mfc2(F.relu(mbn(mfc1(mean)))) if mean_mlp else mean
            exp, exp_aux = y + m, None
        elif kind == "sign":
            exp = torch.sigmoid(y)
            exp_aux = torch.sign(exp - 0.5)
        else:
            exp, exp_aux = y, None
    holder.to("cuda:0")
    cache = pu._FoldedHead()
    code = {"cls": HEAD_CLS, "reg": HEAD_PLAIN, "trans": HEAD_TRANS, "sign": HEAD_SIGN}[kind]
    # x as the models hand it over: the [B,1024] view of the pooled [B,1,1024] buffer
    xd = x.to("cuda:0").view(B, 1, 1024).permute(0, 2, 1).view(B, 1024)
    out, aux = pu.fc_head(xd, (fc1, fc2, fc3), (bn1, bn2), code, cache, mean=mean.to("cuda:0") if kind == "trans" else None,
                          mean_mlp=(mfc1, mbn, mfc2) if (kind == "trans" and mean_mlp) else None)
    assert out.shape == exp.shape
    assert rel_err(out.cpu(), exp) <= TOL_F32
    if kind == "cls":
        # argmax: equal wherever the two largest log-probabilities are not within rounding of each other
        top2 = exp.topk(2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4
        assert aux.dtype == torch.int64 and torch.equal(aux.cpu()[clear], exp_aux[clear])
        assert torch.equal(aux.cpu(), out.cpu().max(1)[1])
    if kind == "sign":
        clear = (exp - 0.5).abs() > 1e-5
        assert torch.equal(aux.cpu()[clear], exp_aux[clear])
    # a weight update is followed (derived cache keyed on the parameter versions)
    with torch.no_grad():
        fc3.bias.add_(1.0)
    out2, _ = pu.fc_head(xd, (fc1, fc2, fc3), (bn1, bn2), code, cache, mean=mean.to("cuda:0") if kind == "trans" else None,
                         mean_mlp=(mfc1, mbn, mfc2) if (kind == "trans" and mean_mlp) else None)
    if kind in ("reg", "trans"):
        assert torch.allclose(out2, out + 1.0, atol=1e-5)


def test_fc_head_propagates_nan_and_rejects_bad_widths(nat):
    import pointnet2_utils as pu
    x = torch.zeros(3, 1024, device="cuda:0")
    x[1, 5] = float("nan")
    w1p = torch.randn(256, 512, 4, device="cuda:0")
    w2p = torch.randn(128, 256, 4, device="cuda:0")
    b1, b2 = torch.zeros(512, device="cuda:0"), torch.zeros(256, device="cuda:0")
    w3, b3 = torch.randn(3, 256, device="cuda:0"), torch.zeros(3, device="cuda:0")
    out, _ = nat.fc_head(x, w1p, b1, w2p, b2, w3, b3, 0)
    assert torch.isnan(out[1]).all() and torch.isfinite(out[0]).all() and torch.isfinite(out[2]).all()
    with pytest.raises(RuntimeError):
        nat.fc_head(x, w1p, b1, w2p, torch.zeros(100, device="cuda:0"), torch.randn(3, 100, device="cuda:0"), b3, 0)


# ---- row GEMM on the tensor cores (MSG factorised first layer, training convolutions) ------------------------------------------
@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (1000, 320, 128), (5000, 320, 64), (70001, 128, 256), (4096, 512, 1024), (333, 16, 96),
                                   (20000, 332, 40), (257, 1024, 512)])
@pytest.mark.parametrize("mode", ["bf16", "bf16_from_f32", "tf32"])
def test_rows_gemm_vs_torch(nat, M, K, N, mode):
    """pn2_rows_gemm against fp64 products of the SAME rounded operands (bf16 modes) / of the fp32 operands (TF32 mode)."""
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    if mode == "tf32":
        if K % 4:
            pytest.skip("TF32 rows are multiples of 16 bytes")
        got = nat.rows_gemm(a.cuda(), w.cuda())
        assert got.dtype == torch.float32
        exp = a.double() @ w.double().t()
        assert rel_err(got.cpu(), exp) <= 2e-3                      # 10-bit mantissas
        # strided rows (a column slice of a wider buffer) and a preallocated strided output
        wide = torch.zeros(M, K + 8).cuda()
        wide[:, 4:4 + K] = a.cuda()
        out = torch.full((M, N + 4), -3.0).cuda()
        nat.rows_gemm(wide[:, 4:4 + K], w.cuda(), out=out[:, :N])
        assert torch.equal(out[:, :N], got) and bool((out[:, N:] == -3.0).all())
        return
    if K % 8:
        pytest.skip("bf16 rows are multiples of 16 bytes")
    wb = w.bfloat16()
    ab = a.bfloat16()
    got = nat.rows_gemm(ab.cuda() if mode == "bf16" else a.cuda(), wb.cuda(), torch.bfloat16)
    assert got.dtype == torch.bfloat16
    exp = (ab.double() @ wb.double().t())
    assert rel_err(got.float().cpu(), exp) <= 2.0 ** -8             # one bf16 rounding of the fp32-accumulated result
    got32 = nat.rows_gemm(ab.cuda(), wb.cuda(), torch.float32)
    assert rel_err(got32.cpu(), exp) <= 1e-5


@pytest.mark.parametrize("M,Cout,Cin", [(4096, 64, 32), (100000, 96, 64), (33, 32, 12), (70001, 128, 128), (5000, 256, 128), (3000, 1024, 512),
                                        (20000, 128, 332), (257, 64, 16), (1, 32, 4)])
def test_rows_wgrad_vs_torch(nat, M, Cout, Cin):
    """pn2_rows_wgrad_tf32 (dy^T x, MN-major tcgen05 operands straight from TMA) against the fp64 product; run twice: bitwise
    repeatable (partials are added in a fixed order)."""
    g = torch.Generator().manual_seed(M + Cout + Cin)
    dy = torch.randn(M, Cout, generator=g)
    x = torch.randn(M, Cin, generator=g)
    exp = dy.double().t() @ x.double()
    got = nat.rows_wgrad(dy.cuda(), x.cuda())
    assert got.shape == (Cout, Cin)
    scale = float(exp.abs().max()) + M ** 0.5
    assert float((got.cpu().double() - exp).abs().max()) <= 2e-3 * scale
    assert torch.equal(got, nat.rows_wgrad(dy.cuda(), x.cuda()))


@pytest.mark.parametrize("B,N,D,S,K", [(3, 512, 128, 64, 16), (2, 300, 7, 41, 5), (4, 1024, 320, 128, 32), (1, 64, 64, 64, 64)])
def test_gather_and_scatter_add_rows_vs_torch(nat, B, N, D, S, K):
    """Training mode's feature gather (index_points(feature, idx), :112 / :208) and its backward against torch.gather /
    scatter_add_ in fp32 on the CPU; the empty-ball index N gathers zeros and scatters nothing."""
    g = torch.Generator().manual_seed(B + N + D)
    feat = torch.randn(B, N, D, generator=g)
    idx = torch.randint(0, N, (B, S, K), generator=g, dtype=torch.int32)
    idx[0, 0, :] = N                                             # an empty ball
    ok = (idx < N).unsqueeze(-1)
    flat = idx.reshape(B, -1, 1).long().clamp(0, N - 1).expand(-1, -1, D)
    exp = torch.gather(feat, 1, flat).view(B, S, K, D) * ok
    got = nat.gather_rows(feat.cuda(), idx.cuda())
    assert torch.equal(got.cpu(), exp)
    # strided feature rows (a channel slice of a wider tensor)
    wide = torch.randn(B, N, D + 8, generator=g)
    got2 = nat.gather_rows(wide.cuda()[:, :, 4:4 + D], idx.cuda())
    assert torch.equal(got2.cpu(), torch.gather(wide[:, :, 4:4 + D], 1, flat).view(B, S, K, D) * ok)
    grad = torch.randn(B, S, K, D, generator=g)
    exp_g = torch.zeros(B, N, D).scatter_add_(1, flat, (grad * ok).reshape(B, -1, D))
    got_g = nat.scatter_add_rows(grad.cuda(), idx.cuda(), N)
    assert rel_err(got_g.cpu(), exp_g) <= 1e-5                   # fp32 reductions in a different order


@pytest.mark.parametrize("env", [{"PN2_F32_FFMA": "1"}, {"PN2_F32_PASSES": "3"}, {"PN2_F32_CLUSTER": "1"}, {"PN2_F32_CLUSTER": "4"}, {"PN2_F32_UNITS": "0"}, {"PN2_F32_U2": "1"}])
def test_fp32_mode_variants_keep_parity(env):
    """fp32 mode runs on the tensor cores with split (bf16 hi + bf16 lo) operands, four passes per K step (pn2_sa_tcs.cu).  The
    process-wide switches select the FFMA kernel instead (PN2_F32_FFMA), three passes, and the cluster size of the weight
    multicast; every variant must stay within the fp32 tolerance on the fp32 parity cases."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_kernels.py"), os.path.join(root, "tests", "test_gpu_models.py"),
                        "-m", "gpu", "-q", "-x", "-k", "fp32 and not variants"], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
