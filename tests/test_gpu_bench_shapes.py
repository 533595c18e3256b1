"""-m gpu: oracle parity at the shapes that bench.py times ("next round" item 1).

  * the bench workload itself: pointnet2_cls_ssg, B = 2048, N = 1024 -- FPS and ball-query indices of BOTH sampling
    levels bit-exact against oracle/pn2_oracle.c over the whole batch (this is the `cpw = 16` ball-query launch shape and
    the 8-points-per-thread FPS shape), fused features of every level and the model output on clouds taken from the
    start, the middle and the end of the batch (32-bit row indexing) in both precision modes;
  * configs[0] (B = 24) and configs[1] (pointnet2_cls_msg, B = 32): indices against the C oracle, features against the
    op-for-op CPU port of the reference path;
  * configs[2]: FPS + ball query at B = 16, N = 8192 -> 2048, r = 0.2, K = 32;
  * every ball-query launch shape (`PN2_BQ_CPW` = 4 / 8 / 16 forced in child processes, and batches that select 8 and
    16 on their own);
  * the public sample_and_group / sample_and_group_all on CUDA tensors (pointnet2_utils.py:92-141).

Indices are compared bit-exactly; features as max|a-b| / max|b| <= 1e-5 (fp32 mode) / 2e-2 (bf16 mode).
"""
import os
import subprocess
import sys

import pytest
import torch

import cloudgen

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 2e-2}
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture()
def pu():
    import pointnet2_utils
    prev = pointnet2_utils.get_precision()
    yield pointnet2_utils
    pointnet2_utils.set_precision(prev)
    pointnet2_utils.set_reference_strides(False)


def build(name):
    import pn2_models
    torch.manual_seed(0)
    m = pn2_models.get_model(name, 7)
    pn2_models.randomize_bn_(m)
    return m.eval()


class FixedStarts:
    """fps_start_source handing out pre-drawn start vectors in call order (one per sampling layer)."""

    def __init__(self, starts):
        self.starts, self.i = starts, 0

    def __call__(self, B, N):
        s = self.starts[self.i % len(self.starts)]
        self.i += 1
        assert s.shape[0] == B and int(s.max()) < N
        return s


def oracle_ssg_chain(oracle, model, x_bcn, starts):
    """The three set-abstraction levels of pointnet2_cls_ssg on the C oracle (CPU): per-level outputs and indices."""
    out, pts, feat = [], x_bcn, None
    for i, sa in enumerate((model.sa1, model.sa2, model.sa3)):
        layers = oracle.layers_of(sa.mlp_convs, sa.mlp_bns)
        new_pts, new_feat, fps_idx, gidx = oracle.set_abstraction(
            pts, feat, sa.point_number, sa.sample_number, sa.radius, layers, sa.group_all, None if sa.group_all else starts[i])
        out.append((new_pts, new_feat, fps_idx, gidx))
        pts, feat = new_pts, new_feat
    return out


# ---- the bench workload: cls_ssg, B = 2048, N = 1024 ------------------------------------------------------------------
@pytest.fixture(scope="module")
def bench_batch():
    B, N = 2048, 1024
    x = cloudgen.channels_first(cloudgen.uniform_cloud(B, N, 1001))          # bench.py's rank-0 step-0 input
    g = torch.Generator().manual_seed(42)
    starts = [torch.randint(0, n, (B,), generator=g, dtype=torch.long) for n in (1024, 512)]
    return x, starts


def test_bench_workload_indices_bit_exact_over_the_whole_batch(pu, oracle, bench_batch):
    """Both sampling levels of the B = 2048 step: FPS indices, gathered centres and ball-query indices equal the oracle's
    for every one of the 2048 clouds (ball query runs its 16-centres-per-warp launch shape here)."""
    x, starts = bench_batch
    be = pu._backend()
    pts = x.permute(0, 2, 1)                                                  # strided [B,N,3]: sa1's layout
    exp_fps = oracle.farthest_point_sample(pts, 512, starts[0])
    i64, i32, cen = be.fps(pts.to(DEV), starts[0].to(DEV), 512, want_idx64=True, want_idx32=True, want_new_points=True)
    assert torch.equal(i64.cpu(), exp_fps)
    assert torch.equal(i32.cpu().long(), exp_fps)
    exp_cen = oracle.index_points(pts, exp_fps).contiguous()
    assert torch.equal(cen.cpu(), exp_cen)
    exp_ball = oracle.query_ball_point(0.2, 32, pts, exp_cen)
    b64, b32 = be.ball_query(cen, pts.to(DEV), pu._radius_sq(0.2), 32, want_idx64=True, want_idx32=True)
    assert torch.equal(b64.cpu(), exp_ball)
    assert torch.equal(b32.cpu().long(), exp_ball)
    # level 2: contiguous [B,512,3] centres of level 1 are the cloud
    exp_fps2 = oracle.farthest_point_sample(exp_cen, 128, starts[1])
    j64, _, cen2 = be.fps(cen, starts[1].to(DEV), 128, want_idx64=True, want_new_points=True)
    assert torch.equal(j64.cpu(), exp_fps2)
    exp_cen2 = oracle.index_points(exp_cen, exp_fps2).contiguous()
    assert torch.equal(cen2.cpu(), exp_cen2)
    exp_ball2 = oracle.query_ball_point(0.4, 64, exp_cen, exp_cen2)
    c64, _ = be.ball_query(cen2, cen, pu._radius_sq(0.4), 64, want_idx64=True)
    assert torch.equal(c64.cpu(), exp_ball2)


PROBE_CLOUDS = [0, 1, 2, 1023, 1024, 2045, 2046, 2047]          # first / middle / last rows of the 2048-cloud launch


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_bench_workload_features_vs_oracle(pu, oracle, bench_batch, prec):
    """The whole B = 2048 forward; the pooled features of every level and the log-probabilities of clouds from the start,
    the middle and the end of the batch against the oracle chain run on exactly those clouds."""
    x, starts = bench_batch
    pu.set_precision(prec)
    model = build("pointnet2_cls_ssg")
    sel = torch.tensor(PROBE_CLOUDS)
    exp = oracle_ssg_chain(oracle, model, x[sel], [s[sel] for s in starts])
    with torch.no_grad():
        head_in = exp[2][1].reshape(len(sel), 1024)
        h = torch.relu(model.bn1(model.fc1(head_in)))
        h = torch.relu(model.bn2(model.fc2(h)))
        exp_logp = torch.log_softmax(model.fc3(h), -1)
    model = model.to(DEV)
    xd = x.to(DEV)
    got = []
    with torch.no_grad(), pu.fps_start_source(FixedStarts([s for s in starts])):
        p1, f1 = model.sa1(xd, None)
        p2, f2 = model.sa2(p1, f1)
        p3, f3 = model.sa3(p2, f2)
        got = [(p1, f1), (p2, f2), (p3, f3)]
    with torch.no_grad(), pu.fps_start_source(FixedStarts([s for s in starts])):
        logp, l3, _ = model(xd)
    tol = TOL[prec]
    for lvl, ((gp, gf), (ep, ef, _, _)) in enumerate(zip(got, exp)):
        assert torch.equal(gp[sel.to(DEV)].cpu(), ep), f"level {lvl}: centres differ"
        err = rel_err(gf[sel.to(DEV)].cpu(), ef)
        assert err <= tol * (1 if lvl == 0 else 2), (lvl, err)          # bf16 mode re-rounds the features between levels
    assert torch.equal(l3, got[2][1])
    assert rel_err(logp[sel.to(DEV)].cpu(), exp_logp) <= 5 * tol


# ---- configs[0] and configs[1]: whole batches against the CPU port, indices against the C oracle ----------------------
def port_forward(name, x, starts):
    import pn2_models
    import pn2_torch_port as port
    torch.manual_seed(0)
    m = pn2_models.get_model(name, 7, sa_impl=port)
    pn2_models.randomize_bn_(m)
    m.eval()
    port.set_start_source(FixedStarts(starts))
    try:
        with torch.no_grad():
            return m(x)
    finally:
        port.set_start_source(None)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name,B,seed", [("pointnet2_cls_ssg", 24, 1001), ("pointnet2_cls_msg", 32, 1002)])
def test_config_batches_vs_cpu_port(pu, oracle, name, B, seed, prec):
    """configs[0] (cls_ssg, B = 24) and configs[1] (cls_msg, B = 32, three-radius grouping) at their literal batch sizes:
    level-1 and level-2 indices against the C oracle, l3 features and log-probabilities against the CPU port."""
    x = cloudgen.channels_first(cloudgen.uniform_cloud(B, 1024, seed))
    g = torch.Generator().manual_seed(42)
    starts = [torch.randint(0, n, (B,), generator=g, dtype=torch.long) for n in (1024, 512)]
    exp_logp, exp_l3, exp_pred = port_forward(name, x, starts)
    pu.set_precision(prec)
    model = build(name).to(DEV)
    be = pu._backend()
    with torch.no_grad(), pu.fps_start_source(FixedStarts(starts)):
        logp, l3, pred = model(x.to(DEV))
    tol = TOL[prec]
    assert rel_err(l3.cpu(), exp_l3) <= 2 * tol
    assert rel_err(logp.cpu(), exp_logp) <= 5 * tol
    if prec == "fp32":
        assert torch.equal(pred.cpu(), exp_pred)
    # indices of both sampling levels, every radius of the layer
    pts = x.permute(0, 2, 1)
    if name.endswith("ssg"):
        radii = [([0.2], [32]), ([0.4], [64])]
    else:
        radii = [([0.1, 0.2, 0.4], [16, 32, 128]), ([0.2, 0.4, 0.8], [32, 64, 128])]
    cloud_cpu, cloud_dev = pts, pts.to(DEV)
    for lvl, (S, (rs, Ks)) in enumerate(zip((512, 128), radii)):
        exp_fps = oracle.farthest_point_sample(cloud_cpu, S, starts[lvl])
        i64, _, cen = be.fps(cloud_dev, starts[lvl].to(DEV), S, want_idx64=True, want_new_points=True)
        assert torch.equal(i64.cpu(), exp_fps)
        exp_cen = oracle.index_points(cloud_cpu, exp_fps).contiguous()
        assert torch.equal(cen.cpu(), exp_cen)
        outs = be.ball_query_multi(cen, cloud_dev, [pu._radius_sq(r) for r in rs], Ks, want_idx64=True, want_idx32=True)
        for (b64, b32), r, K in zip(outs, rs, Ks):
            e = oracle.query_ball_point(r, K, cloud_cpu, exp_cen)
            assert torch.equal(b64.cpu(), e), (lvl, r, K)
            assert torch.equal(b32.cpu().long(), e)
        cloud_cpu, cloud_dev = exp_cen, cen


# ---- configs[2]: FPS + ball query microbenchmark shape ----------------------------------------------------------------
def test_config2_fps_and_ball_query_at_b16_n8192(pu, oracle):
    x = cloudgen.uniform_cloud(16, 8192, 1003)
    start = torch.randint(0, 8192, (16,), generator=torch.Generator().manual_seed(42), dtype=torch.long)
    be = pu._backend()
    exp_fps = oracle.farthest_point_sample(x, 2048, start)
    i64, _, cen = be.fps(x.to(DEV), start.to(DEV), 2048, want_idx64=True, want_new_points=True)
    assert torch.equal(i64.cpu(), exp_fps)
    exp_cen = oracle.index_points(x, exp_fps).contiguous()
    assert torch.equal(cen.cpu(), exp_cen)
    exp = oracle.query_ball_point(0.2, 32, x, exp_cen)
    b64, b32 = be.ball_query(cen, x.to(DEV), pu._radius_sq(0.2), 32, want_idx64=True, want_idx32=True)
    assert torch.equal(b64.cpu(), exp)
    assert torch.equal(b32.cpu().long(), exp)
    # through the public functions, seeded like the reference (:59 draws on the CPU default generator)
    torch.manual_seed(42)
    want_start = torch.randint(0, 8192, (16,), dtype=torch.long)
    torch.manual_seed(42)
    f = pu.farthest_point_sample(x.to(DEV), 2048)
    assert torch.equal(f.cpu(), oracle.farthest_point_sample(x, 2048, want_start))
    q = pu.query_ball_point(0.2, 32, x.to(DEV), pu.index_points(x.to(DEV), f))
    assert q.dtype == torch.int64 and torch.equal(q.cpu(), oracle.query_ball_point(0.2, 32, x, oracle.index_points(x, f.cpu()).contiguous()))


# ---- every ball-query launch shape -----------------------------------------------------------------------------------
BQ_BATCH_CASES = [  # (B, N, C, S, K, r, layout): B x ceil(S / (8 warps x cpw)) >= 592 picks cpw = 8 from 74 clouds, 16 from 148
    (80, 1024, 3, 512, 32, 0.2, "strided"), (160, 1024, 3, 512, 32, 0.2, "strided"), (640, 512, 3, 128, 64, 0.4, "contig"),
    (80, 1024, 10, 512, 32, 0.2, "strided"), (160, 1024, 10, 512, 16, 0.1, "strided"), (640, 512, 10, 128, 64, 0.4, "contig"),
    (300, 1000, 3, 500, 24, 0.15, "contig"), (150, 777, 10, 333, 48, 0.35, "strided"),
]


@pytest.mark.parametrize("B,N,C,S,K,r,layout", BQ_BATCH_CASES)
def test_ball_query_batched_launch_shapes_vs_oracle(pu, oracle, B, N, C, S, K, r, layout):
    """Batches large enough that the launcher picks 8 and 16 centres per warp on its own (the shapes every benchmark run
    uses), ragged S and N included."""
    base = cloudgen.uniform_cloud(B, N, 5000 + B + C)
    x = cloudgen.with_category(base) if C == 10 else base
    x = x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "strided" else x.contiguous()
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
    fps = oracle.farthest_point_sample(x, S, start)
    cen = oracle.index_points(x, fps).contiguous()
    exp = oracle.query_ball_point(r, K, x, cen)
    be = pu._backend()
    i64, i32 = be.ball_query(cen.to(DEV), x.to(DEV), oracle.radius_sq_f32(r), K, want_idx64=True, want_idx32=True)
    assert torch.equal(i64.cpu(), exp)
    assert torch.equal(i32.cpu().long(), exp)
    multi = be.ball_query_multi(cen.to(DEV), x.to(DEV), [oracle.radius_sq_f32(r * 0.5), oracle.radius_sq_f32(r)], [max(1, K // 2), K],
                                want_idx64=True)
    assert torch.equal(multi[1][0].cpu(), exp)
    assert torch.equal(multi[0][0].cpu(), oracle.query_ball_point(r * 0.5, max(1, K // 2), x, cen))


@pytest.mark.parametrize("cpw", ["4", "8", "16"])
def test_ball_query_centres_per_warp_variants_keep_parity(cpw):
    """PN2_BQ_CPW (read once per process) forces the number of centres a warp walks per staged cloud: every value the
    launcher can pick must return the oracle's rows on every ball-query case of the suite."""
    e = dict(os.environ, PN2_BQ_CPW=cpw)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"),
                        os.path.join(ROOT, "tests", "test_gpu_bench_shapes.py"), "-m", "gpu", "-q", "-x",
                        "-k", "test_ball_query_vs_oracle or test_golden_fps_ball_sqdist or test_ball_query_batched_launch_shapes"],
                       env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


# ---- the public grouping functions on CUDA tensors (a5 / a6) -----------------------------------------------------------
@pytest.mark.parametrize("B,N,C,D,S,K,r,layout", [(24, 1024, 3, 0, 512, 32, 0.2, "strided"), (24, 512, 3, 128, 128, 64, 0.4, "contig"),
                                                  (6, 1024, 10, 0, 512, 32, 0.2, "strided"), (6, 512, 10, 128, 128, 64, 0.4, "contig"),
                                                  (3, 300, 5, 7, 50, 12, 0.5, "contig")])
def test_public_sample_and_group_on_cuda(pu, oracle, B, N, C, D, S, K, r, layout):
    """sample_and_group(points, feature, S, K, r, returnfps) called as a user would (CUDA tensors, the FPS start drawn by the
    function itself from the seeded CPU generator, pointnet2_utils.py:59): all four results bit-equal to the oracle's."""
    base = cloudgen.uniform_cloud(B, N, 600 + N + C)
    if C == 10:
        x = cloudgen.with_category(base)
    elif C == 3:
        x = base
    else:
        x = torch.cat([base, torch.rand(B, N, C - 3, generator=torch.Generator().manual_seed(2))], -1)
    x = x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "strided" else x.contiguous()
    f = torch.randn(B, N, D, generator=torch.Generator().manual_seed(3)) if D else None
    torch.manual_seed(123)
    start = torch.randint(0, N, (B,), dtype=torch.long)
    fps = oracle.farthest_point_sample(x, S, start)
    cen = oracle.index_points(x, fps).contiguous()
    idx = oracle.query_ball_point(r, K, x, cen)
    exp_rows = oracle.group(x, f, idx, cen, 0)
    exp_grouped = oracle.index_points(f if f is not None else x, idx)
    torch.manual_seed(123)
    new_points, new_feature, grouped, fps_idx = pu.sample_and_group(x.to(DEV), None if f is None else f.to(DEV), S, K, r, returnfps=True)
    assert fps_idx.dtype == torch.int64 and torch.equal(fps_idx.cpu(), fps)
    assert torch.equal(new_points.cpu(), cen)
    assert new_feature.shape == (B, S, K, C + D) and torch.equal(new_feature.cpu(), exp_rows)
    assert torch.equal(grouped.cpu(), exp_grouped)
    torch.manual_seed(123)
    two = pu.sample_and_group(x.to(DEV), None if f is None else f.to(DEV), S, K, r)
    assert len(two) == 2 and torch.equal(two[1].cpu(), exp_rows)
    # group_all (:122-141): zero centre, raw coordinates first
    z, rows = pu.sample_and_group_all(x.to(DEV), None if f is None else f.to(DEV))
    assert z.shape == (B, 1, C) and not z.any()
    assert rows.shape == (B, 1, N, C + D) and torch.equal(rows.cpu(), oracle.group(x, f, None, None, 2))


@pytest.mark.parametrize("cells", ["0", "1"])
def test_ball_query_grid_cull_variants_keep_parity(cells):
    """PN2_BQ_CELLS (read once per process): 1 forces the exact spatial cull (grid of cells >= the radius, 27-cell
    neighbourhood, hit bit mask) for every radius and cloud -- non-normalised, duplicate-heavy, C = 10, few cells included --
    0 forces the scan kernel; both must return the oracle's rows on every ball-query case of the suite."""
    e = dict(os.environ, PN2_BQ_CELLS=cells)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"),
                        os.path.join(ROOT, "tests", "test_gpu_bench_shapes.py"), "-m", "gpu", "-q", "-x",
                        "-k", "test_ball_query_vs_oracle or test_golden_fps_ball_sqdist or test_ball_query_batched_launch_shapes or "
                              "test_config2 or test_ball_query_empty or test_ball_query_odd_clouds or test_bench_workload_indices"],
                       env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


def test_ball_query_per_trip_bookkeeping_kernel_keeps_parity():
    """PN2_BQ_V1=1 (read once per process) selects the round-1 scan kernel (slot arithmetic inside every trip) instead of the
    deferred-bookkeeping kernel that is the default for C <= 4 (hit masks parked in lanes, lists resolved once per 1024
    points): the same cases must pass on both, so the switch stays usable for re-measurement."""
    e = dict(os.environ, PN2_BQ_V1="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"),
                        os.path.join(ROOT, "tests", "test_gpu_bench_shapes.py"), "-m", "gpu", "-q", "-x",
                        "-k", "test_ball_query_vs_oracle or test_golden_fps_ball_sqdist or test_ball_query_batched_launch_shapes or "
                              "test_config2 or test_ball_query_empty or test_ball_query_odd_clouds or test_ball_query_few_points or test_multi_radius"],
                       env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


@pytest.mark.parametrize("N,S,K,r,kind", [(20, 8, 20, 0.5, "plain"), (20, 8, 20, 0.5, "nan_centre"), (130, 40, 130, 3.0, "plain"),
                                          (1025, 64, 40, 0.3, "plain"), (1024, 64, 1024, 3.0, "plain"), (2049, 33, 64, 0.25, "nan_centre"),
                                          (3000, 50, 3000, 3.0, "plain"), (1056, 40, 16, 0.02, "plain")])
def test_ball_query_few_points_and_window_edges_vs_oracle(pu, oracle, N, S, K, r, kind):
    """Edges of the deferred-bookkeeping kernel: lists as long as the cloud (K = N is the reference's limit, :86 slices the
    sorted [B,S,N] tensor; the padding slots of the staged cloud must never be reported, not even for a NaN centre whose every
    distance is NaN = in radius, :85), lists longer than one 1024-point window, clouds one point past a window / a 128-point trip, and balls that are nearly always empty."""
    B = 3
    x = cloudgen.uniform_cloud(B, N, 77 + N)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
    cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
    if kind == "nan_centre":
        cen[1, 2, 0] = float("nan")
    exp = oracle.query_ball_point(r, K, x, cen, allow_empty=True)
    be = pu._backend()
    i64, i32 = be.ball_query(cen.to(DEV), x.to(DEV), oracle.radius_sq_f32(r), K, want_idx64=True, want_idx32=True)
    assert torch.equal(i64.cpu(), exp)
    assert torch.equal(i32.cpu().long(), exp)


@pytest.mark.parametrize("seed", range(6))
def test_ball_query_random_shapes_vs_oracle(pu, oracle, seed):
    """Seeded random sweep over cloud sizes, list lengths, radii, layouts and batch sizes (every launch shape: 4 / 8 / 16 centres
    per warp, one and several 1024-point windows, ragged last trips), single- and multi-radius: rows bit-equal to the oracle's."""
    g = torch.Generator().manual_seed(9000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    be = pu._backend()
    for _ in range(7):
        N = ri(1, 3000) if ri(0, 3) else ri(1, 200)
        S = ri(1, min(N, 600))
        B = [1, 2, 3, 40, 150][ri(0, 4)] if N * S < 200_000 else ri(1, 3)
        C = 3 if ri(0, 2) else 10
        K = ri(1, min(N, 160))
        r = [0.05, 0.1, 0.2, 0.4, 0.8, 3.0][ri(0, 5)]
        base = cloudgen.uniform_cloud(B, N, 100 * seed + N) if ri(0, 1) else cloudgen.surface_cloud(B, N, 100 * seed + N)
        x = cloudgen.with_category(base) if C == 10 else base
        x = x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if ri(0, 1) else x.contiguous()
        start = torch.randint(0, N, (B,), generator=g, dtype=torch.long)
        cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
        exp = oracle.query_ball_point(r, K, x, cen, allow_empty=True)
        i64, i32 = be.ball_query(cen.to(DEV), x.to(DEV), oracle.radius_sq_f32(r), K, want_idx64=True, want_idx32=True)
        assert torch.equal(i64.cpu(), exp), (B, N, S, C, K, r)
        assert torch.equal(i32.cpu().long(), exp), (B, N, S, C, K, r)
        radii = sorted([r * 0.5, r, r * 1.7][: ri(2, 3)])
        Ks = [ri(1, min(N, 140)) for _ in radii]
        outs = be.ball_query_multi(cen.to(DEV), x.to(DEV), [oracle.radius_sq_f32(q) for q in radii], Ks, want_idx64=True)
        for (o64, _), q, k in zip(outs, radii, Ks):
            assert torch.equal(o64.cpu(), oracle.query_ball_point(q, k, x, cen, allow_empty=True)), (B, N, S, C, radii, Ks, q)


@pytest.mark.parametrize("kind", ["scaled", "shifted", "nan_point", "nan_centre", "far_centre", "flat", "inf_point"])
def test_ball_query_odd_clouds_vs_oracle(pu, oracle, kind):
    """Clouds the grid cull must not get wrong: un-normalised coordinates (x1000, +5000: the fp32 rounding of the expanded form
    grows with |p|^2), a NaN / Inf point (NaN distances count as in radius, :85), NaN and far-away centres, a degenerate
    (planar) bounding box."""
    B, N, S, K, r = 3, 700, 90, 24, 0.2
    x = cloudgen.uniform_cloud(B, N, 321)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
    scale = 1.0
    if kind == "scaled":
        x, scale = x * 1000.0, 1000.0
    elif kind == "shifted":
        x = x + 5000.0
    elif kind == "flat":
        x[..., 2] = 0.25
    cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
    if kind == "nan_point":
        x[1, 17, 1] = float("nan")
    elif kind == "inf_point":
        x[2, 5, 0] = float("inf")
    elif kind == "nan_centre":
        cen[0, 3, 2] = float("nan")
    elif kind == "far_centre":
        cen[1, 7] = torch.tensor([30.0, -40.0, 5.0])
    be = pu._backend()
    for rr in (r * scale, 0.05 * scale):
        exp = oracle.query_ball_point(rr, K, x, cen, allow_empty=True)         # a far centre has no neighbour: row of N
        i64, i32 = be.ball_query(cen.to(DEV), x.to(DEV), oracle.radius_sq_f32(rr), K, want_idx64=True, want_idx32=True)
        assert torch.equal(i64.cpu(), exp), (kind, rr)
        assert torch.equal(i32.cpu().long(), exp)
