"""Input preparation (pn2_preprocess.py) against the restatement of provider.py:5-21,166-180."""
import numpy as np
import pytest
import torch

import cloudgen


def _raw(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, N, 3, generator=g) * 3.0 - 1.0 + torch.randn(B, 1, 3, generator=g)).float()     # off-centre, unnormalised


def test_oracle_restatement_properties(oracle):
    x = _raw(3, 200, 1)
    y = oracle.normalization(x.numpy())
    assert y.shape == (3, 200, 3) and np.allclose(y.mean(1), 0.0, atol=1e-6)
    assert np.allclose(np.sqrt((y ** 2).sum(-1)).max(1), 1.0, atol=1e-6)
    cat = torch.tensor([0, 6, 3])
    z = oracle.splice_torch(torch.from_numpy(y).float(), cat, 7)
    assert z.shape == (3, 200, 10) and torch.equal(z[..., 3:].sum(-1), torch.ones(3, 200))
    assert all(bool((z[b, :, 3 + int(cat[b])] == 1).all()) for b in range(3))


def test_splice_without_the_batch_loop_equals_the_reference_form(oracle):
    import pn2_preprocess
    x = _raw(4, 50, 2)
    cat = torch.tensor([1, 0, 6, 2])
    assert torch.equal(pn2_preprocess.splice_torch(x, cat, 7), oracle.splice_torch(x, cat, 7))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,layout", [(5, 1024, "contig"), (3, 777, "strided"), (1, 1, "contig"), (64, 2048, "contig")])
def test_prepare_pose_input_on_the_gpu(oracle, B, N, layout):
    import pn2_preprocess
    x = _raw(B, N, 10 + N)
    if N == 1:
        x = x + 1.0       # a single point: centroid = the point, norm 0 -> 0/0 = nan in the reference too
    cat = torch.arange(B) % 7
    xin = x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "strided" else x
    got = pn2_preprocess.prepare_pose_input(xin.cuda(), cat.cuda(), 7).cpu()
    exp = oracle.splice_torch(torch.from_numpy(oracle.normalization(x.numpy())).float(), cat, 7).permute(0, 2, 1)
    assert got.shape == (B, 10, N)
    if N == 1:
        assert torch.isnan(got[:, :3]).all() and torch.isnan(exp[:, :3]).all() and torch.equal(got[:, 3:], exp[:, 3:])
        return
    assert torch.equal(got[:, 3:], exp[:, 3:])                                   # one-hot rows: exact
    assert float((got[:, :3] - exp[:, :3]).abs().max()) <= 1e-5                  # normalised coordinates lie in [-1, 1]: the fp32 tolerance of the path
    plain = pn2_preprocess.normalization(xin.cuda()).cpu()
    assert plain.shape == (B, N, 3) and torch.equal(plain, got[:, :3].permute(0, 2, 1))


# ---- the numpy copies of farthest point sampling + random_sample (provider.py:183-204, ModelDataLoader.py:10-46) ----

import hashlib

import golden_io

NPFPS = golden_io.npfps_cases()


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _case(name):
    d = NPFPS[name]
    if not d["input_ok"]:
        pytest.skip("this numpy's default_rng stream differs from the build container's; input not reproducible")
    return d


@pytest.mark.parametrize("name", sorted(NPFPS))
def test_numpy_fps_restatement_matches_the_golden_vectors(oracle, name):
    d = _case(name)
    rows, idx = oracle.farthest_point_sample_np(d["points"], d["number"], d["start"])
    assert np.array_equal(idx, d["idx"])
    assert _sha(rows) == str(d["rows_sha"][0]) and rows.dtype == d["points"].dtype


@pytest.mark.parametrize("name", ["f64_small", "f64_cols6", "f64_dup", "f32_small"])
def test_numpy_fps_drop_in_draws_like_the_reference(fake_native, name):
    import pn2_preprocess
    d = _case(name)
    seed = 1000 + [s[5] for s in _specs() if s[0] == name][0]
    np.random.seed(seed)
    rows = pn2_preprocess.farthest_point_sample(d["points"], d["number"], device="cpu")
    after = np.random.get_state()[1].copy()
    np.random.seed(seed)
    np.random.randint(0, d["points"].shape[0])                          # exactly one draw, the reference's (:193)
    assert np.array_equal(after, np.random.get_state()[1])
    assert _sha(rows) == str(d["rows_sha"][0])
    assert fake_native._native.calls == ["fps_xyz"]


def _specs():
    src = open(golden_io.os.path.join(golden_io.os.path.dirname(golden_io.GOLDEN), "..", "oracle", "make_golden.py")).read()
    ns = {"np": np}
    exec(src[src.index("NPFPS_SPECS = ["):src.index("def gen_npfps(")], ns)
    return ns["NPFPS_SPECS"]


def test_random_sample_keeps_the_reference_draw():
    import pn2_preprocess
    pc = np.random.default_rng(0).standard_normal((50, 4))
    np.random.seed(3)
    got = pn2_preprocess.random_sample(pc, 20)
    np.random.seed(3)
    exp = pc[np.random.choice(50, 20, replace=False)]                    # ModelDataLoader.py:44-45
    assert np.array_equal(got, exp)
    assert pn2_preprocess.random_sample(pc, 50) is pc and pn2_preprocess.random_sample(pc, 64) is pc      # :42-43
    np.random.seed(3)
    assert np.array_equal(pn2_preprocess.random_sample(torch.from_numpy(pc), 20).numpy(), exp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(NPFPS))
def test_numpy_fps_on_the_gpu(name):
    import pn2_preprocess
    import pointnet2_utils as pu
    d = _case(name)
    pc, number = d["points"], d["number"]
    seed = 1000 + [s[5] for s in _specs() if s[0] == name][0]
    np.random.seed(seed)
    rows = pn2_preprocess.farthest_point_sample(pc, number)
    assert rows.dtype == pc.dtype and rows.shape == (number, pc.shape[1])
    assert _sha(rows) == str(d["rows_sha"][0])
    # the kernel on the array as it lies (row stride C, no xyz copy), three clouds per launch with different starts
    N = pc.shape[0]
    starts = torch.tensor([d["start"], 0, N - 1])
    stack = torch.from_numpy(np.stack([pc, pc[::-1].copy(), pc])).cuda()
    idx = pu._backend().fps_xyz(stack, starts.cuda(), number).cpu().numpy()
    assert np.array_equal(idx[0], d["idx"])
    import pn2_oracle
    for b, cloud in enumerate([pc, pc[::-1].copy(), pc]):
        assert np.array_equal(idx[b], pn2_oracle.farthest_point_sample_np(cloud, number, int(starts[b]))[1]), (name, b)
