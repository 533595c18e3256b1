"""CPU suite, build container only (-m reference is implied by the skip rule in conftest): the reference's
own, unmodified model files run on top of this repo's pointnet2_utils (drop-in boundary = "a module named
pointnet2_utils first on sys.path", train_rotation.py:16-18,32), and the oracle tracks the live reference on
randomly drawn shapes."""
import importlib
import sys

import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import cloudgen
import golden_io
from conftest import REFERENCE_MODEL_DIR

pytestmark = pytest.mark.reference


def import_reference(name):
    """Import /root/reference/model/<name>.py the way the reference's scripts do (bare module name)."""
    sys.path.append(REFERENCE_MODEL_DIR)           # appended: our package dir stays first, like sys.path.append in the scripts
    try:
        sys.modules.pop(name, None)
        m = importlib.import_module(name)
    finally:
        sys.path.remove(REFERENCE_MODEL_DIR)
    assert m.__file__.startswith(REFERENCE_MODEL_DIR)
    if not hasattr(m, "torch"):
        m.torch = torch                             # sign_ssg.py / sign_msg.py forget `import torch` (sign_ssg.py:34)
    return m


@pytest.mark.parametrize("name", sorted(golden_io.model_cases()))
def test_unmodified_reference_model_files_run_on_our_pointnet2_utils(fake_native, name):
    pu = fake_native
    mod = import_reference(name)
    assert sys.modules["pointnet2_utils"] is pu, "the reference model file must have picked up OUR pointnet2_utils"
    d = golden_io.model_cases()[name]
    torch.manual_seed(0)
    model = mod.get_model(7).eval()
    import pn2_models
    pn2_models.randomize_bn_(model, 7)
    assert isinstance(model.sa1, (pu.PointNetSetAbstraction, pu.PointNetSetAbstractionMsg))
    xyz = torch.from_numpy(d["xyz"])
    pose = not name.startswith("pointnet2_cls")
    x = cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz)
    args = (x, torch.from_numpy(d["mean"])) if name.startswith("translation") else (x,)
    with torch.no_grad():
        torch.manual_seed(42)
        out = model(*args)
    outs = out if isinstance(out, tuple) else (out,)
    for j, o in enumerate(outs):
        ref = torch.from_numpy(d[f"out{j}"])
        if ref.dtype.is_floating_point and ref.abs().max() > 0:
            err = float((o - ref).abs().max() / ref.abs().max())
            assert err <= 5e-5, (name, j, err)
        else:
            assert torch.equal(o, ref)
    # same parameter names / shapes as our table-driven restatement: checkpoints are interchangeable
    ours = pn2_models.get_model(name, 7)
    assert [(k, tuple(v.shape)) for k, v in ours.state_dict().items()] == [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    ours.load_state_dict(model.state_dict())


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(B=st.integers(1, 3), N=st.integers(20, 300), C=st.sampled_from([3, 10]), S=st.integers(8, 96), K=st.integers(1, 20),
       r=st.floats(0.15, 1.5), strided=st.booleans(), dup=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_oracle_tracks_live_reference_on_random_shapes(oracle, B, N, C, S, K, r, strided, dup, seed):
    ref = import_reference_utils()
    if S * N * C < 400:        # below this torch.bmm skips MKL and uses a non-FMA kernel (SURVEY.md 8a trap)
        S = 400 // (N * C) + 1
    K = min(K, N)
    x = cloudgen.uniform_cloud(B, N, seed % 1000)
    if dup:
        x = cloudgen.duplicate_heavy(x, 0.5, seed % 997)
    if C == 10:
        x = cloudgen.with_category(x)
    pts = x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if strided else x.contiguous()
    torch.manual_seed(seed)
    start = torch.randint(0, N, (B,), dtype=torch.long)
    torch.manual_seed(seed)
    fps = ref.farthest_point_sample(pts, S)
    assert torch.equal(oracle.farthest_point_sample(pts, S, start), fps)
    cen = ref.index_points(pts, fps)
    assert torch.equal(oracle.query_ball_point(r, K, pts, cen), ref.query_ball_point(r, K, pts, cen))
    assert torch.equal(oracle.square_distance(cen, pts), ref.square_distance(cen, pts))


_ref_utils = None


def import_reference_utils():
    """The reference's own pointnet2_utils under an alias (so that it does not shadow ours)."""
    global _ref_utils
    if _ref_utils is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("reference_pointnet2_utils", REFERENCE_MODEL_DIR + "/pointnet2_utils.py")
        _ref_utils = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ref_utils)
    return _ref_utils


@pytest.mark.parametrize("kind", ["ssg", "msg", "all"])
def test_training_mode_matches_the_reference_modules(fake_native, kind):
    """Training mode (batch-statistics BatchNorm, autograd): our modules -- sampling / grouping through the native layer,
    MLP on ATen, MSG branches with the factorised first convolution -- against the reference's own modules with the same
    parameters, inputs and FPS draws: outputs, parameter gradients, feature gradients and running statistics."""
    pu = fake_native
    ref = import_reference_utils()
    torch.manual_seed(3)
    if kind == "ssg":
        mk = lambda m: m.PointNetSetAbstraction(24, 8, 0.6, 20 + 3, [16, 16, 32], False)
    elif kind == "msg":
        mk = lambda m: m.PointNetSetAbstractionMsg(24, [4, 8], [0.4, 0.8], 20, [[8, 8, 16], [16, 16, 32]], num_category=0)
    else:
        mk = lambda m: m.PointNetSetAbstraction(None, None, None, 20 + 3, [16, 32], True)
    torch.manual_seed(1)
    ours = mk(pu).train()
    torch.manual_seed(1)
    theirs = mk(ref).train()
    theirs.load_state_dict(ours.state_dict())
    g = torch.Generator().manual_seed(9)
    pts = torch.rand(3, 3, 60, generator=g) * 2 - 1
    fa = torch.randn(3, 20, 60, generator=g, requires_grad=True)
    fb = fa.detach().clone().requires_grad_(True)
    torch.manual_seed(77)
    pa, ya = ours(pts, fa)
    torch.manual_seed(77)
    pb, yb = theirs(pts, fb)
    assert torch.equal(pa, pb) and ya.shape == yb.shape
    assert float((ya - yb).abs().max() / yb.abs().max()) < 1e-5
    w = torch.randn(ya.shape, generator=g)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    scale = max(float(p.grad.abs().max()) for p in theirs.parameters())
    for (na, a), (nb, b) in zip(ours.named_parameters(), theirs.named_parameters()):
        assert na == nb and float((a.grad - b.grad).abs().max()) / scale < 1e-5, na
    assert float((fa.grad - fb.grad).abs().max() / fb.grad.abs().max()) < 1e-5
    for (na, a), (nb, b) in zip(ours.named_buffers(), theirs.named_buffers()):
        assert na == nb and torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-6), na      # running statistics
