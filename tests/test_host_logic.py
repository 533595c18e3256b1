"""CPU suite: host logic of the drop-in pointnet2_utils.py with its native layer replaced by an oracle-backed
fake that lives in tests/ (RNG call order, plumbing, channel orders, concat, strides, caching, errors)."""
import numpy as np
import pytest
import torch

import cloudgen
import golden_io
from test_gpu_models import MODULE_CTORS


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", sorted(MODULE_CTORS))
def test_modules_match_reference_golden_through_the_host_path(fake_native, name):
    pu = fake_native
    d = golden_io.module_cases()[name]
    m = MODULE_CTORS[name](pu)
    m.load_state_dict(d["sd"])
    m.eval()
    pts = cloudgen.channels_first(torch.from_numpy(d["points"]))
    feat = torch.from_numpy(d["feature"]) if "feature" in d else None
    with torch.no_grad():
        torch.manual_seed(42)
        new_pts, new_feat = m(pts, feat)
    assert torch.equal(new_pts, torch.from_numpy(d["new_points"]))
    assert rel_err(new_feat, torch.from_numpy(d["new_feature"])) <= 1e-5
    assert new_feat.shape == tuple(d["new_feature"].shape) and new_feat.dtype == torch.float32


def test_one_rng_draw_per_layer_in_module_order(fake_native, monkeypatch):
    """The reference draws torch.randint once per farthest_point_sample call (:59): once per SA layer, once per
    MSG layer (shared by all radii), never for group_all."""
    pu = fake_native
    draws = []
    real = torch.randint

    def spy(*a, **k):
        draws.append(a[:2])
        return real(*a, **k)

    monkeypatch.setattr(torch, "randint", spy)
    x = cloudgen.channels_first(cloudgen.uniform_cloud(2, 128, 1))
    sa = pu.PointNetSetAbstraction(16, 8, 0.5, 3, [8], False).eval()
    msg = pu.PointNetSetAbstractionMsg(16, [4, 8, 8], [0.3, 0.5, 0.8], 0, [[8], [8], [8]]).eval()
    ga = pu.PointNetSetAbstraction(None, None, None, 3 + 8, [8], True).eval()
    with torch.no_grad():
        p1, f1 = sa(x, None)
        assert len(draws) == 1
        msg(x, None)
        assert len(draws) == 2
        ga(p1, f1)
        assert len(draws) == 2
    assert all(d == (0, 128) for d in draws)


def test_seeded_forward_is_reproducible_and_seed_sensitive(fake_native):
    pu = fake_native
    x = cloudgen.channels_first(cloudgen.uniform_cloud(2, 200, 2))
    sa = pu.PointNetSetAbstraction(32, 8, 0.5, 3, [8], False).eval()
    with torch.no_grad():
        torch.manual_seed(1); a = sa(x, None)[0]
        torch.manual_seed(1); b = sa(x, None)[0]
        torch.manual_seed(2); c = sa(x, None)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_output_strides_and_reference_strides_switch(fake_native):
    pu = fake_native
    x = cloudgen.channels_first(cloudgen.uniform_cloud(2, 64, 3))
    sa = pu.PointNetSetAbstraction(16, 8, 0.6, 3, [8, 12], False).eval()
    with torch.no_grad():
        torch.manual_seed(5)
        pts, feat = sa(x, None)
        assert pts.shape == (2, 3, 16) and pts.stride() == (48, 1, 3)       # view of contiguous [B,S,C] (:173)
        assert feat.shape == (2, 12, 16) and feat.permute(0, 2, 1).is_contiguous()
        pu.set_reference_strides(True)
        torch.manual_seed(5)
        _, feat2 = sa(x, None)
        assert feat2.is_contiguous() and torch.equal(feat2, feat)


def test_state_dict_names_match_the_reference_layout(fake_native):
    pu = fake_native
    sa = pu.PointNetSetAbstraction(16, 8, 0.6, 3, [8, 12], False)
    keys = list(sa.state_dict().keys())
    assert keys[:2] == ["mlp_convs.0.weight", "mlp_convs.0.bias"]
    assert "mlp_bns.1.running_var" in keys and "mlp_bns.1.num_batches_tracked" in keys
    assert sa.mlp_convs[0].weight.shape == (8, 3, 1, 1)
    msg = pu.PointNetSetAbstractionMsg(16, [4, 8], [0.3, 0.5], 5, [[8], [8, 4]], num_category=7)
    keys = list(msg.state_dict().keys())
    assert "conv_blocks.1.1.weight" in keys and "bn_blocks.0.0.running_mean" in keys
    assert msg.conv_blocks[0][0].weight.shape == (8, 5 + 3 + 7, 1, 1)      # in_channel + 3 + num_category (:187)
    assert not any("folded" in k or "packed" in k for k in keys)           # derived caches never enter state_dict


def test_functions_follow_the_reference_signatures(fake_native, oracle):
    pu = fake_native
    x = cloudgen.uniform_cloud(2, 64, 4)
    f = torch.randn(2, 64, 5, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(3)
    idx = pu.farthest_point_sample(x, 16)
    torch.manual_seed(3)
    start = torch.randint(0, 64, (2,), dtype=torch.long)
    assert idx.dtype == torch.int64 and torch.equal(idx, oracle.farthest_point_sample(x, 16, start))
    cen = pu.index_points(x, idx)
    assert torch.equal(cen, oracle.index_points(x, idx))
    ball = pu.query_ball_point(0.5, 8, x, cen)
    assert ball.dtype == torch.int64 and ball.shape == (2, 16, 8)
    assert torch.equal(pu.square_distance(cen, x), oracle.square_distance(cen, x))
    torch.manual_seed(3)
    new_pts, grouped, grouped_pts, fps_idx = pu.sample_and_group(x, f, 16, 8, 0.5, returnfps=True)
    assert torch.equal(fps_idx, idx) and torch.equal(new_pts, cen)
    assert grouped.shape == (2, 16, 8, 3 + 5) and grouped_pts.shape == (2, 16, 8, 5)
    assert torch.equal(grouped[..., :3], oracle.index_points(x, ball) - cen.unsqueeze(2))     # [xyz_norm, feat] (:113-114)
    assert torch.equal(grouped[..., 3:], oracle.index_points(f, ball))
    zeros, rows = pu.sample_and_group_all(x, f)
    assert torch.equal(zeros, torch.zeros(2, 1, 3)) and torch.equal(rows, torch.cat([x, f], -1).view(2, 1, 64, 8))
    with pytest.raises(RuntimeError):
        pu.query_ball_point(0.5, 65, x, cen)


def test_radius_squared_is_double_then_one_rounding():
    import pointnet2_utils as pu
    assert pu._radius_sq(0.2) == float(np.float32(0.2 ** 2))
    assert pu._radius_sq(0.2) != float(np.float32(0.2) * np.float32(0.2))


def test_precision_switch_validates():
    import pointnet2_utils as pu
    prev = pu.get_precision()
    with pytest.raises(ValueError):
        pu.set_precision("fp16")
    pu.set_precision("bf16"); assert pu.get_precision() == "bf16"
    pu.set_precision(prev)


def test_fold_cache_rebuilds_when_parameters_change(fake_native):
    pu = fake_native
    x = cloudgen.channels_first(cloudgen.uniform_cloud(2, 64, 5))
    sa = pu.PointNetSetAbstraction(16, 8, 0.6, 3, [8], False).eval()
    with torch.no_grad():
        torch.manual_seed(0); sa(x, None)
        torch.manual_seed(0); sa(x, None)
        n0 = pu._native.calls.count("pack")
        sa.mlp_bns[0].running_mean.add_(1.0)
        torch.manual_seed(0); sa(x, None)
    assert n0 == 1 and pu._native.calls.count("pack") == 2


def test_real_native_layer_refuses_cpu_tensors():
    import _pn2_native
    nat = _pn2_native.Native()
    with pytest.raises(RuntimeError, match="no CPU path"):
        nat.fps(torch.zeros(1, 8, 3), torch.zeros(1, dtype=torch.long), 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        nat.ball_query(torch.zeros(1, 2, 3), torch.zeros(1, 8, 3), 0.04, 2)


def test_training_mode_runs_autograd_through_the_host_path(fake_native):
    pu = fake_native
    x = cloudgen.channels_first(cloudgen.uniform_cloud(4, 64, 6))
    sa1 = pu.PointNetSetAbstractionMsg(16, [4, 8], [0.4, 0.8], 0, [[8], [8]]).train()
    sa2 = pu.PointNetSetAbstraction(None, None, None, 16 + 3, [8], True).train()
    p1, f1 = sa1(x, None)
    _, f2 = sa2(p1, f1)
    f2.sum().backward()
    assert sa1.conv_blocks[0][0].weight.grad is not None and sa2.mlp_convs[0].weight.grad.abs().sum() > 0
    assert int(sa1.bn_blocks[0][0].num_batches_tracked) == 1


def test_training_mode_factorised_msg_branch_matches_the_direct_form(fake_native):
    """MSG training forward / backward with the first convolution factorised (feature part once per point) equals the
    direct form (gather the D-wide rows, convolve every grouped copy) to fp32 rounding."""
    import torch
    pu = fake_native
    torch.manual_seed(0)
    mod = pu.PointNetSetAbstractionMsg(16, [4, 8], [0.4, 0.8], 24, [[8, 8, 16], [16, 16, 32]], num_category=0).train()
    pts = (torch.rand(3, 3, 40) * 2 - 1)
    feat = torch.randn(3, 24, 40, requires_grad=True)
    outs, grads = [], []
    for flag in (True, False):
        pu._factorise_first_layer = flag
        mod.zero_grad()
        if feat.grad is not None:
            feat.grad = None
        torch.manual_seed(5)
        _, f = mod(pts, feat)
        f.square().sum().backward()
        outs.append(f.detach().clone())
        grads.append([p.grad.clone() for p in mod.parameters()] + [feat.grad.clone()])
    pu._factorise_first_layer = True
    assert float((outs[0] - outs[1]).abs().max() / outs[1].abs().max()) < 1e-5
    scale = max(float(b.abs().max()) for b in grads[1])     # a convolution bias in front of BatchNorm has zero gradient: noise only
    for a, b in zip(*grads):
        assert float((a - b).abs().max()) / scale < 1e-5


def test_data_writes_need_invalidate_caches(fake_native):
    """the derived-weight caches key on tensor._version, which a write through `.data` does not bump -- the
    documented hook `invalidate_caches()` makes the next forward repack."""
    pu = fake_native
    m = pu.PointNetSetAbstraction(8, 4, 0.8, 3, [8, 8], False).eval()
    pts = cloudgen.channels_first(cloudgen.uniform_cloud(2, 64, 3))
    with torch.no_grad():
        torch.manual_seed(1)
        a = m(pts, None)[1].clone()
        m.mlp_convs[0].weight.data.mul_(3.0)           # invisible to the version counter
        torch.manual_seed(1)
        b = m(pts, None)[1].clone()
        pu.invalidate_caches()
        torch.manual_seed(1)
        c = m(pts, None)[1].clone()
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_fused_head_is_used_in_eval_and_matches_the_op_sequence(fake_native):
    """Eval-mode models hand the FC head to ONE fc_head call (BatchNorm1d folded, weights packed [K/4][H][4]); the result equals
    the reference's op sequence; training mode keeps the ATen sequence (autograd)."""
    import pn2_models
    import torch.nn.functional as F
    pu = fake_native
    for name in ("pointnet2_cls_ssg", "translation_ssg", "sign_ssg", "rotation_ssg"):
        torch.manual_seed(0)
        model = pn2_models.get_model(name, 7)
        pn2_models.randomize_bn_(model)
        model.eval()
        xyz = cloudgen.uniform_cloud(2, 1024, 5)
        pose = not name.startswith("pointnet2_cls")
        x = cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz)
        mean = xyz[:, :3, :].mean(1)
        args = (x, mean) if name.startswith("translation") else (x,)
        pu._native.calls.clear()
        with torch.no_grad():
            torch.manual_seed(4)
            out = model(*args)
            assert pu._native.calls.count("fc_head") == 1
            # the same forward with the head computed by the reference's op sequence
            torch.manual_seed(4)
            feat, pts = None, x
            for i in range(1, model.n_sa + 1):
                pts, feat = getattr(model, f"sa{i}")(pts, feat)
            h = F.relu(model.bn1(model.fc1(feat.view(2, 1024))))
            h = F.relu(model.bn2(model.fc2(h)))
            y = model.fc3(h)
            if name.startswith("pointnet2_cls"):
                exp = F.log_softmax(y, -1)
            elif name.startswith("translation"):
                exp = y + model.mean_fc2(F.relu(model.mean_bn1(model.mean_fc1(mean))))
            elif name.startswith("sign"):
                exp = torch.sigmoid(y)
            else:
                exp = y
        got = out[0] if isinstance(out, tuple) else out
        assert torch.allclose(got, exp, rtol=1e-5, atol=1e-6), name


def test_bf16_rows_are_handed_to_every_next_layer(fake_native, monkeypatch):
    """bf16 mode: a layer whose kernel writes the bf16 copy of its pooled rows attaches it to the tensor it returns, and the NEXT
    layer passes it on as `feature_bf16` whether or not its own kernel writes one (the streaming kernel of the wide group_all
    stacks reads it too); a tensor modified in between loses the copy."""
    pu = fake_native
    be = pu._backend()
    seen = []
    real = be.sa_fused

    def sa_fused(points, feature, centres, idx32, order, dims, packed, prec, out, co_off=0, feature_bf16=None, out_bf16=None, algo_dims=None):
        seen.append((order, feature_bf16, out_bf16))
        if out_bf16 is not None:
            real(points, feature, centres, idx32, order, dims, packed, prec, out, co_off)
            out_bf16.copy_(out.to(out_bf16.dtype))
            return out
        return real(points, feature, centres, idx32, order, dims, packed, prec, out, co_off)

    monkeypatch.setattr(be, "sa_fused", sa_fused)
    monkeypatch.setattr(be, "sa_bf16_side_io", lambda K, C, order, dims, prec: order != 2)      # group_all stacks: no bf16 output
    pu.set_precision("bf16")
    try:
        sa1 = pu.PointNetSetAbstraction(16, 8, 0.4, 3, [8, 16], False).eval()
        sa2 = pu.PointNetSetAbstraction(None, None, None, 16 + 3, [16, 32], True).eval()
        x = cloudgen.channels_first(cloudgen.uniform_cloud(2, 64, 3))
        with torch.no_grad():
            p1, f1 = sa1(x, None)
            assert seen[-1][1] is None and seen[-1][2] is not None              # no features in, bf16 rows out
            twin = seen[-1][2]
            sa2(p1, f1)
            assert seen[-1][0] == 2 and seen[-1][1] is twin and seen[-1][2] is None
            f1.add_(1.0)                                                         # the copy no longer describes the tensor
            sa2(p1, f1)
            assert seen[-1][1] is None
    finally:
        pu.set_precision("fp32")
