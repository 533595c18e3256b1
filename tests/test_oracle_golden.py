"""CPU suite, part 1: the oracle (C restatement) and the torch timing port against the golden vectors that
oracle/make_golden.py recorded from the live reference.  This is the pin of the oracle on machines without
/root/reference."""
import numpy as np
import pytest
import torch

import cloudgen
import golden_io

INDEX_CASES = golden_io.index_cases()
MODULE_CASES = golden_io.module_cases()


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("case", sorted(INDEX_CASES))
def test_oracle_indices_match_reference(oracle, case):
    d = INDEX_CASES[case]
    pts = golden_io.as_layout(d["points"], d["layout"])
    start = torch.from_numpy(d["start"])
    fps = oracle.farthest_point_sample(pts, d["S"], start)
    assert torch.equal(fps, torch.from_numpy(d["fps"]).long())
    cen = oracle.index_points(pts, fps)
    ball = oracle.query_ball_point(d["r"], d["K"], pts, cen)
    assert torch.equal(ball, torch.from_numpy(d["ball"]).long())
    sd = oracle.square_distance(cen[:, :4].contiguous(), pts)
    assert torch.equal(sd, torch.from_numpy(d["sqdist_sample"]))


@pytest.mark.parametrize("case", sorted(INDEX_CASES))
def test_torch_port_indices_match_reference(case):
    import pn2_torch_port as port
    d = INDEX_CASES[case]
    pts = golden_io.as_layout(d["points"], d["layout"])
    fps = port.fps(pts, d["S"], torch.from_numpy(d["start"]))
    assert torch.equal(fps, torch.from_numpy(d["fps"]).long())
    cen = port.take_rows(pts, fps)
    assert torch.equal(port.ball_query(d["r"], d["K"], pts, cen), torch.from_numpy(d["ball"]).long())


def _oracle_module(oracle, name, d):
    import pn2_torch_port as port
    from test_gpu_models import MODULE_CTORS
    m = MODULE_CTORS[name](port)
    m.load_state_dict(d["sd"])
    m.eval()
    pts = cloudgen.channels_first(torch.from_numpy(d["points"]))
    feat = torch.from_numpy(d["feature"]) if "feature" in d else None
    B, N = d["points"].shape[:2]
    torch.manual_seed(42)
    start = torch.randint(0, N, (B,), dtype=torch.long)
    if isinstance(m, port.PointNetSetAbstraction):
        o_pts, o_feat, _, _ = oracle.set_abstraction(pts, feat, m.point_number, m.sample_number, m.radius,
                                                     oracle.layers_of(m.mlp_convs, m.mlp_bns), m.group_all, start)
    else:
        o_pts, o_feat, _, _ = oracle.set_abstraction_msg(pts, feat, m.point_number, m.sample_number_list, m.radius_list,
                                                         [oracle.layers_of(c, b) for c, b in zip(m.conv_blocks, m.bn_blocks)], start)
    port.set_start_source(lambda B_, N_: start)
    try:
        with torch.no_grad():
            p_pts, p_feat = m(pts, feat)
    finally:
        port.set_start_source(None)
    return (o_pts, o_feat), (p_pts, p_feat)


@pytest.mark.parametrize("name", sorted(MODULE_CASES))
def test_oracle_and_port_modules_match_reference(oracle, name):
    d = MODULE_CASES[name]
    (o_pts, o_feat), (p_pts, p_feat) = _oracle_module(oracle, name, d)
    ref_pts, ref_feat = torch.from_numpy(d["new_points"]), torch.from_numpy(d["new_feature"])
    assert torch.equal(o_pts, ref_pts) and torch.equal(p_pts, ref_pts)
    assert rel_err(o_feat, ref_feat) <= 1e-5
    assert rel_err(p_feat, ref_feat) <= 1e-5


@pytest.mark.parametrize("C", [3, 5, 8, 10, 13, 16])
def test_summation_order_rule_matches_this_hosts_aten(oracle, C):
    """Self-check of the layout-dependent sum order against torch.sum on THIS host: documents an
    ATen / CPU-capability drift instead of letting index parity fail silently somewhere else.
    mode 1 = contiguous operand; strided operand ([B,C,N] permuted): mode 0 for the points ATen vectorises
    over (n < 32*floor(N/32)), mode 2 for the leftover points."""
    import ctypes
    g = torch.Generator().manual_seed(C)
    N = 32 * 5 + 19
    v = torch.rand(1, N, C, generator=g) ** 2
    strided = v.permute(0, 2, 1).contiguous().permute(0, 2, 1)

    def rule(rows, modes):
        return torch.tensor([oracle.lib().pn2o_sum_order(ctypes.c_void_p(rows[i].data_ptr()), C, modes[i]) for i in range(rows.shape[0])])

    rows = v[0].contiguous()
    assert torch.equal(rule(rows, [1] * N), v.sum(-1)[0]), (C, "contiguous")
    assert torch.equal(rule(rows, [0 if n < 32 * 5 else 2 for n in range(N)]), strided.sum(-1)[0]), (C, "strided")


def test_fps_edge_cases(oracle):
    x = cloudgen.uniform_cloud(2, 16, 3)
    start = torch.tensor([3, 15])
    idx = oracle.farthest_point_sample(x, 40, start)          # S > N: after N distinct picks index 0 repeats
    assert idx.shape == (2, 40)
    for b in range(2):
        assert sorted(idx[b, :16].tolist()) == list(range(16))
        assert (idx[b, 16:] == 0).all()
    assert oracle.farthest_point_sample(x, 0, start).shape == (2, 0)
    import pn2_torch_port as port
    nan = x.clone()
    nan[0, 5] = float("nan")        # a NaN distance never replaces (:65-66): point 5 keeps 1e10 and, once picked, sticks
    idx = oracle.farthest_point_sample(nan, 16, start)
    assert torch.equal(idx, port.fps(nan, 16, start))
    assert (idx[0, 1:] == 5).all()


def test_ball_query_edge_cases(oracle):
    x = cloudgen.uniform_cloud(1, 64, 4)
    cen = x[:, :8].contiguous()
    all_in = oracle.query_ball_point(100.0, 64, x, cen)
    assert torch.equal(all_in, torch.arange(64).expand(1, 8, 64))
    tiny = oracle.query_ball_point(1e-3, 4, x, cen)            # only the centre itself: padded with it
    assert torch.equal(tiny, torch.arange(8).view(1, 8, 1).expand(1, 8, 4))
    with pytest.raises(RuntimeError):
        oracle.query_ball_point(0.1, 65, x, cen)               # K > N: the reference raises at :89
    far = torch.full((1, 2, 3), 9.0)
    with pytest.raises(RuntimeError):
        oracle.query_ball_point(0.1, 4, x, far)                # empty ball: the reference fails in index_points
    assert (oracle.query_ball_point(0.1, 4, x, far, allow_empty=True) == 64).all()
