"""The PointNet-v1-like pose heads (pn2_v1.py; SURVEY.md §8(f) row 1: width / rotation / translation / sign).

CPU: the fused entry point's plumbing through the oracle-backed fake of the native layer, and (build container only)
state_dict / forward interchangeability with the reference's own model files.  GPU: the fused group_all kernel against the
head's own ATen op sequence."""
import importlib
import sys

import pytest
import torch

import cloudgen
import pn2_models
import pn2_v1
from conftest import REFERENCE_MODEL_DIR


def _input(name, B, N, seed):
    xyz = cloudgen.uniform_cloud(B, N, seed)
    x = cloudgen.with_category(xyz)                                  # [B, N, 3 + 7]
    if name == "width":                                              # width.py: normal_channel=True -> 6 + 7 channels
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.cat([xyz, torch.randn(B, N, 3, generator=g), x[..., 3:]], -1)
    return cloudgen.channels_first(x), xyz[:, :3, :].mean(1)


def _aten_pooled(model, x):
    import torch.nn.functional as F
    for conv, bn in zip(model.conv, model.bn_conv):
        x = F.relu(bn(conv(x)))
    return torch.max(x, 2)[0]


@pytest.mark.parametrize("name", sorted(pn2_v1.V1_ARCH))
def test_fused_entry_point_matches_the_op_sequence_on_the_oracle(fake_native, name):
    torch.manual_seed(0)
    model = pn2_v1.get_v1_model(name, 7).eval()
    pn2_models.randomize_bn_(model, 7)
    x, _ = _input(name, 3, 200, 31)
    with torch.no_grad():
        got = pn2_v1.pointwise_mlp_max(x, model.conv, model.bn_conv)
        exp = _aten_pooled(model, x)
    assert got.shape == (3, 1024)
    assert float((got - exp).abs().max() / exp.abs().max()) <= 1e-5
    assert fake_native._native.calls.count("sa_fused") == 1


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(pn2_v1.V1_ARCH))
def test_state_dicts_and_outputs_interchange_with_the_reference_files(fake_native, name):
    sys.path.append(REFERENCE_MODEL_DIR)
    try:
        for m in (name, "pointnet_utils"):
            sys.modules.pop(m, None)
        ref_mod = importlib.import_module(name)
    finally:
        sys.path.remove(REFERENCE_MODEL_DIR)
    assert ref_mod.__file__.startswith(REFERENCE_MODEL_DIR)
    torch.manual_seed(0)
    ref = ref_mod.get_model(num_category=7).eval()
    torch.manual_seed(0)
    ours = pn2_v1.get_v1_model(name, 7).eval()
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref) == list(sd_ours)                                             # same names, same order
    assert all(sd_ref[k].shape == sd_ours[k].shape for k in sd_ref)
    assert all(torch.equal(sd_ref[k], sd_ours[k]) for k in sd_ref)                   # same seeded initialisation
    pn2_models.randomize_bn_(ref, 7)
    ours.load_state_dict(ref.state_dict())                                           # the reference's checkpoint loads
    x, mean = _input(name, 4, 256, 77)
    args = (x, mean) if name == "translation" else (x,)
    with torch.no_grad():
        a, b = ref(*args), ours(*args)
    a, b = (a if isinstance(a, tuple) else (a,)), (b if isinstance(b, tuple) else (b,))
    assert all(torch.allclose(u, v, rtol=1e-5, atol=1e-6) for u, v in zip(a, b))


@pytest.mark.gpu
@pytest.mark.parametrize("name,prec,tol", [("width", "fp32", 1e-5), ("rotation", "fp32", 1e-5), ("width", "bf16", 2e-2),
                                           ("translation", "bf16", 2e-2), ("sign", "bf16", 2e-2)])
def test_fused_head_on_the_gpu(name, prec, tol):
    import pointnet2_utils as pu
    dev = "cuda:0"
    torch.manual_seed(0)
    model = pn2_v1.get_v1_model(name, 7)
    pn2_models.randomize_bn_(model, 7)
    model = model.eval()
    x, mean = _input(name, 5, 1024, 13)
    with torch.no_grad():
        exp = _aten_pooled(model, x)          # CPU ATen: true fp32 (cuDNN would run the convolutions in TF32)
    model = model.to(dev)
    x, mean = x.to(dev), mean.to(dev)
    prev = pu.get_precision()
    pu.set_precision(prec)
    try:
        with torch.no_grad():
            got = pn2_v1.pointwise_mlp_max(x, model.conv, model.bn_conv).cpu()
            out = model(x, mean) if name == "translation" else model(x)
    finally:
        pu.set_precision(prev)
    err = float((got - exp).abs().max() / exp.abs().max())
    assert err <= tol, (name, prec, err)                     # max|a-b| / max|b|, the tolerance of DESIGN.md §2
    outs = out if isinstance(out, tuple) else (out,)
    assert all(torch.isfinite(o).all() for o in outs)


# ---- golden vectors from the live reference (tests/golden/v1.npz, written by oracle/make_golden.py:gen_v1) ----

def _golden_head(name):
    import golden_io
    d = golden_io.v1_cases()[name]
    torch.manual_seed(0)
    model = pn2_v1.get_v1_model(name, 7).eval()
    pn2_models.randomize_bn_(model, 7)
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in d["sd_keys"]]
    checksum = float(sum(v.double().abs().sum() for v in sd.values()))
    if abs(checksum - float(d["sd_checksum"][0])) > 1e-6 * checksum:
        pytest.skip("seeded initialisation differs from the build container's torch; weights not reproducible")
    x = cloudgen.channels_first(torch.from_numpy(d["points"]))
    return model, d, x, torch.from_numpy(d["mean"])


def _check_outputs(name, outs, d, tol):
    exp0 = torch.from_numpy(d["out0"])
    assert float((outs[0] - exp0).abs().max() / exp0.abs().max()) <= 5 * tol, name
    if name == "sign" and tol <= 1e-5:
        assert torch.equal(outs[1], torch.from_numpy(d["out1"]))


@pytest.mark.parametrize("name", sorted(pn2_v1.V1_ARCH))
def test_golden_head_through_the_oracle_backed_entry_point(fake_native, name):
    model, d, x, mean = _golden_head(name)
    with torch.no_grad():
        pooled = pn2_v1.pointwise_mlp_max(x, model.conv, model.bn_conv)
        out = model(x, mean) if name == "translation" else model(x)         # CPU tensor: the ATen op sequence
    exp = torch.from_numpy(d["pooled"])
    assert float((pooled - exp).abs().max() / exp.abs().max()) <= 1e-5
    _check_outputs(name, out if isinstance(out, tuple) else (out,), d, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("name", sorted(pn2_v1.V1_ARCH))
def test_golden_head_on_the_gpu(name, prec, tol):
    import pointnet2_utils as pu
    model, d, x, mean = _golden_head(name)
    model, x, mean = model.to("cuda:0"), x.to("cuda:0"), mean.to("cuda:0")
    prev = pu.get_precision()
    pu.set_precision(prec)
    try:
        with torch.no_grad():
            pooled = pn2_v1.pointwise_mlp_max(x, model.conv, model.bn_conv).cpu()
            out = model(x, mean) if name == "translation" else model(x)
    finally:
        pu.set_precision(prev)
    exp = torch.from_numpy(d["pooled"])
    err = float((pooled - exp).abs().max() / exp.abs().max())
    assert err <= tol, (name, prec, err)
    outs = tuple(o.cpu() for o in (out if isinstance(out, tuple) else (out,)))
    _check_outputs(name, outs, d, tol)
