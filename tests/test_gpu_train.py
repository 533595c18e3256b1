"""GPU suite: the training step of pn2_train.py on one B200 (sampling / grouping on the product kernels, MLP + BatchNorm on
ATen under autograd): gradients land in the flat bucket, the loss on a fixed batch goes down, forward in training mode agrees
with the ATen restatement of the same step run on the oracle's indices."""
import pytest
import torch

import cloudgen
import pn2_models
import pn2_shard
import pn2_train
import pointnet2_utils as pu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rotation_ssg", "rotation_msg", "translation_msg"])
def test_training_steps_reduce_the_loss(name):
    dev = "cuda:0"
    torch.manual_seed(0)
    model = pn2_models.get_model(name, 7).to(dev)
    bucket = pn2_train.GradBucket(model.parameters())
    opt = pn2_train.make_optimizer(model, learning_rate=1e-3)
    crit = pn2_train.get_loss("L2_loss", "sum")
    B = 6
    xyz = cloudgen.uniform_cloud(B, 1024, 21)
    x = cloudgen.channels_first(cloudgen.with_category(xyz)).to(dev)
    mean = xyz[:, :3, :].mean(1).to(dev) if name.startswith("translation") else None
    tgt = torch.randn(B, 3, generator=torch.Generator().manual_seed(3)).to(dev)
    before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    losses = []
    for i in range(8):
        with pn2_shard.sharded_starts(B, 1, 0, seed=50, device=dev):      # same centroids every step: a fixed problem
            losses.append(float(pn2_train.train_step(model, crit, opt, bucket, x, tgt, mean=mean)))
        assert bucket.attached() and torch.isfinite(bucket.flat).all() and float(bucket.flat.abs().sum()) > 0
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert not torch.equal(before, after)
    assert all(l == l for l in losses) and min(losses[4:]) < losses[0], losses


# ---- fused training-mode BatchNorm + ReLU (+ max over K) kernels, csrc/pn2_bn_train.cu ----

def _bn_reference(y, gamma, beta, eps, K, dz):
    """ATen in float64: relu(batch_norm(y)) (+ max over groups of K rows) and its gradients."""
    y = y.double().requires_grad_(True)
    gamma, beta = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    z = torch.relu(torch.nn.functional.batch_norm(y, None, None, gamma, beta, True, 0.0, eps))
    if K:
        z = z.view(-1, K, y.shape[1]).max(1)[0]
    z.backward(dz.double())
    return z.detach(), y.grad, gamma.grad, beta.grad, y.detach().mean(0), y.detach().var(0, unbiased=False)


@pytest.mark.parametrize("M,C,K", [(1000, 8, 0), (4096, 96, 16), (2048, 1024, 128), (999, 20, 0), (990, 20, 33),
                                   (512, 2048, 0), (600, 300, 4), (2, 4, 0), (300000, 64, 32), (64, 32, 64)])
def test_bn_relu_train_kernels_match_aten(M, C, K):
    be = pu._backend()
    g = torch.Generator().manual_seed(M + C)
    y = (torch.randn(M, C, generator=g) * 2.0 + torch.randn(C, generator=g) * 3.0).cuda()       # off-centre channels
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.3).cuda()
    dz = torch.randn(M // K if K else M, C, generator=g).cuda()
    z, stats, arg = be.bn_relu_train_fwd(y, gamma, beta, 1e-5, K)
    dy, sums = be.bn_relu_train_bwd(dz, y, arg, stats, K)
    zr, dyr, dgr, dbr, mean, var = _bn_reference(y, gamma, beta, 1e-5, K, dz)

    def rel(a, b):
        return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))

    assert rel(stats[0], mean) <= 1e-6 and rel(stats[1], var) <= 1e-6
    assert rel(z, zr) <= 1e-5, rel(z, zr)
    # dy is a difference of O(|a| |dz|) terms (exactly zero for M = 2): measure it on that scale
    dy_scale = max(float(dyr.abs().max()), 1e-2 * float(stats[3].abs().max() * dz.abs().max()))
    assert float((dy.double() - dyr).abs().max()) / dy_scale <= 2e-5
    assert rel(sums[1], dgr) <= 1e-5 and rel(sums[0], dbr) <= 1e-5
    if K:      # the argmax rows reproduce the pooled values
        full = torch.relu(torch.addcmul(stats[4], y, stats[3])).view(M // K, K, C)
        assert torch.equal(full.gather(1, arg.long().unsqueeze(1)).squeeze(1), z)
        assert torch.equal(arg.long(), full.max(1)[1])                                          # first maximal row, like torch.max
    # deterministic: a second launch gives the same bits
    z2, stats2, _ = be.bn_relu_train_fwd(y, gamma, beta, 1e-5, K)
    dy2, sums2 = be.bn_relu_train_bwd(dz, y, arg, stats2, K)
    assert torch.equal(z, z2) and torch.equal(stats, stats2) and torch.equal(dy, dy2) and torch.equal(sums, sums2)


def test_bn_relu_train_propagates_nan_like_aten():
    be = pu._backend()
    y = torch.randn(64, 8).cuda()
    y[5, 3] = float("nan")
    z, stats, _ = be.bn_relu_train_fwd(y, torch.ones(8).cuda(), torch.zeros(8).cuda(), 1e-5, 0)
    assert torch.isnan(z[:, 3]).all() and torch.isfinite(z[:, [0, 1, 2, 4, 5, 6, 7]]).all()
    with pytest.raises(ValueError):
        be.bn_relu_train_fwd(torch.randn(1, 8).cuda(), torch.ones(8).cuda(), torch.zeros(8).cuda(), 1e-5, 0)


@pytest.mark.parametrize("kind", ["ssg", "msg", "msg_plain", "all"])
def test_fused_training_path_equals_the_aten_op_sequence(kind):
    """Same module, same centroids: the fused rows path (GEMM + product BatchNorm/ReLU/max kernels) against the reference's
    op sequence on ATen (conv2d, batch_norm, relu, max) -- outputs, parameter / input gradients and running statistics."""
    dev = "cuda:0"
    torch.manual_seed(0)
    if kind == "ssg":
        mod = pu.PointNetSetAbstraction(64, 16, 0.4, 24 + 3, [32, 32, 64], False)
    elif kind == "all":
        mod = pu.PointNetSetAbstraction(None, None, None, 24 + 3, [32, 64, 128], True)
    else:
        mod = pu.PointNetSetAbstractionMsg(64, [8, 16], [0.3, 0.6], 24 if kind == "msg" else 4, [[16, 16, 32], [16, 24, 48]])
    mod = mod.to(dev).train()
    D = 4 if kind == "msg_plain" else 24            # D = 4 <= first width: the factorised first layer is not used
    pts = (torch.rand(5, 3, 256, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev)
    feat0 = torch.randn(5, D, 256, generator=torch.Generator().manual_seed(2)).to(dev)
    state0 = {k: v.clone() for k, v in mod.state_dict().items()}
    prev_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False         # both sides in true fp32
    res = {}
    try:
        for fused in (True, False):
            pu._train_fused = fused
            mod.load_state_dict(state0)
            mod.zero_grad()
            feat = feat0.clone().requires_grad_(True)
            with pn2_shard.sharded_starts(5, 1, 0, seed=9, device=dev):
                _, f = mod(pts, feat)
            (f * torch.linspace(0.5, 1.5, f.shape[1], device=dev)[None, :, None]).square().sum().backward()
            res[fused] = (f.detach().clone(), [p.grad.clone() for p in mod.parameters()], feat.grad.clone(),
                          {k: v.clone() for k, v in mod.state_dict().items() if "running" in k or "tracked" in k})
    finally:
        pu._train_fused = True
        torch.backends.cudnn.allow_tf32 = prev_tf32
    (fa, ga, xa, sa), (fb, gb, xb, sb) = res[True], res[False]
    assert fa.shape == fb.shape and float((fa - fb).abs().max() / fb.abs().max()) <= 1e-5
    scale = max(float(g.abs().max()) for g in gb)
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) / scale <= 2e-5
    assert float((xa - xb).abs().max() / xb.abs().max()) <= 2e-5
    for k in sa:
        assert torch.allclose(sa[k].float(), sb[k].float(), rtol=1e-5, atol=1e-6), k
