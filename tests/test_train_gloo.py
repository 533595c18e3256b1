"""CPU suite: the data-parallel training step (pn2_train.py, BASELINE configs[4]) at world_size 2 over gloo.

Host logic only -- the native layer is the oracle-backed fake of tests/fake_native.py.  Checks what the design
promises: every gradient lives in ONE flat bucket that is all-reduced once per step; after the step all replicas hold
identical parameters; the averaged gradient equals the mean of the per-shard gradients of a single process; the FPS
start indices of the shards are slices of one global draw."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import pn2_train


def test_bucket_aliases_every_gradient_and_zero_keeps_the_views():
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(5, 7), nn.BatchNorm1d(7), nn.ReLU(), nn.Linear(7, 3))
    bucket = pn2_train.GradBucket(net.parameters())
    assert bucket.numel == sum(p.numel() for p in net.parameters()) and bucket.attached()
    net(torch.randn(4, 5)).sum().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.equal(flat, bucket.flat) and bucket.flat.abs().sum() > 0      # backward wrote into the bucket
    bucket.zero()
    assert bucket.attached() and all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters())
    net.zero_grad(set_to_none=True)                                             # the way to break it: detected
    assert not bucket.attached()
    with pytest.raises(RuntimeError):
        pn2_train.train_step(net, pn2_train.get_loss("L2_loss", "sum"), pn2_train.make_optimizer(net), bucket,
                             torch.randn(4, 5), torch.randn(4, 3))


def test_optimizer_and_loss_follow_the_reference_defaults():
    net = nn.Linear(3, 3)
    opt = pn2_train.make_optimizer(net)
    g = opt.param_groups[0]
    assert isinstance(opt, torch.optim.Adam) and g["lr"] == 1e-4 and g["betas"] == (0.9, 0.999) and g["eps"] == 1e-8 and g["weight_decay"] == 1e-4
    a, b = torch.randn(6, 3), torch.randn(6, 3)
    assert torch.equal(pn2_train.get_loss("L2_loss", "sum")(a, b), nn.MSELoss(reduction="sum")(a, b))
    assert torch.equal(pn2_train.get_loss("L1_loss", "mean")(a, b), nn.L1Loss(reduction="mean")(a, b))


def _setup_paths():
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (os.path.join(root, "pointnet-like-pose-estimation_b200"), os.path.join(root, "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)


def _worker(rank, world, port, out_dir):
    _setup_paths()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import cloudgen
    import pn2_models
    import pn2_shard
    import pointnet2_utils as pu
    from fake_native import OracleNative
    pu._native = OracleNative()
    torch.manual_seed(0)
    model = pn2_models.get_model("rotation_ssg", 7)           # same seed -> identical replicas
    bucket = pn2_train.GradBucket(model.parameters())
    opt = pn2_train.make_optimizer(model)
    crit = pn2_train.get_loss("L2_loss", "sum")
    B = 4
    x = cloudgen.channels_first(cloudgen.with_category(cloudgen.uniform_cloud(B, 96, 11)))
    tgt = torch.randn(B, 3, generator=torch.Generator().manual_seed(5))
    lo, hi = pn2_shard.shard_bounds(B, world, rank)
    grads, losses, draws = [], [], 0
    for step in range(2):
        with pn2_shard.sharded_starts(B, world, rank, seed=100 + step) as src:
            losses.append(float(pn2_train.train_step(model, crit, opt, bucket, x[lo:hi], tgt[lo:hi])))
        draws += src.draws
        grads.append(bucket.flat.clone())
    state = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    parts = [None] * world
    dist.all_gather_object(parts, (state, grads, losses, draws))
    if rank == 0:
        torch.save(parts, os.path.join(out_dir, f"train_world{world}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _local_grads(out_dir):
    """Single process, no collective: the gradient of each 2-cloud shard at the initial parameters."""
    _setup_paths()
    import cloudgen
    import pn2_models
    import pn2_shard
    import pointnet2_utils as pu
    from fake_native import OracleNative
    prev, threads = pu._native, torch.get_num_threads()
    pu._native = OracleNative()
    torch.set_num_threads(2)            # same ATen blocking as the workers
    try:
        B, out = 4, []
        x = cloudgen.channels_first(cloudgen.with_category(cloudgen.uniform_cloud(B, 96, 11)))
        tgt = torch.randn(B, 3, generator=torch.Generator().manual_seed(5))
        for rank in range(2):
            torch.manual_seed(0)
            model = pn2_models.get_model("rotation_ssg", 7).train()
            lo, hi = pn2_shard.shard_bounds(B, 2, rank)
            with pn2_shard.sharded_starts(B, 2, rank, seed=100):
                loss = pn2_train.get_loss("L2_loss", "sum")(model(x[lo:hi]).float(), tgt[lo:hi])
            loss.backward()
            out.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]))
        return out
    finally:
        pu._native = prev
        torch.set_num_threads(threads)


@pytest.mark.timeout(600)
def test_two_replicas_take_the_same_step(tmp_path):
    port = 29700 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = torch.load(os.path.join(tmp_path, "train_world2.pt"))
    (s0, g0, l0, d0), (s1, g1, l1, d1) = parts
    assert torch.equal(s0, s1)                                   # replicas stay identical after two steps
    assert all(torch.equal(a, b) for a, b in zip(g0, g1))        # both hold the same averaged gradient
    assert l0 != l1                                              # ... computed from different shards
    assert d0 == d1 == 2 * 2                                     # one global draw per sampling layer per step (sa1, sa2; sa3 is group_all)
    local = _local_grads(str(tmp_path))
    mean = (local[0] + local[1]) / 2
    # step-0 bucket = mean of the per-shard gradients (to rounding: different thread counts reorder the BatchNorm sums)
    assert float((g0[0] - mean).abs().max() / mean.abs().max()) < 1e-4
