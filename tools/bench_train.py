#!/usr/bin/env python
"""BASELINE.json configs[4]: rotation_msg training step (N = 2048, train_rotation.py hyper-parameters) with the flat-bucket
NCCL gradient all-reduce, 32 clouds per GPU (256 at 8 GPUs), weak scaling.

    python tools/bench_train.py [--steps 10 --warmup 3 --batch 32 --model rotation_msg]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_train.py

Prints one JSON line on rank 0: clouds/s over all ranks (device time, max over ranks), ms per step and the all-reduce's
share.  Sampling, grouping and the BatchNorm / ReLU / max-pool of the shared MLP run on the product kernels, the 1x1
convolutions as cuBLAS GEMMs on the grouped rows (PN2_TRAIN_FUSED=0: the reference's op sequence on ATen, for comparison)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="clouds per GPU")
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--model", default="rotation_msg")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    import cloudgen, pn2_models, pn2_shard, pn2_train
    torch.manual_seed(0)
    model = pn2_models.get_model(args.model, 7).to(dev)
    bucket = pn2_train.GradBucket(model.parameters())
    opt = pn2_train.make_optimizer(model)
    crit = pn2_train.get_loss("L2_loss", "sum")
    B = args.batch * world
    lo, hi = pn2_shard.shard_bounds(B, world, rank)
    xyz = cloudgen.uniform_cloud(B, args.points, 1005)[lo:hi]
    x = cloudgen.channels_first(cloudgen.with_category(xyz)).to(dev)
    tgt = torch.randn(B, 3, generator=torch.Generator().manual_seed(5))[lo:hi].to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ar_ms = 0.0

    def step(i, timed):
        nonlocal ar_ms
        with pn2_shard.sharded_starts(B, world, rank, seed=100 + i, device=dev):
            loss = pn2_train.train_step(model, crit, opt, bucket, x, tgt)
        if timed and world > 1:      # all-reduce cost alone, on the same buffer (outside the step's own call)
            ev[2].record(); bucket.all_reduce_mean(); ev[3].record(); torch.cuda.synchronize()
            ar_ms += ev[2].elapsed_time(ev[3])
        return loss

    for i in range(args.warmup):
        loss0 = step(i, False)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = 0.0
    for i in range(args.steps):
        ev[0].record()
        loss = step(args.warmup + i, False)
        ev[1].record(); torch.cuda.synchronize()
        t += ev[0].elapsed_time(ev[1])
    for i in range(3):
        step(0, True)
    tt = torch.tensor([t], device=dev)
    if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(tt) / args.steps
        print(json.dumps({"metric": f"{args.model} training step clouds/s (N={args.points})", "value": B / ms * 1e3, "unit": "clouds/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "weak",
                          "config": {"workload": f"{args.model} train step, {args.batch} clouds per GPU, Adam lr 1e-4 wd 1e-4, MSE sum",
                                     "grad_bucket_elems": bucket.numel, "grad_bucket_mb": bucket.numel * 4 / 1e6},
                          "allreduce_ms": ar_ms / 3 if world > 1 else 0.0, "loss_first": float(loss0), "loss_last": float(loss),
                          "mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
