#!/usr/bin/env python
"""BASELINE.json configs[4]: rotation_msg training step (N = 2048, train_rotation.py hyper-parameters) with the flat-bucket
NCCL gradient all-reduce, 32 clouds per GPU (256 at 8 GPUs), weak scaling.

Driven by ``python bench.py --config 4 [--gpus N --steps K --warmup W --batch 32]`` (under torchrun for N > 1); ``run`` /
``run_reference`` below are its two arms.

A step = train_rotation.py:107-128 on this rank's shard: zero the gradient bucket, forward in training mode (batch-statistics
BatchNorm), MSE(sum) loss, backward into the flat bucket, ONE NCCL all-reduce of the bucket, Adam.  `value` keeps the shard's
input resident; `e2e` copies the shard's input and target from pinned host memory and reads the loss back every step.
Prints clouds/s over all ranks (device time, max over ranks), the all-reduce's own time and the kernel shares by name."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    p = os.path.join(ROOT, p)
    if p not in sys.path:
        sys.path.insert(0, p)
import torch

MODEL = "rotation_msg"
N_POINTS = 2048
METRIC = "rotation_msg training step clouds/s (N=2048)"
CPU_SAMPLE = 4


def config(batch, world):
    return {"workload": f"configs[4]: {MODEL} train step (forward with batch-statistics BatchNorm, MSE sum, backward, Adam lr 1e-4 wd 1e-4), "
                        f"N={N_POINTS}, {batch} clouds per GPU",
            "batch_per_gpu": batch, "global_batch": batch * world,
            "parallelism": f"data-parallel x{world}, local BatchNorm statistics, one flat-bucket NCCL all-reduce per step",
            "reference_arm_sample": f"{CPU_SAMPLE}-cloud train steps of the same model on the CPU port"}


def run(args, R):
    import bench
    import cloudgen
    import pn2_models
    import pn2_shard
    import pn2_train
    import pointnet2_utils as pu
    dev = R.dev
    batch = args.batch or 32
    torch.manual_seed(0)
    model = pn2_models.get_model(MODEL, 7).to(dev)
    bucket = pn2_train.GradBucket(model.parameters())
    opt = pn2_train.make_optimizer(model)
    crit = pn2_train.get_loss("L2_loss", "sum")
    B = batch * R.world
    lo, hi = pn2_shard.shard_bounds(B, R.world, R.rank)
    xyz = cloudgen.uniform_cloud(B, N_POINTS, 1005)[lo:hi]
    x_h = cloudgen.channels_first(cloudgen.with_category(xyz)).pin_memory()
    t_h = torch.randn(B, 3, generator=torch.Generator().manual_seed(5))[lo:hi].contiguous().pin_memory()
    x_d, t_d = x_h.to(dev), t_h.to(dev)
    loss_h = torch.zeros(1).pin_memory()
    be = pu._backend()
    K, W = args.steps, max(args.warmup, 3)
    step_no = [0]

    def step(e2e):
        x, t = (x_h.to(dev, non_blocking=True), t_h.to(dev, non_blocking=True)) if e2e else (x_d, t_d)
        with pn2_shard.sharded_starts(B, R.world, R.rank, seed=100 + step_no[0], device=dev):
            loss = pn2_train.train_step(model, crit, opt, bucket, x, t)
        step_no[0] += 1
        if e2e:
            loss_h.copy_(loss.reshape(1), non_blocking=True)
        return loss

    clocks = bench.ClockSampler(R.local) if R.rank == 0 else None
    for _ in range(W):
        loss0 = step(False)
    R.barrier()
    l0 = be.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        loss = step(False)
    ev1.record()
    R.barrier()
    launches = be.launches - l0
    ms = R.reduce_max(ev0.elapsed_time(ev1))
    ev0.record()
    for _ in range(K):
        step(True)
    ev1.record()
    R.barrier()
    ms_e2e = R.reduce_max(ev0.elapsed_time(ev1))
    ar_ms = 0.0
    if R.world > 1:                    # the all-reduce alone, on the same bucket
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(5):
            R.barrier()
            e2.record(); bucket.all_reduce_mean(); e3.record()
            torch.cuda.synchronize(dev)
            if i >= 2:
                ar_ms += e2.elapsed_time(e3) / 3
    # kernel shares of one step (torch profiler, outside the timed regions)
    shares = None
    if not args.quick:
        # every rank takes the step (it contains the all-reduce); only rank 0 records it
        if R.rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step(False)
                torch.cuda.synchronize(dev)
        else:
            step(False)
            torch.cuda.synchronize(dev)
    if R.rank == 0 and not args.quick:
        rows = sorted(((e.key, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[1])
        tot = sum(r[1] for r in rows) or 1.0
        shares = {"kernel_ms_total": round(tot, 3), "top": [{"name": k[:80], "ms": round(v, 3), "share": round(v / tot, 3)} for k, v in rows[:12]],
                  "library_gemm_ms": round(sum(v for k, v in rows if any(s in k.lower() for s in ("cutlass", "cublas", "gemm", "xmma", "nvjet")) and "pn2" not in k.lower()), 3)}
    clk = clocks.stop() if clocks else None
    if R.rank != 0:
        R.finish()
        return
    cpu = None
    if R.world == 1 and not args.no_cpu_baseline and not args.quick:
        v, threads = cpu_train_throughput(1)
        cpu = {"value": v, "unit": "clouds/s", "cores": threads, "kind": "port",
               "sample": f"one {CPU_SAMPLE}-cloud train step of {MODEL} (N={N_POINTS}) on the torch CPU port of the reference path"}
    fwd_flops = bench.model_flops_per_cloud(model, N_POINTS)
    value = B * K / (1e-3 * ms)
    pk = bench.peaks()
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": R.world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (GEMMs in TF32, the policy of the reference's cuDNN convolutions)",
        "data": "synthetic", "config": config(batch, R.world),
        "e2e": {"value": B * K / (1e-3 * ms_e2e), "unit": "clouds/s", "h2d_bytes_per_step": (x_h.numel() + t_h.numel()) * 4 * R.world,
                "d2h_bytes_per_step": 4 * R.world},
        "gpu_launches": launches, "clocks": clk,
        "roofline": {"kernel": "whole step (forward + backward ~ 3x forward FLOPs)", "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": None, "traffic": None,
                     "note": "the step is bound by the BatchNorm / ReLU / max-pool passes over the grouped activations (HBM), see kernel_shares; "
                             f"model TFLOP/s = {3 * value / R.world * fwd_flops / 1e12:.1f} per GPU"},
        "allreduce_ms": ar_ms, "grad_bucket_mb": bucket.numel * 4 / 1e6, "loss_first": float(loss0), "loss_last": float(loss),
        "mem_gb": torch.cuda.max_memory_allocated() / 1e9, "kernel_shares": shares, "cpu_baseline": cpu}))
    R.finish()


def cpu_train_throughput(steps):
    import bench
    import cloudgen
    import pn2_models
    import pn2_train
    P = bench.port()
    threads = bench.cpu_threads()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = pn2_models.get_model(MODEL, 7, sa_impl=P).train()
    opt = pn2_train.make_optimizer(model)
    crit = pn2_train.get_loss("L2_loss", "sum")
    x = cloudgen.channels_first(cloudgen.with_category(cloudgen.uniform_cloud(CPU_SAMPLE, N_POINTS, 1005)))
    tgt = torch.randn(CPU_SAMPLE, 3, generator=torch.Generator().manual_seed(5))

    def one():
        opt.zero_grad()
        loss = crit(model(x), tgt)
        loss.backward()
        opt.step()

    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    return CPU_SAMPLE * steps / (time.perf_counter() - t0), threads


def run_reference(args, R):
    if R.rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    v, threads = cpu_train_throughput(steps)
    sample = f"{steps} train step(s) of {CPU_SAMPLE} clouds (N={N_POINTS}), torch CPU port of the reference path + autograd + Adam, {threads} host threads"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "clouds/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
                      "ms_per_step": 1e3 * CPU_SAMPLE / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": config(args.batch or 32, R.world),
                      "cpu_baseline": {"value": v, "unit": "clouds/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


if __name__ == "__main__":
    sys.argv[1:1] = ["--config", "4"]
    sys.path.insert(0, ROOT)
    import bench
    bench.main()
