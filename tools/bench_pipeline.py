#!/usr/bin/env python
"""BASELINE.json configs[3]: the full pose pipeline -- rotation / translation / sign heads, SSG and MSG, plus the v1 width
head -- on B = 1024 clouds (N = 1024), batch-sharded over the GPUs of one box (strong scaling: 1024 / world clouds per GPU).

Driven by ``python bench.py --config 3 [--gpus N --steps K --warmup W --graph {0,1,2} --precision bf16|fp32]`` (under torchrun for
N > 1); ``run`` / ``run_reference`` below are its two arms.

A step = for this rank's shard: host -> device copy of the raw xyz, normals (pinned) and the category ids, device-side input
preparation (provider.normalization + splice_torch + transpose, one launch), the seven heads (pn2_pipeline.PosePipeline:
--graph 2 = one CUDA graph per head on one stream per head, 1 = the graphs back to back, 0 = eager), device -> host copy of
every head's output.  No collective: the shards are independent; FPS start indices are drawn once for the global batch per
sampling layer and sliced per rank (pn2_shard.GlobalStarts), so the concatenated outputs do not depend on the number of GPUs.
value = clouds that went through ALL seven heads per second (device time, max over ranks); `value` keeps the raw inputs
resident on the device, `e2e` includes the copies."""
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

METRIC = "full pose pipeline clouds/s (6 PointNet++ heads + v1 width head, N=1024)"
CPU_SAMPLE = 16


def config(B, world, graph=None):
    return {"workload": f"configs[3]: rotation/translation/sign SSG+MSG heads + v1 width head on B={B} clouds (N=1024), host buffers in and out, "
                        "device-side input preparation",
            "global_batch": B, "parallelism": f"batch-sharded x{world} ({B // max(world, 1)} clouds per GPU), no collective",
            "reference_arm_sample": f"{CPU_SAMPLE}-cloud slices through the same seven heads on the CPU port"}


def make_raw(B, n_points):
    g = torch.Generator().manual_seed(1004)
    raw = torch.rand(B, n_points, 3, generator=g) * 2 - 1 + torch.randn(B, 1, 3, generator=g)
    normals = torch.randn(B, n_points, 3, generator=g)               # width.py: xyz + normals + one-hot
    category = torch.arange(B) % 7
    return raw, normals, category


def run(args, R):
    import bench
    import pn2_pipeline
    import pn2_shard
    import pointnet2_utils as pu
    dev = R.dev
    pu.set_precision(args.precision)
    B = args.batch or 1024
    n_points = 1024
    lo, hi = pn2_shard.shard_bounds(B, R.world, R.rank)
    heads = pn2_pipeline.build_heads(dev)
    raw, normals, category = make_raw(B, n_points)
    raw_h, nrm_h, cat_h = (t[lo:hi].contiguous().pin_memory() for t in (raw, normals, category))
    raw_d, nrm_d, cat_d = raw_h.to(dev), nrm_h.to(dev), cat_h.to(dev)
    mode = {0: "eager", 1: "graphs", 2: "streams"}[args.graph]
    src = pn2_shard.GlobalStarts(B, R.world, R.rank, seed=7)          # global draws, this rank's slice
    pipe = pn2_pipeline.PosePipeline(heads, mode=mode, start_source=src)
    be = pu._backend()
    outs_host = {}

    def step(e2e):
        if e2e:
            xyz, nrm, cat = raw_h.to(dev, non_blocking=True), nrm_h.to(dev, non_blocking=True), cat_h.to(dev, non_blocking=True)
        else:
            xyz, nrm, cat = raw_d, nrm_d, cat_d
        outs = pipe(xyz, nrm, cat)
        if e2e:
            for n, o in outs.items():
                if n not in outs_host:
                    outs_host[n] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in o]
                for h, t in zip(outs_host[n], o):
                    h.copy_(t, non_blocking=True)
        return outs

    K, W = args.steps, max(args.warmup, 3)
    clocks = bench.ClockSampler(R.local) if R.rank == 0 else None
    for _ in range(W):
        step(True)
    R.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        step(False)
    ev1.record()
    R.barrier()
    ms = R.reduce_max(ev0.elapsed_time(ev1))
    ev0.record()
    for _ in range(K):
        step(True)
    ev1.record()
    R.barrier()
    ms_e2e = R.reduce_max(ev0.elapsed_time(ev1))
    # per-head device time, heads run alone one after the other (eager, rank 0's shard)
    head_ms = {}
    eager = pn2_pipeline.PosePipeline(heads, mode="eager", start_source=src)
    l0 = be.launches
    eager(raw_d, nrm_d, cat_d)
    launches = (be.launches - l0) * K                 # the graphs replay exactly the eager sequence's kernels
    x, mean, xw = eager.prepare(raw_d, nrm_d, cat_d)
    for n in eager.names:
        one = pn2_pipeline.PosePipeline({n: heads[n]}, mode="eager", start_source=src)
        for _ in range(2):
            one.forward_prepared(x, mean, xw)
        torch.cuda.synchronize(dev)
        ev0.record()
        for _ in range(5):
            one.forward_prepared(x, mean, xw)
        ev1.record()
        torch.cuda.synchronize(dev)
        head_ms[n] = round(ev0.elapsed_time(ev1) / 5, 3)
    clk = clocks.stop() if clocks else None
    if R.rank != 0:
        R.finish()
        return
    flops = sum(bench.model_flops_per_cloud(heads[n], n_points) for n in heads if not n.startswith("v1/"))
    flops += 2.0 * n_points * (13 * 64 + 64 * 64 * 2 + 64 * 128 + 128 * 1024) + 2.0 * (1024 * 512 + 512 * 256 + 256)
    pk = bench.peaks()
    value = B * K / (1e-3 * ms)
    cpu = None
    if R.world == 1 and not args.no_cpu_baseline:
        v, threads = cpu_pipeline_throughput(1)
        cpu = {"value": v, "unit": "clouds/s", "cores": threads, "kind": "port",
               "sample": f"one pass of {CPU_SAMPLE} clouds through the seven heads on the torch CPU port of the reference path"}
    h2d = (raw_h.numel() + nrm_h.numel()) * 4 + cat_h.numel() * 8
    d2h = sum(t.numel() * t.element_size() for v in outs_host.values() for t in v)
    peak, src_s = bench.pick_peak(pk, 1e-3 * ms, clk)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": R.world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic", "config": config(B, R.world), "precision_mode": args.precision, "schedule": mode,
        "e2e": {"value": B * K / (1e-3 * ms_e2e), "unit": "clouds/s", "h2d_bytes_per_step": h2d * R.world, "d2h_bytes_per_step": d2h * R.world},
        "gpu_launches": launches, "clocks": clk,
        "roofline": {"kernel": "whole pipeline (seven heads)", "bound": "tensor", "achieved": value / R.world * flops / 1e12, "peak": peak,
                     "unit": "TFLOP/s", "frac": value / R.world * flops / 1e12 / peak, "traffic": None, "peak_source": src_s,
                     "algorithmic_flops_per_cloud": flops},
        "head_ms_alone_rank0": head_ms, "head_ms_sum": round(sum(head_ms.values()), 3), "cpu_baseline": cpu}))
    R.finish()


def cpu_pipeline_throughput(steps):
    import bench
    import pn2_models
    import pn2_v1
    P = bench.port()
    threads = bench.cpu_threads()
    torch.set_num_threads(threads)
    import cloudgen
    raw, normals, category = make_raw(CPU_SAMPLE, 1024)
    xyz = cloudgen.normalize(raw)
    x10 = cloudgen.with_category(xyz)                                    # [B,N,10], category b % 7
    x = cloudgen.channels_first(x10)
    xw = cloudgen.channels_first(torch.cat([xyz, normals, x10[..., 3:]], -1))
    mean = raw[:, :3, :].mean(1)
    models = {}
    import pn2_pipeline
    for name in pn2_pipeline.HEADS:
        torch.manual_seed(0)
        if name.startswith("v1/"):
            m = pn2_v1.get_v1_model(name[3:], 7)
        else:
            m = pn2_models.get_model(name, 7, sa_impl=P)
        pn2_models.randomize_bn_(m)
        models[name] = m.eval()

    def one_pass():
        with torch.no_grad():
            for name, m in models.items():
                if name.startswith("v1/"):
                    P.v1_forward(m, xw)
                elif name.startswith("translation"):
                    m(x, mean)
                else:
                    m(x)

    one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    return CPU_SAMPLE * steps / (time.perf_counter() - t0), threads


def run_reference(args, R):
    if R.rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    v, threads = cpu_pipeline_throughput(steps)
    B = args.batch or 1024
    sample = f"{steps} pass(es) of {CPU_SAMPLE} clouds through the seven heads, torch CPU port of the reference path, {threads} host threads"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "clouds/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
                      "ms_per_step": 1e3 * CPU_SAMPLE / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": config(B, R.world),
                      "cpu_baseline": {"value": v, "unit": "clouds/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


if __name__ == "__main__":
    sys.argv[1:1] = ["--config", "3"]
    sys.path.insert(0, ROOT)
    import bench
    bench.main()
