#!/usr/bin/env python
"""BASELINE.json configs[3]: the full pose pipeline -- rotation / translation / sign heads, SSG and MSG, plus the v1 width
head -- on B = 1024 clouds (N = 1024), batch-sharded over the GPUs of one box (strong scaling: 1024 / world clouds per GPU).

    python tools/bench_pipeline.py [--batch 1024 --steps 10 --warmup 3 --precision bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_pipeline.py

A step = for this rank's shard: host -> device copy of the raw xyz (pinned) and the category ids, device-side input
preparation (provider.normalization + splice_torch + transpose, pn2_preprocess.prepare_pose_input), the seven heads one
after the other like mutilthreading/predict_test.py runs them, device -> host copy of every head's output.  No collective:
the shards are independent; FPS start indices are drawn once for the global batch per sampling layer and sliced per
rank (pn2_shard), so the concatenated outputs do not depend on the number of GPUs.  Prints one JSON line on rank 0:
pipeline clouds/s (clouds that went through ALL seven heads; device time, max over ranks) and per-head milliseconds."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import torch.distributed as dist

HEADS = ["rotation_ssg", "rotation_msg", "translation_ssg", "translation_msg", "sign_ssg", "sign_msg", "v1/width"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024, help="global batch")
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--graph", type=int, default=0, help="1: every head replayed as one CUDA graph (pn2_graph.GraphedForward)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    import cloudgen, pn2_models, pn2_preprocess, pn2_shard, pn2_v1, pointnet2_utils as pu
    pu.set_precision(args.precision)
    B = args.batch
    lo, hi = pn2_shard.shard_bounds(B, world, rank)
    models = {}
    for name in HEADS:
        torch.manual_seed(0)
        m = pn2_v1.get_v1_model(name[3:], 7) if name.startswith("v1/") else pn2_models.get_model(name, 7)
        pn2_models.randomize_bn_(m)
        models[name] = m.eval().to(dev)
    g = torch.Generator().manual_seed(1004)
    raw = (torch.rand(B, args.points, 3, generator=g) * 2 - 1 + torch.randn(B, 1, 3, generator=g))[lo:hi].contiguous().pin_memory()
    normals = torch.randn(B, args.points, 3, generator=g)[lo:hi].contiguous().pin_memory()       # width.py: xyz + normals + one-hot
    category = (torch.arange(B) % 7)[lo:hi].contiguous().pin_memory()
    outs_host = {}
    ev = {n: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for n in HEADS}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    head_ms = {n: 0.0 for n in HEADS}

    graphs = {}
    if args.graph:
        from pn2_graph import GraphedForward
        src = pn2_shard.GlobalStarts(B, world, rank, seed=7)             # global draws, this rank's slice: as in the eager path
        xyz0 = raw.to(dev)
        x0 = pn2_preprocess.prepare_pose_input(xyz0, category.to(dev), 7)
        ex = {"x": x0, "mean": xyz0[:, :3, :].mean(1), "xw": torch.cat([x0[:, :3], normals.to(dev).permute(0, 2, 1), x0[:, 3:]], 1)}
        for n in HEADS:
            a = (ex["xw"],) if n == "v1/width" else ((ex["x"], ex["mean"]) if n.startswith("translation") else (ex["x"],))
            graphs[n] = GraphedForward(models[n], *a, start_source=src)

    def step(i, timed):
        xyz = raw.to(dev, non_blocking=True)
        nrm = normals.to(dev, non_blocking=True)
        cat = category.to(dev, non_blocking=True)
        x = pn2_preprocess.prepare_pose_input(xyz, cat, 7)                          # [b, 10, N]
        mean = xyz[:, :3, :].mean(1)                                                 # train_translation.py:113
        xw = torch.cat([x[:, :3], nrm.permute(0, 2, 1), x[:, 3:]], 1)               # [b, 13, N]
        with torch.no_grad(), pn2_shard.sharded_starts(B, world, rank, seed=42 + i, device=dev):
            for n in HEADS:
                ev[n][0].record()
                run = graphs[n] if args.graph else models[n]
                if n == "v1/width":
                    out = run(xw)
                elif n.startswith("translation"):
                    out = run(x, mean)
                else:
                    out = run(x)
                out = out if isinstance(out, tuple) else (out,)
                outs_host[n] = [o.to("cpu", non_blocking=True) for o in out]
                ev[n][1].record()
        if timed:
            torch.cuda.synchronize()
            for n in HEADS:
                head_ms[n] += ev[n][0].elapsed_time(ev[n][1])

    for i in range(args.warmup):
        step(i, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = 0.0
    for i in range(args.steps):
        e0.record()
        step(args.warmup + i, True)
        e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    tt = torch.tensor([t], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(tt) / args.steps
        print(json.dumps({"metric": f"full pose pipeline clouds/s (7 heads, N={args.points})", "value": B / ms * 1e3, "unit": "clouds/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong",
                          "dtype": args.precision, "data": "synthetic",
                          "config": {"workload": f"configs[3]: {', '.join(HEADS)} on B={B} clouds, {hi - lo} per GPU, host buffers in and out",
                                     "parallelism": f"batch-sharded x{world}, no collective",
                                     "cuda_graphs": bool(args.graph)},
                          "head_ms_rank0": {n: round(head_ms[n] / args.steps, 3) for n in HEADS},
                          "h2d_bytes_per_step": (raw.numel() + normals.numel()) * 4 + category.numel() * 8,
                          "d2h_bytes_per_step": sum(o.numel() * o.element_size() for v in outs_host.values() for o in v)}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
