#!/usr/bin/env python
"""Forward throughput of the eight PointNet++ models (random weights, synthetic clouds, N = 1024) on one GPU:
clouds/s in bf16 and fp32 mode.  Secondary numbers for profiles/; the contract benchmark is bench.py."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda:0"
res = {}
for name in pn2_models.ARCH:
    torch.manual_seed(0)
    model = pn2_models.get_model(name, 7); pn2_models.randomize_bn_(model); model = model.eval().to(dev)
    xyz = cloudgen.uniform_cloud(B, 1024, 5)
    pose = not name.startswith("pointnet2_cls")
    x = cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz).to(dev)
    args = (x, xyz[:, :3, :].mean(1).to(dev)) if name.startswith("translation") else (x,)
    for prec in ("bf16", "fp32"):
        pu.set_precision(prec)
        with torch.no_grad():
            for _ in range(3):
                model(*args)
            torch.cuda.synchronize()
            n = 5 if prec == "bf16" else 2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                model(*args)
            e1.record(); torch.cuda.synchronize()
        res[f"{name}/{prec}"] = round(B * n / (1e-3 * e0.elapsed_time(e1)))
        print(f"{name:20s} {prec}: {res[f'{name}/{prec}']:>9d} clouds/s  (B={B})", flush=True)
print(json.dumps({"batch": B, "clouds_per_s": res}))
