#!/usr/bin/env python
"""Forward throughput of the eight PointNet++ models and the four PointNet-v1-like pose heads (random weights, synthetic
clouds, N = 1024) on one GPU: clouds/s in bf16 and fp32 mode.  Secondary numbers for profiles/ (BASELINE configs[3]: the
full pose pipeline = rotation / translation / sign SSG + MSG heads + the v1 width head); the contract benchmark is bench.py."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pn2_v1, pointnet2_utils as pu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda:0"
res = {}
for name in list(pn2_models.ARCH) + [f"v1/{n}" for n in pn2_v1.V1_ARCH]:
    torch.manual_seed(0)
    v1 = name.startswith("v1/")
    model = pn2_v1.get_v1_model(name[3:], 7) if v1 else pn2_models.get_model(name, 7)
    pn2_models.randomize_bn_(model); model = model.eval().to(dev)
    xyz = cloudgen.uniform_cloud(B, 1024, 5)
    pose = not name.startswith("pointnet2_cls")
    pts = cloudgen.with_category(xyz) if pose else xyz
    if name == "v1/width":      # width.py: 6 + num_category input channels
        pts = torch.cat([xyz, torch.randn(B, 1024, 3, generator=torch.Generator().manual_seed(6)), pts[..., 3:]], -1)
    x = cloudgen.channels_first(pts).to(dev)
    args = (x, xyz[:, :3, :].mean(1).to(dev)) if "translation" in name else (x,)
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
