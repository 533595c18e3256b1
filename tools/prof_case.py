#!/usr/bin/env python
"""A short forward for runs under ncu: python tools/prof_case.py <precision> <batch> [model] [n_forwards]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu
prec, B = sys.argv[1], int(sys.argv[2])
name = sys.argv[3] if len(sys.argv) > 3 else "pointnet2_cls_ssg"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
pu.set_precision(prec)
torch.manual_seed(0)
m = pn2_models.get_model(name, 7); pn2_models.randomize_bn_(m); m = m.eval().cuda()
xyz = cloudgen.uniform_cloud(B, 1024, 1001)
pose = not name.startswith("pointnet2_cls")
x = cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz).cuda()
args = (x, xyz[:, :3, :].mean(1).cuda()) if name.startswith("translation") else (x,)
with torch.no_grad():
    for _ in range(n):
        out = m(*args)
torch.cuda.synchronize()
print("ok", float((out[0] if isinstance(out, tuple) else out).sum()))
