#!/usr/bin/env python
"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities): which kernels the configs[4] step spends
its time in.  python tools/profile_train.py [--batch 32 --points 2048 --model rotation_msg]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--points", type=int, default=2048)
ap.add_argument("--model", default="rotation_msg")
ap.add_argument("--rows", type=int, default=30)
args = ap.parse_args()
import cloudgen, pn2_models, pn2_train
dev = "cuda:0"
torch.manual_seed(0)
model = pn2_models.get_model(args.model, 7).to(dev)
bucket = pn2_train.GradBucket(model.parameters())
opt = pn2_train.make_optimizer(model)
crit = pn2_train.get_loss("L2_loss", "sum")
xyz = cloudgen.uniform_cloud(args.batch, args.points, 1005)
x = cloudgen.channels_first(cloudgen.with_category(xyz)).to(dev)
tgt = torch.randn(args.batch, 3, generator=torch.Generator().manual_seed(5)).to(dev)
for _ in range(3):
    pn2_train.train_step(model, crit, opt, bucket, x, tgt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        pn2_train.train_step(model, crit, opt, bucket, x, tgt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=90))
