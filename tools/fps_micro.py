#!/usr/bin/env python
"""configs[2] FPS microbench (B = 16, N = 8192 -> 2048): median device time of 30 launches in us (bench.device_time_us: no host sync between launches, L2 flushed).  PN2_FPS_CLUSTER selects the variant
(0 = one CTA per cloud, 1 = 8-CTA cluster with per-CTA slots, 2 = 8-CTA cluster with per-warp slots, 4 = 4-CTA cluster with per-warp slots; default: per-warp slots, 8 CTAs up to four clouds, 4 beyond)."""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pointnet-like-pose-estimation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import cloudgen
import pointnet2_utils as pu
from bench import device_time_us
be = pu._backend()
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device="cuda")
for B, N, S in [(16, 8192, 2048), (16, 4096, 1024), (4, 8192, 2048), (12, 8192, 2048), (14, 8192, 2048), (32, 8192, 2048), (48, 8192, 2048), (64, 8192, 2048), (74, 8192, 2048), (64, 4096, 1024)]:
    x = cloudgen.uniform_cloud(B, N, 1003).cuda()
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(42), dtype=torch.long).cuda()
    idx, _, _ = be.fps(x, start, S, want_idx64=True, want_new_points=True)
    us = device_time_us(lambda: be.fps(x, start, S, want_idx64=True, want_new_points=True), flush=flush)
    print(f"PN2_FPS_CLUSTER={os.environ.get('PN2_FPS_CLUSTER', '(default)')} B={B} N={N} S={S}: {us:.1f} us  checksum {int(idx.sum())}")
