#!/usr/bin/env python
"""configs[2] ball-query microbench (B = 16, N = 8192, 2048 centres, r = 0.2, K = 32): median device time of 30 launches in us (bench.device_time_us)."""
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
for B, N, S, K, r in [(16, 8192, 2048, 32, 0.2), (8, 8192, 2048, 32, 0.2), (32, 8192, 2048, 32, 0.2), (16, 4096, 1024, 32, 0.2),
                      (2048, 1024, 512, 32, 0.2), (2048, 512, 128, 64, 0.4), (512, 1024, 512, 32, 0.2), (512, 1024, 512, 16, 0.1), (24, 1024, 512, 32, 0.2)]:
    x = cloudgen.uniform_cloud(B, N, 1003).cuda()
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(42), dtype=torch.long).cuda()
    _, _, ctr = be.fps(x, start, S, want_idx64=True, want_new_points=True)
    i64, _ = be.ball_query(ctr, x, pu._radius_sq(r), K, want_idx64=True, want_idx32=False)
    us = device_time_us(lambda: be.ball_query(ctr, x, pu._radius_sq(r), K, want_idx64=True, want_idx32=False), flush=flush)
    print(f"PN2_BQ_CPW={os.environ.get('PN2_BQ_CPW', '(default)')} PN2_BQ_CELLS={os.environ.get('PN2_BQ_CELLS', '(default)')} B={B} N={N} S={S} K={K} r={r}: {us:.1f} us  checksum {int(i64.sum())}")
