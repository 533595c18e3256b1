#!/usr/bin/env python
"""One layer of the fused training-mode BatchNorm + ReLU (+ max over K) kernels at the size of rotation_msg's widest branch
(32 clouds x 512 groups x 128 rows, 128 channels): device time and achieved bandwidth per kernel; also the target of the
ncu capture of tools/gpu_profile_bn.sh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pointnet-like-pose-estimation_b200"))
import torch, pointnet2_utils as pu

be = pu._backend()
M, C, K = 32 * 512 * 128, 128, 128
g = torch.Generator().manual_seed(0)
y = torch.randn(M, C, generator=g).cuda()
gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
dz = torch.randn(M, C, generator=g).cuda()
dzp = torch.randn(M // K, C, generator=g).cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for pooled in (0, K):
    for _ in range(2):
        z, stats, arg = be.bn_relu_train_fwd(y, gamma, beta, 1e-5, pooled)
        be.bn_relu_train_bwd(dzp if pooled else dz, y, arg, stats, pooled)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        z, stats, arg = be.bn_relu_train_fwd(y, gamma, beta, 1e-5, pooled)
    e[1].record()
    for _ in range(reps):
        be.bn_relu_train_bwd(dzp if pooled else dz, y, arg, stats, pooled)
    e[2].record(); torch.cuda.synchronize()
    gb = M * C * 4 / 1e9
    fwd_b = gb * (2 if pooled else 3)             # stats reads y; apply reads y (and writes z unless pooled)
    bwd_b = gb * (2 if pooled else 5)             # pooled: apply reads y, writes dy (stats touch one row per group); else 2 + 3
    tf, tb = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    print(f"K={pooled}: forward {tf:.3f} ms ({fwd_b / tf * 1e3:.0f} GB/s algorithmic), backward {tb:.3f} ms ({bwd_b / tb * 1e3:.0f} GB/s)")
