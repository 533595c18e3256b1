#!/usr/bin/env python
"""tcgen05.mma (M=128, K=16, bf16) cycles per instruction on one SM for several N and operand layouts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pointnet-like-pose-estimation_b200"))
import torch, _pn2_native
nat = _pn2_native.Native()
def run(n, variant, count=512):
    err = torch.zeros(1, device="cuda:0"); out = torch.zeros(16, device="cuda:0")
    rc = nat.lib.pn2_selftest_umma(n, count, variant, err.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        print('N', n, 'variant', variant, 'error', nat.lib.pn2_last_error_string()); return -1, -1
    torch.cuda.synchronize()
    return out[0].item(), out[1].item()
for n in (32, 64, 128, 256):
    for name, v in (("A,B no-swizzle", 100), ("A sw128", 110), ("A sw128, two accumulators", 160), ("unrolled, fixed descriptors", 300)) + ((("B sw128", 101), ("A,B sw128", 111)) if n == 128 else ()):
        total, issue = run(n, v)
        print(f"N={n:3d} {name:28s}: {total:6.1f} cycles/MMA to completion, {issue:6.1f} issuing   (tensor floor {n/2:.0f})")
