#!/usr/bin/env python
"""TMEM -> register read bandwidth on one SM (tcgen05.ld.32x32b.x32), bytes per clock vs number of warps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pointnet-like-pose-estimation_b200"))
import torch, _pn2_native
nat = _pn2_native.Native()
for warps in (1, 2, 4, 8, 16):
    for batch in (1, 2, 4):
        err = torch.zeros(1, device="cuda:0")
        out = torch.zeros(128, 16, device="cuda:0")
        rc = nat.lib.pn2_selftest_umma(warps, 4096, 10 + batch, err.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        print(f"warps {warps:2d} loads/wait {batch}: {out.view(-1)[0].item():8.1f} B/clk  ({out.view(-1)[1].item():.0f} cycles)")
