#!/usr/bin/env python
"""Per-kernel device time of one forward of any model of the zoo (CUDA events around every C-ABI launch).
   python tools/kernel_table.py pointnet2_cls_msg [B] [bf16|fp32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu

name = sys.argv[1] if len(sys.argv) > 1 else "pointnet2_cls_msg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
pu.set_precision(sys.argv[3] if len(sys.argv) > 3 else "bf16")
dev = "cuda:0"
torch.manual_seed(0)
model = pn2_models.get_model(name, 7); pn2_models.randomize_bn_(model); model = model.eval().to(dev)
xyz = cloudgen.uniform_cloud(B, 1024, 5)
pose = not name.startswith("pointnet2_cls")
x = cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz).to(dev)
args = (x, xyz[:, :3, :].mean(1).to(dev)) if name.startswith("translation") else (x,)
be = pu._backend()
with torch.no_grad():
    for _ in range(3): model(*args)
    torch.cuda.synchronize()
    be.trace = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model(*args); e1.record(); torch.cuda.synchronize()
tr, be.trace = be.trace, None
tot = e0.elapsed_time(e1)
print(f"{name} B={B} {pu.get_precision()}: forward {tot:.3f} ms = {B / tot * 1e3:.0f} clouds/s")
acc = 0.0
for nm, meta, s, e in tr:
    ms = s.elapsed_time(e); acc += ms
    d = f"S={meta.get('S')} K={meta.get('K')} N={meta.get('N')}" + (f" dims={'x'.join(map(str, meta['dims']))}" if 'dims' in meta else "")
    fl = ""
    if nm == "sa_fused":
        dims = list(meta["dims"]); f = 2 * meta["B"] * meta["S"] * meta["K"] * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        fl = f"  {f / ms / 1e9:7.1f} TFLOP/s"
    print(f"  {nm:12s} {ms:8.3f} ms  {d}{fl}")
print(f"  kernels {acc:.3f} ms of {tot:.3f}")
