#!/usr/bin/env python
"""Merged timeline (CTA 0) of the warp-specialised fused SA kernel: worker warp 0, MMA warp, producer warp 0.
   python tools/tc4_trace.py sa1|sa2 [B] [first_event] [n_events]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu

which = sys.argv[1] if len(sys.argv) > 1 else "sa2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 600
count = int(sys.argv[4]) if len(sys.argv) > 4 else 120
dev = "cuda:0"
pu.set_precision("bf16")
torch.manual_seed(0)
model = pn2_models.get_model("pointnet2_cls_ssg", 7); pn2_models.randomize_bn_(model); model = model.eval().to(dev)
x = cloudgen.channels_first(cloudgen.uniform_cloud(B, 1024, 1)).to(dev)
be = pu._backend()
buf = torch.zeros(5 * 1024, dtype=torch.int64, device=dev)
with torch.no_grad():
    model(x)
    l1p, l1f = model.sa1(x, None)
    torch.cuda.synchronize()
    be.lib.pn2_debug_trace(buf.data_ptr())
    if which == "sa1": model.sa1(x, None)
    else: model.sa2(l1p, l1f)
    torch.cuda.synchronize()
    be.lib.pn2_debug_trace(None)
t = buf.cpu().view(5, 1024)
ev = []
for r, name in ((0, "W"), (4, "M"), (1, "P")):
    for v in t[r].tolist():
        v = int(v)
        if v: ev.append((v & 0xFFFFFFFFFFFF, name, (v >> 48) & 0xFFFF))
ev.sort()
def nm(role, e):
    k, l, u = e & 0xF00, (e >> 4) & 0xF, e & 0xF
    return {0x100: f"issue   l{l} u{u}", 0x200: f"deliver    u{u}", 0x400: f"acc_seen l{l} u{u}", 0x500: f"phase_done l{l} u{u}"}.get(k, hex(e))
t0 = ev[0][0]
prev = {}
for clk, role, e in ev[skip:skip + count]:
    d = clk - prev.get(role, clk); prev[role] = clk
    print(f"{clk - t0:9d} {'  ' * 'WMP'.index(role) * 12}{role} (+{d:5d}) {nm(role, e)}")
# per-role steady-state tile period
for r, name in ((0, "W"), (4, "M"), (1, "P")):
    clks = [int(v) & 0xFFFFFFFFFFFF for v in t[r].tolist() if int(v)]
    if len(clks) > 100: print(name, "events", len(clks), "cycles/event", (clks[-1] - clks[100]) / (len(clks) - 101))

