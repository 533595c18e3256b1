#!/usr/bin/env python
"""Timeline (CTA 0) of the streamed-weight fused SA kernels -- pn2_sa_tcs.cu (fp32 mode) / pn2_sa_tc.cu (bf16 ring):
compute warp 0 (C), producer warp (P), and for the bf16 pipelined last layer the pooling warp 4 (L).
   python tools/ring_trace.py fp32|bf16 sa2|sa3 [B] [first_event] [n_events]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
which = sys.argv[2] if len(sys.argv) > 2 else "sa2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 300
count = int(sys.argv[5]) if len(sys.argv) > 5 else 140
dev = "cuda:0"
pu.set_precision(prec)
torch.manual_seed(0)
model = pn2_models.get_model("pointnet2_cls_ssg", 7); pn2_models.randomize_bn_(model); model = model.eval().to(dev)
x = cloudgen.channels_first(cloudgen.uniform_cloud(B, 1024, 1)).to(dev)
be = pu._backend()
buf = torch.zeros(5 * 1024, dtype=torch.int64, device=dev)
with torch.no_grad():
    model(x)
    l1p, l1f = model.sa1(x, None)
    l2p, l2f = model.sa2(l1p, l1f)
    torch.cuda.synchronize()
    be.lib.pn2_debug_trace(buf.data_ptr())
    if which == "sa2": model.sa2(l1p, l1f)
    else: model.sa3(l2p, l2f)
    torch.cuda.synchronize()
    be.lib.pn2_debug_trace(None)
t = buf.cpu().view(5, 1024)
ev = []
for r, name in ((0, "C"), (1, "P"), (2, "L")):
    for v in t[r].tolist():
        v = int(v)
        if v: ev.append((v & 0xFFFFFFFFFFFF, name, (v >> 48) & 0xFFFF))
ev.sort()
NAMES = {0x100: "tile start", 0x200: "gather done", 0x300: "layer sync  l", 0x400: "mma issued  l", 0x500: "acc ready   l", 0x600: "epi done    l",
         0x700: "slab issued s", 0x800: "chunk full", 0x900: "slot free", 0xA00: "slab pooled s"}
def nm(e):
    k, l = e & 0xF00, e & 0xFF
    return NAMES.get(k, hex(e)) + (str(l) if NAMES.get(k, "").endswith((" l", " s")) else "")
t0 = ev[0][0]
prev = {}
for clk, role, e in ev[skip:skip + count]:
    d = clk - prev.get(role, clk); prev[role] = clk
    print(f"{clk - t0:9d} {' ' * 26 * 'CPL'.index(role)}{role} (+{d:5d}) {nm(e)}")
# per-role summary: cycles between consecutive tile starts, and where they go (compute warp 0)
c = [(clk, e) for clk, role, e in ev if role == "C"]
starts = [i for i, (_, e) in enumerate(c) if e == 0x100]
if len(starts) > 3:
    per = [c[starts[i + 1]][0] - c[starts[i]][0] for i in range(1, len(starts) - 1)]
    print("tiles traced", len(starts), "cycles per tile (this CTA): mean", sum(per) / len(per), "min", min(per), "max", max(per))
    acc = {}
    for i in range(starts[1], starts[-1]):
        k = nm(c[i + 1][1])
        acc[k] = acc.get(k, 0) + c[i + 1][0] - c[i][0]
    n = len(starts) - 2
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"   {k:18s} {v / n:9.0f} cycles per tile (time spent reaching this mark)")
