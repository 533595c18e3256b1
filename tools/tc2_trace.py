#!/usr/bin/env python
"""Stage timeline of the pipelined fused SA kernel (CTA 0) for one cls_ssg layer: prints per-stage durations.
   python tools/tc2_trace.py sa1|sa2 [B]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu

which = sys.argv[1] if len(sys.argv) > 1 else "sa2"
COOP = which == "sa2" and os.environ.get("PN2_TC_COOP", "1") != "0"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = "cuda:0"
pu.set_precision("bf16")
torch.manual_seed(0)
model = pn2_models.get_model("pointnet2_cls_ssg", 7); pn2_models.randomize_bn_(model); model = model.eval().to(dev)
x = cloudgen.channels_first(cloudgen.uniform_cloud(B, 1024, 1)).to(dev)
be = pu._backend()
buf = torch.zeros(5 * 1024, dtype=torch.int64, device=dev)
with torch.no_grad():
    model(x)                                   # warm-up (packs weights)
    l1p, l1f = model.sa1(x, None)
    torch.cuda.synchronize()
    be.lib.pn2_debug_trace(buf.data_ptr())
    if which == "sa1":
        model.sa1(x, None)
    else:
        model.sa2(l1p, l1f)
    torch.cuda.synchronize()
    be.lib.pn2_debug_trace(None)
t = buf.cpu().view(5, 1024)
names = {0x300: "gather_done", 0x600: "pool_done"} if not (len(sys.argv) > 1 and sys.argv[1] == 'sa2' and os.environ.get('PN2_TC_COOP', '1') != '0') else {}
def decode(v):
    ev, clk = (v >> 48) & 0xFFFF, v & 0xFFFFFFFFFFFF
    return ev, clk
roles = []
for r in range(5):
    evs = [decode(int(v)) for v in t[r].tolist() if int(v) != 0]
    roles.append(evs)
t0 = min(e[0][1] for e in roles if e)
def name(ev):
    if COOP:
        if 0x400 <= ev < 0x410: return f"acc_ready l{(ev & 0xF) >> 1} u{ev & 1}"
        if 0x500 <= ev < 0x510: return f"epi_done  l{(ev & 0xF) >> 1} u{ev & 1}"
        if 0x600 <= ev < 0x610: return f"pool+gather done u{ev & 1}"
    if ev in names: return names[ev]
    k, l, u = ev & 0xF00, (ev >> 4) & 0xF, ev & 0xF
    if 0x610 <= ev < 0x640: return {0x610: "twin_issued", 0x620: "pool_ld", 0x630: "pool_blk"}[ev & 0xFF0] + (f" {ev & 0xF}" if ev != 0x610 else "")
    if 0x700 <= ev < 0x710: return f"mma_issued l{ev & 0xF}"
    if 0x710 <= ev < 0x720: return f"unit_barrier l{ev & 0xF}"
    if 0x520 <= ev < 0x550: return {0x520: "epi_ldwait", 0x530: "epi_blocks", 0x540: "epi_tcfence"}[ev & 0xFF0] + f" l{ev & 0xF}"
    return {0x100: f"mma_wait_done l{l} u{u}", 0x200: f"mma_issued   l{l} u{u}", 0x400: f"acc_ready l{ev & 0xF}", 0x500: f"epi_done  l{ev & 0xF}"}.get(k, hex(ev))
for r, evs in enumerate(roles):
    if not evs: continue
    print(f"--- role {r} ({'MMA warp' if r == 4 else 'unit ' + str(r)}): {len(evs)} events")
    prev = None
    lo, hi = (80, 80 + 70)
    for ev, clk in evs[lo:hi]:
        print(f"   {clk - t0:9d}  (+{(clk - prev) if prev is not None else 0:6d})  {name(ev)}")
        prev = clk
# steady-state averages per unit: time between consecutive pool_done
for r in range(4):
    pd = [clk for ev, clk in roles[r] if ev == 0x600]
    if COOP and r == 0: print('coop: period per tile pair =', (pd[-1] - pd[4]) / max(1, len(pd) - 5))
    if len(pd) > 10:
        d = [b - a for a, b in zip(pd[4:-1], pd[5:])]
        print(f"unit {r}: tile period mean {sum(d)/len(d):.0f} cycles over {len(d)} tiles")
    # stage means
    stages = {}
    evs = roles[r]
    for (e0, c0), (e1, c1) in zip(evs[8:-1], evs[9:]):
        stages.setdefault((name(e0), name(e1)), []).append(c1 - c0)
    for k, v in stages.items():
        print(f"      {k[0]:14s} -> {k[1]:14s}: mean {sum(v)/len(v):7.0f}  n={len(v)}")
