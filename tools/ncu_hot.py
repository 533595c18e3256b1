#!/usr/bin/env python
"""Hot regions of one launch from `ncu --page source --csv`: instruction share, stall-sample share and the
dominant opcodes per chunk of SASS; then the top individual instructions by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot_i = sum(int(r[ia]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
print("total warp-instructions", tot_i, "samples", tot_s, "sass lines", len(data))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
def opname(t):
    f = t.split()
    return f[1] if f[0].startswith('@') else f[0]
for k in range(0, len(data), chunk):
    seg = data[k:k + chunk]
    si = sum(int(r[ia]) for r in seg); ss = sum(int(r[isamp]) for r in seg)
    if si / tot_i < 0.004 and ss / tot_s < 0.004:
        continue
    ops = {}
    for r in seg:
        ops[opname(r[isrc])] = ops.get(opname(r[isrc]), 0) + int(r[ia])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print(f"{k:5d} inst {100*si/tot_i:5.1f}% samp {100*ss/tot_s:5.1f}%  ", ' '.join(f"{o}:{100*c/tot_i:.1f}" for o, c in top))
print("--- top stall instructions")
for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][isamp]))[:25]:
    print(f"{i:5d} samp {100*int(r[isamp])/tot_s:5.1f}% inst {100*int(r[ia])/tot_i:5.2f}%  {r[isrc].strip()[:90]}")
