#!/usr/bin/env python
"""Per CUDA source line: share of executed warp-instructions and stall samples of one kernel launch.
   python tools/ncu_lines.py <ncu --page source --csv (sass)> <cubin> <kernel substring> [top]
Joins ncu's SASS-level counts with the line table that `nvdisasm -g` prints for the cubin (-lineinfo build)."""
import csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
# the report lists the function twice (two views); keep the first copy
addrs = [int(r[0], 16) for r in data]
n = len(addrs)
for k in range(1, n):
    if addrs[k] <= addrs[k - 1] and addrs[k] == addrs[0]:
        n = k; break
data = data[:n]; base = int(data[0][0], 16)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith(".text."): infn = kname in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m and infn: line_of[int(m.group(1), 16)] = cur
agg = {}
for r in data:
    key = line_of.get(int(r[0], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0, 0]); a[0] += int(r[ia]); a[1] += int(r[isamp])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values()) or 1
print(f"kernel {kname}: {ti} warp-instructions, {ts} samples, {len(data)} SASS lines")
srcs = {}
for (f, l), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    txt = ""
    if f != "?":
        if f not in srcs:
            try: srcs[f] = open(f"pointnet-like-pose-estimation_b200/csrc/{f}").read().splitlines()
            except OSError: srcs[f] = []
        if 0 < l <= len(srcs[f]): txt = srcs[f][l - 1].strip()[:90]
    print(f"{100*i/ti:5.1f}% inst {100*s/ts:5.1f}% samp  {f}:{l}  {txt}")
