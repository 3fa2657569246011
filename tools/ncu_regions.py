#!/usr/bin/env python
"""Split `ncu --page source --csv` output at BAR.SYNC instructions: per-region instruction counts,
sample share and top stall reasons.  usage: python tools/ncu_regions.py src.csv [n_frames]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seg = 0
acc = collections.defaultdict(float); samp = collections.defaultdict(float)
st = collections.defaultdict(collections.Counter); ops = collections.defaultdict(collections.Counter)
for r in body:
    src = r[ix["Source"]]
    n = float(r[ix["Instructions Executed"]] or 0); s = float(r[ix["# Samples"]] or 0)
    acc[seg] += n; samp[seg] += s
    for h in stalls: st[seg][h[6:]] += float(r[ix[h]] or 0)
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[seg][op.split(".")[0]] += n
    if "BAR.SYNC" in src: seg += 1
tot = sum(acc.values()); ts = sum(samp.values())
for k in sorted(acc):
    if acc[k] / tot < 0.002: continue
    top = ", ".join(f"{a}:{b / max(samp[k], 1) * 100:.0f}%" for a, b in st[k].most_common(5))
    print(f"region {k}: inst {acc[k] / tot * 100:5.1f}% ({acc[k] / frames:7.0f}/frame)  samples {samp[k] / ts * 100:5.1f}%  stalls [{top}]")
    print("      ops:", ", ".join(f"{a}:{b / frames:.0f}" for a, b in ops[k].most_common(10)))
