#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: instruction mix by opcode, stall reasons, hottest SASS.
usage: ncu -i rep --page source --csv --kernel-name regex:X > src.csv; python tools/ncu_src_summary.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
print(f"SASS lines {len(body)}  warp-inst {tot_inst:.0f}  samples {tot_samp:.0f}")
ops = collections.Counter(); samp = collections.Counter()
for r in body:
    s = r[ix["Source"]].strip()
    toks = s.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDG", "STG", "MUFU", "BAR", "SHFL")) else ".".join(op.split(".")[:2])
    ops[op] += f(r, "Instructions Executed"); samp[op] += f(r, "# Samples")
print("opcode            inst%   samples%")
for op, n in ops.most_common(28):
    print(f"{op:16s} {100*n/tot_inst:6.2f}  {100*samp[op]/max(tot_samp,1):6.2f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h: sum(f(r, h) for r in body) for h in stalls}
tot = sum(st.values())
print("stall reasons (all samples):")
for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:24s} {100*v/max(tot,1):6.2f}%")
print("hottest SASS by samples:")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = max(stalls, key=lambda h: f(r, h))
    print(f"  {f(r,'# Samples'):7.0f} {top:18s} {r[ix['Source']].strip()[:90]}")
