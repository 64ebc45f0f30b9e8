#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: contiguous SASS ranges with (nearly) equal execution counts = basic blocks /
loops, with their share of the kernel's executed warp instructions and their top stall reasons.
usage: tools/ncu_blocks.py file.csv [kernel-substring] [min-share-%]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sel = sys.argv[2] if len(sys.argv) > 2 else ""
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
kern = None; hdr = None; cur = []
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if kern: kernels.append((kern, hdr, cur))
        kern = r[1]; hdr = None; cur = []
    elif r and r[0] == "Address": hdr = r
    elif hdr and len(r) >= len(hdr) - 2: cur.append(r)
if kern: kernels.append((kern, hdr, cur))
for kern, hdr, cur in kernels:
    if sel not in kern: continue
    ix = {h: i for i, h in enumerate(hdr)}
    ie = ix["Instructions Executed"]; isrc = ix["Source"]; isamp = ix["# Samples"]
    tot = sum(int(r[ie]) for r in cur); tots = sum(int(r[isamp]) for r in cur)
    print(f"== {kern}: {tot:.3e} warp instructions, {len(cur)} SASS lines, {tots} samples")
    blocks = []; b = None
    for n, r in enumerate(cur):
        e = int(r[ie])
        if b and (abs(e - b["e"]) <= 0.02 * max(e, b["e"], 1)): b["n"] += 1; b["sum"] += e; b["samp"] += int(r[isamp]); b["rows"].append(r)
        else:
            b = dict(start=n, e=e, n=1, sum=e, samp=int(r[isamp]), rows=[r]); blocks.append(b)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for b in blocks:
        sh = 100.0 * b["sum"] / max(tot, 1)
        if sh < minshare: continue
        ops = {}
        for r in b["rows"]:
            op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
            op = op.split(".")[0]; ops[op] = ops.get(op, 0) + 1
        st = {c: sum(int(r[ix[c]]) for r in b["rows"]) for c in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print(f" lines {b['start']:5d}+{b['n']:4d}  exec/line {b['e']:10d}  share {sh:5.1f}%  samples {100.0*b['samp']/max(tots,1):5.1f}%  "
              + " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8]) + "  | " + " ".join(f"{k[6:]}:{v}" for k, v in top))
