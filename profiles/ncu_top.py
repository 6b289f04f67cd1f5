#!/usr/bin/env python
"""Print the hottest SASS lines (by warp-stall samples) of an ncu report: python profiles/ncu_top.py rep.ncu-rep [kernel-regex] [n]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None and row and row[0] == "Address": cur["hdr"] = row; continue
    if cur["hdr"] is not None and row: cur["rows"].append(row)
seen = set()
for b in blocks:
    short = b["name"].split("(")[0]
    if not re.search(pat, short) or short in seen: continue
    seen.add(short)
    h = b["hdr"]; si = h.index("Warp Stall Sampling (All Samples)"); so = h.index("Source")
    rows = [r for r in b["rows"] if r[si].isdigit()]
    tot = sum(int(r[si]) for r in rows) or 1
    print(f"== {short}: {tot} samples, {len(rows)} instructions")
    stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_")]
    for r in sorted(rows, key=lambda r: -int(r[si]))[:n]:
        top = sorted(((int(r[i]) if r[i].isdigit() else 0, h[i]) for i in stall_cols), reverse=True)[:2]
        print(f"{100*int(r[si])/tot:5.1f}%  {r[so].strip()[:90]:90s} {top}")
