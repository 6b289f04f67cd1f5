#!/usr/bin/env python
"""Warp-stall samples of an ncu report aggregated per CUDA source line: python profiles/ncu_lines.py rep.ncu-rep kernel-regex [n]"""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
hdr, agg, src, cur_line, cur_file, fname = None, collections.Counter(), {}, None, None, None
done_first = False
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = r.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0].strip():  # a CUDA source line row
        cur_line = (fname, int(r[0]))
        src[cur_line] = r[1].strip()
    elif r[si].isdigit() and cur_line:
        agg[cur_line] += int(r[si])
tot = sum(agg.values()) or 1
print(f"{tot} samples")
for (f, ln), c in agg.most_common(n):
    print(f"{100*c/tot:5.1f}%  {f}:{ln:<4d} {src.get((f, ln), '')[:110]}")
