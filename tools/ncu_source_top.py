#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` by CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
file = None; hdr = None
inst = collections.Counter(); samp = collections.Counter(); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": file = r[1].split('/')[-1]; continue
    if r[0] in ("Function Name",): continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[2] == "-" and r[0].isdigit():
        ln = int(r[0])
        try:
            ie = int(float(r[7])); s = int(float(r[6]))
        except ValueError:
            continue
        inst[(file, ln)] += ie; samp[(file, ln)] += s; src[(file, ln)] = r[1][:90]
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-instructions", ti, "stall samples", ts)
for (f, l), c in inst.most_common(topn):
    print(f"{c:10d} {100*c/ti:5.1f}% samp {100*samp[(f,l)]/max(1,ts):5.1f}%  {f}:{l}  {src[(f,l)]}")
