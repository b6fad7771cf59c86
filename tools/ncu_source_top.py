#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per kernel and CUDA source line.

    python tools/ncu_source_top.py src.csv [topn] [kernel-substring]
Prints, per kernel, executed warp-instructions, mean active threads and stall samples of the hottest lines."""
import collections
import csv
import sys

csv.field_size_limit(1 << 30)
rows = csv.reader(open(sys.argv[1]))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
file = kern = None
hdr = None
inst = collections.defaultdict(collections.Counter)
tinst = collections.defaultdict(collections.Counter)
samp = collections.defaultdict(collections.Counter)
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        file = r[1].split('/')[-1]
        continue
    if r[0] == "Function Name":
        kern = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_tinst, i_samp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit() and r[2] == "-":
        try:
            ie, te, s = int(float(r[i_inst])), int(float(r[i_tinst])), int(float(r[i_samp]))
        except ValueError:
            continue
        k = (file, int(r[0]))
        inst[kern][k] += ie
        tinst[kern][k] += te
        samp[kern][k] += s
        src[k] = r[1][:100]
for kern in inst:
    if want and want not in kern:
        continue
    ti, ts = sum(inst[kern].values()), sum(samp[kern].values())
    print(f"== {kern}: {ti} warp-instructions, {ts} stall samples, {sum(tinst[kern].values()) / max(1, ti):.1f} threads/instr")
    by_file = collections.Counter()
    for (f, l), c in inst[kern].items():
        by_file[f] += c
    print("   by file:", {f: f"{100 * c / ti:.1f}%" for f, c in by_file.most_common()})
    for (f, l), c in inst[kern].most_common(topn):
        print(f"{c:10d} {100 * c / ti:5.1f}% thr {tinst[kern][(f, l)] / max(1, c):4.1f} samp {100 * samp[kern][(f, l)] / max(1, ts):5.1f}%  {f}:{l}  {src[(f, l)]}")
