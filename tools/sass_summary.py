#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libbgym.so (cuobjdump -sass): python tools/sass_summary.py > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = ""
for name in ("libbgym.so", "libbgym_policy.so"):
    txt += subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "balatro_gym_b200", name)], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
COLS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("LDS.128", r"\bLDS\S*\.128"), ("STS.128", r"\bSTS\S*\.128"),
        ("LDG.E.128", r"\bLDG\.E\S*\.128"), ("STG.E.128", r"\bSTG\.E\S*\.128"), ("ATOMG/RED", r"\b(ATOMG|RED)\b"), ("SHFL", r"\bSHFL"),
        ("VOTE", r"\bVOTE"), ("DMUL/DFMA/DADD", r"\b(DMUL|DFMA|DADD)\b"), ("UTCHMMA", r"\bUTC\w*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("HMMA", r"\bHMMA")]
kern, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", demangle(m.group(1)).replace("(int)", "")).replace("bgym::", "").replace("void ", "").replace("(anonymous namespace)::", "")
        continue
    if kern and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        total[kern] += 1
        for name, pat in COLS:
            if re.search(pat, line):
                counts[kern][name] += 1
tag = sys.argv[1] if len(sys.argv) > 1 else "2"
print(f"# Round {tag} — SASS mnemonic counts per kernel of balatro_gym_b200/libbgym.so and libbgym_policy.so (cuobjdump -sass, sm_100a; tools/sass_summary.py)\n")
print("UBLKCP = cp.async.bulk (the TMA engine's 1-D bulk copy), SYNCS = mbarrier arrive/expect_tx/try_wait, LDGSTS = cp.async,\nUTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld.\n")
print("| kernel | instr | " + " | ".join(c[0] for c in COLS) + " |")
print("|---" * (len(COLS) + 2) + "|")
for k, n in total.most_common():
    print(f"| {k} | {n} | " + " | ".join(str(counts[k][c[0]]) for c in COLS) + " |")
