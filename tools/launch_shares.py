#!/usr/bin/env python
"""profiles/rNN_launch_shares.md from the ncu launch list (gpurun_out/launches.csv written by tools/gpu_prof_part.sh):
python tools/launch_shares.py [launches.csv] > profiles/r02_launch_shares.md"""
import collections
import csv
import io
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("nsecond", "ns"):
        v /= 1e3
    elif r["Metric Unit"] in ("msecond", "ms"):
        v *= 1e3
    tot[r["Kernel Name"]] += v
    cnt[r["Kernel Name"]] += 1
total = sum(tot.values())
print("# Launch list of `python bench.py --steps 20 --warmup 3 --burn-in 30 --no-cpu-baseline --e2e-steps 3 --no-facade` under\n"
      "# `ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400` (400 launches from the 300th on; cold-cache, serialised: compare SHARES)\n"
      "# torch kernels in the list are the bench's state statistics and bookkeeping, outside the timed step loop\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| `{k[:90]}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% | {v / cnt[k]:.1f} |")
step = {k: v / cnt[k] for k, v in tot.items() if "env_step_" in k or "sample_actions" in k}
ssum = sum(step.values())
print("\nShare of one rollout step (sampler + average launch of each step kernel, serialised; in production the seven level-1\n"
      "list kernels run concurrently, then the two level-2 kernels):\n")
for k, v in sorted(step.items(), key=lambda kv: -kv[1]):
    print(f"* `{k[:70]}`: {v:.1f} us = {100 * v / ssum:.1f}%")
