#!/usr/bin/env python
"""Rebuild the tracked profiles/rNN_* files from the scratch outputs of tools/gpu_round.sh
(gpurun_out/: bench.json, bench_reference.json, pytest_gpu.log, launches.csv, prof_step.ncu-rep,
prof_hands.ncu-rep).  Usage: python tools/refresh_profiles.py [round-tag, default r01]"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def raw_summary(rep):
    csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", os.path.basename(rep) + ".raw.csv")
    open(tmp, "w").write(csv_txt)
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), tmp],
                          capture_output=True, text=True).stdout


def main():
    bench = json.load(open(os.path.join(OUT, "bench.json")))
    ref = json.load(open(os.path.join(OUT, "bench_reference.json")))
    json.dump(bench, open(os.path.join(PROF, f"{tag}_bench_line.json"), "w"))
    json.dump(ref, open(os.path.join(PROF, f"{tag}_bench_reference_line.json"), "w"))
    shutil.copy(os.path.join(OUT, "pytest_gpu.log"), os.path.join(PROF, f"{tag}_pytest_gpu.log"))

    # ---- launch list and shares
    lines = [l for l in open(os.path.join(OUT, "launches.csv")) if l.startswith('"')]
    shutil.copy(os.path.join(OUT, "launches.csv"), os.path.join(PROF, f"{tag}_launches.csv"))
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = collections.Counter(); cnt = collections.Counter()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("nsecond", "ns"):
            v /= 1e3
        elif r["Metric Unit"] in ("msecond", "ms"):
            v *= 1e3
        tot[r["Kernel Name"]] += v; cnt[r["Kernel Name"]] += 1
    total = sum(tot.values())
    with open(os.path.join(PROF, f"{tag}_launch_shares.md"), "w") as f:
        f.write("# Launch list of `python bench.py --steps 20 --warmup 3 --burn-in 30 --no-cpu-baseline --e2e-steps 3` under\n"
                "# `ncu --metrics gpu__time_duration.sum --clock-control none -c 600` (first 600 launches; cold-cache, serialised: compare SHARES)\n"
                "# torch kernels in the list are the bench's synthetic-state generator and the e2e leg, outside the timed step loop\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k[:90]}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% | {v / cnt[k]:.1f} |\n")
        step = {k: v / cnt[k] for k, v in tot.items() if "env_step_" in k}
        ssum = sum(step.values())
        f.write("\nShare of one env-step (average launch of each step kernel, serialised):\n\n")
        for k, v in sorted(step.items(), key=lambda kv: -kv[1]):
            f.write(f"* `{k[:70]}`: {v:.1f} us = {100 * v / ssum:.1f}%\n")

    # ---- ncu --set full summaries
    r = bench["roofline"]
    with open(os.path.join(PROF, f"{tag}_ncu_summary_final.md"), "w") as f:
        f.write(f"# ncu summaries, round {tag[1:]}, final build (B200; `ncu --set full --clock-control none --import-source on`, "
                "raw metrics via `ncu -i X.ncu-rep --page raw --csv`)\n\n"
                "Command profiled: `python bench.py --steps 20 --warmup 3 --burn-in 100 --no-cpu-baseline --no-hands --e2e-steps 3` "
                "(2^20 envs, BASELINE configs[3] workload).\n"
                "Per-launch times under ncu are cold-cache and serialised (the three gather passes run concurrently in the real step); "
                "the bench line's CUDA-event timing is the performance number.\n\n")
        f.write(f"Bench line of the same build (not under a profiler): value {bench['value']:.3e} env-steps/s, step launches "
                f"{r['kernel_ms']:.3f} ms, roofline.frac {r['frac']:.3f}, fused rollout {bench['fused_rollout']['value']:.3e}, "
                f"hands {bench['hands']['value']:.3e} (frac {bench['hands']['roofline']['frac']:.3f}), clocks {bench['clocks']}\n\n")
        f.write("### Step (hot/cold arrays, category-partitioned): main pass + the three gather passes of one step\n```\n")
        f.write(raw_summary(os.path.join(OUT, "prof_step.ncu-rep")))
        f.write("```\n\n### score_hands5_kernel, 2^24 hands (BASELINE configs[1])\n```\n")
        f.write(raw_summary(os.path.join(OUT, "prof_hands.ncu-rep")))
        f.write("```\n")
    print("profiles refreshed:", sorted(x for x in os.listdir(PROF) if x.startswith(tag)))


if __name__ == "__main__":
    main()
