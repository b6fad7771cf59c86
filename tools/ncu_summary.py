#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` files.

    python tools/ncu_summary.py raw.csv [raw2.csv ...]                      # the handful of metrics we track, per launch
    python tools/ncu_summary.py --traffic-json profiles/r02_traffic.json --envs 1048576 --hands 16777216 raw.csv ...
The second form writes the DRAM traffic bench.py reports in `roofline.traffic` (dram__bytes_read.sum +
dram__bytes_write.sum per launch, summed over the launches of one env-step), with its provenance: git head, hash of the
kernel sources (balatro_gym_b200._lib.source_hash), kernel names and per-launch numbers.  bench.py refuses to use a file
whose source hash differs from the code it runs.
"""
import csv
import json
import os
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__sass_average_branch_targets_threads_uniform.pct',
        'smsp__sass_average_branch_targets_threads_uniform.pct', 'sm__icc_request_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']

_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
          "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}


def launches(path):
    """[(kernel name, {metric: (value string, unit)})] of one raw-page csv."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append((r[hdr.index('Kernel Name')], {h: (r[i], units[i]) for i, h in enumerate(hdr)}))
    return out


def _num(m, key):
    v, u = m[key]
    return float(v) * _SCALE.get(u, 1.0)


def main():
    args = sys.argv[1:]
    traffic_json, envs, hands = None, 1 << 20, 1 << 24
    files = []
    while args:
        a = args.pop(0)
        if a == "--traffic-json":
            traffic_json = args.pop(0)
        elif a == "--envs":
            envs = int(args.pop(0))
        elif a == "--hands":
            hands = int(args.pop(0))
        else:
            files.append(a)
    all_launches = []
    for f in files:
        for name, m in launches(f):
            all_launches.append((os.path.basename(f), name, m))
            if traffic_json is None:
                print('==', name)
                for w in WANT:
                    if w in m:
                        print(f"  {w:86s} {m[w][0]} {m[w][1]}")
    if traffic_json is None:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from balatro_gym_b200 import _lib
    per = []
    for f, name, m in all_launches:
        per.append({"report": f, "kernel": name.split("(")[0], "us": round(_num(m, 'gpu__time_duration.sum'), 3),
                    "dram_read_bytes": _num(m, 'dram__bytes_read.sum'), "dram_write_bytes": _num(m, 'dram__bytes_write.sum')})
    step = [p for p in per if "env_step_" in p["kernel"]]
    kinds = sorted(set(p["kernel"] for p in step))
    # exactly one launch of each step kernel (main pass + one kernel per list): the launch set of ONE env-step
    assert len(step) == len(kinds), f"expected one launch per step kernel, got {[p['kernel'] for p in step]}"
    hands5 = [p for p in per if p["kernel"].endswith("score_hands5_kernel")]
    head = subprocess.run(["git", "rev-parse", "--short=12", "HEAD"], cwd=root, capture_output=True, text=True).stdout.strip()
    dirty = bool(subprocess.run(["git", "status", "--porcelain", "--", "balatro_gym_b200/csrc", "include"], cwd=root,
                                capture_output=True, text=True).stdout.strip())
    out = {
        "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none (serialised, cold cache)",
        "git_head": head + ("+uncommitted kernel edits" if dirty else ""), "source_hash": _lib.source_hash(),
        "envs": envs, "hands": hands,
        "step": {"kernels": [p["kernel"] for p in step], "launches": step,
                 "traffic_bytes": sum(p["dram_read_bytes"] + p["dram_write_bytes"] for p in step),
                 "bytes_per_env_step": sum(p["dram_read_bytes"] + p["dram_write_bytes"] for p in step) / envs},
        "hands5": ({"launch": hands5[0], "traffic_bytes": hands5[0]["dram_read_bytes"] + hands5[0]["dram_write_bytes"]} if hands5 else None),
        "other": [p for p in per if p not in step and p not in hands5],
    }
    json.dump(out, open(traffic_json, "w"), indent=1)
    print(f"wrote {traffic_json}: step {out['step']['traffic_bytes'] / 1e6:.1f} MB over {len(step)} launches "
          f"({out['step']['bytes_per_env_step']:.0f} B per env-step), source hash {out['source_hash']}")


if __name__ == "__main__":
    main()
