#!/usr/bin/env python
"""bench.py — env-steps/s (and hands-scored/s) of the B200-native Balatro step path.

    python bench.py --gpus N --steps K --warmup W            # our arm, one rank per GPU (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU path

A "step" is one pass of the hot path over one batch of synthetic input: every env of the slab
takes one uniform-random LEGAL action (sampler kernel + fused step kernel).  The workload is
BASELINE.json configs[3] ("full run incl. shop, rerolls, planets and consumables, 2^20 envs/GPU",
state generator = configs[2]: 5 random jokers, enhancements/editions/seals, boss blinds, autoreset);
configs[1] (2^24-hand scoring microbench) is reported in the same line under "hands".

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each number is taken.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time


REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

L_OBS = 176    # bytes of one observation record (include/bgym.h BgymObs)
B_STEP = 872   # canonical algorithmic bytes per env-step
B_HAND = 32    # canonical algorithmic bytes per scored hand
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def measured_traffic(kind):
    """DRAM bytes per launch set from the committed ncu capture (profiles/r02_traffic.json, written by
    tools/ncu_summary.py --traffic-json), or (None, why) when the capture was taken on other kernel sources than the
    ones this run executes: a stale figure is not reported."""
    path = os.path.join(REPO, "profiles", "r02_traffic.json")
    try:
        d = json.load(open(path))
        from balatro_gym_b200 import _lib
        prov = {"file": "profiles/r02_traffic.json", "git_head": d["git_head"], "source_hash": d["source_hash"]}
        if d["source_hash"] != _lib.source_hash():
            sys.stderr.write(f"bench.py: {path} was captured on other kernel sources (hash {d['source_hash']} != "
                             f"{_lib.source_hash()}): roofline.traffic is not reported; re-run tools/gpu_prof_part.sh + "
                             "tools/ncu_summary.py --traffic-json\n")
            prov["stale"] = True
            return None, None, prov
        if kind == "step":
            prov["kernels"] = d["step"]["kernels"]
            return d["step"]["traffic_bytes"], d["envs"], prov
        prov["kernels"] = [d["hands5"]["launch"]["kernel"]]
        return d["hands5"]["traffic_bytes"], d["hands"], prov
    except Exception as e:      # no capture committed
        return None, None, {"file": "profiles/r02_traffic.json", "unavailable": repr(e)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def bind_host_memory_near_gpu(torch, dev):
    """Best effort: run this rank and place its future host allocations (the pinned staging buffers) on the NUMA node
    the GPU hangs off.  With eight ranks on one box, pinned buffers that all land on one socket send half of the result
    copies across the inter-socket link (SCALE_r01: e2e efficiency 0.24 at 8 GPUs).  Returns what was done."""
    out = {"gpu_numa_node": None, "cpu_affinity": None, "mempolicy": None}
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
        if bus is None:
            import subprocess as sp
            q = sp.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev.index or 0)],
                       capture_output=True, text=True).stdout.strip()
            bus = q[-12:].lower() if q else None            # 00000000:1B:00.0 -> 0000:1b:00.0
        else:
            bus = f"0000:{bus:02x}:00.0" if isinstance(bus, int) else str(bus).lower()
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read())
        out["gpu_numa_node"] = node
        if node < 0:
            return out
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            out["cpu_affinity"] = f"{len(allowed)} cpus of node {node}"
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED = 1
        rc = libc.syscall(238, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))     # set_mempolicy (x86-64)
        out["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
    except Exception as e:
        out["error"] = repr(e)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the one place bench.py executes oracle/)
# -------------------------------------------------------------------------------------------------
def cpu_env_baseline(config, rounds, warmup_rounds, steps_per_round=1024):
    from oracle import refenv, refbaseline
    root = refenv.reference_available()
    if root is None:
        # no reference on this box: time the C port instead (kind "port", 1 core)
        import numpy as np
        from oracle import coracle
        from balatro_gym_b200 import layout as L
        n = 4096
        ov = coracle.OracleVec(n)
        ov.reset(np.arange(1, n + 1))
        act = np.zeros(n, np.int32)
        ts = []
        for r in range(rounds + warmup_rounds):
            t0 = time.perf_counter()
            for _ in range(16):
                coracle.step(ov.state, act, ov.obs, ov.reward, ov.terminated, ov.truncated, ov.info, None, flags=L.FLAG_AUTORESET | 4)
            ts.append(time.perf_counter() - t0)
        ts = ts[warmup_rounds:]
        return {"value": n * 16 * len(ts) / sum(ts), "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"C oracle, {n} envs x 16 steps x {len(ts)} rounds, fused random-legal policy, autoreset"}, ts
    res = refbaseline.run_env_baseline(config, steps_per_round=steps_per_round, rounds=rounds, warmup_rounds=warmup_rounds)
    total = res["steps_per_round_total"] * len(res["per_round_s"])
    return {"value": total / sum(res["per_round_s"]), "unit": UNIT, "cores": res["cores"], "kind": "reference",
            "sample": (f"unmodified reference BalatroEnv ({'byte-compiled oracle/_ref' if '_ref' in root else root}), "
                       f"{res['cores']} worker processes x {steps_per_round} env-steps x {len(res['per_round_s'])} rounds, "
                       f"config {config} state generator, random legal actions, no per-step IPC (upper bound of an AsyncVectorEnv)")}, res["per_round_s"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = "c4"
    base, per_round = cpu_env_baseline(cfg, rounds=max(1, args.steps), warmup_rounds=max(0, args.warmup), steps_per_round=1024)
    ms = 1000.0 * sum(per_round) / max(1, len(per_round))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
        # the SAME config dict as our arm: both arms run this workload; the reference arm runs it on a bounded
        # sample (cpu_baseline.sample: one env per host core instead of envs_per_gpu envs)
        "config": workload_config(args, args.envs),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_envs):
    return {"workload": "BASELINE configs[3]: full run incl. shop/rerolls/planets/consumables; EVERY episode (first reset and "
                        "every autoreset) starts from the configs[2] state generator (5 random jokers, card enhancements/"
                        "editions/seals) plus 2 random consumables over all 52 names; random legal actions, autoreset",
            "envs_per_gpu": n_envs, "policy": "uniform random legal action per env",
            "generator": "c4 = BGYM_FLAG_GEN_C3 | BGYM_FLAG_GEN_CONS (include/bgym.h); reference arm: oracle/refbaseline.py::_inject_c3 per episode",
            "l2": "inputs larger than L2 (state+obs records of one step = %.0f MB per GPU)" % (n_envs * (320 + L_OBS) / 1e6)}


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import balatro_gym_b200 as b
    from balatro_gym_b200 import dist as bdist
    from balatro_gym_b200 import layout as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    # rank 0's stdout carries exactly one JSON line.  NCCL prints its version banner to stdout when the first
    # communicator is created (whatever NCCL_DEBUG says from VERSION up), so the process group is brought up —
    # first collective included — with fd 1 pointing at stderr.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, local_rank, ws = bdist.init_process_group("nccl")
        dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(dev)
        bdist.max_over_ranks(0.0, dev)
        torch.cuda.synchronize(dev)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    n = args.envs
    K, W = args.steps, max(args.warmup, 3)
    peak, peak_src = measured_peak()

    # the state generator runs ON THE DEVICE in reset and in every in-kernel autoreset (round 1 scattered it once
    # from the host and autoresets then produced vanilla episodes: 14 % generator state left in the timed window)
    env = b.BalatroVecEnv(n, device=dev, seed=1, autoreset=True, env_offset=rank * n, generator="c4")
    env.reset()
    launches = 0

    def rollout_step():
        env.sample_actions(seed=2024)
        env.step(env.actions, want_info=False)

    def state_stats():
        """Workload description computed on the device from the env records (not timed)."""
        deck = env.state_field("deck").to(torch.int32)
        phase = env.state_field("phase").long()
        pm = torch.bincount(phase, minlength=4).double() / n
        return {"generator_state_frac": float(((deck & 0xFFC0) != 0).any(dim=1).double().mean()),
                "mean_joker_n": float(env.state_field("joker_n").double().mean()),
                "mean_cons_n": float(env.state_field("cons_n").double().mean()),
                "mean_ante": float(env.state_field("ante").double().mean()),
                "episodes_per_env": float(env.state_field("episode").double().mean()),
                "phase_frac": {"play": float(pm[0]), "shop": float(pm[1]), "blind_select": float(pm[2])}}

    # spread the envs over game phases before timing (episodes desynchronise within ~100 steps)
    for _ in range(args.burn_in):
        rollout_step()
    for _ in range(W):
        rollout_step()
    torch.cuda.synchronize(dev)
    window_start = state_stats()
    # the actions of the whole timed window are kept (K x n int32) so that its action mix is exact, at no cost
    # inside the timed region: the sampler writes step k's actions to acts[k], the step reads them there
    acts = torch.empty((K, n), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    bdist.barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 2)]
    torch.cuda.synchronize(dev)
    ev[0].record()
    for k in range(K):
        env.sample_actions(seed=2024, out=acts[k])
        ev[1 + 2 * k].record()          # step kernel bracket (same stream as the launches)
        env.step(acts[k], want_info=False)
        ev[2 + 2 * k].record()
        launches += 10  # sampler + main pass + seven level-1 list kernels + the level-2 kernel (round advances and resets)
    ev[2 * K + 1].record()
    torch.cuda.synchronize(dev)
    bdist.barrier()
    window_end = state_stats()
    ah = torch.bincount(acts.reshape(-1).long().clamp(0, 63), minlength=64).double()
    ah = (ah / ah.sum()).cpu().tolist()
    action_mix = {"select_card": sum(ah[2:10]), "play_hand": ah[0], "discard": ah[1], "use_consumable": sum(ah[10:15]),
                  "shop_buy": sum(ah[20:30]), "shop_reroll": ah[30], "shop_end": ah[31], "sell_joker": sum(ah[32:37]),
                  "select_blind": sum(ah[45:48]), "skip_blind": ah[48]}
    # every env takes exactly one legal action per step and the legal ids of the three phases are disjoint,
    # so the phase mix of the window follows from the action histogram
    phase_mix = {"play": sum(ah[0:15]), "shop": sum(ah[20:37]), "blind_select": sum(ah[45:49])}
    del acts
    total_ms = ev[0].elapsed_time(ev[2 * K + 1])
    step_kernel_ms = sum(ev[1 + 2 * k].elapsed_time(ev[2 + 2 * k]) for k in range(K)) / K
    total_ms = bdist.max_over_ranks(total_ms, dev)
    step_kernel_ms_max = bdist.max_over_ranks(step_kernel_ms, dev)
    value = ws * n * K / (total_ms / 1000.0)
    achieved = n * B_STEP / (step_kernel_ms_max / 1000.0) / 1e9      # GB/s, algorithmic bytes

    # fused rollout (policy inside the step kernel): one launch per step
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        env.step(random_policy=True, want_info=False)
    e1.record()
    torch.cuda.synchronize(dev)
    fused_ms = bdist.max_over_ranks(e0.elapsed_time(e1), dev)
    fused_value = ws * n * K / (fused_ms / 1000.0)
    # the same two loops replayed from CUDA graphs (BalatroVecEnv.graphed_rollout_step): no launch gaps
    graph_vals = {}
    for pol in ("sampler", "fused"):
        replay = env.graphed_rollout_step(pol, seed=2024)
        for _ in range(W):
            replay()
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(K):
            replay()
        g1.record()
        torch.cuda.synchronize(dev)
        graph_vals[pol] = ws * n * K / (bdist.max_over_ranks(g0.elapsed_time(g1), dev) / 1000.0)
    # clocks were sampled (nvidia-smi, 20 ms period) from the start of the timed steps to here: all loops run
    # the same step kernels back to back
    clk = clocks.stop() if rank == 0 else None

    # episode statistics (K6): folded per step on the device over a short untimed rollout, then ONE tiny
    # all-reduce per rollout, off the step path
    env.stats.zero_()
    for _ in range(64):
        env.step(random_policy=True, want_info=False)
        env.accumulate_stats()
    stats = bdist.allreduce_stats(env.stats.clone())

    # ---- e2e: the public API with HOST buffers (pinned), copies inside the timed region ----
    # balatro_gym_b200.HostMirror: every step this step's actions come from pinned host memory (H2D) and the step's
    # results reach pinned host memory (D2H): reward, terminated and the selection records (selected_cards as one flag byte, mask word)
    # of every env, and the 176-byte observation record of every env whose record the step rewrote (an env whose
    # action was a card toggle keeps its record) — packed on the device and written into the host mirror by the GPU.
    # The copies of step t run on a second stream while step t+1 is launched; the action path is synchronous.  The
    # timed region ends when the last step's results are in host memory; the mirror is then compared with the
    # device arrays, whole.
    Ke = max(3, min(K, args.e2e_steps))
    numa = bind_host_memory_near_gpu(torch, dev)
    mirror = b.HostMirror(env)
    main_stream = torch.cuda.current_stream(dev)
    mirror.pull_all()
    dirty = []

    def e2e_step(t):
        env.sample_actions(seed=7)                               # the agent's decision, made on the device
        mirror.actions.copy_(env.actions, non_blocking=True)     # ... and handed to the host, as a host-driven loop has it
        main_stream.synchronize()
        mirror.step()                                            # H2D actions, step, D2H deltas (second stream)

    for t in range(2):
        e2e_step(t)
    mirror.wait()
    bdist.barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        e2e_step(t)
    mirror.wait()
    e2e_s = time.perf_counter() - t0
    dc = [mirror.delta_counts(k) for k in range(2)]
    dirty_frac, shop_frac = sum(c[0] for c in dc) / (2.0 * n), sum(c[1] for c in dc) / (2.0 * n)
    e2e_d2h = (mirror.dense_d2h_bytes_per_step + dirty_frac * n * L.MIRROR_CORE_BYTES + shop_frac * n * L.MIRROR_SHOP_BYTES
               + 4 * n)    # + the action hand-off
    e2e_rank_gbs = (e2e_d2h + mirror.h2d_bytes_per_step) * Ke / e2e_s / 1e9     # this rank's PCIe traffic, both directions
    e2e_s = bdist.max_over_ranks(e2e_s, dev)
    e2e_value = ws * n * Ke / e2e_s
    # the host mirror equals the device arrays (whole observation records: selection records folded in)
    host_rec, dev_rec = mirror.obs_records(), env.obs_numpy()
    for name in L.OBS_DTYPE.names:
        assert np.array_equal(host_rec[name], dev_rec[name]), f"host mirror differs from the device observations in {name}"
    del host_rec, dev_rec
    assert torch.equal(mirror.terminated, env.terminated.cpu()) and torch.equal(mirror.reward, env.reward.cpu())
    del mirror

    # ---- hands microbench (configs[1]) ----
    hands = None
    if rank == 0 and not args.no_hands:
        hands = bench_hands(torch, b, dev, peak, args)

    # ---- PPO rollout collection (configs[4]; every rank runs its env slab) ----
    ppo = None
    if not args.no_ppo:
        ppo = bench_ppo_rollout(torch, bdist, dev, args, rank, ws)

    if rank != 0:
        return 0
    tbytes, tenvs, traffic_prov = measured_traffic("step")
    traffic = None if tbytes is None else tbytes * n / float(tenvs)
    frac_physical = None if traffic is None else traffic / (step_kernel_ms_max / 1000.0) / 1e9 / peak
    host_facing = None if args.no_facade else bench_host_facing()
    cpu_base = None
    if not args.no_cpu_baseline and ws == 1:
        cpu_base, _ = cpu_env_baseline("c4", rounds=3, warmup_rounds=1, steps_per_round=4096)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64+f64", "data": "synthetic", "config": workload_config(args, n),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_note": "frac = ALGORITHMIC bytes (872 B per env-step, SURVEY 8d) / kernel_ms / peak; frac_physical = DRAM "
                                  "bytes the launches really move (ncu) / kernel_ms / peak",
                     "traffic": traffic, "frac_physical": frac_physical, "traffic_provenance": traffic_prov,
                     "traffic_unit": "dram__bytes_read.sum + dram__bytes_write.sum over the launches of one env-step (ncu --set full)",
                     "kernel": "one env-step = env_step_main_kernel + 7 env_step_list_kernel launches + env_step_level_kernel<2> (all launches of the step are inside the timed bracket)", "bytes_per_unit": B_STEP,
                     "units_per_launch": n, "kernel_ms": step_kernel_ms_max, "peak_source": peak_src,
                     "physical_bytes_per_unit": "main pass 32 (toggle record) + 4 (action) read, 32 + 16 (selection record) + 10 written = 94 B per env; list kernels add the hot / toggle / cold / obs records of the ~25 % deferred envs at 64-byte DRAM granularity"},
        "timed_window": {"start": window_start, "end": window_end, "action_mix": action_mix, "phase_mix": phase_mix,
                         "note": "state statistics computed on the device from the env records right before / after the timed "
                                 "steps; action and phase mix are exact over all envs x steps of the timed window (rank 0 slab)"},
        "previous_headline": {"round": 1, "value": 4.87e9, "unit": UNIT,
                              "note": "round 1 applied the generator once from the host and autoreset built vanilla episodes: "
                                      "14 % generator state in its timed window (VERDICT r01 weak #1)"},
        "cpu_baseline": cpu_base,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": e2e_d2h,
                "steps": Ke, "pcie_gbs_rank0": e2e_rank_gbs, "host_memory": numa, "rewritten_record_frac": dirty_frac, "shop_chunk_frac": shop_frac,
                "whole_array_d2h_bytes_per_step": (L.OBS_BYTES + 8 + 1 + 4) * n,
                "note": "HostMirror: reward + terminated + packed selection (flag byte + mask word: 18 B) of every env and the observation record "
                        "(one aligned 128-byte line, + 32 B of shop chunks where they can have changed) of every env whose record the step "
                        "rewrote reach pinned host memory each step, written by the GPU itself (zero-copy stores); the mirror is "
                        "checked against the device arrays after the timed region; bound by the PCIe link"},
        "gpu_launches": launches,
        "clocks": clk,
        "fused_rollout": {"value": fused_value, "unit": UNIT, "ms_per_step": fused_ms / K,
                          "note": "policy sampled inside the step kernel (one launch per step)"},
        "graph_replay": {"sampler_plus_step": graph_vals["sampler"], "fused": graph_vals["fused"], "unit": UNIT,
                         "note": "the value / fused_rollout loops replayed from a CUDA graph per step"},
        "hands": hands,
        "ppo_rollout": ppo,
        "host_facing": host_facing,
        "episode_stats": {"episodes": float(stats[0]), "mean_return": float(stats[1] / max(1.0, float(stats[0]))),
                          "mean_length": float(stats[2] / max(1.0, float(stats[0])))},
    }
    print(json.dumps(line))
    return 0


def bench_ppo_rollout(torch, bdist, dev, args, rank, ws):
    """configs[4]: PPO rollout collection with the policy on the device — 2^19 envs per GPU
    (2^22 over 8), every episode from the c4 state generator, observation records -> policy (one tcgen05 kernel) -> masked
    categorical -> env step, GAE at the end; nothing leaves the device.  Reported as whole-job env-steps/s, max over ranks."""
    from balatro_gym_b200 import BalatroVecEnv
    from balatro_gym_b200.rollout import RolloutCollector, make_policy, featurize, masked_sample, gae
    n, T = args.ppo_envs, args.ppo_steps
    # the same per-episode state generator as the headline loop (reset AND every in-kernel autoreset)
    vec = BalatroVecEnv(n, device=dev, seed=1, env_offset=rank * n, generator="c4")
    vec.reset()
    for _ in range(32):
        vec.step(random_policy=True)
    policy = make_policy(device=dev, seed=0)
    roll = RolloutCollector(vec, policy, n_steps=T, seed=1)
    roll.collect()                                   # warm-up: cuBLAS heuristics, allocator
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    bdist.barrier(); torch.cuda.synchronize(dev)
    ev[0].record(); roll.collect(); ev[1].record()
    torch.cuda.synchronize(dev)
    ms = bdist.max_over_ranks(ev[0].elapsed_time(ev[1]), dev)

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps
    obs0 = roll.obs[0]
    with torch.no_grad():
        roll.refresh_inference_weights()
        t_fwd = timed(lambda: roll._forward(obs0))          # the whole policy: one tcgen05 kernel straight from the records
        logits = roll._forward(obs0)[0]
        t_samp = timed(lambda: masked_sample(logits, obs0, seed=1, step=0, actions=roll.actions[0], logp=roll.logp[0], entropy=roll.entropy[0]))
        t_step = timed(lambda: vec.step(roll.actions[0], want_info=False))
        t_copy = timed(lambda: roll.obs[1].copy_(vec.obs_buf))
        t_gae = timed(lambda: gae(roll.rewards, roll.values, roll.dones, 0.99, 0.95, roll.advantages, roll.returns))
        # the library-GEMM path of the same forward (first-layer kernel + eleven cuBLASLt GEMMs), for comparison
        roll.fused = False
        t_lib = timed(lambda: roll._forward(obs0))
        roll.fused = True
    if rank != 0:
        return None
    flops = 2 * n * (416 * 256 + 10 * 128 + 21 * 64 + 256 * 128 + 128 * 64 + 64 * 32 + 224 * 512 + 512 * 512 + 2 * (512 * 256 + 256 * 256) + 256 * 61)
    return {"value": ws * n * T / (ms / 1e3), "unit": "env-steps/s", "envs_per_gpu": n, "rollout_steps": T,
            "ms_per_step": ms / T,
            "breakdown_ms": {"policy_forward_fused_tcgen05": t_fwd, "masked_sample": t_samp, "env_step": t_step,
                             "obs_copy": t_copy, "gae_whole_rollout": t_gae},
            "policy_forward_library_gemms_ms": t_lib,
            "policy_forward_tflops": flops / t_fwd / 1e9,
            "policy": "BalatroFeaturesExtractor topology (416|10|21 -> 224 -> 512 -> 512) + pi/vf [256,256] heads, bf16 weights: all 14 "
                      "layers in ONE tcgen05 kernel straight from the observation records (libbgym_policy.so; activations in "
                      "swizzled shared memory, fp32 accumulators in tensor memory, weights streamed by bulk copies)",
            "note": "policy_forward_library_gemms_ms = the same forward as first-layer kernel + eleven cuBLASLt GEMMs with bias+ReLU epilogues "
                    "(what round 1 and the first half of round 2 ran)"}


def bench_hands(torch, b, dev, peak, args):
    """configs[1]: 2^24 random 5-card plays, no jokers, base hand levels."""
    from balatro_gym_b200.score import score_hands
    n = args.hands
    g = torch.Generator(device=dev); g.manual_seed(0x5EED)
    cards = torch.zeros((n, 8), dtype=torch.uint8, device=dev)
    chunk = 1 << 22
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        cards[s:s + m, :5] = torch.rand((m, 52), device=dev, generator=g).topk(5, dim=1).indices.to(torch.uint8)
    out = score_hands(cards, want_x_mult=False, want_money=False)
    for _ in range(5):
        score_hands(cards, want_x_mult=False, want_money=False, out=out)
    torch.cuda.synchronize(dev)
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        score_hands(cards, want_x_mult=False, want_money=False, out=out)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    value = n / (ms / 1000.0)
    ach = n * B_HAND / (ms / 1000.0) / 1e9
    hb, hn, hprov = measured_traffic("hands5")
    htraffic = None if hb is None else hb * n / float(hn)
    # e2e: host cards in pinned memory -> device -> scores back
    h_cards = torch.empty((n, 8), dtype=torch.uint8, pin_memory=True); h_cards.copy_(cards)
    h_score = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_type = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        cards.copy_(h_cards, non_blocking=True)
        score_hands(cards, want_x_mult=False, want_money=False, out=out)
        h_score.copy_(out["score"], non_blocking=True); h_type.copy_(out["hand_type"], non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e = 3 * n / (time.perf_counter() - t0)
    res = {"metric": "hands_scored_per_sec", "value": value, "unit": "hands/s", "n_hands": n, "ms_per_launch": ms,
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "bytes_per_unit": B_HAND, "kernel": "score_hands5_kernel",
                        "traffic": htraffic, "frac_physical": None if htraffic is None else htraffic / (ms / 1000.0) / 1e9 / peak,
                        "traffic_provenance": hprov,
                        "note": "the kernel reads 8 B and writes 17 B per hand: it is bound by instruction issue (ncu: sm throughput 77 %, "
                                "issue slots 62 %), not by HBM; frac uses the canonical 32 B per hand"},
           "e2e": {"value": e2e, "unit": "hands/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 9 * n}}
    res["with_jokers"] = bench_hands_jokers(torch, dev, peak)
    if not args.no_cpu_baseline:
        try:
            from oracle import refenv, refbaseline
            if refenv.reference_available():
                sample = cards[: 1 << 15].cpu().numpy()
                r = refbaseline.run_hands_baseline(sample)
                res["cpu_baseline"] = {"value": r["hands"] / r["seconds"], "unit": "hands/s", "cores": r["cores"], "kind": "reference",
                                       "sample": f"first {r['hands']} hands of the same input through classify + to_scoring_format + UnifiedScorer.score_hand"}
                assert r["checksum"] == int(out["score"][: 1 << 15].sum().item()), "hands checksum differs from the reference"
                res["cpu_baseline"]["checksum_matches_gpu"] = True
        except AssertionError:
            raise
        except Exception as e:  # baseline is a reported number, never a reason to lose the bench line
            res["cpu_baseline"] = {"unavailable": repr(e)}
    return res


def bench_host_facing():
    """The reference-shaped entry points a Python caller drives from the host, random legal play: the N = 1
    Gymnasium facade (one C call per step through bgym_vec_step_host) and the SB3 VecEnv adapter at the
    reference trainers' slab sizes.  Wall clock, everything included (observation dicts built on the host)."""
    import numpy as np
    from balatro_gym_b200.env import BalatroEnv
    from balatro_gym_b200.sb3_vec_env import BalatroSB3VecEnv
    rng = np.random.default_rng(0)
    env = BalatroEnv(seed=3)
    obs, _ = env.reset(seed=3)

    def run(k):
        nonlocal obs
        t0 = time.perf_counter()
        for _ in range(k):
            obs, r, term, trunc, info = env.step(int(rng.choice(np.flatnonzero(obs["action_mask"]))))
            if term:
                obs, _ = env.reset()
        return k / (time.perf_counter() - t0)
    run(200)
    out = {"gym_facade_n1": {"value": run(4000), "unit": "env-steps/s"}}
    env.close()
    for n in (8, 64):
        v = BalatroSB3VecEnv(n, seed=1)
        o = v.reset()

        def vrun(k):
            nonlocal o
            t0 = time.perf_counter()
            for _ in range(k):
                m = o["action_mask"]
                o, r, d, i = v.step((rng.random(m.shape) * m).argmax(axis=1))
            return k * n / (time.perf_counter() - t0)
        vrun(20)
        out[f"sb3_vec_env_n{n}"] = {"value": vrun(300), "unit": "env-steps/s"}
        v.close()
    out["note"] = "reference on one host core: ~1.3e4 env-steps/s per BalatroEnv process (SURVEY 6)"
    return out


def bench_hands_jokers(torch, dev, peak):
    """The general scoring kernel (joker interpreter K4): 2^22 plays of 1..8 cards, 5 distinct random jokers,
    config-3 modifiers, random hand levels.  64 B per hand (+8 joker ids +16 mods +8 x_mult)."""
    from balatro_gym_b200.score import score_hands
    n = 1 << 22
    g = torch.Generator(device=dev); g.manual_seed(0xC3)
    cards = torch.rand((n, 52), device=dev, generator=g).topk(8, dim=1).indices.to(torch.uint8).contiguous()
    ncards = torch.randint(1, 9, (n,), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
    jokers = torch.zeros((n, 8), dtype=torch.uint8, device=dev)
    jokers[:, :5] = (torch.rand((n, 145), device=dev, generator=g).topk(5, dim=1).indices + 1).to(torch.uint8)
    z = torch.zeros((n, 8), dtype=torch.int64, device=dev)
    enh = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.25, torch.randint(1, 9, (n, 8), device=dev, generator=g), z)
    ed = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.1, torch.randint(1, 4, (n, 8), device=dev, generator=g), z)
    seal = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.1, torch.randint(1, 5, (n, 8), device=dev, generator=g), z)
    m = (enh << 6) | (ed << 10) | (seal << 13)
    mods = torch.where(m >= 2 ** 15, m - 2 ** 16, m).to(torch.int16).contiguous()
    levels = torch.randint(1, 6, (n, 12), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
    out = score_hands(cards, mods8=mods, n_cards=ncards, jokers8=jokers, levels12=levels)
    for _ in range(3):
        score_hands(cards, mods8=mods, n_cards=ncards, jokers8=jokers, levels12=levels, out=out)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        score_hands(cards, mods8=mods, n_cards=ncards, jokers8=jokers, levels12=levels, out=out)
    e1.record(); torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 10
    ach = n * 64 / (ms / 1e3) / 1e9
    return {"value": n / (ms / 1e3), "unit": "hands/s", "n_hands": n, "ms_per_launch": ms,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "bytes_per_unit": 64,
                         "kernel": "score_hands_kernel (joker interpreter)",
                         "note": "bound by instruction issue of a divergent interpreter (per-lane card counts; see profiles/r02_ncu_summary.md for active lanes per instruction), not by HBM"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--hands", type=int, default=1 << 24)
    ap.add_argument("--burn-in", type=int, default=150)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-hands", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--no-facade", action="store_true")
    ap.add_argument("--ppo-envs", type=int, default=1 << 19, help="envs per GPU of the PPO rollout block (configs[4]: 2^22 over 8)")
    ap.add_argument("--ppo-steps", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    rc = run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
