"""CPU baseline: the UNMODIFIED reference env on the host cores — TEST / BENCH INFRASTRUCTURE.

One BalatroEnv per worker process (that is the reference's own vectorisation: SB3 SubprocVecEnv,
hpc_train.py:62), free-running without per-step IPC, which is the upper bound for an
AsyncVectorEnv.  Random legal actions from obs['action_mask'].
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


_ALL_CONSUMABLES = None


def _inject_c3(env, R, rng, consumables=False):
    """SURVEY 8(d) C3 state generator (Appendix E recipe), per episode; consumables=True adds config 4's two random
    consumables over all 52 names (the device-side counterpart is BGYM_FLAG_GEN_C3 | BGYM_FLAG_GEN_CONS)."""
    global _ALL_CONSUMABLES
    if consumables:
        if _ALL_CONSUMABLES is None:
            from balatro_gym_b200 import layout as L
            _ALL_CONSUMABLES = L.TAROT_NAMES + L.PLANET_NAMES + L.SPECTRAL_NAMES
        env.state.consumables = [_ALL_CONSUMABLES[int(rng.integers(0, 52))] for _ in range(2)]
    costs = [j for j in R.jokers.JOKER_LIBRARY if j.base_cost > 0]
    idx = rng.choice(len(costs), size=5, replace=False)
    env.state.jokers = [costs[i] for i in idx]
    C = R.cards
    for i in range(52):
        enh = int(rng.integers(1, 9)) if rng.random() < 0.25 else 0
        ed = int(rng.integers(1, 4)) if rng.random() < 0.1 else 0
        seal = int(rng.integers(1, 5)) if rng.random() < 0.1 else 0
        if enh or ed or seal:
            env.state.card_states[i] = C.CardState(i, C.Enhancement(enh), C.Edition(ed), C.Seal(seal))


def _env_worker(rank, config, steps_per_round, rounds, q):
    from oracle.refenv import load_reference
    R = load_reference()
    rng = np.random.default_rng(1000 + rank)
    seed = 1 + rank * 100003
    env = R.BalatroEnv(seed=seed)

    def new_episode():
        nonlocal seed
        seed += 1
        obs, _ = env.reset(seed=seed)
        if config != "c1":
            _inject_c3(env, R, rng, consumables=(config == "c4"))
        return obs

    obs = new_episode()
    first = True
    out = []
    for r in range(rounds):
        t0 = time.perf_counter()
        for _ in range(steps_per_round):
            if first and config == "c1":
                a = 45
            else:
                a = int(rng.choice(np.flatnonzero(obs["action_mask"])))
            first = False
            try:
                obs, rew, term, trunc, info = env.step(a)
            except Exception:      # the reference raises on some consumables: SafeBalatroEnv ends the episode
                term = True
            if term:
                obs = new_episode()
                first = True
        out.append(time.perf_counter() - t0)
    q.put((rank, out))


def run_env_baseline(config="c4", cores=None, steps_per_round=1024, rounds=4, warmup_rounds=1):
    """Returns dict(per_round_s=[...], steps_per_round_total, cores).  A 'round' is steps_per_round
    env-steps on every worker."""
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_env_worker, args=(i, config, steps_per_round, rounds + warmup_rounds, q)) for i in range(cores)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    # a round completes when the slowest worker finishes it
    times = np.array([r[1] for r in sorted(res)])        # [cores, rounds]
    per_round = times.max(axis=0)[warmup_rounds:]
    return dict(per_round_s=per_round.tolist(), steps_per_round_total=steps_per_round * cores, cores=cores)


def _hands_worker(rank, cards, q):
    from oracle.refscore import RefScorer
    rs = RefScorer()
    t0 = time.perf_counter()
    acc = 0
    for row in cards:
        acc += rs.score(row[:5].tolist())["score"]
    q.put((rank, time.perf_counter() - t0, acc))


def run_hands_baseline(cards8, cores=None):
    """Score the given hands (5-card plays, no jokers) with the reference's O2 path split over cores."""
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    chunks = np.array_split(cards8, cores)
    procs = [ctx.Process(target=_hands_worker, args=(i, chunks[i], q)) for i in range(cores)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    return dict(seconds=max(r[1] for r in res), hands=len(cards8), cores=cores, checksum=sum(r[2] for r in res))
