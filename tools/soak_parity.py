#!/usr/bin/env python
"""Long CUDA-vs-oracle soak (native Philox mode): N envs x T steps, every record of every step compared
bit for bit (state, observation, reward, terminated, info), with the fused random policy, in-kernel
autoreset, the config-3 state generator, and periodic bursts of arbitrary (masked / out-of-range) actions.
Runs on the GPU box; writes a summary to gpurun_out/soak_parity.txt (copied to profiles/ by hand).

    python tools/soak_parity.py --envs 16384 --steps 1000
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--one-launch", action="store_true", help="exercise the small-slab kernel instead of the multi-pass step")
    ap.add_argument("--generator", default="c4", choices=["none", "c3", "c4"],
                    help="device-side state generator applied by reset and by every in-kernel autoreset (mirrored by the oracle)")
    args = ap.parse_args()
    import torch
    from balatro_gym_b200 import BalatroVecEnv, layout as L
    from oracle import coracle
    from conftest import assert_records_equal

    n = args.envs
    import balatro_gym_b200
    assert balatro_gym_b200.load().bgym_set_option(1, (1 << 40) if args.one_launch else 0) == 0
    gen = None if args.generator == "none" else args.generator
    gflags = L.GENERATORS[gen]
    v = BalatroVecEnv(n, seed=args.seed, autoreset=True, generator=gen)
    v.reset()
    ov = coracle.OracleVec(n)
    coracle.reset(ov.state, ov.obs, np.arange(args.seed, n + args.seed), flags=gflags)
    assert_records_equal(ov.state, v.state_numpy(), L.STATE_DTYPE, (), "reset state")
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    episodes = 0
    ulp_cases = 0
    phases = np.zeros(4, dtype=np.int64)
    max_ante = 1
    for t in range(args.steps):
        if t % 11 == 5:
            act = rng.integers(-2, 62, size=n).astype(np.int32)
            v.step(torch.from_numpy(act).cuda())
            ov.step(act, flags=L.FLAG_AUTORESET | gflags)
        else:
            v.step(random_policy=True)
            oact = np.zeros(n, np.int32)
            coracle.step(ov.state, oact, ov.obs, ov.reward, ov.terminated, ov.truncated, ov.info, None,
                         flags=L.FLAG_AUTORESET | 4 | gflags)
            assert np.array_equal(oact, v.actions.cpu().numpy()), f"policy actions differ at step {t}"
        st = v.state_numpy()
        assert_records_equal(ov.state, st, L.STATE_DTYPE, (), f"step {t} state")
        assert_records_equal(ov.obs, v.obs_numpy(), L.OBS_DTYPE, (), f"step {t} obs")
        # rewards: bit-exact, except the 3*log10(score) term of a played hand at ante > 3, where CUDA's
        # log10 and the host libm may differ in the last place (tests/conftest.py reward_close): 1e-12 relative
        r, inf = v.reward.cpu().numpy(), v.info_numpy()
        same = ov.reward.view(np.uint64) == r.view(np.uint64)
        ok = same | (((inf["flags"] & L.F_PLAYED) != 0) & (np.abs(ov.reward - r) <= 1e-12 * np.maximum(1, np.abs(r))))
        assert ok.all(), f"step {t} reward"
        ulp_cases += int((~same).sum())
        assert np.array_equal(ov.terminated, v.terminated.cpu().numpy()), f"step {t} terminated"
        assert_records_equal(ov.info, inf, L.INFO_DTYPE, (), f"step {t} info")
        episodes += int(ov.terminated.sum())
        phases += np.bincount(st["phase"] & 3, minlength=4)
        max_ante = max(max_ante, int(st["ante"].max()))
    msg = (f"soak ({'one-launch small-slab kernel' if args.one_launch else 'multi-pass step: main pass on the toggle records + list kernels'}, generator {args.generator}): {n} envs x {args.steps} steps = {n * args.steps} env-steps, 0 mismatches "
           f"(state, obs, terminated, info bit-exact every step; rewards bit-exact except {ulp_cases} last-place log10 cases "
           f"within 1e-12); episodes finished {episodes}; "
           f"env-steps by phase PLAY/SHOP/BLIND_SELECT/PACK_OPEN = {phases.tolist()}; max ante reached {max_ante}; "
           f"wall {time.time() - t0:.1f} s")
    print(msg)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/soak_parity.txt", "a").write(msg + "\n")


if __name__ == "__main__":
    main()
