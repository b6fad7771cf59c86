"""10^6 fresh reference steps replayed on the CUDA path, under the driver's eyes (BASELINE north_star: "bit-exact
parity with balatro_gym on 10^6 replayed steps").  The unmodified reference (oracle/_ref on the GPU box) plays
episodes of the four configurations in worker processes with every RNG tapped; CUDA replays actions + draws and every
state record, observation, reward, termination and score breakdown is compared (tools/lockstep_cuda.py).
BGYM_LOCKSTEP_STEPS overrides the step count."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_one_million_reference_steps_replay_on_cuda(reference, capsys):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import balatro_gym_b200
    import lockstep_cuda
    steps = int(os.environ.get("BGYM_LOCKSTEP_STEPS", "1000000"))
    lines = []
    try:
        done = lockstep_cuda.run(steps, seed0=700001, log=lines.append)
    finally:
        assert balatro_gym_b200.load().bgym_set_option(1, 65536) == 0
    total = sum(done.values())
    with capsys.disabled():
        print("\n" + "\n".join(lines))
        print(f"LOCKSTEP reference-vs-CUDA: {total} steps, 0 mismatches")
    assert total >= steps and all(v > 0 for v in done.values())
