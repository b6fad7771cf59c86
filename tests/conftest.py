import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (from /root/reference here, or oracle/_ref when that was built)."""
    from oracle import refenv
    if refenv.reference_available() is None:
        pytest.skip("reference not available (neither /root/reference nor oracle/_ref)")
    return refenv.load_reference()


def load_trace(name):
    import numpy as np
    from balatro_gym_b200 import layout as L
    z = np.load(os.path.join(REPO, "tests", "golden", f"trace_{name}.npz"))
    out = {k: z[k] for k in z.files}
    T, E = out["action"].shape
    out["draws"] = out["draws"].reshape(T, -1).view(L.DRAWS_DTYPE).reshape(T, E)
    out["state"] = out["state"].reshape(T, -1).view(L.STATE_DTYPE).reshape(T, E)
    out["obs"] = out["obs"].reshape(T, -1).view(L.OBS_DTYPE).reshape(T, E)
    out["init_state"] = out["init_state"].reshape(E, -1).view(L.STATE_DTYPE).reshape(E)
    return out


STATE_SKIP = ("rng_seed", "rng_ctr", "ep_len", "episode")


def assert_records_equal(a, b, dtype, skip=(), where=""):
    import numpy as np
    for name in dtype.names:
        if name in skip:
            continue
        if not np.array_equal(a[name], b[name]):
            bad = np.argwhere(np.asarray(a[name] != b[name]).reshape(len(a), -1).any(axis=1)).ravel()
            i = int(bad[0])
            raise AssertionError(f"{where}: field {name} differs for {len(bad)} records; first idx {i}: "
                                 f"{a[name][i].tolist()} vs {b[name][i].tolist()}")


def reward_close(r_ref, r_got, ante):
    """Rewards are bit-exact except the ante>3 play branch (np.log10 is SVML on AVX512 hosts,
    1 ulp off libm / CUDA on ~2% of arguments): 1e-12 relative there."""
    import numpy as np
    exact = r_ref == r_got
    close = np.abs(r_ref - r_got) <= 1e-12 * np.maximum(1.0, np.abs(r_ref))
    return np.where(ante > 3, close, exact)
