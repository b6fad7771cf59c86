"""balatro_gym_b200 — B200-native batched Balatro environment (drop-in for the step path of
cassiusfive/balatro-gym).  Public surface:

    BalatroVecEnv      vector-env entry point (device-resident state, sm_100a kernels)
    HostMirror         pinned-host copy of a BalatroVecEnv's step results, kept current by observation deltas
    BalatroEnv         Gymnasium facade over one env; make("BalatroGym-v0")
    score_hands        batched hand scoring microkernel
    BalatroSB3VecEnv   Stable-Baselines3 VecEnv protocol over the device env (sb3_vec_env.py)
    validate           BalatroEnvValidator port + batched determinism / masking / checkpoint checks
    build              compile the CUDA library in-tree
"""
from . import layout
from ._lib import build, load, BgymError, SO_PATH

__all__ = ["BalatroVecEnv", "HostMirror", "BalatroEnv", "BalatroSB3VecEnv", "make_sb3_vec_env", "make", "make_balatro_env", "score_hands", "build", "load",
           "BgymError", "layout", "SO_PATH"]


def __getattr__(name):  # lazy: importing the package must not need torch/CUDA
    if name in ("BalatroVecEnv", "HostMirror"):
        from . import vec_env
        return getattr(vec_env, name)
    if name in ("BalatroEnv", "make", "make_balatro_env", "register_envs", "reference_deck"):
        from . import env
        return getattr(env, name)
    if name in ("BalatroSB3VecEnv", "make_sb3_vec_env"):
        from . import sb3_vec_env
        return getattr(sb3_vec_env, name)
    if name == "score_hands":
        from .score import score_hands
        return score_hands
    raise AttributeError(name)
