"""Minimal stand-in for the `gymnasium` package (not installed in this image, no network).

TEST INFRASTRUCTURE ONLY.  It exists so the unmodified reference
(`balatro_gym/balatro_env_2.py:24-25`, `balatro_gym/env.py`) can be imported as the parity
oracle and as the CPU baseline.  Only the surface the reference touches at import/construct
time is provided: Env, Wrapper, spaces.{Space,Discrete,Box,MultiBinary,MultiDiscrete,Dict,Tuple},
make, register.  If the real gymnasium is importable it is used instead (see oracle/refenv.py).
"""
from . import spaces  # noqa: F401

__version__ = "0.0-shim"


class Env:
    metadata = {}
    render_mode = None
    action_space = None
    observation_space = None

    def __init__(self, *a, **k):
        pass

    def reset(self, *, seed=None, options=None):
        return None, {}

    def step(self, action):
        raise NotImplementedError

    def render(self):
        return None

    def close(self):
        return None

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


_REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, kw = _REGISTRY[id]
    kw = dict(kw.get("kwargs", {}), **kwargs)
    if isinstance(entry_point, str):
        import importlib
        mod, _, attr = entry_point.partition(":")
        entry_point = getattr(importlib.import_module(mod), attr)
    return entry_point(**kw)
