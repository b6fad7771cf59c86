"""ctypes binding of the C-ABI shared library (include/bgym.h -> libbgym.so).

The library is built in-tree by `build()` (nvcc, sm_100a).  There is NO fallback: if the
library is missing or a CUDA device is absent, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

from . import layout as L

_PKG = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "libbgym.so")
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("bgym_kernels.cu", "bgym_step_part.cuh", "bgym_rollout.cuh", "bgym_env.cuh", "bgym_device.cuh")]
HEADERS = [os.path.join(_REPO, "include", f) for f in ("bgym.h", "bgym_tables.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # Python-float semantics: no fused multiply-add contraction anywhere on the path
              "-fmad=false", "-shared", "-Xcompiler", "-fPIC"]

_lib = None
_policy_lib = None
POLICY_SOURCE = os.path.join(_PKG, "csrc", "bgym_policy.cu")
POLICY_HEADER = os.path.join(_REPO, "include", "bgym_policy.h")
POLICY_SO_PATH = os.path.join(_PKG, "libbgym_policy.so")


class BgymError(RuntimeError):
    pass


def source_hash() -> str:
    """sha256 over the kernel sources and headers the library is built from: ties measured artefacts (the ncu DRAM
    traffic bench.py reports, profiles/*_traffic.json) to the code they were measured on."""
    import hashlib
    h = hashlib.sha256()
    # (the rollout / policy kernels are not part of what those captures measure: env step and hand scoring)
    for p in [q for q in SOURCES if not q.endswith(("bgym_rollout.cuh", "bgym_policy.cuh"))] + HEADERS:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/bgym_kernels.cu for sm_100a into balatro_gym_b200/libbgym.so (in-tree)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise BgymError("nvcc not found: cannot build libbgym.so")
    extra = os.environ.get("BGYM_NVCC_EXTRA", "").split()        # experiments only (e.g. -DBGYM_GATHER_CTAS=6)
    # several ranks of one job may get here at once (torchrun): build under a file lock, into a temporary file that
    # replaces the library atomically, so that nobody ever loads a half-written .so
    import fcntl
    with open(SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another process built it while we waited
                return SO_PATH
            tmp = f"{SO_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SOURCES[0]]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise BgymError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, SO_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    if verbose:
        print(res.stderr)
    return SO_PATH


_vp, _i64, _i32, _u32, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_uint64

# every symbol include/bgym.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "bgym_abi_version": (_i32, []),
    "bgym_last_error": (C.c_char_p, []),
    "bgym_device_count": (_i32, []),
    "bgym_set_option": (_i32, [_i32, _i64]),
    "bgym_reset": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "bgym_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "bgym_sync_state": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "bgym_sync_obs": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "bgym_pack_dirty_obs": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "bgym_scatter_dirty_obs": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "bgym_release_stream": (_i32, [_vp]),
    "bgym_action_mask": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "bgym_sample_actions": (_i32, [_vp, _i64, _vp, _u32, _u64, _i64, _vp]),
    "bgym_sample_actions_ctr": (_i32, [_vp, _i64, _vp, _u32, _vp, _i64, _vp]),
    "bgym_score_hands": (_i32, [_vp] * 12 + [_u32, _i64, _i32, _vp]),
    "bgym_episode_stats": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bgym_featurize": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "bgym_policy_first_layer": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bgym_masked_sample": (_i32, [_vp, _i32, _vp, _i64, _vp, _u32, _u64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "bgym_gae": (_i32, [_vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _i64, _i64, _vp]),
    "bgym_vec_create": (_i32, [C.POINTER(_vp), _i64, _i32]),
    "bgym_vec_destroy": (_i32, [_vp]),
    "bgym_vec_reset_host": (_i32, [_vp, _vp, _vp, _vp]),
    "bgym_vec_reset_masked_host": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "bgym_vec_step_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32]),
    "bgym_vec_pointers": (_i32, [_vp] + [C.POINTER(_vp)] * 7),
    "bgym_vec_get_state": (_i32, [_vp, _vp]),
    "bgym_vec_set_state": (_i32, [_vp, _vp]),
}


def load():
    """Load libbgym.so (building it first if sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
        elif not os.path.exists(SO_PATH):
            raise BgymError(f"{SO_PATH} is missing and nvcc is not available; run __graft_entry__.build()")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.bgym_abi_version() != 2:
        raise BgymError("libbgym.so ABI version mismatch")
    _lib = lib
    return lib


def build_policy(force: bool = False) -> str:
    """Compile csrc/bgym_policy.cu (the fused tcgen05 policy forward, include/bgym_policy.h) for sm_100a into
    balatro_gym_b200/libbgym_policy.so.  Its own library: the env path does not depend on it."""
    if not force and os.path.exists(POLICY_SO_PATH) and all(os.path.getmtime(p) <= os.path.getmtime(POLICY_SO_PATH)
                                                             for p in (POLICY_SOURCE, POLICY_HEADER)):
        return POLICY_SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise BgymError("nvcc not found: cannot build libbgym_policy.so")
    import fcntl
    with open(POLICY_SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            tmp = f"{POLICY_SO_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                   "-I", os.path.join(_REPO, "include"), "-o", tmp, POLICY_SOURCE]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise BgymError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, POLICY_SO_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return POLICY_SO_PATH


POLICY_SYMBOLS = {
    "bgym_policy_program": (_i32, [_vp, _vp, _vp]),
    "bgym_policy_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "bgym_policy_last_error": (C.c_char_p, []),
}


def load_policy():
    """Load libbgym_policy.so (building it first when the source is newer and nvcc is present)."""
    global _policy_lib
    if _policy_lib is not None:
        return _policy_lib
    stale = not os.path.exists(POLICY_SO_PATH) or any(os.path.getmtime(p) > os.path.getmtime(POLICY_SO_PATH) for p in (POLICY_SOURCE, POLICY_HEADER))
    if stale:
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build_policy(force=True)
        elif not os.path.exists(POLICY_SO_PATH):
            raise BgymError(f"{POLICY_SO_PATH} is missing and nvcc is not available; run __graft_entry__.build()")
    lib = C.CDLL(POLICY_SO_PATH)
    for name, (res, args) in POLICY_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _policy_lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().bgym_last_error().decode()
        raise BgymError(f"{what} failed (rc={rc}): {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise BgymError("balatro_gym_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
