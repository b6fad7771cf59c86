"""The C-ABI library loads, exports every symbol include/bgym.h declares, and the numpy dtypes in
balatro_gym_b200/layout.py match the C structs byte for byte.  No GPU needed, no compute calls."""
import os
import re
import subprocess
import tempfile

import numpy as np

from conftest import REPO
from balatro_gym_b200 import layout as L
from balatro_gym_b200 import _lib


def test_library_exports_all_declared_symbols():
    lib = _lib.load()
    hdr = open(os.path.join(REPO, "include", "bgym.h")).read()
    declared = set(re.findall(r"\b(bgym_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bgym_abi_version() == 2
    assert lib.bgym_device_count() >= 0


def test_policy_library_exports_its_header_and_program():
    """libbgym_policy.so (include/bgym_policy.h, the fused tcgen05 policy forward): loads, exports every declared symbol,
    and its weight-tile program is self-consistent (no GPU needed: the program is host code)."""
    import ctypes as C
    lib = _lib.load_policy()
    hdr = open(os.path.join(REPO, "include", "bgym_policy.h")).read()
    declared = set(re.findall(r"\b(bgym_policy_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.POLICY_SYMBOLS), declared ^ set(_lib.POLICY_SYMBOLS)
    step_dt = np.dtype([(k, "<i4") for k in ("offset", "bytes", "layer", "n0", "n", "kb", "a_kb", "col", "first", "last", "group", "_pad")])
    steps = np.zeros(80, dtype=step_dt)
    wb, bf = C.c_int64(0), C.c_int64(0)
    n = lib.bgym_policy_program(steps.ctypes.data, C.addressof(wb), C.addressof(bf))
    steps = steps[:n]
    assert n == 72 and steps.itemsize == 48
    assert np.array_equal(steps["offset"], np.concatenate([[0], np.cumsum(steps["bytes"])[:-1]])) and wb.value == int(steps["bytes"].sum())
    assert (steps["bytes"] == 128 * steps["n"]).all() and (steps["n"] % 16 == 0).all() and steps["n"].max() == 256
    assert (steps["col"] + steps["n"] <= 512).all() and (steps["a_kb"] < 8).all()          # tensor memory columns, activation K-blocks
    assert int(steps["last"].sum()) == 7 and steps["last"][-1] == 1 and bf.value == 512 * 7
    assert (np.diff(steps["group"]) >= 0).all()
    assert lib.bgym_policy_forward(None, None, None, None, None, 4, None) < 0 and b"bad arguments" in lib.bgym_policy_last_error()


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    assert lib.bgym_step(None, None, None, None, None, None, None, None, None, None, None, None, 4, 0, None) < 0
    assert b"bgym_step" in lib.bgym_last_error()
    assert lib.bgym_reset(None, None, None, None, None, None, None, None, 4, 0, None) < 0
    assert lib.bgym_sync_state(None, None, 4, 0, None) < 0 and lib.bgym_sync_obs(None, None, 4, 0, None) < 0
    assert lib.bgym_pack_dirty_obs(None, None, None, None, 4, 4, None) < 0 and lib.bgym_scatter_dirty_obs(None, 4, None, None, None) < 0


def _c_offsets(struct, fields):
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "bgym.h"\nint main(){\n'
    src += f'printf("%zu\\n", sizeof({struct}));\n'
    for f in fields:
        src += f'printf("%zu\\n", offsetof({struct}, {f}));\n'
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "o.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "o")
        subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), "-o", exe, c])
        vals = [int(x) for x in subprocess.check_output([exe]).split()]
    return vals[0], vals[1:]


def test_struct_layouts_match_numpy_dtypes():
    for struct, dt in (("BgymState", L.STATE_DTYPE), ("BgymHot", L.HOT_DTYPE), ("BgymCold", L.COLD_DTYPE),
                       ("BgymObs", L.OBS_DTYPE), ("BgymInfo", L.INFO_DTYPE),
                       ("BgymDraws", L.DRAWS_DTYPE), ("BgymScoreCtx", L.SCORE_CTX_DTYPE),
                       ("BgymTog", L.TOG_DTYPE), ("BgymSel", L.SEL_DTYPE)):
        size, offs = _c_offsets(struct, dt.names)
        assert size == dt.itemsize, struct
        assert offs == [dt.fields[n][1] for n in dt.names], struct
    # device record strides are odd multiples of 16 B (bank-conflict-free 128-bit access per lane)
    for dt, size in ((L.HOT_DTYPE, 144), (L.COLD_DTYPE, 176), (L.OBS_DTYPE, 176)):
        assert dt.itemsize == size and dt.itemsize % 32 == 16
    assert L.STATE_DTYPE.itemsize == 320 == L.HOT_BYTES + L.COLD_BYTES
    # the toggle record's first half is hot bytes 16..31, field for field; the selection record's fields are the
    # observation record's
    for name in L.TOG_OWNED_FIELDS:
        assert L.TOG_DTYPE.fields[name][1] + 16 == L.HOT_DTYPE.fields[name][1], name
        assert L.TOG_DTYPE.fields[name][0] == L.HOT_DTYPE.fields[name][0], name
    assert sorted(L.HOT_DTYPE.fields[n][1] for n in L.TOG_OWNED_FIELDS)[0] == 16
    assert sum(L.HOT_DTYPE.fields[n][0].itemsize for n in L.TOG_OWNED_FIELDS) == 16
    for name in L.SEL_DTYPE.names:
        assert L.SEL_DTYPE.fields[name][0] == L.OBS_DTYPE.fields[name][0], name
    assert L.TOG_DTYPE.itemsize == 32 and L.SEL_DTYPE.itemsize == 16


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "balatro_gym_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import balatro_gym_b200 as b
    with pytest.raises(b.BgymError):
        b.BalatroVecEnv(4)


def test_mask_word_expands_to_the_reference_action_mask():
    """obs['action_mask'] of the reference (int8[60]) is carried as one 64-bit word; the Python layers expand it."""
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 1 << 60, size=257, dtype=np.uint64)
    m = L.mask_from_bits(bits)
    assert m.shape == (257, 60) and m.dtype == np.int8
    back = (m.astype(np.uint64) << np.arange(60, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
    assert np.array_equal(back, bits)
    assert L.mask_from_bits(np.uint64(0b1011)).tolist()[:5] == [1, 1, 0, 1, 0]
    rec = np.zeros(3, dtype=L.OBS_DTYPE)
    rec["action_mask_bits"] = [1, 2, (1 << 59)]
    assert L.obs_value(rec, "action_mask")[:, [0, 1, 59]].tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    assert L.obs_value(rec, "money") is not None and "action_mask" in L.OBS_KEYS and "action_mask" not in L.OBS_DTYPE.names
    assert len(L.OBS_KEYS) == 31


def test_sb3_seed_chain_matches_the_kernels_constants():
    """next_episode_seed (host mirror of the in-kernel autoreset seed chain): fixed vectors."""
    from balatro_gym_b200.sb3_vec_env import next_episode_seed
    def ref(x):
        x = (x + 0x9E3779B9) & 0xFFFFFFFF
        x ^= x >> 16; x = (x * 0x85EBCA6B) & 0xFFFFFFFF
        x ^= x >> 13; x = (x * 0xC2B2AE35) & 0xFFFFFFFF
        x ^= x >> 16
        return x or 1
    xs = np.array([1, 2, 12345, 0xFFFFFFFF, 0x61C88647], dtype=np.uint32)
    assert next_episode_seed(xs).tolist() == [ref(int(x)) for x in xs]


def test_plain_c_caller_compiles_and_links():
    """examples/host_loop.c: the boundary is usable from C with nothing but include/bgym.h and libbgym.so."""
    _lib.load()
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "host_loop")
        subprocess.check_call(["gcc", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), "-o", exe,
                               os.path.join(REPO, "examples", "host_loop.c"), "-L", os.path.dirname(_lib.SO_PATH), "-lbgym",
                               "-Wl,-rpath," + os.path.dirname(_lib.SO_PATH)])
        rc = subprocess.call([exe, "16", "2"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert rc in (0, 2)     # 2 = "no CUDA device" (this container); 0 on a GPU box


def test_committed_ncu_traffic_matches_the_kernel_sources():
    """profiles/r02_traffic.json (the DRAM traffic bench.py reports in roofline.traffic) must have been captured on the
    kernel sources in the tree: a kernel edit without a fresh `ncu` capture fails here, loudly, instead of leaving a
    stale constant in the bench line."""
    import json
    import os
    from balatro_gym_b200 import _lib
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
    d = json.load(open(path))
    assert d["source_hash"] == _lib.source_hash(), \
        "kernel sources changed since the ncu capture: run tools/gpu_prof_part.sh on the GPU box, then tools/ncu_summary.py --traffic-json"
    assert d["step"]["traffic_bytes"] > 0 and len(d["step"]["kernels"]) == 9      # main pass + seven list kernels + the level-2 kernel


def test_host_mirror_flag_gather_constant():
    """HostMirror packs the eight selected_cards flags of a selection record into one byte with a 64-bit multiply: the
    product's top byte must be sum(flag_i << i) for every flag pattern (no carries between partial products)."""
    import numpy as np
    from balatro_gym_b200.vec_env import HostMirror
    flags = ((np.arange(256)[:, None] >> np.arange(8)) & 1).astype(np.uint8)          # all 256 patterns, little-endian bytes
    words = flags.view(np.uint64)[:, 0]
    with np.errstate(over="ignore"):
        top = ((words * np.uint64(HostMirror._GATHER_BITS)) >> np.uint64(56)).astype(np.uint8)
    assert np.array_equal(top, np.arange(256, dtype=np.uint8))
    assert np.array_equal(np.unpackbits(top[:, None], axis=1, bitorder="little"), flags)   # and the host-side expansion inverts it
