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
    assert lib.bgym_abi_version() == 1
    assert lib.bgym_device_count() >= 0


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    assert lib.bgym_step(None, None, None, None, None, None, None, None, None, 4, 0, None) < 0
    assert b"bgym_step" in lib.bgym_last_error()
    assert lib.bgym_reset(None, None, None, None, None, None, 4, 0, None) < 0


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
                       ("BgymDraws", L.DRAWS_DTYPE), ("BgymScoreCtx", L.SCORE_CTX_DTYPE)):
        size, offs = _c_offsets(struct, dt.names)
        assert size == dt.itemsize, struct
        assert offs == [dt.fields[n][1] for n in dt.names], struct
    # device record strides are odd multiples of 16 B (bank-conflict-free 128-bit access per lane)
    for dt, size in ((L.HOT_DTYPE, 144), (L.COLD_DTYPE, 176), (L.OBS_DTYPE, 176)):
        assert dt.itemsize == size and dt.itemsize % 32 == 16
    assert L.STATE_DTYPE.itemsize == 320 == L.HOT_BYTES + L.COLD_BYTES


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
