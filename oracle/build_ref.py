#!/usr/bin/env python
"""Byte-compile the reference's hot-path modules into oracle/_ref/ (git-ignored, travels to the GPU box).

The reference is pure Python, so "building" it means compiling its sources — where they lie under
/root/reference — to sourceless .pyc files.  No reference SOURCE is copied into the repo; the output
directory is git-ignored and only used by the CPU baseline (`bench.py --impl reference`) and by
tests that want the real reference beside the C oracle on the GPU box.
"""
import os
import py_compile
import sys

SRC = "/root/reference/balatro_gym"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "balatro_gym")
MODULES = ["__init__", "env", "balatro_game", "scoring_engine", "cards", "constants", "shop", "jokers",
           "planets", "consumables", "unified_scoring", "complete_joker_effects", "boss_blinds",
           "balatro_env_2", "balatro_sim"]


def main():
    if not os.path.isdir(SRC):
        print("build_ref: /root/reference not mounted; keeping whatever oracle/_ref already holds")
        return 0
    os.makedirs(DST, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(SRC, m + ".py"), cfile=os.path.join(DST, m + ".pyc"),
                           dfile=f"balatro_gym/{m}.py", doraise=True, optimize=0)
    # the same files as one zip archive (zipimport loads sourceless .pyc): a second carrier in case
    # a sync tool filters *.pyc files
    import zipfile
    zpath = os.path.join(HERE, "_ref", "balatro_gym_ref.zip")
    with zipfile.ZipFile(zpath, "w", zipfile.ZIP_DEFLATED) as z:
        for m in MODULES:
            z.write(os.path.join(DST, m + ".pyc"), f"balatro_gym/{m}.pyc")
    print(f"build_ref: {len(MODULES)} modules -> {DST} and {zpath}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
