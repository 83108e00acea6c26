"""TEST INFRASTRUCTURE ONLY -- stage an UNTOUCHED copy of the reference for the GPU box.

`/root/reference` exists only in the build container; `gpurun` ships `/root/repo` and
nothing else.  This script copies `/root/reference/{sbp_env,maps,tests}` byte for byte
into `oracle/_ref/` (git-ignored: it never enters the history; NOT gpurun-ignored: it
travels with the snapshot like the built `.so` files), so that on the B200

  * the `-m gpu` planner tests can run the UNMODIFIED reference planners twice -- with
    the reference's own checker and with the `sbp_env_b200` checker + hooks on real CUDA --
    and compare whole seeded trajectories (tests/test_gpu_reference_planners.py), and
  * `bench.py --impl reference` / the `cpu_baseline` leg can time the reference's own
    NumPy / Python functions on the box's host cores (`cpu_baseline.kind = "reference"`).

Nothing under `oracle/_ref/` is imported by the product package.  `__graft_entry__.build()`
runs this when `/root/reference` is present; on the GPU box the staged copy is used as is.

    python oracle/stage_ref.py [--check]
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SBP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PARTS = ("sbp_env", "maps", "tests")


def _same_tree(a: str, b: str) -> bool:
    if not os.path.isdir(b):
        return False
    cmp = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if cmp.left_only or cmp.right_only or cmp.funny_files:
        return False
    _, mismatch, errors = filecmp.cmpfiles(a, b, cmp.common_files, shallow=False)
    if mismatch or errors:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in cmp.common_dirs)


def staged() -> bool:
    return os.path.isdir(os.path.join(DST, "sbp_env")) and os.path.isdir(os.path.join(DST, "maps"))


def stage(verbose: bool = True) -> bool:
    """Copy the reference parts that are not yet identical; returns True when a staged copy
    exists afterwards."""
    if not os.path.isdir(os.path.join(SRC, "sbp_env")):
        if verbose:
            print(f"stage_ref: {SRC} absent; using the staged copy as is" if staged()
                  else f"stage_ref: {SRC} absent and nothing staged")
        return staged()
    os.makedirs(DST, exist_ok=True)
    for part in PARTS:
        a, b = os.path.join(SRC, part), os.path.join(DST, part)
        if not os.path.isdir(a):
            continue
        if _same_tree(a, b):
            continue
        if os.path.isdir(b):
            shutil.rmtree(b)
        shutil.copytree(a, b, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if verbose:
            print(f"stage_ref: {a} -> {b}")
    return staged()


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = all(_same_tree(os.path.join(SRC, p), os.path.join(DST, p)) for p in PARTS
                 if os.path.isdir(os.path.join(SRC, p)))
        print("staged copy identical to the reference" if ok else "staged copy DIFFERS / missing")
        sys.exit(0 if ok else 1)
    sys.exit(0 if stage() else 1)
