"""TEST INFRASTRUCTURE ONLY -- container-side loader for the *unmodified* reference.

`/root/reference` (soraxas/sbp-env v2.0.2) is pure Python but imports five
third-party modules that are absent from this image (rtree, pygame, overrides,
docopt, SALib).  None of them takes part in the collision / nearest-neighbour
arithmetic, so they are replaced by inert stand-ins in ``sys.modules`` and the
reference is then imported from where it lies (nothing is copied).

Used by ``oracle/gen_golden.py`` (fixture generation), by the cross-checks in
``tests/`` and by the reference arm of ``bench.py``.  On the GPU box ``/root/reference``
does not exist: there the untouched copy staged by ``oracle/stage_ref.py`` under
``oracle/_ref/`` (git-ignored, shipped by gpurun) is imported instead; tests that need
the reference are skipped when neither is present.  Never imported by the product package.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")


def _pick_root() -> str:
    env = os.environ.get("SBP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/sbp_env"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sbp_env"))


def _install_stubs() -> None:
    if "overrides" not in sys.modules:
        m = types.ModuleType("overrides")
        m.overrides = lambda f: f
        sys.modules["overrides"] = m

    if "pygame" not in sys.modules:
        pg = mock.MagicMock(name="pygame")
        loc = types.ModuleType("pygame.locals")
        loc.__all__ = []
        pg.locals = loc
        sys.modules["pygame"] = pg
        sys.modules["pygame.locals"] = loc

    if "rtree" not in sys.modules:
        # The planners never insert into the R-tree; a brute
        # force stand-in keeps `utils/common.py:7` importable.
        rt = types.ModuleType("rtree")
        idx = types.ModuleType("rtree.index")

        class Property:  # noqa: D401 - attribute bag
            dimension = 2

        class Index:
            def __init__(self, interleaved=True, properties=None):
                self._items = []

            def insert(self, id_, coords, obj=None):
                self._items.append((id_, tuple(coords), obj))

            def nearest(self, coords, num_results=1, objects=False):
                d = len(coords) // 2
                q = coords[:d]

                def key(it):
                    c = it[1][:d]
                    return sum((a - b) ** 2 for a, b in zip(c, q))

                for it in sorted(self._items, key=key)[:num_results]:
                    yield it[2] if objects == "raw" else it[0]

        idx.Property = Property
        idx.Index = Index
        rt.index = idx
        sys.modules["rtree"] = rt
        sys.modules["rtree.index"] = idx

    if "SALib" not in sys.modules:
        sal = types.ModuleType("SALib")
        smp = types.ModuleType("SALib.sample")
        for name in ("saltelli", "sobol_sequence", "latin", "finite_diff", "fast_sampler"):
            setattr(smp, name, mock.MagicMock(name=f"SALib.sample.{name}"))
        sal.sample = smp
        sys.modules["SALib"] = sal
        sys.modules["SALib.sample"] = smp

    if "docopt" not in sys.modules:
        dm = types.ModuleType("docopt")

        class DocoptExit(SystemExit):
            pass

        def docopt(*a, **k):
            raise DocoptExit("docopt is stubbed in this image")

        dm.docopt = docopt
        dm.DocoptExit = DocoptExit
        sys.modules["docopt"] = dm


def import_reference():
    """Return the imported, unmodified ``sbp_env`` package of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the tree is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sbp_env  # noqa: F401

    return sbp_env


def reference_map_path(name: str) -> str:
    return os.path.join(REFERENCE_ROOT, "maps", name)
