"""
ref_shims.py — make the UNMODIFIED reference driver importable in this container and on the GPU box.

The reference imports seven third-party packages that are absent here and cannot be installed
(no network): Levenshtein, pysam, pyfaidx, pathos, natsort, magic, matplotlib.
install() puts the stand-in packages of baseline/shims/ (documented in its README) and the reference tree on
sys.path and registers oracle/_ref/SlidingWindow.*.so (the reference's own compiled Cython) as
scarmapper.SlidingWindow, so that `import scarmapper.INDEL_Processing` runs the reference code itself.
The reference tree is /root/reference in the build container and baseline/_ref/ScarMapper - a byte-for-byte,
git-ignored copy made by oracle/build_ref.build_baseline - where /root/reference does not exist (GPU box).
Used by tests/golden/make_golden.py (fixture generation), the tests that drive the reference, and bench.py's
reference arm.  TEST / BENCH INFRASTRUCTURE ONLY.

pathos.multiprocessing.Pool: install(real_pool=True) -> multiprocess.Pool, the class pathos itself re-exports
(forked workers, dill pickling); install() -> an in-process serial pool (fixture generation, quick tests).
"""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(_ROOT, "baseline", "shims")
BASELINE = os.path.join(_ROOT, "baseline", "_ref", "ScarMapper")   # byte-for-byte copy made by build_ref.build_baseline


def _resolve() -> str:
    """The reference tree: $SCARMAPPER_REFERENCE, else /root/reference (build container), else the git-ignored copy
    that travels to the GPU box (baseline/_ref/ScarMapper)."""
    for cand in (os.environ.get("SCARMAPPER_REFERENCE", ""), "/root/reference", BASELINE):
        if cand and os.path.isfile(os.path.join(cand, "scarmapper", "INDEL_Processing.py")):
            return cand
    return os.environ.get("SCARMAPPER_REFERENCE", "/root/reference")


REFERENCE = _resolve()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE, "scarmapper", "INDEL_Processing.py"))


def baseline_available() -> bool:
    return os.path.isfile(os.path.join(BASELINE, "scarmapper", "INDEL_Processing.py"))


def install(real_pool: bool = False, reference: str | None = None):
    """real_pool: pathos.multiprocessing.Pool is multiprocess.Pool (what pathos wraps: forked workers, dill
    pickling) instead of the in-process serial stand-in.  reference: tree to import from (default REFERENCE)."""
    global REFERENCE
    if reference:
        REFERENCE = reference
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE)
    from oracle import build_ref

    # the stand-in packages are real importable packages (baseline/shims/*, see its README), so forked and spawned
    # workers resolve them like any installed package
    if SHIMS not in sys.path:
        sys.path.insert(0, SHIMS)
    os.environ["SCAR_SHIM_SERIAL_POOL"] = "0" if real_pool else "1"
    import Levenshtein
    if not Levenshtein.compiled:
        raise RuntimeError("baseline/shims/Levenshtein/_lev.so is not built (gcc missing?)")

    # one reference tree at a time: a previous install() from another tree must not shadow this one
    for other in ("/root/reference", BASELINE):
        if other != REFERENCE and other in sys.path:
            sys.path.remove(other)
    for name in [m for m in sys.modules if m == "scarmapper" or m.startswith("scarmapper.") or m == "Valkyries"
                 or m.startswith("Valkyries.")]:
        mod = sys.modules[name]
        f = getattr(mod, "__file__", None) or ""
        if f and not os.path.abspath(f).startswith(os.path.abspath(REFERENCE) + os.sep) and "scarmapper_b200" not in f \
                and os.sep + "oracle" + os.sep not in f:
            del sys.modules[name]
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    build_ref.build()
    sw = build_ref.load()
    if sw is None:
        raise RuntimeError("oracle/_ref/SlidingWindow.*.so missing")
    sys.modules["scarmapper.SlidingWindow"] = sw
    import scarmapper  # the reference package
    scarmapper.SlidingWindow = sw
    return sw
