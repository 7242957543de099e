"""
build_ref.py — compile the reference's own Cython hot kernel where it lies.

Builds /root/reference/scarmapper/SlidingWindow.pyx (the only native module the reference
builds, scarmapper/setup.py:14) into oracle/_ref/SlidingWindow.<abi>.so.  Nothing is copied:
Cython reads the .pyx in place, the generated C goes to a temporary directory, and only the
shared object lands in oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).

The reference's own scarmapper/setup.py is NOT run (it copies files and rmtree's a directory).

TEST INFRASTRUCTURE ONLY — used to pin oracle/scar_oracle.c and as the "reference" CPU baseline.
"""
from __future__ import annotations

import glob
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("SCARMAPPER_REFERENCE", "/root/reference")
PYX = os.path.join(REFERENCE, "scarmapper", "SlidingWindow.pyx")


def ref_so(name: str = "SlidingWindow") -> str | None:
    hits = sorted(glob.glob(os.path.join(REF_DIR, name + ".*.so")))
    return hits[0] if hits else None


def build(force: bool = False, name: str = "SlidingWindow") -> str | None:
    """Returns the path of the built module, or None when the reference tree is absent
    (e.g. on the GPU box, where the prebuilt file is used as is).  name: SlidingWindow (the module the reference
    builds) or AlignmentProcessing (its legacy sibling, scarmapper/AlignmentProcessing.pyx)."""
    PYX = os.path.join(REFERENCE, "scarmapper", name + ".pyx")
    have = ref_so(name)
    if not os.path.exists(PYX):
        return have
    if have and not force and os.path.getmtime(have) >= os.path.getmtime(PYX):
        return have
    from Cython.Build import cythonize
    from setuptools import Distribution
    from setuptools.command.build_ext import build_ext

    os.makedirs(REF_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="scar_ref_build_")
    try:
        exts = cythonize([PYX], build_dir=tmp, quiet=True, language_level=3)
        for e in exts:
            e.name = name
        dist = Distribution({"name": "scar_ref", "ext_modules": exts})
        cmd = build_ext(dist)
        cmd.build_lib = os.path.join(tmp, "lib")
        cmd.build_temp = os.path.join(tmp, "tmp")
        cmd.ensure_finalized()
        cmd.run()
        built = glob.glob(os.path.join(tmp, "lib", "**", name + "*.so"), recursive=True)
        if not built:
            raise RuntimeError("cythonize produced no {} shared object".format(name))
        for old in glob.glob(os.path.join(REF_DIR, name + ".*.so")):
            os.remove(old)
        dst = os.path.join(REF_DIR, os.path.basename(built[0]))
        shutil.copy2(built[0], dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


BASELINE_DIR = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "ScarMapper")


def baseline_tree() -> str | None:
    """baseline/_ref/ScarMapper when it holds the reference's driver modules, else None."""
    return BASELINE_DIR if os.path.isfile(os.path.join(BASELINE_DIR, "scarmapper", "INDEL_Processing.py")) else None


def build_baseline(force: bool = False) -> str | None:
    """The reference is Python: its driver (DataProcessing.main_loop, INDEL_Processing.py:1038-1061) cannot be
    compiled into a binary, and /root/reference does not exist on the GPU box.  So that the UNMODIFIED driver can
    be run there - as the `--impl reference` arm of bench.py and under the drop-in in the GPU tests - its Python
    modules are copied, byte for byte, into the git-ignored (but gpurun-shipped) baseline/_ref/ScarMapper/, together
    with the compiled SlidingWindow module under the name `from scarmapper import SlidingWindow` resolves.
    Nothing under baseline/_ref/ is committed; tests/test_baseline_tree.py checks the copy against the original."""
    src = REFERENCE
    if not os.path.isfile(os.path.join(src, "scarmapper", "INDEL_Processing.py")):
        return baseline_tree()
    for pkg in ("scarmapper", "Valkyries"):
        dst = os.path.join(BASELINE_DIR, pkg)
        os.makedirs(dst, exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(src, pkg))):
            if not fn.endswith(".py") or fn == "setup.py":
                continue
            a, b = os.path.join(src, pkg, fn), os.path.join(dst, fn)
            if force or not os.path.exists(b) or open(a, "rb").read() != open(b, "rb").read():
                shutil.copyfile(a, b)
                os.chmod(b, 0o644)
    so = build()
    if so:
        dst = os.path.join(BASELINE_DIR, "scarmapper", os.path.basename(so))
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(so):
            shutil.copy2(so, dst)
    return baseline_tree()


def load(name: str = "SlidingWindow"):
    """Import the compiled reference module (oracle/_ref/<name>.*.so) or return None."""
    so = ref_so(name)
    if so is None:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out or "reference tree not present and no prebuilt module")
    print(build(force="--force" in sys.argv, name="AlignmentProcessing") or "no AlignmentProcessing module")
    print(build_baseline(force="--force" in sys.argv) or "no baseline tree")
