"""
CPU-only: the C-ABI library builds, loads, and exports every symbol include/scar_b200.h declares.
No compute call is made here (there is no GPU in the build container).
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "scar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scar_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from scarmapper_b200 import build, _native
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libscar_b200.so does not export " + n
    # the ctypes binding covers exactly the declared surface
    assert sorted(_native.PROTOTYPES) == names
    assert _native.lib().scar_version() >= 100


def test_struct_layouts_match_the_header():
    from scarmapper_b200 import _native as N
    assert ctypes.sizeof(N.Ragged) == 32
    assert N.RECORD_DTYPE.itemsize == 16 and N.ENTRY_DTYPE.itemsize == 32
    from oracle import scar_oracle as orc
    assert orc.RECORD_DTYPE == N.RECORD_DTYPE


def test_cutsite_search_host_entry_matches_oracle():
    # scar_cutsite_search is host-only arithmetic (no device needed)
    import random
    from oracle import scar_oracle as orc
    from scarmapper_b200.engine import cutsite_search
    from scarmapper_b200._native import ScarError
    rng = random.Random(3)
    for _ in range(300):
        L = rng.randint(30, 300)
        t = "".join(rng.choice("ACGT") for _ in range(L))
        n = rng.randint(5, 25)
        p = rng.randint(0, L - n)
        sg = t[p:p + n]
        rc = rng.random() < 0.5
        if rc:
            sg = orc.rcomp(sg)
        want = orc.cutsite_search(t, sg, rc)
        if want < 0:
            with pytest.raises(ScarError):
                cutsite_search(t, sg, rc)
        else:
            assert cutsite_search(t, sg, rc) == want


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the compute entry points must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scarmapper_b200.engine import ScarEngine, ScarProblem
    from scarmapper_b200._native import ScarError
    prob = ScarProblem(targets=["ACGT" * 30], cutsites=[60], sample_target=[0], left_index=["ACGTACGT"],
                       right_index=["ACGTACGT"])
    with pytest.raises(ScarError) as e:
        ScarEngine(prob)
    assert e.value.code == -1


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scarmapper_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), fn + " mentions the oracle"
