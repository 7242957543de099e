"""baseline/_ref/ScarMapper (oracle/build_ref.build_baseline) is the UNMODIFIED reference driver: every module in it
is byte-identical with the file under /root/reference, and none of it is tracked by git.  Runs in the build
container (both trees present)."""
import os
import subprocess

import pytest

from oracle import build_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORIGINAL = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ORIGINAL, "scarmapper")), reason="reference tree not present")


def test_baseline_copy_is_byte_identical_and_untracked():
    tree = build_ref.build_baseline()
    assert tree and os.path.isdir(tree)
    n = 0
    for pkg in ("scarmapper", "Valkyries"):
        for fn in sorted(os.listdir(os.path.join(tree, pkg))):
            if fn.endswith(".py"):
                assert open(os.path.join(tree, pkg, fn), "rb").read() == open(os.path.join(ORIGINAL, pkg, fn), "rb").read(), fn
                n += 1
    assert n >= 10
    tracked = subprocess.run(["git", "ls-files", "baseline/_ref"], cwd=ROOT, stdout=subprocess.PIPE, text=True).stdout.strip()
    assert tracked == "", "baseline/_ref must stay out of the repository"
    ignored = subprocess.run(["git", "check-ignore", "-q", "baseline/_ref/ScarMapper/scarmapper/INDEL_Processing.py"], cwd=ROOT)
    assert ignored.returncode == 0
    gi = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gi) or "baseline" not in open(gi).read()     # it must travel to the GPU box


def test_debug_read_limit_mirrors_the_reference_loop():
    """--Verbose DEBUG (reference INDEL_Processing.py:906-914): `if self.read_count > read_limit: eof = True` is tested
    at the top of an iteration that still reads and counts one more record.  The loop below is that control flow with
    a small limit; the drop-in's READ_LIMIT_DEBUG is the same formula at the reference's read_limit = 1 000 000."""
    from scarmapper_b200 import dropin

    def reference_loop(read_limit, n_available):
        read_count, eof = 0, False
        while not eof:
            if read_count > read_limit:
                eof = True
            if read_count >= n_available:      # StopIteration
                eof = True
                continue
            read_count += 1
        return read_count
    for limit in (0, 1, 5, 40):
        assert reference_loop(limit, 10 ** 6) == limit + 2
        assert reference_loop(limit, 3) == min(3, limit + 2)
    assert dropin.READ_LIMIT_DEBUG == 1000000 + 2
