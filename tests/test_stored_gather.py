"""
scar_stored_gather_host (host, no GPU): the stored sequence per platform (scarmapper/INDEL_Processing.py:946,978,981:
Illumina as read, TruSeq [6:-6], Ramsden rcomp) of a list of reads as one CSR buffer, rows of 3'-strand guides
reverse-complemented (frequency_output, :312-327) - against Sequence_Magic.rcomp's table restated in Python.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scarmapper_b200 import engine                    # noqa: E402
from scarmapper_b200.engine import HostRagged         # noqa: E402

# Valkyries/Sequence_Magic.py:97-118 (unmapped characters pass)
_PAIRS = "ATCGMKRYVBHD"
_TABLE = {}
for _i in range(0, len(_PAIRS), 2):
    for _a, _b in ((_PAIRS[_i], _PAIRS[_i + 1]), (_PAIRS[_i + 1], _PAIRS[_i])):
        _TABLE[_a] = _b
        _TABLE[_a.lower()] = _b.lower()


def rcomp(s):
    return "".join(_TABLE.get(c, c) for c in reversed(s))


def stored(platform, s):
    return s[6:-6] if platform == "TruSeq" else rcomp(s) if platform == "Ramsden" else s


@pytest.mark.parametrize("platform", ["Illumina", "TruSeq", "Ramsden"])
@pytest.mark.parametrize("n_rows", [0, 7, 150000])        # the large case is split over threads
def test_stored_gather_matches_the_definition(platform, n_rows):
    rng = np.random.default_rng(n_rows + len(platform))
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMxz-", dtype=np.uint8)
    reads = ["".join(map(chr, alphabet[rng.integers(0, len(alphabet) if i % 5 == 0 else 5, int(rng.integers(0, 40)))]))
             for i in range(500)]
    ragged = HostRagged.from_strings([r.encode() for r in reads])
    rows = rng.integers(0, len(reads), n_rows).astype(np.int64)
    for flip in (None, rng.integers(0, 2, n_rows).astype(np.uint8)):
        out, off = engine.stored_gather(ragged, rows, platform, flip=flip)
        assert off.shape == (n_rows + 1,) and off[0] == 0
        text = out.tobytes().decode("latin-1")
        check = range(n_rows) if n_rows < 1000 else list(range(0, n_rows, 997)) + [n_rows - 1]
        for k in check:
            want = stored(platform, reads[rows[k]])
            if flip is not None and flip[k]:
                want = rcomp(want)
            assert text[off[k]:off[k + 1]] == want, (platform, k)
        want_len = np.asarray([len(stored(platform, r)) for r in reads], dtype=np.int64)[rows]
        assert (np.diff(off) == want_len).all()
