"""
Pins oracle/scar_oracle.c against the reference ITSELF: the compiled, unmodified
scarmapper/SlidingWindow.pyx (oracle/_ref, built by oracle/build_ref.py).  Randomised
differential test over the cases the survey lists: tandem-repeat targets, reduced alphabets,
N / lowercase bytes, reads of 10-300 nt including the loop-bound lengths, HR donors.
"""
import random

import numpy as np
import pytest

from oracle import scar_oracle as orc


def fresh_summary():
    # INDEL_Processing.py:94
    return ["idx", 0, 0, 0, 0, 0, [0, 0], [0, 0], "junction data", "tgt", [0, 0]]


def ref_call(mod, read, target, cut, hr="", cutwindow=None):
    left, right = orc.window_mapping(target, cut)
    cw = target[cut - 4:cut + 4] if cutwindow is None else cutwindow
    summary = fresh_summary()
    sub, summary2 = mod.sliding_window(read, target, cut, len(target), 15, len(target) - 15, summary,
                                       left, right, cw, hr)
    assert summary2 is summary
    return sub, summary


def summary_from_record(rec):
    s = fresh_summary()
    st, fl = int(rec["status"]), int(rec["flags"])
    if st == orc.ST_NO_JUNCTION:
        s[6][0] += 1
    if st == orc.ST_NO_CUT:
        s[6][1] += 1
    if st == orc.ST_SCAR:
        s[2] += bool(fl & orc.FL_LEFT_DEL)
        s[3] += bool(fl & orc.FL_RIGHT_DEL)
        s[4] += bool(fl & orc.FL_INSERTION)
        s[5] += bool(fl & orc.FL_MICROHOM)
    n = int(rec["hr_n"])
    if n:
        s[10][0] += 1
        s[10][1] += n - 1
    return s


def rand_seq(rng, n, alphabet):
    return "".join(rng.choice(alphabet) for _ in range(n))


def make_case(rng):
    alphabet = rng.choice(["AC", "ACG", "ACGT", "ACGT", "ACGT", "ACGTN"])
    L = rng.choice([60, 80, 120, 235, 248, 251, 472])
    if rng.random() < 0.3:
        unit = rand_seq(rng, rng.randint(1, 12), alphabet)
        target = (unit * (L // len(unit) + 1))[:L]
        target = "".join(c if rng.random() > 0.03 else rng.choice(alphabet) for c in target)
    else:
        target = rand_seq(rng, L, alphabet)
    cut = rng.randint(5, L - 5) if rng.random() < 0.2 else rng.randint(L // 3, 2 * L // 3)
    rl = rng.choice([10, 20, 34, 35, 36, 37, 45, 50, 51, 52, 75, 100, 101, 150, 150, 150, 151, 200, 300])
    mode = rng.random()
    if mode < 0.15:
        read = rand_seq(rng, rl, "ACGT")
    else:
        dl = rng.choice([0, 0, 1, 2, 5, 10, 30])
        dr = rng.choice([0, 0, 1, 3, 7, 12, 40])
        ins = rand_seq(rng, rng.choice([0, 0, 0, 1, 2, 5, 12, 25]), rng.choice(["ACGT", "ACGTN", "acgt"]))
        a = max(0, rl // 2 - rng.randint(0, 5))
        lo = max(0, cut - dl - a)
        left = target[lo:max(lo, cut - dl)]
        right = target[min(L, cut + dr):]
        read = (left + ins + right)[:rl]
        if len(read) < rl:
            read += rand_seq(rng, rl - len(read), "ACGT")
        if rng.random() < 0.2:
            p = rng.randrange(len(read))
            read = read[:p] + rng.choice("ACGTNn") + read[p + 1:]
    hr = ""
    if rng.random() < 0.25:
        if rng.random() < 0.5 and len(read) > 60:
            p = rng.randint(20, len(read) - 40)
            hr = read[p:p + 12]
        else:
            hr = rand_seq(rng, 12, "ACGT")
    return read, target, cut, hr


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_sliding_window_matches_compiled_reference(ref_sliding_window, seed):
    rng = random.Random(seed)
    n_scar = 0
    for _ in range(3000):
        read, target, cut, hr = make_case(rng)
        sub, summary = ref_call(ref_sliding_window, read, target, cut, hr)
        rec = orc.sliding_window(read, target, cut, 15, None, hr)
        mine = orc.record_to_sublist(rec, read, target, cut)
        assert mine == sub, (read, target, cut, hr, rec)
        assert summary_from_record(rec) == summary, (read, target, cut, hr, rec)
        n_scar += bool(sub)
    assert n_scar > 300


def test_cutwindow_branch(ref_sliding_window):
    """The :56-60 branch is dead in the driver (8-nt cutwindow) but reachable through the
    per-read signature with a 10-nt cutwindow; the oracle follows it."""
    rng = random.Random(9)
    hits = 0
    for _ in range(400):
        target = rand_seq(rng, 200, "ACGT")
        cut = 100
        read = target[25:175]
        i = rng.randint(0, 40)
        cw = target[cut - 10 - i:cut - i]
        sub, summary = ref_call(ref_sliding_window, read, target, cut, "", cw)
        rec = orc.sliding_window(read, target, cut, 15, cw, "")
        assert orc.record_to_sublist(rec, read, target, cut) == sub
        assert summary_from_record(rec) == summary
        hits += bool(rec["flags"] & orc.FL_CUTWINDOW)
    assert hits > 0


def test_levenshtein_known_answers():
    # parity unpinned (python-Levenshtein absent): textbook values of the unit-cost distance
    kat = [("", "", 0), ("A", "", 1), ("", "ACG", 3), ("kitten", "sitting", 3), ("flaw", "lawn", 2),
           ("ATTACTCG", "ATTACTCG", 0), ("ATTACTCG", "ATTACTCA", 1), ("ATTACTCG", "TTACTCGA", 2),
           ("GATTACA", "GCATGCU", 4), ("intention", "execution", 5)]
    for a, b, d in kat:
        assert orc.levenshtein(a, b) == d
        assert orc.levenshtein(b, a) == d
    assert orc.match_maker("ATTACTCG", "ATTACTCGA") == 0   # Sequence_Magic.py:133-134
    assert orc.match_maker("ATTACTCG", "ATTACTC") == 2


def test_levenshtein_against_bruteforce_recursion():
    from functools import lru_cache
    rng = random.Random(5)

    def lev(a, b):
        @lru_cache(None)
        def f(i, j):
            if i == 0:
                return j
            if j == 0:
                return i
            return min(f(i - 1, j) + 1, f(i, j - 1) + 1, f(i - 1, j - 1) + (a[i - 1] != b[j - 1]))
        return f(len(a), len(b))

    for _ in range(500):
        a = rand_seq(rng, rng.randint(0, 12), "ACGTN")
        b = rand_seq(rng, rng.randint(0, 12), "ACGTN")
        assert orc.levenshtein(a, b) == lev(a, b)
