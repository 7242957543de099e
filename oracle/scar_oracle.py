"""
scar_oracle.py — ctypes wrapper around oracle/scar_oracle.c plus the dict-based scar-table
aggregation, written the way the reference drives it.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (scarmapper_b200/) never imports this.

Reference followed (all under /root/reference):
  scarmapper/SlidingWindow.pyx:20-171           -> orc_sliding_window
  scarmapper/INDEL_Processing.py:57-80,759-795  -> orc_window_counts, orc_cutsite_search
  scarmapper/INDEL_Processing.py:135-196        -> classify_batch + aggregate (below)
  scarmapper/INDEL_Processing.py:945-975,1126-1201 -> orc_phasing, orc_index_matching
  Valkyries/Sequence_Magic.py:52-136            -> orc_rcomp, orc_match_maker
Levenshtein (python-Levenshtein =0.12.2, not vendored): parity unpinned, see scar_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import OrderedDict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "scar_oracle.c")
_LIB = os.path.join(_HERE, "libscar_oracle.so")

RECORD_DTYPE = np.dtype([
    ("status", "u1"), ("flags", "u1"), ("sample", "<u2"),
    ("clj", "<u2"), ("crj", "<u2"), ("tlj", "<u2"), ("trj", "<u2"),
    ("hr_n", "<u2"), ("phase_r1", "i1"), ("phase_r2", "i1"),
])
assert RECORD_DTYPE.itemsize == 16

ST_UNASSIGNED, ST_FILTERED, ST_NO_JUNCTION, ST_NO_CUT, ST_N_INSERTION, ST_SCAR = range(6)
FL_LEFT_DEL, FL_RIGHT_DEL, FL_INSERTION, FL_MICROHOM, FL_HR, FL_CUTWINDOW = 1, 2, 4, 8, 16, 32
SAMPLE_NONE = 0xFFFF
PLATFORMS = {"Illumina": 0, "TruSeq": 1, "Ramsden": 2}
DEFAULT_MISMATCH = {0: 1, 1: 0, 2: 3}  # INDEL_Processing.py:1142-1147


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (-O2, no fast-math)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", _LIB, _SRC])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        u8p, i64p, i32p = C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_levenshtein.argtypes = [u8p, C.c_int, u8p, C.c_int]
        L.orc_levenshtein.restype = C.c_int
        L.orc_match_maker.argtypes = [u8p, C.c_int, u8p, C.c_int]
        L.orc_match_maker.restype = C.c_int
        L.orc_rcomp.argtypes = [u8p, C.c_int, u8p]
        L.orc_cutsite_search.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_cutsite_search.restype = C.c_int
        L.orc_window_counts.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_sliding_window.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int,
                                         u8p, C.c_int, C.c_void_p]
        L.orc_index_matching.argtypes = [C.c_int, C.c_int, C.c_int, u8p, i64p, u8p, i64p,
                                         u8p, C.c_int, u8p, C.c_int, u8p, C.c_int]
        L.orc_index_matching.restype = C.c_int
        L.orc_phasing.argtypes = [u8p, C.c_int, C.c_int, u8p, i64p, u8p, i64p, C.c_void_p]
        L.orc_is_filtered.argtypes = [u8p, C.c_int, C.c_double, C.c_int]
        L.orc_is_filtered.restype = C.c_int
        L.orc_classify_batch.argtypes = (
            [C.c_int64, u8p, i64p, u8p, i64p, u8p, i64p, i32p,
             C.c_int, C.c_int, C.c_int, u8p, i64p, u8p, i64p, i32p,
             C.c_int, u8p, i64p, i32p, i32p, u8p, i64p, u8p, i64p,
             u8p, C.c_int, C.c_double, C.c_int, C.c_void_p])
        _lib = L
    return _lib


def _b(x) -> bytes:
    return x.encode("latin-1") if isinstance(x, str) else bytes(x)


def csr(strings):
    """list of str/bytes -> (uint8 array, int64 offsets[n+1])"""
    bs = [_b(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if data.size == 0:
        data = np.zeros(1, np.uint8)  # keep a valid pointer
    return data, off


def _p(a):
    return None if a is None else a.ctypes.data


def levenshtein(a, b) -> int:
    a, b = _b(a), _b(b)
    return lib().orc_levenshtein(a, len(a), b, len(b))


def match_maker(query, unknown) -> int:
    q, u = _b(query), _b(unknown)
    return lib().orc_match_maker(q, len(q), u, len(u))


def rcomp(seq):
    s = _b(seq)
    out = C.create_string_buffer(len(s) + 1)
    lib().orc_rcomp(s, len(s), out)
    r = out.raw[:len(s)]
    return r.decode("latin-1") if isinstance(seq, str) else r


def cutsite_search(target, sgrna, reverse_comp: bool) -> int:
    t, s = _b(target), _b(sgrna)
    return lib().orc_cutsite_search(t, len(t), s, len(s), int(bool(reverse_comp)))


def window_mapping(target, cutsite):
    """(left_target_windows, right_target_windows) exactly as INDEL_Processing.py:57-80."""
    t = target
    nl, nr = C.c_int(), C.c_int()
    lib().orc_window_counts(len(t), cutsite, C.byref(nl), C.byref(nr))
    left = [t[cutsite - 10 - i:cutsite - i] for i in range(nl.value)]
    right = [t[cutsite + i:cutsite + i + 10] for i in range(nr.value)]
    return left, right


def sliding_window(read, target, cutsite, lower_limit=15, cutwindow=None, hr_donor=""):
    """One read -> numpy record (RECORD_DTYPE scalar)."""
    r, t, h = _b(read), _b(target), _b(hr_donor)
    cw = _b(target[max(cutsite - 4, 0):cutsite + 4]) if cutwindow is None else _b(cutwindow)
    rec = np.zeros(1, dtype=RECORD_DTYPE)
    lib().orc_sliding_window(r, len(r), t, len(t), cutsite, lower_limit, cw, len(cw), h, len(h),
                             rec.ctypes.data)
    return rec[0]


def record_to_sublist(rec, read, target, cutsite):
    """What the reference returns for this read: [] or the 10-field list (SlidingWindow.pyx:171)."""
    if rec["status"] != ST_SCAR:
        return []
    clj, crj, tlj, trj = int(rec["clj"]), int(rec["crj"]), int(rec["tlj"]), int(rec["trj"])
    ins = read[clj:crj] if rec["flags"] & FL_INSERTION else ""
    mh = read[crj:clj] if rec["flags"] & FL_MICROHOM else ""
    return [target[tlj:cutsite], target[cutsite:trj], ins, mh, read, clj, crj, tlj, trj,
            "HR" if rec["flags"] & FL_HR else ""]


class Problem:
    """Static inputs of a run: targets, cutsites, samples, indices, phasing, options."""

    def __init__(self, targets, cutsites, sample_target, left_index=None, right_index=None,
                 platform="Illumina", mismatch=None, phasing=None, hr_donor="", n_limit=0.01,
                 min_length=100):
        self.targets = [t if isinstance(t, str) else t.decode("latin-1") for t in targets]
        self.cutsites = np.asarray(cutsites, dtype=np.int32)
        self.sample_target = np.asarray(sample_target, dtype=np.int32)
        self.n_samples = len(self.sample_target)
        self.left_index = list(left_index) if left_index is not None else [""] * self.n_samples
        self.right_index = list(right_index) if right_index is not None else [""] * self.n_samples
        self.platform = PLATFORMS[platform] if isinstance(platform, str) else int(platform)
        self.mismatch = DEFAULT_MISMATCH[self.platform] if mismatch is None else int(mismatch)
        # phasing: per target, (list of R1 strings, list of R2 strings) or None
        self.phasing = phasing if phasing is not None else [([], []) for _ in self.targets]
        self.hr_donor = hr_donor
        self.n_limit = float(n_limit)
        self.min_length = int(min_length)


def classify_batch(prob: Problem, reads, f0=None, f1=None, sample_in=None) -> np.ndarray:
    """reads: list of str/bytes, or (uint8 array, int64 offsets).  Returns records[n]."""
    L = lib()
    seq_b, seq_o = reads if isinstance(reads, tuple) else csr(reads)
    n = len(seq_o) - 1
    f0b = f0o = f1b = f1o = None
    if f0 is not None:
        f0b, f0o = f0 if isinstance(f0, tuple) else csr(f0)
    if f1 is not None:
        f1b, f1o = f1 if isinstance(f1, tuple) else csr(f1)
    sin = None if sample_in is None else np.ascontiguousarray(sample_in, dtype=np.int32)
    lb, lo = csr(prob.left_index)
    rb, ro = csr(prob.right_index)
    tb, to = csr(prob.targets)
    r1, r2, poff = [], [], [0]
    for p1, p2 in prob.phasing:
        r1 += list(p1); r2 += list(p2); poff.append(len(r1))
    r1b, r1o = csr(r1)
    r2b, r2o = csr(r2)
    poff = np.asarray(poff, dtype=np.int32)
    hr = _b(prob.hr_donor)
    out = np.zeros(n, dtype=RECORD_DTYPE)
    seq_b = np.ascontiguousarray(seq_b, dtype=np.uint8)
    seq_o = np.ascontiguousarray(seq_o, dtype=np.int64)
    L.orc_classify_batch(
        n, _p(seq_b), _p(seq_o), _p(f0b), _p(f0o), _p(f1b), _p(f1o), _p(sin),
        prob.platform, prob.mismatch, prob.n_samples, _p(lb), _p(lo), _p(rb), _p(ro),
        _p(prob.sample_target), len(prob.targets), _p(tb), _p(to), _p(prob.cutsites), _p(poff),
        _p(r1b), _p(r1o), _p(r2b), _p(r2o), hr, len(hr), prob.n_limit, prob.min_length,
        out.ctypes.data)
    return out


def platform_seq(platform: int, seq: bytes) -> bytes:
    """Sequence stored per platform (INDEL_Processing.py:946,978,981)."""
    if platform == 1:
        return seq[6:-6]
    if platform == 2:
        return rcomp(seq)
    return seq


def aggregate(prob: Problem, records: np.ndarray, reads) -> "list[OrderedDict]":
    """Scar table per sample, the reference's way (INDEL_Processing.py:186-196): a dict keyed by
    'ldel|rdel|ins|mh|hr' holding [count, first read's global index], in first-seen order."""
    seq_b, seq_o = reads if isinstance(reads, tuple) else csr(reads)
    raw = seq_b.tobytes()
    tables = [OrderedDict() for _ in range(prob.n_samples)]
    idx = np.nonzero(records["status"] == ST_SCAR)[0]
    for r in idx:
        rec = records[r]
        s = int(rec["sample"])
        t = int(prob.sample_target[s])
        target, cut = prob.targets[t], int(prob.cutsites[t])
        seq = platform_seq(prob.platform, raw[seq_o[r]:seq_o[r + 1]]).decode("latin-1")
        sub = record_to_sublist(rec, seq, target, cut)
        key = "{}|{}|{}|{}|{}".format(sub[0], sub[1], sub[2], sub[3], sub[9])
        d = tables[s]
        if key in d:
            d[key][0] += 1
        else:
            d[key] = [1, int(r)]
    return tables


def counters(prob: Problem, records: np.ndarray) -> np.ndarray:
    """[n_samples, 16] int64, columns as include/scar_records.h SCAR_CT_*; plus the phase
    histograms are left to the caller (np.bincount over phase_r1/phase_r2)."""
    out = np.zeros((prob.n_samples, 16), dtype=np.int64)
    st, fl, sm = records["status"], records["flags"], records["sample"].astype(np.int64)
    ok = st != ST_UNASSIGNED

    def add(col, mask, weights=None):
        m = mask & ok
        out[:, col] += np.bincount(sm[m], weights=None if weights is None else weights[m],
                                   minlength=prob.n_samples).astype(np.int64)[:prob.n_samples]

    scar = st == ST_SCAR
    add(0, ok)
    add(1, st >= ST_NO_JUNCTION)
    add(2, scar & ((fl & FL_LEFT_DEL) != 0))
    add(3, scar & ((fl & FL_RIGHT_DEL) != 0))
    add(4, scar & ((fl & FL_INSERTION) != 0))
    add(5, scar & ((fl & FL_MICROHOM) != 0))
    add(6, st == ST_NO_JUNCTION)
    add(7, st == ST_NO_CUT)
    add(8, st == ST_FILTERED)
    hr = records["hr_n"].astype(np.int64)
    add(9, hr > 0)
    add(10, hr > 0, weights=(hr - 1).astype(np.float64))
    add(11, st == ST_N_INSERTION)
    add(12, scar)
    return out


def parse_illumina_header(name: str):
    """name.split(':')[-1].split('+')[0..1] (INDEL_Processing.py:1155-1156)."""
    parts = name.split(":")[-1].split("+")
    return parts[0], parts[1]
