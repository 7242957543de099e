"""
synth — deterministic synthetic inputs for the five BASELINE.json configs.

Input manufacturing only (tests, bench.py, golden-vector generation); not on the product path.
The shipped fixture VALUES used here (four target rows of docs/ScarMapper_Targets.csv:2-5, the
8x12 Illumina dual indices of Indices/Illumina_Dual_Indices.csv:2-97) are restated as constants
because /root/reference does not exist on the GPU box; the real genome is unavailable, so each
locus is uniform-random ACGT of the row's length with the sgRNA implanted at the centre.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "scar_synth.c")
_LIB = os.path.join(_HERE, "libscar_synth.so")

# docs/ScarMapper_Targets.csv:2-5  (name, stop-start, sgRNA, ReverseComp)
SHIPPED_TARGETS = [
    ("Rosa26a", 472, "ACTCCAGTCTTTCTAGAAGA", True),
    ("AAVS1.1", 248, "GGGGCCACTAGGGACAGGAT", True),
    ("AAVS1.2", 251, "CAGGTAAAACTGACGCACGG", False),
    ("LBR2.1", 235, "GCCGATGGTGAAGTGGTAAG", True),
]
# Indices/Illumina_Dual_Indices.csv:2-97, row order D501+D701 .. D508+D712
ILLUMINA_D7 = [("D701", "ATTACTCG"), ("D702", "TCCGGAGA"), ("D703", "CGCTCATT"), ("D704", "GAGATTCC"),
               ("D705", "ATTCAGAA"), ("D706", "GAATTCGT"), ("D707", "CTGAAGCT"), ("D708", "TAATGCGC"),
               ("D709", "CGGCTATG"), ("D710a", "TCGATGAC"), ("D711", "TCTCGCGC"), ("D712", "AGCGATAG")]
ILLUMINA_D5 = [("D501", "TATAGCCT"), ("D502", "ATAGAGGC"), ("D503", "CCTATCCT"), ("D504", "GGCTCTGA"),
               ("D505", "AGGCGAAG"), ("D506", "TAATCTTA"), ("D507", "CAGGACGT"), ("D508", "GTACTGAC")]


def illumina_dual_indices():
    """[(Index_ID, Forward(D7), Reverse(D5))] x 96 in file order."""
    return [("{}+{}".format(n5, n7), s7, s5) for n5, s5 in ILLUMINA_D5 for n7, s7 in ILLUMINA_D7]


_COMP = bytes.maketrans(b"ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", b"TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn")


def rcomp(s: str) -> str:
    return s.encode().translate(_COMP)[::-1].decode()


def find_cutsite(target: str, sgrna: str, reverse_comp: bool) -> int:
    """First occurrence with the reference's loop bound (INDEL_Processing.py:759-795)."""
    work = rcomp(sgrna) if reverse_comp else sgrna
    lft, rt = 0, len(sgrna)
    while rt < len(target) - 1:
        if target[lft:rt] == work:
            return lft + 3 if reverse_comp else rt - 3
        lft += 1
        rt += 1
    raise ValueError("sgRNA does not map")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", _LIB, _SRC])
    return _LIB


class _Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_int32), ("min_len", C.c_int32),
                ("n_targets", C.c_int32), ("n_samples", C.c_int32), ("idx_len", C.c_int32),
                ("mode", C.c_int32), ("n_pool", C.c_int32), ("pool_stride", C.c_int32),
                ("p_wt", C.c_double), ("p_n", C.c_double), ("p_sub", C.c_double),
                ("p_idx1", C.c_double), ("p_idx2", C.c_double)]


_lib = None


def _get_lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.synth_reads.argtypes = [C.POINTER(_Params), C.c_int64, C.c_int64] + [C.c_void_p] * 14
    return _lib


@dataclass
class SynthConfig:
    name: str
    n_reads: int
    seed: int
    read_len: int = 150
    min_len: int = 150
    mode: int = 0
    platform: str = "Illumina"
    target_names: list = field(default_factory=list)
    targets: list = field(default_factory=list)        # locus sequences (str)
    sgrnas: list = field(default_factory=list)
    reverse_comp: list = field(default_factory=list)
    cutsites: list = field(default_factory=list)
    sample_ids: list = field(default_factory=list)     # Index_ID per sample, manifest order
    index_d7: list = field(default_factory=list)       # Master_Index_File column 2 per sample
    index_d5: list = field(default_factory=list)       # Master_Index_File column 3 per sample
    sample_target: list = field(default_factory=list)  # target id per sample
    phasing_seqs: list = field(default_factory=list)   # per target (forward 10-mer, reverse 10-mer)
    p_wt: float = 0.40
    p_n: float = 0.005
    p_sub: float = 0.002
    p_idx1: float = 0.10
    p_idx2: float = 0.01
    hr_donor: str = ""
    n_limit: float = 0.01
    min_length: int = 100
    pool: np.ndarray = None
    mh_pairs: np.ndarray = None
    mh_off: np.ndarray = None

    # ---- static problem description ---------------------------------------------------------
    @property
    def n_samples(self):
        return len(self.sample_ids)

    def phasing_lists(self):
        """Per target ([F0..Fh], [R0..Rh]) built as TargetMapper.phasing does (TargetMapper.py:54-69)."""
        out = []
        for fwd, rev in self.phasing_seqs:
            fl, rl = len(fwd), len(rev)
            h = int(fl * 0.5)
            rh = int(rl * 0.5)
            r1 = [fwd[fl - h - i:fl - i] for i in range(h + 1)]
            r2 = [rev[rl - rh - i:rl - i] for i in range(h + 1)]
            out.append((r1, r2))
        return out

    # ---- reads ------------------------------------------------------------------------------
    def generate(self, first: int = 0, n: int | None = None, want_index: bool = True):
        """Reads [first, first+n) -> dict(seq [n,read_len] u8, len [n] i32, f0/f1 [n,8] u8, truth [n] i32)."""
        n = self.n_reads - first if n is None else n
        L = _get_lib()
        tb = np.frombuffer("".join(self.targets).encode(), dtype=np.uint8).copy()
        to = np.zeros(len(self.targets) + 1, dtype=np.int64)
        to[1:] = np.cumsum([len(t) for t in self.targets])
        cs = np.asarray(self.cutsites, dtype=np.int32)
        ns = self.n_samples
        st = np.asarray(self.sample_target, dtype=np.int32) if ns else None
        ia = np.frombuffer("".join(self.index_d7).encode(), dtype=np.uint8).copy() if ns else None
        ib = np.frombuffer("".join(self.index_d5).encode(), dtype=np.uint8).copy() if ns else None
        idx_len = len(self.index_d7[0]) if ns else 8
        P = _Params(self.seed, self.read_len, self.min_len, len(self.targets), ns, idx_len, self.mode,
                    self.pool.shape[0], self.pool.shape[1], self.p_wt, self.p_n, self.p_sub,
                    self.p_idx1, self.p_idx2)
        seq = np.empty((n, self.read_len), dtype=np.uint8)
        ln = np.empty(n, dtype=np.int32)
        use_idx = bool(ns) and want_index
        f0 = np.empty((n, idx_len), dtype=np.uint8) if use_idx else None
        f1 = np.empty((n, idx_len), dtype=np.uint8) if use_idx else None
        truth = np.empty(n, dtype=np.int32)
        p = lambda a: None if a is None else a.ctypes.data
        L.synth_reads(C.byref(P), first, n, p(tb), p(to), p(cs), p(st), p(ia), p(ib), p(self.pool),
                      p(self.mh_pairs), p(self.mh_off), p(seq), p(ln), p(f0), p(f1), p(truth))
        return {"seq": seq, "len": ln, "f0": f0, "f1": f1, "truth": truth}

    def reads_as_strings(self, batch):
        seq, ln = batch["seq"], batch["len"]
        raw = seq.tobytes()
        rl = self.read_len
        return [raw[i * rl:i * rl + int(ln[i])].decode("latin-1") for i in range(len(ln))]

    def fields_as_strings(self, batch):
        if batch["f0"] is None:
            return None, None
        w = batch["f0"].shape[1]
        a, b = batch["f0"].tobytes(), batch["f1"].tobytes()
        n = len(batch["len"])
        return ([a[i * w:(i + 1) * w].decode() for i in range(n)],
                [b[i * w:(i + 1) * w].decode() for i in range(n)])


def _random_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def make_target(rng, length: int, sgrna: str, reverse_comp: bool, allow_repeats: bool = False) -> str:
    """Uniform-random locus with the sgRNA (or its reverse complement) implanted at the centre.
    Unless allow_repeats, redraw until every 10-mer of the locus is unique (a chance repeat, about
    0.3 expected in 800 nt, makes wild-type reads hit a far window; kept for one parity fixture)."""
    work = rcomp(sgrna) if reverse_comp else sgrna
    while True:
        body = _random_seq(rng, length)
        p = length // 2 - len(work) // 2
        t = body[:p] + work + body[p + len(work):]
        kmers = [t[i:i + 10] for i in range(length - 9)]
        if allow_repeats or len(set(kmers)) == len(kmers):
            return t


def _insertion_pool(rng, n_pool=256, stride=24):
    """Fixed pool: lower half 1-3 nt, upper half 5-20 nt (first byte of a row = length)."""
    pool = np.zeros((n_pool, stride), dtype=np.uint8)
    for e in range(n_pool):
        ln = int(rng.integers(1, 4)) if e < n_pool // 2 else int(rng.integers(5, 21))
        pool[e, 0] = ln
        pool[e, 1:1 + ln] = np.frombuffer(_random_seq(rng, ln).encode(), dtype=np.uint8)
    return pool


def _mh_pairs(targets, cutsites, reach=30, max_per_m=64):
    """Per target: (x, y) with x < cut <= y and target[x:x+m] == target[y:y+m] for some m in 1..6;
    deleting target[x:y] leaves an m-nt microhomology at the junction."""
    pairs, off = [], [0]
    for t, cut in zip(targets, cutsites):
        by_m = {m: [] for m in range(1, 7)}
        for x in range(max(16, cut - reach), cut):
            for y in range(cut, min(len(t) - 16, cut + reach)):
                m = 0
                while m < 6 and y + m < len(t) and t[x + m] == t[y + m]:
                    m += 1
                if m:
                    by_m[m].append((x, y))
        for m in range(1, 7):
            pairs += by_m[m][:max_per_m]
        off.append(len(pairs))
    return (np.asarray(pairs, dtype=np.int32).reshape(-1, 2) if pairs else np.zeros((0, 2), np.int32),
            np.asarray(off, dtype=np.int64))


def make_config(name: str, n_reads: int | None = None, seed: int = 20260101, allow_repeats: bool = False,
                **over) -> SynthConfig:
    """name in {C1, C2, C3, C4, C5}: the BASELINE.json configs, in order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    defaults = {"C1": 1_000_000, "C2": 10_000_000, "C3": 10_000_000, "C4": 10_000_000, "C5": 200_000_000}
    cfg = SynthConfig(name=name, n_reads=defaults[name] if n_reads is None else n_reads, seed=seed)
    rows = list(SHIPPED_TARGETS[:1])
    if name == "C3":
        cfg.mode = 1
        rows = list(SHIPPED_TARGETS)
        for k in range(8):
            rows.append(("SYN{:02d}".format(k + 1), int(rng.integers(230, 481)), _random_seq(rng, 20), k % 2 == 0))
    if name == "C4":
        cfg.mode = 2
        cfg.read_len = cfg.min_len = 300
        rows = [("LONG800", 800, _random_seq(rng, 20), False)]
    for nm, ln, sg, rc in rows:
        t = make_target(rng, ln, sg, rc, allow_repeats)
        cfg.target_names.append(nm)
        cfg.targets.append(t)
        cfg.sgrnas.append(sg)
        cfg.reverse_comp.append(rc)
        cfg.cutsites.append(find_cutsite(t, sg, rc))
    idx = illumina_dual_indices()
    if name == "C1":
        idx = idx[:1]
    for s, (iid, d7, d5) in enumerate(idx):
        cfg.sample_ids.append(iid)
        cfg.index_d7.append(d7)
        cfg.index_d5.append(d5)
        cfg.sample_target.append(s % len(cfg.targets))
    # phasing primers chosen so that phase k of a wild-type read starting at cut - len/2 + phi matches
    for t, cut in zip(cfg.targets, cfg.cutsites):
        s0 = cut - cfg.read_len // 2
        e0 = s0 + cfg.read_len
        if s0 < 0 or e0 + 5 > len(t) or s0 + 10 > len(t):
            cfg.phasing_seqs.append(("", ""))
            continue
        rct = rcomp(t)
        q0 = len(t) - e0 - 5
        cfg.phasing_seqs.append((t[s0:s0 + 10], rct[q0:q0 + 10]))
    cfg.pool = _insertion_pool(rng)
    cfg.mh_pairs, cfg.mh_off = _mh_pairs(cfg.targets, cfg.cutsites)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg
