"""
ref_shims.py — make the UNMODIFIED reference driver importable in this container.

The reference imports seven third-party packages that are absent here and cannot be installed
(no network): Levenshtein, pysam, pyfaidx, pathos, natsort, magic, matplotlib.
install() registers minimal stand-ins in sys.modules, puts /root/reference on sys.path and
registers oracle/_ref/SlidingWindow.*.so (the reference's own compiled Cython) as
scarmapper.SlidingWindow, so that `import scarmapper.INDEL_Processing` runs the reference code
itself.  Used only by tests/golden/make_golden.py (fixture generation) and by tests that are
skipped when /root/reference is absent.  TEST INFRASTRUCTURE ONLY.

Stand-in semantics:
  Levenshtein.distance  -> oracle C unit-cost distance (parity unpinned at this boundary)
  natsort.natsorted     -> natural sort; for the (int, int) tuple keys and 'Phase F10'-style
                           labels the reference sorts, numeric-aware tuple ordering
  pathos.multiprocessing.Pool -> in-process serial starmap (same results, no pickling)
  magic.from_file       -> gzip magic bytes -> 'application/gzip', else 'text/plain'
  pysam.FastaFile.fetch -> plain FASTA reader raising KeyError on unknown contigs
  pyfaidx.Fasta         -> first record
  matplotlib            -> inert mocks (plots are not part of the parity surface)
"""
from __future__ import annotations

import os
import re
import sys
import types
from unittest import mock

REFERENCE = os.environ.get("SCARMAPPER_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "scarmapper"))


def _natural_key(x):
    if isinstance(x, (tuple, list)):
        return tuple(_natural_key(v) for v in x)
    if isinstance(x, (int, float)):
        return (0, x, "")
    parts = re.split(r"(\d+)", str(x))
    return tuple((0, int(p), "") if p.isdigit() else (1, 0, p) for p in parts)


def _natsorted(seq, key=None, reverse=False, alg=None):
    if key is None:
        return sorted(seq, key=_natural_key, reverse=reverse)
    return sorted(seq, key=lambda v: _natural_key(key(v)), reverse=reverse)


def _read_fasta(path):
    import gzip
    opener = gzip.open if path.endswith((".gz", ".bgz")) else open
    recs, name, chunks = [], None, []
    with opener(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(chunks)))
                name, chunks = line[1:].split()[0], []
            elif line:
                chunks.append(line)
    if name is not None:
        recs.append((name, "".join(chunks)))
    return recs


def install():
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE)
    from oracle import build_ref, scar_oracle

    lev = types.ModuleType("Levenshtein")
    lev.distance = lambda a, b: scar_oracle.levenshtein(a, b)
    sys.modules["Levenshtein"] = lev

    ns = types.ModuleType("natsort")
    ns.natsorted = _natsorted
    ns_inner = types.ModuleType("natsort.natsort")
    ns_inner.natsorted = _natsorted
    ns.natsort = ns_inner
    sys.modules["natsort"] = ns
    sys.modules["natsort.natsort"] = ns_inner

    class _SerialPool:
        def __init__(self, *a, **k):
            pass

        def starmap(self, fn, it):
            return [fn(*args) for args in it]

        def map(self, fn, it):
            return [fn(a) for a in it]

    pathos = types.ModuleType("pathos")
    pmp = types.ModuleType("pathos.multiprocessing")
    pmp.Pool = _SerialPool
    pathos.multiprocessing = pmp
    sys.modules["pathos"] = pathos
    sys.modules["pathos.multiprocessing"] = pmp

    magic = types.ModuleType("magic")

    def from_file(path, mime=True):
        with open(path, "rb") as f:
            return "application/gzip" if f.read(2) == b"\x1f\x8b" else "text/plain"
    magic.from_file = from_file
    sys.modules["magic"] = magic

    pysam = types.ModuleType("pysam")

    class FastaFile:
        def __init__(self, path):
            self._recs = dict(_read_fasta(path))

        def fetch(self, chrom, start=None, stop=None):
            if chrom not in self._recs:
                raise KeyError(chrom)
            return self._recs[chrom][start:stop]
    pysam.FastaFile = FastaFile
    pysam.depth = lambda *a, **k: []
    sys.modules["pysam"] = pysam

    pyfaidx = types.ModuleType("pyfaidx")

    class Fasta:
        def __init__(self, path):
            self._recs = _read_fasta(path)

        def __getitem__(self, i):
            return self._recs[i][1]
    pyfaidx.Fasta = Fasta
    sys.modules["pyfaidx"] = pyfaidx

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf"):
        sys.modules[name] = mock.MagicMock()

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
