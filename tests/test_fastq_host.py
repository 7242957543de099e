"""CPU-only: the host FASTQ indexer (scar_fastq_index) against a plain Python restatement of
FASTQ_Reader.seq_read + the header split (FASTQ_Tools.py:623-653, INDEL_Processing.py:1155-1156)."""
import random

import numpy as np
import pytest

from scarmapper_b200 import fastq
from scarmapper_b200._native import ScarError


def python_parse(text: str):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    out = []
    for i in range(0, len(lines) - 3, 4):
        name = lines[i].rstrip("\r").strip("@")
        seq, qual = lines[i + 1].strip(), lines[i + 3].strip()
        assert len(seq) == len(qual)
        parts = name.split(":")[-1].split("+")
        out.append((seq, parts[0], parts[1]))
    return out


def test_indexer_matches_python_parser():
    rng = random.Random(2)
    recs = []
    for i in range(3000):
        n = rng.choice([0, 1, 30, 150, 151, 300])
        seq = "".join(rng.choice("ACGTN") for _ in range(n))
        a = "".join(rng.choice("ACGTN") for _ in range(rng.choice([0, 6, 8, 8, 8, 10])))
        b = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 8, 8, 9])))
        extra = rng.choice(["", "+TAIL", "+X+Y"])
        pad = rng.choice(["", " ", "\t"])
        eol = rng.choice(["\n", "\r\n"])
        recs.append("@{}M0:{}:x y:N:0:{}+{}{}{}{}{}{}+{}{}{}".format("@" * rng.randint(0, 2), i, a, b, extra, eol,
                                                                  seq, pad, eol, eol, "I" * n, eol))
    text = "".join(recs)
    want = python_parse(text.replace("\r\n", "\n"))
    fb = fastq.index_text(np.frombuffer(text.encode(), dtype=np.uint8))
    assert fb.n == len(want) == 3000
    for i, (seq, a, b) in enumerate(want):
        assert fb.read(i) == seq
        assert fb.f0.item(i).decode() == a and fb.f1.item(i).decode() == b
    # no trailing newline and a truncated last record
    fb2 = fastq.index_text(np.frombuffer(text.rstrip("\r\n").encode(), dtype=np.uint8))
    assert fb2.n == 3000
    cut = text[:len(text) - 40]
    assert fastq.index_text(np.frombuffer(cut.encode(), dtype=np.uint8), limit=10).n == 10


def test_indexer_errors_where_the_reference_raises():
    with pytest.raises(ScarError):      # len(seq) != len(qual): FASTQ_Tools.py:645-649 raises ValueError
        fastq.index_text(np.frombuffer(b"@a:AC+GT\nACGT\n+\nII\n", dtype=np.uint8))
    with pytest.raises(ScarError):      # no '+': split("+")[1] raises IndexError (INDEL_Processing.py:1156)
        fastq.index_text(np.frombuffer(b"@a:ACGT\nACGT\n+\nIIII\n", dtype=np.uint8))
    ok = fastq.index_text(np.frombuffer(b"@a:ACGT\nACGT\n+\nIIII\n", dtype=np.uint8), want_index_fields=False)
    assert ok.n == 1 and ok.read(0) == "ACGT"


def _odd_header_corpus(n=2000, seed=5):
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        k = rng.choice([0, 1, 30, 150, 151, 300])
        seq = "".join(rng.choice("ACGTN") for _ in range(k))
        a = "".join(rng.choice("ACGTN") for _ in range(rng.choice([0, 6, 8, 8, 8, 10])))
        b = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 8, 8, 9])))
        extra = rng.choice(["", "+TAIL", "+X+Y", " "])
        pad = rng.choice(["", " ", "\t"])
        eol = rng.choice(["\n", "\r\n"])
        recs.append("@{}M0:{}:x y:N:0:{}+{}{}{}{}{}{}+{}{}{}".format("@" * rng.randint(0, 2), i, a, b, extra, eol,
                                                                  seq, pad, eol, eol, "I" * k, eol))
    return "".join(recs)


def test_indexer_matches_the_reference_fastq_reader_on_gzip(tmp_path):
    """The importable reference itself: Valkyries.FASTQ_Tools.FASTQ_Reader (gzip branch - its text branch opens the
    file in mode 'rU', which Python 3.11+ rejects) driven the way consensus_demultiplex drives it
    (next(fastq1.seq_read()), INDEL_Processing.py:900-903), plus the header split of index_matching (:1155-1156), on
    a corpus of odd headers, CRLF line ends, padded sequence lines and empty reads.  scarmapper_b200.fastq.load must
    deliver the same sequences and index fields."""
    import gzip
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("reference tree not present")
    ref_shims.install()
    from Valkyries import FASTQ_Tools
    for final_newline in (True, False):
        text = _odd_header_corpus()
        if not final_newline:
            text = text.rstrip("\r\n")
        path = str(tmp_path / "odd_{}.fastq.gz".format(int(final_newline)))
        with gzip.open(path, "wb") as f:
            f.write(text.encode())
        reader = FASTQ_Tools.FASTQ_Reader(path, None)
        want = []
        while True:
            try:
                r = next(reader.seq_read())
            except StopIteration:
                break
            parts = r.name.split(":")[-1].split("+")
            want.append((r.seq, parts[0], parts[1]))
        fb = fastq.load(path)
        assert fb.n == len(want) == 2000
        for i, (seq, a, b) in enumerate(want):
            assert fb.read(i) == seq, i
            assert fb.f0.item(i).decode() == a and fb.f1.item(i).decode() == b, i
