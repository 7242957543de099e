"""
CPU tests of the parallel gzip inflate (csrc/scar_pgzip.h through scar_gunzip_host): what the
reference gets from Python's gzip module in FASTQ_Reader (Valkyries/FASTQ_Tools.py:571-575) - here zlib / gzip of the
standard library is the oracle.  Exactness never depends on the guessed block starts: the cases below force wrong
guesses, absent guesses, stored and fixed blocks, members, padding and damage.
"""
import gzip
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scarmapper_b200 import _native as N          # noqa: E402
from scarmapper_b200 import fastq                 # noqa: E402


def fastq_text(n, seed=3, quality="mixed"):
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    tmpl = bases[rng.integers(0, 4, 150)]
    out = []
    for i in range(n):
        s = tmpl.copy()
        m = rng.random(150) < 0.04
        s[m] = bases[rng.integers(0, 4, int(m.sum()))]
        q = (b"I" * 150) if quality == "const" else bytes(rng.choice(np.frombuffer(b"#,:FFFFF", dtype=np.uint8), 150))
        out.append(b"@M01:7:000-AB:1:%d:%d:%d 1:N:0:ACGTACGT+TTGCAATG\n" % (1101 + i // 3000, rng.integers(1000, 30000),
                                                                           rng.integers(1000, 30000)))
        out.append(s.tobytes() + b"\n+\n" + q + b"\n")
    return b"".join(out)


_DEFLATED = {}


def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, mem=8):
    key = (id(data), len(data), level, strategy, wbits)
    if key not in _DEFLATED:
        co = zlib.compressobj(level, zlib.DEFLATED, wbits, mem, strategy)
        _DEFLATED[key] = (data, co.compress(data) + co.flush())
    return _DEFLATED[key][1]


TEXT = fastq_text(40000)


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("threads,chunk", [(1, 0), (4, 0), (4, 1 << 16), (8, 20000), (3, 4099)])
def test_single_stream_matches_zlib(level, threads, chunk):
    gz = deflate(TEXT, level)
    st = {}
    got = fastq.gunzip(gz, threads=threads, chunk_bytes=chunk, stats=st)
    assert got.tobytes() == TEXT
    assert st["threads"] == threads
    if chunk and chunk <= 1 << 16:
        assert st["chunks"] > 8 and st["guessed"] >= 8, st      # the second pass really ran on guesses


def test_members_padding_and_header_fields(tmp_path):
    k = len(TEXT) // 3
    import io
    buf = io.BytesIO()
    with gzip.GzipFile(filename="a name.fastq", mode="wb", fileobj=buf, compresslevel=5, mtime=0) as g:   # FNAME
        g.write(TEXT[:k])
    extra = bytearray(deflate(TEXT[k:2 * k + 7], 9))
    extra[3] |= 4                                                   # FEXTRA: a 6-byte field nobody reads
    extra[10:10] = b"\x06\x00" + b"XY\x02\x00\x01\x02"
    gz = buf.getvalue() + bytes(extra) + deflate(b"", 6) + deflate(TEXT[2 * k + 7:], 1) + b"\0" * 700
    for threads, chunk in ((1, 0), (4, 30000), (6, 1 << 20)):
        assert fastq.gunzip(gz, threads=threads, chunk_bytes=chunk).tobytes() == TEXT
    assert gzip.decompress(gz[:-700]) == TEXT                # (the oracle agrees that this is the text)


def test_stored_fixed_and_tiny_streams():
    cases = [b"", b"A", b"ACGT\n" * 3, TEXT[:1000], os.urandom(70000)]
    for data in cases:
        for level, strategy in ((0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_HUFFMAN_ONLY),
                                (9, zlib.Z_RLE)):
            gz = deflate(data, level, strategy)
            assert fastq.gunzip(gz, threads=3, chunk_bytes=4096, capacity=len(data)).tobytes() == data


def test_binary_data_has_no_guessable_blocks_and_still_inflates():
    rng = np.random.default_rng(11)
    data = (rng.integers(0, 256, 600000, dtype=np.uint8) & rng.integers(0, 256, 600000, dtype=np.uint8)).tobytes() + TEXT[:300000]
    gz = deflate(data, 6)
    st = {}
    assert fastq.gunzip(gz, threads=4, chunk_bytes=50000, stats=st).tobytes() == data
    assert st["chunks"] > 5


def test_flush_points_and_mixed_block_types():
    # a writer that flushes: stored empty blocks (sync), full flushes (window reset), levels changed in mid-stream
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    out, step = [], len(TEXT) // 23
    for i, a in enumerate(range(0, len(TEXT), step)):
        out.append(co.compress(TEXT[a:a + step]))
        out.append(co.flush((zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH, zlib.Z_BLOCK if hasattr(zlib, "Z_BLOCK") else zlib.Z_SYNC_FLUSH)[i % 4]))
    out.append(co.flush())
    gz = b"".join(out)
    for chunk in (0, 10007, 1 << 17):
        assert fastq.gunzip(gz, threads=5, chunk_bytes=chunk).tobytes() == TEXT


def test_a_decoy_block_header_inside_the_data_does_not_matter():
    # text that CONTAINS a valid dynamic block (a stored block carrying compressed bytes): a guess may land on the
    # decoy; the chain of exact block boundaries throws it away
    inner = deflate(TEXT[:200000], 6, wbits=-15)
    data = TEXT[:50000] + inner + TEXT[50000:120000]
    for level in (0, 1, 6):
        gz = deflate(data, level)
        for chunk in (8192, 30000):
            assert fastq.gunzip(gz, threads=4, chunk_bytes=chunk).tobytes() == data


def test_damage_is_reported():
    gz = deflate(TEXT, 6)
    L = N.lib()
    bad = {
        "truncated": gz[:len(gz) // 2],
        "trailer_crc": gz[:-8] + bytes([gz[-8] ^ 1]) + gz[-7:],
        "trailer_len": gz[:-4] + bytes([gz[-4] ^ 1]) + gz[-3:],
        "garbage_mid": gz[:len(gz) // 2] + b"\xff" * 64 + gz[len(gz) // 2 + 64:],
        "bit_flip": gz[:len(gz) // 3] + bytes([gz[len(gz) // 3] ^ 0x10]) + gz[len(gz) // 3 + 1:],
        "not_gzip": b"@r1\nACGT\n+\nIIII\n" * 20,
        "junk_after_member": gz + b"junk that is no member header",
    }
    for name, data in bad.items():
        for threads, chunk in ((1, 0), (4, 20000)):
            with pytest.raises(N.ScarError) as e:
                fastq.gunzip(data, threads=threads, chunk_bytes=chunk, capacity=len(TEXT) + 100)
            assert "gzip" in str(e.value), name
    # a buffer that is too small is an error, not an overrun
    with pytest.raises(N.ScarError):
        fastq.gunzip(gz, threads=2, capacity=len(TEXT) - 1)
    assert L is not None


def test_read_text_uses_it_for_gzip_files(tmp_path):
    p = str(tmp_path / "r.fastq.gz")
    with open(p, "wb") as f:
        f.write(deflate(TEXT[:500000], 4))
    assert fastq.read_text(p).tobytes() == TEXT[:500000]


def test_large_window_distances_and_long_runs():
    # distances up to 32 KiB across chunk borders, runs (distance 1) and short periods (2, 3)
    rng = np.random.default_rng(5)
    block = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), 32768))
    data = block + b"I" * 100000 + block + b"AB" * 30000 + b"XYZ" * 20000 + block[::-1] + block + TEXT[:100000] + block
    for level in (1, 9):
        gz = deflate(data, level)
        for chunk in (2048, 16384):
            assert fastq.gunzip(gz, threads=4, chunk_bytes=chunk).tobytes() == data


def test_text_that_inflates_beyond_the_chunk_budget_is_handed_back(monkeypatch):
    # very compressible text: a chunk that has produced more than its budget stops at a block boundary and the reader
    # decodes on from there (memory stays bounded); the result is the same
    data = (b"@r\n" + b"A" * 150 + b"\n+\n" + b"I" * 150 + b"\n") * 60000 + TEXT[:200000]
    gz = deflate(data, 6)
    monkeypatch.setenv("SCAR_PGZIP_BUDGET", "300000")
    st = {}
    for threads, chunk in ((1, 0), (4, 2048), (4, 1 << 14)):
        assert fastq.gunzip(gz, threads=threads, chunk_bytes=chunk, stats=st).tobytes() == data
    assert st["redone"] > 0


def test_seeded_fuzz_against_zlib():
    """Random mixtures (text-like runs, repeats at random distances, binary noise), random deflate parameters, flush
    points, member splits, chunk sizes and thread counts: always the bytes zlib gives."""
    rng = np.random.default_rng(2026)
    letters = np.frombuffer(b"ACGTN@+:IF#,0123456789\n", dtype=np.uint8)
    for case in range(60):
        parts = []
        for _ in range(int(rng.integers(1, 12))):
            kind = int(rng.integers(0, 4))
            n = int(rng.integers(1, 60000))
            if kind == 0:
                parts.append(bytes(letters[rng.integers(0, len(letters), n)]))
            elif kind == 1 and parts:
                src = b"".join(parts)
                a = int(rng.integers(0, len(src)))
                parts.append(src[a:a + n] * int(rng.integers(1, 4)))
            elif kind == 2:
                parts.append(bytes(rng.integers(0, 256, n, dtype=np.uint8)))
            else:
                parts.append(bytes([int(rng.integers(32, 127))]) * n)
        data = b"".join(parts)
        level = int(rng.choice([1, 3, 6, 9]))
        strategy = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        n_members = int(rng.integers(1, 4))
        cuts = sorted(int(v) for v in rng.integers(0, len(data) + 1, n_members - 1))
        gz = b""
        for a, b in zip([0] + cuts, cuts + [len(data)]):
            co = zlib.compressobj(level, zlib.DEFLATED, 31, int(rng.integers(1, 10)), strategy)
            piece, out = data[a:b], []
            step = max(1, len(piece) // int(rng.integers(1, 6)))
            for q in range(0, len(piece), step):
                out.append(co.compress(piece[q:q + step]))
                if rng.random() < 0.3:
                    out.append(co.flush(int(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))))
            out.append(co.flush())
            gz += b"".join(out) + b"\0" * int(rng.integers(0, 3) * 8)
        threads = int(rng.integers(1, 7))
        chunk = int(rng.choice([0, 1024, 4096, 30000, 1 << 17]))
        got = fastq.gunzip(gz, threads=threads, chunk_bytes=chunk, capacity=len(data))
        assert got.tobytes() == data, (case, level, strategy, n_members, threads, chunk)
