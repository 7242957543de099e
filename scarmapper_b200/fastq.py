"""
fastq.py — the batched loader's front end: a whole FASTQ file -> views for the GPU path.

Replaces the per-read generator FASTQ_Reader.seq_read (reference Valkyries/FASTQ_Tools.py:562-653:
text or gzip chosen by libmagic MIME type; here by the gzip magic bytes) and the per-index header
split of index_matching (scarmapper/INDEL_Processing.py:1155-1156).  The file text is read (and
inflated) once; scar_fastq_index (C, include/scar_b200.h) emits (offset, length) views into it for
the sequence and the two Illumina index fields, and the text itself is what gets copied to the GPU.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os

import numpy as np

from . import _native as N
from .engine import HostRagged


def gunzip(data, threads: int = 0, chunk_bytes: int = 0, capacity: int | None = None, stats: dict | None = None) -> np.ndarray:
    """A gzip stream (one member or several) -> text, inflated by several threads (scar_gunzip_host,
    csrc/scar_pgzip.h: speculative two-pass decoding, CRC-checked).  Replaces gzip.open() of FASTQ_Reader
    (Valkyries/FASTQ_Tools.py:571-575).  `capacity`: size of the text when known; otherwise the length field of the
    last member's trailer is tried first (exact for a single member under 4 GiB) and the buffer grows as needed."""
    gz = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    L = N.lib()
    tail = gz[-4:].tobytes() if gz.size >= 18 else b""
    cap = int(capacity) if capacity is not None else (int.from_bytes(tail, "little") if tail else 0)
    st = np.zeros(4, dtype=np.int64)
    while True:
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_int64(0)
        rc = L.scar_gunzip_host(gz.ctypes.data, gz.size, int(threads), int(chunk_bytes), out.ctypes.data, cap, C.byref(n),
                                st.ctypes.data)
        if rc != 0 and capacity is None and b"longer than the output buffer" in L.scar_last_error():
            cap = max(2 * cap, 16 * gz.size, 1 << 20)
            continue
        N.check(rc)
        if stats is not None:
            stats.update(chunks=int(st[0]), guessed=int(st[1]), redone=int(st[2]), threads=int(st[3]))
        return out[:n.value]


def read_text(path: str) -> np.ndarray:
    """File bytes (inflated when gzip) as a uint8 array."""
    if len(path) < 3 or not os.path.isfile(path):
        # FASTQ_Tools.py:596-602 logs a warning and exits; here the caller decides
        raise FileNotFoundError("FASTQ file {} not found".format(path))
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        return gunzip(np.fromfile(path, dtype=np.uint8))
    else:
        with open(path, "rb") as f:
            data = f.read()
    return np.frombuffer(data, dtype=np.uint8)


class FastqBatch:
    """Views into one FASTQ text buffer."""

    def __init__(self, text, seq_off, seq_len, f0_off=None, f0_len=None, f1_off=None, f1_len=None):
        self.text = text
        self.seq = HostRagged.from_views(text, seq_off, seq_len)
        self.f0 = HostRagged.from_views(text, f0_off, f0_len) if f0_off is not None else None
        self.f1 = HostRagged.from_views(text, f1_off, f1_len) if f1_off is not None else None
        self.n = len(seq_len)

    def read(self, i: int) -> str:
        return self.seq.item(i).decode("latin-1")


def index_text(text: np.ndarray, want_index_fields: bool = True, limit: int | None = None) -> FastqBatch:
    """Index every complete record of `text` (limit: at most that many records)."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    # 4 lines per record, at least 7 bytes ("@\nA\n+\nI"): an upper bound without a counting pass
    cap = int(np.count_nonzero(text == 10)) // 4 + 1
    if limit is not None:
        cap = min(cap, int(limit))
    so, sl = np.empty(cap, np.int64), np.empty(cap, np.int32)
    if want_index_fields:
        ao, al, bo, bl = np.empty(cap, np.int64), np.empty(cap, np.int32), np.empty(cap, np.int64), np.empty(cap, np.int32)
    else:
        ao = al = bo = bl = None
    n, used = C.c_int64(0), C.c_int64(0)
    p = lambda a: None if a is None else a.ctypes.data
    N.check(N.lib().scar_fastq_index(text.ctypes.data, text.size, 1, int(want_index_fields), cap, p(so), p(sl),
                                     p(ao), p(al), p(bo), p(bl), C.byref(n), C.byref(used)))
    k = n.value
    if want_index_fields:
        return FastqBatch(text, so[:k], sl[:k], ao[:k], al[:k], bo[:k], bl[:k])
    return FastqBatch(text, so[:k], sl[:k])


def load(path: str, want_index_fields: bool = True, limit: int | None = None) -> FastqBatch:
    return index_text(read_text(path), want_index_fields, limit)
