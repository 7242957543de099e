"""Host-side checks of the arithmetic the kernels share (no GPU): the headers scar_swar.h and
scar_demux_core.h compile for the host, a small g++ harness exposes them through ctypes, and every formula
is compared with its plain definition.

  * SWAR byte classification (K1 pack, K2 demux): 2-bit codes, ACGT mask, the multiply gathers;
  * the Myers/Hyyro bit-vector GLOBAL edit distance against a plain dynamic program, and accepts()
    against match_maker(q, u) = lev(q, u) - (len(u) - len(q)) <= t  (reference
    Valkyries/Sequence_Magic.py:123-136, scarmapper/INDEL_Processing.py:1173-1192).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scarmapper_b200", "csrc")

HARNESS = r"""
#include <cstring>
#include <vector>
#include "scar_swar.h"
#include "scar_demux_core.h"

extern "C" {
// every formula of scar_swar.h on n words
void swar_eval(const uint32_t *w, int64_t n, uint32_t *code8, uint32_t *valid4, uint32_t *zb, uint32_t *top)
{
    for (int64_t i = 0; i < n; ++i) {
        classify4(w[i], code8[i], valid4[i]);
        zb[i] = flags_to_nibble(zero_bytes(w[i]));
        top[i] = codes_to_top_byte(w[i]) >> 24;
    }
}

// the 16-byte predicate of the fused pack: non-zero iff a byte is not A, C, G, T
void swar_any16(const uint32_t *w, int64_t n_cells, uint32_t *any)
{
    for (int64_t i = 0; i < n_cells; ++i) any[i] = acgt_any16(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]) != 0u;
}

// pattern q (length m <= 64), text u (length n): the bit-vector distance and the accepts() verdict for t
void demux_eval(const uint8_t *q, int m, const uint8_t *u, int n, int t, int *lev, int *ok_codes, int *ok_plain)
{
    uint8_t cmap[256];
    std::memset(cmap, 0, sizeof cmap);
    int n_classes = 1;
    for (int i = 0; i < m; ++i) if (!cmap[q[i]]) cmap[q[i]] = (uint8_t)n_classes++;
    std::vector<uint64_t> peq(n_classes, 0ull);
    DemuxString d; d.len = m; d.peq_off = 0; d.code[0] = d.code[1] = 0;
    for (int i = 0; i < m; ++i) {
        peq[cmap[q[i]]] |= 1ull << i;
        if (i < 32) d.code[i >> 4] |= (uint64_t)cmap[q[i]] << (4 * (i & 15));
    }
    *lev = myers_global(peq.data(), m, u, n, cmap);
    uint64_t uc[2];
    class_codes(u, n, cmap, uc);
    *ok_codes = accepts(d, peq.data(), u, n, uc, cmap, t, true) ? 1 : 0;
    *ok_plain = accepts(d, peq.data(), u, n, uc, cmap, t, false) ? 1 : 0;
}
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostlogic")
    src = d / "harness.cpp"
    src.write_text(HARNESS)
    so = d / "harness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, str(src), "-o", str(so)])
    return C.CDLL(str(so))


def _naive_classify(words):
    b = words.view(np.uint8).reshape(-1, 4)                      # little endian: byte k of the word
    code = (b >> 1) & 3
    code8 = code[:, 0] | (code[:, 1] << 2) | (code[:, 2] << 4) | (code[:, 3] << 6)
    ok = np.isin(b, np.frombuffer(b"ACGT", dtype=np.uint8))
    valid4 = ok[:, 0] | (ok[:, 1].astype(np.uint8) << 1) | (ok[:, 2].astype(np.uint8) << 2) | (ok[:, 3].astype(np.uint8) << 3)
    z = b == 0
    zb = z[:, 0] | (z[:, 1].astype(np.uint8) << 1) | (z[:, 2].astype(np.uint8) << 2) | (z[:, 3].astype(np.uint8) << 3)
    return code8.astype(np.uint32), valid4.astype(np.uint32), zb.astype(np.uint32)


def _swar(lib, words):
    n = words.size
    out = [np.empty(n, dtype=np.uint32) for _ in range(4)]
    lib.swar_eval(words.ctypes.data_as(C.c_void_p), C.c_int64(n), *[o.ctypes.data_as(C.c_void_p) for o in out])
    return out


def test_swar_classification_every_byte_in_every_lane(lib):
    # every byte value in every lane, against every byte value in one other lane, the rest random
    rng = np.random.default_rng(7)
    blocks = []
    for lane in range(4):
        for other in range(4):
            if other == lane:
                continue
            a, b = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32), indexing="ij")
            w = (a << (8 * lane)) | (b << (8 * other))
            rest = [k for k in range(4) if k not in (lane, other)]
            for k in rest:
                w = w | (rng.integers(0, 256, size=w.shape, dtype=np.uint32) << (8 * k))
            blocks.append(w.reshape(-1))
    words = np.ascontiguousarray(np.concatenate(blocks))
    code8, valid4, zb, top = _swar(lib, words)
    want_code, want_valid, want_zb = _naive_classify(words)
    assert np.array_equal(code8, want_code)
    assert np.array_equal(top, want_code)
    assert np.array_equal(valid4, want_valid)
    assert np.array_equal(zb, want_zb)


def test_swar_sixteen_byte_predicate(lib):
    """acgt_any16 (what the fused pack tests per cell): every byte value in every one of the 16 positions of an otherwise
    valid cell, and random read-like cells."""
    rng = np.random.default_rng(4)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    cells = letters[rng.integers(0, 4, (16 * 256, 16))].copy()
    for pos in range(16):
        cells[pos * 256:(pos + 1) * 256, pos] = np.arange(256, dtype=np.uint8)
    noisy = letters[rng.integers(0, 4, (20000, 16))].copy()
    hit = rng.random((20000, 16)) < 0.02
    noisy[hit] = rng.integers(0, 256, int(hit.sum()), dtype=np.uint8)
    allc = np.ascontiguousarray(np.concatenate([cells, noisy]))
    words = allc.view(np.uint32).reshape(-1)
    got = np.empty(allc.shape[0], dtype=np.uint32)
    lib.swar_any16(words.ctypes.data_as(C.c_void_p), C.c_int64(allc.shape[0]), got.ctypes.data_as(C.c_void_p))
    want = (~np.isin(allc, letters)).any(axis=1)
    assert (got.astype(bool) == want).all()


def test_swar_classification_read_like_words(lib):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTNnacgtRYKM\x00\n@+-.EUFBDHSVW", dtype=np.uint8)
    words = np.ascontiguousarray(alphabet[rng.integers(0, alphabet.size, size=(400_000, 4))]).view(np.uint32).reshape(-1)
    code8, valid4, zb, top = _swar(lib, words)
    want_code, want_valid, want_zb = _naive_classify(words)
    assert np.array_equal(code8, want_code) and np.array_equal(top, want_code)
    assert np.array_equal(valid4, want_valid) and np.array_equal(zb, want_zb)


def _lev(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _demux(lib, q, u, t):
    lev, okc, okp = C.c_int(0), C.c_int(0), C.c_int(0)
    lib.demux_eval(q, len(q), u, len(u), t, C.byref(lev), C.byref(okc), C.byref(okp))
    return lev.value, okc.value, okp.value


def test_bitvector_distance_and_accepts_match_the_definition(lib):
    rng = np.random.default_rng(3)
    alpha = b"ACGTN"
    cases = [(b"", b""), (b"", b"ACG"), (b"ACGT", b""), (b"A" * 64, b"A" * 64), (b"ACGT" * 16, b"TGCA" * 16)]
    for _ in range(4000):
        m, n = int(rng.integers(0, 13)), int(rng.integers(0, 13))
        q = bytes(alpha[i] for i in rng.integers(0, 4, size=m))
        if rng.random() < 0.6 and m:                                 # a near copy: few edits
            u = bytearray(q)
            for _ in range(int(rng.integers(0, 4))):
                op = rng.integers(0, 3)
                pos = int(rng.integers(0, len(u) + 1))
                if op == 0 and len(u):
                    u[min(pos, len(u) - 1)] = alpha[int(rng.integers(0, 5))]
                elif op == 1:
                    u.insert(pos, alpha[int(rng.integers(0, 5))])
                elif len(u):
                    del u[min(pos, len(u) - 1)]
            u = bytes(u)
        else:
            u = bytes(alpha[i] for i in rng.integers(0, 5, size=n))
        cases.append((q, u))
    for _ in range(300):                                             # long patterns up to the 64-nt limit
        m, n = int(rng.integers(30, 65)), int(rng.integers(20, 80))
        cases.append((bytes(alpha[i] for i in rng.integers(0, 4, size=m)), bytes(alpha[i] for i in rng.integers(0, 5, size=n))))
    for q, u in cases:
        want = _lev(q, u)
        for t in (0, 1, 3):
            lev, okc, okp = _demux(lib, q, u, t)
            assert lev == want, (q, u)
            verdict = int(want - (len(u) - len(q)) <= t)             # match_maker(q, u) <= t
            assert okc == verdict and okp == verdict, (q, u, t, want)
