"""AlignmentProcessing.alignment_processing (reference scarmapper/AlignmentProcessing.pyx:18-149):
  CPU : the oracle restatement and scar_alignment_core.h (compiled for the host) against the golden vectors written
        by the compiled reference, and against the compiled reference itself on fresh random alignments (build
        container only);
  GPU : the batched kernel through the C ABI (scar_alignment_batch) and the drop-in module against the golden
        vectors and the oracle."""
import ctypes as C
import gzip
import json
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import alignment_oracle as ao, build_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scarmapper_b200", "csrc")


def golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "alignment_cases.json.gz"), "rt") as f:
        return json.load(f)


def test_oracle_reproduces_the_golden_vectors():
    cases = golden()
    assert len(cases) >= 500
    kinds = [0, 0, 0]
    for k in cases:
        res, summary = ao.alignment_processing(k["cut"], len(k["consensus"]), k["gap_genome_len"], k["genomic"], k["consensus"], [0] * 11)
        assert res[:7] == k["results"] and summary == k["summary"]
        for i in range(3):
            kinds[i] += bool(res[4 + i])
    assert min(kinds) > 50          # the generator really exercises microhomology, insertions and mismatch runs


@pytest.mark.skipif(build_ref.ref_so("AlignmentProcessing") is None, reason="compiled reference module not present")
def test_oracle_equals_the_compiled_reference_on_random_alignments():
    ref = build_ref.load("AlignmentProcessing")
    rng = random.Random(7)
    for _ in range(3000):
        g, c = ao.random_alignment(rng)
        cut, glen = rng.randint(0, len(g)), rng.choice([len(g), len(g), rng.randint(1, len(g) + 30)])
        want = ref.alignment_processing(cut, len(c), glen, g, c, [0] * 11)
        got = ao.alignment_processing(cut, len(c), glen, g, c, [0] * 11)
        assert got[0] == want[0] and got[1] == want[1]


HARNESS = r"""
#include "scar_alignment_core.h"
extern "C" void aln_eval(const uint8_t *g, int ng, const uint8_t *c, int nc, int cut, int gap_con_len, int gap_genome_len,
                         int max_spans, int32_t *counts, int32_t *spans)
{
    ScarAlignmentCounts out;
    scar_alignment_walk(g, ng, c, nc, cut, gap_con_len, gap_genome_len, max_spans, out, spans);
    const int32_t *p = reinterpret_cast<const int32_t *>(&out);
    for (int i = 0; i < 8; ++i) counts[i] = p[i];
}
"""


def test_core_header_compiled_for_the_host_reproduces_the_golden_vectors(tmp_path):
    from scarmapper_b200.AlignmentProcessing import results_of
    (tmp_path / "h.cpp").write_text(HARNESS)
    so = tmp_path / "h.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, str(tmp_path / "h.cpp"), "-o", str(so)])
    lib = C.CDLL(str(so))
    ms = 64
    for k in golden():
        g, c = k["genomic"].encode(), k["consensus"].encode()
        counts = np.zeros(8, np.int32)
        spans = np.zeros((3, ms, 2), np.int32)
        lib.aln_eval(g, len(g), c, len(c), k["cut"], len(c), k["gap_genome_len"], ms, counts.ctypes.data_as(C.c_void_p),
                     spans.ctypes.data_as(C.c_void_p))
        assert counts[7] == 0
        assert results_of(counts, spans, k["genomic"], k["consensus"])[:7] == k["results"]


@pytest.mark.gpu
def test_gpu_alignment_batch_matches_golden_and_oracle():
    from scarmapper_b200 import AlignmentProcessing as AP
    cases = golden()
    rng = random.Random(99)
    extra = []
    for _ in range(5000):
        g, c = ao.random_alignment(rng, rng.choice([None, 1, 5, 400, 900]))
        extra.append({"cut": rng.randint(0, len(g)), "gap_genome_len": rng.choice([len(g), len(g) + 7]), "genomic": g, "consensus": c})
    allc = cases + extra
    counts, spans = AP.alignment_processing_batch([k["cut"] for k in allc], [len(k["consensus"]) for k in allc],
                                                  [k["gap_genome_len"] for k in allc], [k["genomic"] for k in allc],
                                                  [k["consensus"] for k in allc])
    for i, k in enumerate(allc):
        got = AP.results_of(counts[i], spans[i], k["genomic"], k["consensus"])
        want = k["results"] if "results" in k else ao.alignment_processing(k["cut"], len(k["consensus"]), k["gap_genome_len"],
                                                                           k["genomic"], k["consensus"], [0] * 11)[0][:7]
        assert got[:7] == want, i
    # the drop-in signature: same return value and in-place summary_data increments
    k = cases[3]
    summary = [0] * 11
    res, s2 = AP.alignment_processing(k["cut"], len(k["consensus"]), k["gap_genome_len"], k["genomic"], k["consensus"], summary)
    assert res[:7] == k["results"] and s2 is summary and summary == k["summary"] and res[7] == k["consensus"] and res[8] == k["genomic"]
    # malformed alignments fail loudly
    from scarmapper_b200._native import ScarError
    with pytest.raises(ScarError):
        AP.alignment_processing_batch([1], [5], [5], ["ACGTA"], ["ACG"])
