"""
AlignmentProcessing.py — drop-in for the reference's `scarmapper.AlignmentProcessing` module
(scarmapper/AlignmentProcessing.pyx:18-149), backed by scar_alignment_batch (one thread per alignment on the GPU).

    alignment_processing(cutposition, gap_con_len, gap_genome_len, gapped_genomic, gapped_consensus, summary_data)
        same signature, same return value ([left_deletion, right_deletion, left_insertion, right_insertion,
        micro_homology, insertion_seq, variants, gapped_con, gapped_ref], summary_data) and the same in-place
        summary_data increments as the Cython function; one alignment per call.
    alignment_processing_batch(...)
        the batched form: many alignments in one kernel launch.

The reference module is legacy code: scarmapper/setup.py does not build it and nothing imports it; it is served here
because the north star names it.  Well-formed alignments only (two gapped strings of one length, gap_con_len equal to
it) - anything else fails loudly, where the Cython function would raise IndexError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .engine import as_ragged

MAX_SPANS = 64


def alignment_processing_batch(cutpositions, gap_con_lens, gap_genome_lens, gapped_genomics, gapped_consensuses,
                               device: int = 0, max_spans: int = MAX_SPANS):
    """Returns (counts int32[n, 8], spans int32[n, 3, max_spans, 2]); see include/scar_b200.h."""
    g, c = as_ragged(gapped_genomics), as_ragged(gapped_consensuses)
    n = g.n
    assert c.n == n
    ints = [np.ascontiguousarray(v, dtype=np.int32) for v in (cutpositions, gap_con_lens, gap_genome_lens)]
    assert all(v.size == n for v in ints)
    counts = np.zeros((n, 8), dtype=np.int32)
    spans = np.zeros((n, 3, max_spans, 2), dtype=np.int32)
    gs, cs = g.c_struct(), c.c_struct()
    N.check(N.lib().scar_alignment_batch(device, C.byref(gs), C.byref(cs), *[v.ctypes.data for v in ints], n, max_spans,
                                         counts.ctypes.data, spans.ctypes.data))
    if counts[:, 7].any():
        raise N.ScarError(-3, "an alignment holds more than {} events of one kind".format(max_spans))
    return counts, spans


def results_of(counts, spans, gapped_genomic: str, gapped_consensus: str):
    """The 9-item results_list of the reference for one alignment (AlignmentProcessing.pyx:124-138)."""
    def strings(kind, n, src):
        return [src[int(a):int(a) + int(b)] for a, b in spans[kind, :n]]
    mh = strings(0, int(counts[4]), gapped_genomic)
    ins = strings(1, int(counts[5]), gapped_consensus)
    snv = strings(2, int(counts[6]), gapped_consensus)
    return [int(counts[0]), int(counts[1]), int(counts[2]), int(counts[3]), mh if mh else "", ins if ins else "",
            snv if snv else "", gapped_consensus, gapped_genomic]


def alignment_processing(cutposition, gap_con_len, gap_genome_len, gapped_genomic, gapped_consensus, summary_data, device: int = 0):
    counts, spans = alignment_processing_batch([cutposition], [gap_con_len], [gap_genome_len], [gapped_genomic],
                                               [gapped_consensus], device)
    res = results_of(counts[0], spans[0], gapped_genomic, gapped_consensus)
    if res[4]:
        summary_data[8] += 1
    if res[0] > 0:
        summary_data[2] += 1
    if res[1] > 0:
        summary_data[3] += 1
    if res[2] > 0:
        summary_data[5] += 1
    if res[3] > 0:
        summary_data[6] += 1
    return res, summary_data
