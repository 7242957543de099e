"""scar_homeology_core.h (the per-unique-scar searches of frequency_output) compiled for the host and compared
with the reference's own ScarSearch.homeology_search / templated_insertion_search
(scarmapper/INDEL_Processing.py:550-700) on randomised inputs, including junctions near both ends of the
strings where Python's slice arithmetic wraps.  Skipped when /root/reference is absent (the GPU box)."""
import ctypes as C
import os
import random
import subprocess
import types

import pytest

from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scarmapper_b200", "csrc")
pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")

HARNESS = r"""
#include "scar_homeology_core.h"
extern "C" void hom_eval(const uint8_t *cons, int n_cons, int clj, int crj, const uint8_t *target, int n_target,
                         const uint8_t *ins, int n_ins, const uint8_t *region, int n_region, int tlj, int trj,
                         int want_templates, int32_t *out)
{
    ScarHomeology h;
    hom_homeology(cons, n_cons, clj, crj, target, n_target, n_ins, tlj, trj, h);
    h.lt_start = h.rt_start = -1;
    if (want_templates) hom_templates(ins, n_ins, region, n_region, tlj, trj, h);
    const int32_t *p = reinterpret_cast<const int32_t *>(&h);
    for (int i = 0; i < 16; ++i) out[i] = p[i];
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("homeology")
    (d / "h.cpp").write_text(HARNESS)
    so = d / "h.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, str(d / "h.cpp"), "-o", str(so)])
    return C.CDLL(str(so))


def evaluate(lib, cons, clj, crj, target, ins, region, tlj, trj, templates):
    out = (C.c_int32 * 16)()
    lib.hom_eval(cons.encode(), len(cons), clj, crj, target.encode(), len(target), ins.encode(), len(ins),
                 region.encode(), len(region), tlj, trj, int(templates), out)
    o = list(out)
    left = ["", "", "", ""] if o[0] < 0 else [o[0], o[1], cons[o[2]:o[2] + o[3]], target[o[4]:o[4] + o[5]]]
    right = ["", "", "", ""] if o[6] < 0 else [o[6], o[7], cons[o[8]:o[8] + o[9]], target[o[10]:o[10] + o[11]]]
    lt = "" if o[12] < 0 else region[o[12]:o[12] + 5]
    rt = "" if o[13] < 0 else region[o[13]:o[13] + 5]
    return left, right, lt, rt


def test_searches_match_the_reference_methods(lib):
    ref_shims.install()
    from scarmapper import INDEL_Processing
    from Valkyries import Sequence_Magic
    rng = random.Random(17)
    n_hom = n_tmpl = n_found = 0
    for case in range(6000):
        alpha = "ACGT" if case % 5 else "ACGTNacgtR"
        L = rng.randint(40, 260)
        target = "".join(rng.choice("ACGT") for _ in range(L))
        cut = rng.randint(5, L - 5)
        # a consensus assembled from the locus flanks so that homologies really occur
        dl, dr = rng.randint(0, min(30, cut)), rng.randint(0, min(30, L - cut))
        ins = "".join(rng.choice(alpha) for _ in range(rng.choice([0, 0, 1, 2, 3, 5, 6, 9, 14, 25])))
        if ins and rng.random() < 0.4 and len(ins) >= 5:      # templated: a copy of nearby locus sequence
            a = rng.randint(max(0, cut - 45), max(0, cut - 6))
            piece = target[a:a + len(ins)]
            ins = Sequence_Magic.rcomp(piece) if rng.random() < 0.5 else piece
        left_flank = target[max(0, cut - dl - rng.randint(5, 60)):cut - dl]
        right_flank = target[cut + dr:cut + dr + rng.randint(5, 60)]
        cons = left_flank + ins + right_flank
        clj, crj = len(left_flank), len(left_flank) + len(ins)
        tlj, trj = cut - dl, cut + dr
        if case % 7 == 0:                                      # junctions pushed to the string ends / odd values
            clj = rng.randint(0, max(0, len(cons)))
            crj = rng.randint(0, max(0, len(cons)))
            tlj = rng.randint(0, L)
            trj = rng.randint(tlj, L)
        rc = case % 3 == 0
        oriented = Sequence_Magic.rcomp(target) if rc else target
        me = types.SimpleNamespace(target_region=target, target_dict={"t": [None] * 5 + ["YES" if rc else "NO"]})
        want_l, want_r = INDEL_Processing.ScarSearch.homeology_search(me, cons, clj, crj, oriented, ins, tlj, trj)
        templates = len(ins) >= 5
        got_l, got_r, got_lt, got_rt = evaluate(lib, cons, clj, crj, oriented, ins, target, tlj, trj, templates)
        assert list(want_l) == got_l and list(want_r) == got_r, (case, cons, clj, crj, ins, tlj, trj)
        n_hom += want_l[0] != "" or want_r[0] != ""
        if templates:
            want_lt, want_rt = INDEL_Processing.ScarSearch.templated_insertion_search(me, ins, tlj, trj, "t")
            if rc:   # the reference reverse-complements a found template for 3'-strand guides; positions are the same
                got_lt, got_rt = (Sequence_Magic.rcomp(got_lt) if got_lt else ""), (Sequence_Magic.rcomp(got_rt) if got_rt else "")
            assert (want_lt, want_rt) == (got_lt, got_rt), (case, ins, tlj, trj)
            n_tmpl += 1
            n_found += bool(want_lt or want_rt)
    assert n_hom > 500 and n_tmpl > 500 and n_found > 50       # the generator really exercises the searches


def test_dropin_serves_batch_results_and_falls_back_to_the_reference_method():
    """dropin.pattern_search_batch + serve_pattern_searches: calls with the arguments frequency_output derives are answered from the
    batch (identical to the reference's methods, 5' and 3' guides); any other call runs the reference's method."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_engine import host_homeology_batch
    ref_shims.install()
    from scarmapper import INDEL_Processing
    from Valkyries import Sequence_Magic
    from scarmapper_b200 import dropin
    rng = random.Random(5)
    region = "".join(rng.choice("ACGT") for _ in range(300))
    cut = 150
    for rc in (False, True):
        class Search(INDEL_Processing.ScarSearch):
            def __init__(self):                      # none of the reference's constructor work is needed here
                self.index_name = "s1"
                self.index_dict = {"s1": [None] * 7 + ["t"]}
                self.target_dict = {"t": [None] * 5 + ["YES" if rc else "NO"]}
                self.target_region = region
        me = Search()
        table = {}
        for k in range(300):
            dl, dr = rng.randint(0, 25), rng.randint(0, 25)
            ins = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 0, 2, 5, 9])))
            if len(ins) >= 5 and k % 2:
                ins = region[cut - 30:cut - 30 + len(ins)]
            lf, rf = region[cut - dl - 40:cut - dl], region[cut + dr:cut + dr + 40]
            sub = ["x" * dl, "y" * dr, ins, "", lf + ins + rf, len(lf), len(lf) + len(ins), cut - dl, cut + dr, ""]
            table["k%d" % k] = [1, sub]
        dropin._HOMEOLOGY_BATCH[0] = host_homeology_batch
        dropin.PATTERN_SEARCH_STATS.update(served=0, missed=0)
        try:
            # what the parent does for the whole run (records + reads -> one batched call) ...
            import numpy as np
            from scarmapper_b200 import _native as N
            rec = np.zeros(len(table), dtype=N.RECORD_DTYPE)
            for k, (_c, sub) in enumerate(table.values()):
                rec[k]["clj"], rec[k]["crj"], rec[k]["tlj"], rec[k]["trj"] = sub[5], sub[6], sub[7], sub[8]
                rec[k]["flags"] = N.FL_INSERTION if sub[2] else 0
                rec[k]["status"] = N.ST_SCAR
            reads = [sub[4] for _c, sub in table.values()]
            out = dropin.pattern_search_batch(rec, reads, "Illumina", np.arange(len(table)), np.zeros(len(table), np.int64),
                                              [0], [region], [rc], 0)
            # ... and what a pool worker does with its sample's slice
            dropin.serve_pattern_searches(me, table, out)
        finally:
            dropin._HOMEOLOGY_BATCH[0] = None
        oriented = Sequence_Magic.rcomp(region) if rc else region
        for _count, sub in table.values():           # the derivation of frequency_output (:299-328)
            ins, cons, a, b, c, d = sub[2], sub[4], sub[5], sub[6], sub[7], sub[8]
            if rc:
                ins, cons = Sequence_Magic.rcomp(ins), Sequence_Magic.rcomp(cons)
                a, b = len(cons) - sub[6], len(cons) - sub[5]
                c, d = len(region) - sub[8], len(region) - sub[7]
            want = INDEL_Processing.ScarSearch.homeology_search(me, cons, a, b, oriented, ins, c, d)
            got = me.homeology_search(cons, a, b, oriented, ins, c, d)
            assert [list(x) for x in want] == [list(x) for x in got]
            if len(ins) >= 5:
                assert INDEL_Processing.ScarSearch.templated_insertion_search(me, ins, c, d, "t") == \
                    me.templated_insertion_search(ins, c, d, "t")
        assert dropin.PATTERN_SEARCH_STATS["missed"] == 0 and dropin.PATTERN_SEARCH_STATS["served"] >= 300
        # arguments that are not in the batch: the reference's own method answers
        want = INDEL_Processing.ScarSearch.homeology_search(me, region[100:180], 40, 40, oriented, "", 140, 160)
        assert [list(x) for x in me.homeology_search(region[100:180], 40, 40, oriented, "", 140, 160)] == [list(x) for x in want]
        assert dropin.PATTERN_SEARCH_STATS["missed"] == 1
