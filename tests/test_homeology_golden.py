"""The per-pattern searches of frequency_output against golden vectors produced by the REFERENCE'S OWN methods
(tests/golden/make_homeology_golden.py -> homeology_cases.json.gz; ScarSearch.homeology_search /
templated_insertion_search, scarmapper/INDEL_Processing.py:550-700).  No reference tree is needed to run these:
  CPU: scar_homeology_core.h compiled for the host;
  GPU: the kernel itself through scar_homeology_batch (an oracle independent of the header's host build)."""
import ctypes as C
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scarmapper_b200", "csrc")
_COMP = bytes.maketrans(b"ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", b"TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn")


def rcomp(s):
    return s.encode().translate(_COMP)[::-1].decode()


def cases():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "homeology_cases.json.gz"), "rb") as f:
        return json.loads(f.read())


def check(case, o):
    """o = the 16 int32 of struct ScarHomeology for this case, decoded the way the drop-in decodes them."""
    cons, region = case["consensus"], case["region"]
    oriented = rcomp(region) if case["rc"] else region
    left = ["", "", "", ""] if o[0] < 0 else [o[0], o[1], cons[o[2]:o[2] + o[3]], oriented[o[4]:o[4] + o[5]]]
    right = ["", "", "", ""] if o[6] < 0 else [o[6], o[7], cons[o[8]:o[8] + o[9]], oriented[o[10]:o[10] + o[11]]]
    assert left == case["left"] and right == case["right"], case
    if case["lt"] is not None:
        lt = "" if o[12] < 0 else region[o[12]:o[12] + 5]
        rt = "" if o[13] < 0 else region[o[13]:o[13] + 5]
        if case["rc"]:      # the reference reverse-complements a found template for 3'-strand guides (:694-697)
            lt, rt = (rcomp(lt) if lt else ""), (rcomp(rt) if rt else "")
        assert (lt, rt) == (case["lt"], case["rt"]), case


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


def test_core_header_on_the_host_reproduces_the_reference_vectors(tmp_path):
    (tmp_path / "h.cpp").write_text(HARNESS)
    so = tmp_path / "h.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, str(tmp_path / "h.cpp"), "-o", str(so)])
    lib = C.CDLL(str(so))
    data = cases()
    assert len(data) == 3000
    for case in data:
        oriented = rcomp(case["region"]) if case["rc"] else case["region"]
        out = (C.c_int32 * 16)()
        lib.hom_eval(case["consensus"].encode(), len(case["consensus"]), case["clj"], case["crj"], oriented.encode(),
                     len(oriented), case["insertion"].encode(), len(case["insertion"]), case["region"].encode(),
                     len(case["region"]), case["tlj"], case["trj"], int(len(case["insertion"]) >= 5), out)
        check(case, list(out))


@pytest.mark.gpu
def test_homeology_kernel_reproduces_the_reference_vectors():
    from scarmapper_b200.engine import homeology_batch
    data = cases()
    regions = [c["region"] for c in data]
    targets = [rcomp(c["region"]) if c["rc"] else c["region"] for c in data]       # what homeology_search is handed
    got = homeology_batch([c["consensus"] for c in data], [c["insertion"] for c in data], [c["clj"] for c in data],
                          [c["crj"] for c in data], [c["tlj"] for c in data], [c["trj"] for c in data],
                          list(range(len(data))), targets, regions)
    assert got.shape == (len(data), 16)
    for case, o in zip(data, got.tolist()):
        if len(case["insertion"]) < 5:
            o = o[:12] + [-1, -1] + o[14:]          # templates are searched for insertions of five or more bases only
        check(case, o)
    assert int((np.asarray(got)[:, 0] >= 0).sum()) > 200
