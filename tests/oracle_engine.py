"""
Test double with ScarEngine's result interface, computed by the CPU oracle.  Lets the drop-in glue
(scarmapper_b200/dropin.py: unpacking GPU results into the reference's objects and files) be tested in
the build container, where the reference is present but no GPU is.  Lives in tests/ only.
"""
import os

import numpy as np

from oracle import scar_oracle as orc


class OracleEngine:
    def __init__(self, prob, device=0):
        self.prob = prob
        self.o = orc.Problem(targets=prob.targets, cutsites=prob.cutsites, sample_target=prob.sample_target,
                             left_index=prob.left_index, right_index=prob.right_index, platform=prob.platform,
                             phasing=prob.phasing, hr_donor=prob.hr_donor, n_limit=prob.n_limit,
                             min_length=prob.min_length)
        self.records = None
        self.reads = None

    def process_host(self, seq, f0=None, f1=None, first_index=0, want_records=True, out=None):
        reads = [seq.item(i).decode("latin-1") for i in range(seq.n)]
        a = [f0.item(i).decode("latin-1") for i in range(f0.n)] if f0 is not None else None
        b = [f1.item(i).decode("latin-1") for i in range(f1.n)] if f1 is not None else None
        self.reads = reads
        self.records = orc.classify_batch(self.o, reads, a, b)
        return self.records

    def counters(self):
        return orc.counters(self.o, self.records), int((self.records["status"] == orc.ST_UNASSIGNED).sum())

    def phase_hist(self):
        h = np.zeros((self.o.n_samples, 2, 34), dtype=np.int64)
        ok = self.records["status"] != orc.ST_UNASSIGNED
        for side, col in enumerate(("phase_r1", "phase_r2")):
            v = self.records[col].astype(np.int64)
            bins = np.where(v == -2, 33, v + 1)
            np.add.at(h[:, side, :], (self.records["sample"][ok].astype(np.int64), bins[ok]), 1)
        return h

    def table(self):
        from scarmapper_b200._native import ENTRY_DTYPE
        tabs = orc.aggregate(self.o, self.records, self.reads)
        rows = [(0, 0, v[1], v[0], s, 0) for s, d in enumerate(tabs) for v in d.values()]
        return np.array(rows, dtype=ENTRY_DTYPE) if rows else np.zeros(0, dtype=ENTRY_DTYPE)

    def close(self):
        pass


_HOM_LIB = [None]


def host_homeology_batch(consensus, insertions, clj, crj, tlj, trj, target_idx, targets, regions, device=0):
    """CPU stand-in for engine.homeology_batch: the same header (scar_homeology_core.h) compiled for the host."""
    import ctypes as C
    import subprocess
    import tempfile
    import numpy as np
    if _HOM_LIB[0] is None:
        d = tempfile.mkdtemp(prefix="scar_hom_")
        src = os.path.join(d, "h.cpp")
        csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scarmapper_b200", "csrc")
        open(src, "w").write(r"""
#include "scar_homeology_core.h"
extern "C" void hom_eval(const uint8_t *cons, int n_cons, int clj, int crj, const uint8_t *target, int n_target,
                         const uint8_t *ins, int n_ins, const uint8_t *region, int n_region, int tlj, int trj, int32_t *out)
{
    ScarHomeology h;
    hom_homeology(cons, n_cons, clj, crj, target, n_target, n_ins, tlj, trj, h);
    h.lt_start = h.rt_start = -1;
    if (n_ins >= 5) hom_templates(ins, n_ins, region, n_region, tlj, trj, h);
    h.pad0 = h.pad1 = 0;
    const int32_t *p = reinterpret_cast<const int32_t *>(&h);
    for (int i = 0; i < 16; ++i) out[i] = p[i];
}
""")
        so = os.path.join(d, "h.so")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", csrc, src, "-o", so])
        _HOM_LIB[0] = C.CDLL(so)
    lib = _HOM_LIB[0]

    def items(x):        # list of str, or a HostRagged (what the drop-in passes)
        if hasattr(x, "item"):
            return [x.item(i) for i in range(x.n)]
        return [v.encode("latin-1") if isinstance(v, str) else bytes(v) for v in x]
    consensus, insertions = items(consensus), items(insertions)
    out = np.zeros((len(consensus), 16), dtype=np.int32)
    row = (C.c_int32 * 16)()
    for i, (c, s) in enumerate(zip(consensus, insertions)):
        t, r = targets[int(target_idx[i])].encode(), regions[int(target_idx[i])].encode()
        lib.hom_eval(c, len(c), int(clj[i]), int(crj[i]), t, len(t), s, len(s), r, len(r),
                     int(tlj[i]), int(trj[i]), row)
        out[i] = list(row)
    return out
