"""
Shared helpers for the golden-vector tests: load a fixture written by tests/golden/make_golden.py
(outputs of the reference itself), rebuild the static problem from its inputs, and compare a set
of per-read records (from the oracle or from the CUDA path) with everything the reference
produced: per-sample scar table in first-seen order, summary_data, read_results_list,
read_count_dict and phase_count.
"""
from __future__ import annotations

import glob
import gzip
import json
import os

import numpy as np

from oracle import scar_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _all_names():
    return sorted(os.path.basename(p)[:-len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))


_VECTOR_FILES = ("alignment_cases", "homeology_cases")       # per-function vectors, not runs of the reference driver


def golden_names():
    """The per-read / per-table fixtures (one reference run each)."""
    return [n for n in _all_names() if n not in _VECTOR_FILES and not n.startswith("demux_")]


def launcher_golden_names():
    """Fixtures whose output FILES are compared through the launcher: the above plus the --Demultiplex True runs."""
    return [n for n in _all_names() if n not in _VECTOR_FILES]


def load(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rt") as f:
        return json.load(f)


def phasing_lists(case):
    """TargetMapper.phasing (TargetMapper.py:30-71): per locus, from the first manifest row naming it."""
    if case["platform"] == "TruSeq":
        return [([], []) for _ in case["targets"]]
    out = []
    for fwd, rev in case["phasing_seqs"]:
        fl, rl = len(fwd), len(rev)
        h, rh = int(fl * 0.5), int(rl * 0.5)
        out.append(([fwd[fl - h - i:fl - i] for i in range(h + 1)],
                    [rev[rl - rh - i:rl - i] for i in range(h + 1)]))
    return out


def problem_inputs(case):
    """dict of the static inputs in the reference's own terms (used for both oracle and CUDA ctx)."""
    targets = [t.upper() for t in case["targets"]]
    cutsites = [orc.cutsite_search(t, sg.upper(), rc) for t, sg, rc in
                zip(targets, case["sgrnas"], case["reverse_comp"])]
    return dict(
        targets=targets, cutsites=cutsites, sample_target=case["sample_target"],
        left_index=[s.upper() for s in case["index_col3"]],    # index_dict[.][0]  (INDEL_Processing.py:1110-1113)
        right_index=[s.upper() for s in case["index_col2"]],   # index_dict[.][2]
        platform=case["platform"], phasing=phasing_lists(case), hr_donor=case["hr_donor"],
        n_limit=case["n_limit"], min_length=case["min_length"])


def index_fields(case):
    if case["platform"] != "Illumina":
        return None, None
    f0, f1 = zip(*(orc.parse_illumina_header(n) for n in case["names"]))
    return list(f0), list(f1)


def compare_with_golden(case, golden, records, table_fn=None):
    """records: np.ndarray[RECORD_DTYPE], one per input read, in input order.
    table_fn(sample_pos) -> list of (count, first_read_global_index) in first-seen order; when None
    the tables are derived from the records with the dict-based aggregation."""
    pin = problem_inputs(case)
    prob = orc.Problem(**pin)
    reads = case["reads"]
    pos = {iid: s for s, iid in enumerate(case["sample_ids"])}
    cnt = orc.counters(prob, records)
    tables = orc.aggregate(prob, records, reads) if table_fn is None else None

    # read_count_dict (INDEL_Processing.py:1173-1198)
    for iid, n in golden["read_count_dict"].items():
        if iid == "unidentified":
            assert n == int((records["status"] == orc.ST_UNASSIGNED).sum())
        else:
            assert n == cnt[pos[iid], 0], ("read_count", iid)
    assert int(cnt[:, 0].sum()) + int((records["status"] == orc.ST_UNASSIGNED).sum()) == golden["read_count"]

    for iid, g in golden["samples"].items():
        s = pos[iid]
        t = case["sample_target"][s]
        target, cut = pin["targets"][t], pin["cutsites"][t]
        assert g["cutsite"] == cut
        c = cnt[s]
        gs = g["summary"]
        mine = [int(c[1]), int(c[2]), int(c[3]), int(c[4]), int(c[5]), [int(c[6]), int(c[7])],
                [int(c[8]), 0], [int(c[9]), int(c[10])]]
        assert mine == gs, ("summary", iid, mine, gs)

        sel = np.nonzero((records["sample"] == s) & (records["status"] == orc.ST_SCAR))[0]
        rr = [[int(records["clj"][i]), int(records["crj"][i]), int(records["tlj"][i]), int(records["trj"][i]),
               "HR" if records["flags"][i] & orc.FL_HR else ""] for i in sel]
        assert rr == g.get("read_results", []), ("read_results", iid)

        if table_fn is None:
            mine_tab = [(k, v[0], v[1]) for k, v in tables[s].items()]
        else:
            mine_tab = []
            for count, first in table_fn(s):
                rec = records[first]
                seq = orc.platform_seq(prob.platform, reads[first].encode("latin-1")).decode("latin-1")
                sub = orc.record_to_sublist(rec, seq, target, cut)
                mine_tab.append(("{}|{}|{}|{}|{}".format(sub[0], sub[1], sub[2], sub[3], sub[9]), count, first))
        gt = g.get("table", [])
        assert len(mine_tab) == len(gt), ("table size", iid, len(mine_tab), len(gt))
        for (key, count, first), row in zip(mine_tab, gt):
            gkey, gcount, ldel, rdel, ins, mh, clj, crj, tlj, trj, hr, consensus = row
            assert key == gkey and count == gcount, ("table row", iid, key, gkey, count, gcount)
            rec = records[first]
            seq = orc.platform_seq(prob.platform, reads[first].encode("latin-1")).decode("latin-1")
            assert orc.record_to_sublist(rec, seq, target, cut) == \
                [ldel, rdel, ins, mh, consensus, clj, crj, tlj, trj, hr], ("first sub_list", iid, key)

    # phase_count (INDEL_Processing.py:948-975): label histograms per 'index+locus'
    if case["platform"] == "Illumina":
        for key, g in golden["phase_count"].items():
            iid = key.rsplit("+", 1)[0] if key.count("+") == 1 else key[:key.rfind("+")]
            # index ids themselves contain '+': the locus is the LAST '+'-separated token
            s = pos[iid]
            sel = (records["sample"] == s) & (records["status"] != orc.ST_UNASSIGNED)
            r1, r2 = prob.phasing[case["sample_target"][s]]
            for lab, n in g.items():
                if n == -1:
                    continue
                if lab == "No Read 1 Phasing":
                    m = int((records["phase_r1"][sel] == -1).sum())
                elif lab == "No Read 2 Phasing":
                    m = int((records["phase_r2"][sel] == -1).sum())
                else:
                    kind, k = lab.split()[1][0], int(lab.split()[1][1:])
                    col = "phase_r1" if kind == "F" else "phase_r2"
                    m = int((records[col][sel] == k).sum())
                assert m == n, ("phase", key, lab, m, n)
    return cnt


WALL_CLOCK = ("# Run Start", "# Run End", "Start:", "End:", "FASTQ1:", "FASTQ2:")
VERSION = ("ScarMapper v", "# ScarMapper Search v")     # the goldens were written with version="golden"


def compare_output_files(wd, golden, ignore_version=False, expect_raw=True):
    """Every Frequency / SNV_Frequency / Summary file the reference wrote for this case equals the file in `wd`
    line for line (wall-clock lines excluded), no such file is extra, and every Raw_Data file has the digest the
    reference's own file had.  Returns the number of files compared."""
    import hashlib
    skip = WALL_CLOCK + (VERSION if ignore_version else ())
    seen = 0
    for fn, want in golden["files"].items():
        got = open(os.path.join(wd, fn)).read().split("\n")
        got = "\n".join(ln for ln in got if not ln.startswith(skip))
        if ignore_version:
            want = "\n".join(ln for ln in want.split("\n") if not ln.startswith(VERSION))
        assert got == want, fn
        seen += 1
    mine = sorted(fn for fn in os.listdir(wd)
                  if fn.endswith(("_ScarMapper_Frequency.txt", "_ScarMapper_Summary.txt", "_SNV_Frequency.txt")))
    assert mine == sorted(golden["files"]), (set(mine) ^ set(golden["files"]))
    if expect_raw:
        for fn, want in golden.get("file_sha256", {}).items():
            data = open(os.path.join(wd, fn), "rb").read()
            if fn.endswith(".fastq"):                  # demultiplexed FASTQ (--Demultiplex True): byte for byte
                assert hashlib.sha256(data).hexdigest() == want, fn
            else:
                keep = [ln for ln in data.split(b"\n") if not ln.startswith((b"# Run Start", b"# Run End", b"# ScarMapper Search v"))]
                assert hashlib.sha256(b"\n".join(keep)).hexdigest() == want, fn
            seen += 1
        mine = sorted(fn for fn in os.listdir(wd) if fn.endswith(("_ScarMapper_Raw_Data.txt", "_R1.fastq")))
        assert mine == sorted(golden.get("file_sha256", {})), "Raw_Data / demultiplexed FASTQ file set differs"
    return seen


def write_options_file(wd, case, spawn=1, verbose="ERROR", raw=True):
    """A tab-delimited options file in the format of docs/run_ScarMapper_IndelProcessing.sh for the inputs that
    make_golden.write_inputs put into `wd`."""
    opts = {
        "IndelProcessing": "True", "FASTQ1": wd + "reads.fastq.gz", "FASTQ2": "", "RefSeq": wd + "refseq.fa",
        "Master_Index_File": wd + "master_index.tsv", "SampleManifest": wd + "manifest.tsv",
        "TargetFile": wd + "targets.tsv", "WorkingFolder": wd, "Verbose": verbose, "Job_Name": "golden",
        "Spawn": str(spawn), "Demultiplex": "True" if case.get("demultiplex") else "False", "DeleteConsensusFASTQ": "True",
        "HR_Donor": case["hr_donor"],
        "Platform": case["platform"], "Search_KMER": "10", "N_Limit": str(case["n_limit"]),
        "Minimum_Length": "{} # Length after trimming".format(case["min_length"]),
        "OutputRawData": "True" if raw else "False", "PatternThreshold": "0.0001 # cutoff", "FigureType": "pdf",
    }
    path = wd + "run.sh"
    with open(path, "w") as f:
        f.write("#!/bin/bash\n#Parameter file\n\npython3 scarmapper.py --options_file run.sh\nexit\n\n")
        for k, v in opts.items():
            f.write("--{}\t{}\n".format(k, v))
    return path
