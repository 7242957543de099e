"""
cpu_reference.py — the reference's CPU path, timed for bench.py's cpu_baseline / --impl reference.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by scarmapper_b200/).

What runs:
  stage 1 (serial, main process, like consensus_demultiplex, INDEL_Processing.py:891-1024):
      a Python loop over reads; per read a loop over manifest rows in order, two match_maker calls
      per row with the header re-split each time exactly as index_matching does (:1149-1186), the
      first accepted row wins.  Levenshtein is the oracle's C function through ctypes
      (python-Levenshtein is a C extension too; it is absent here, parity unpinned).
  stages 2-4 (one process per sample, like the pathos Pool, :1047-1059): per read the N / length
      filters (:147-154), then the REFERENCE'S OWN compiled SlidingWindow.sliding_window
      (oracle/_ref/SlidingWindow.*.so, built from scarmapper/SlidingWindow.pyx by oracle/build_ref.py)
      when it is present -> kind "reference"; otherwise the oracle's C restatement -> kind "port".
      Keys are aggregated in a dict as :186-196.

The loops are restated from the reference's control flow, not copied; the reporting code
(frequency_output etc.) is not part of the timed path.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

from oracle import build_ref
from oracle import scar_oracle as orc


def baseline_kind() -> str:
    return "reference" if build_ref.ref_so() else "port"


def demultiplex(names, seqs, sample_ids, left_index, right_index, platform, mismatch):
    """Serial stage 1: returns {sample position: [stored sequences]} and the unidentified count."""
    match_maker = orc.match_maker
    sequence_dict = {}
    unidentified = 0
    n_samples = len(sample_ids)
    for name, seq in zip(names, seqs):
        found = -1
        for s in range(n_samples):
            if platform == 0:
                right_match = match_maker(right_index[s], name.split(":")[-1].split("+")[0])
                left_match = match_maker(left_index[s], name.split(":")[-1].split("+")[1])
            elif platform == 1:
                left_match = match_maker(right_index[s], seq[:6])
                right_match = match_maker(left_index[s], seq[-6:])
            else:
                left_match = match_maker(left_index[s], seq[-len(left_index[s]):])
                right_match = match_maker(right_index[s], seq[:len(right_index[s])])
            if left_match <= mismatch and right_match <= mismatch:
                found = s
                break
        if found < 0:
            unidentified += 1
            continue
        if platform == 1:
            stored = seq[6:-6]
        elif platform == 2:
            stored = orc.rcomp(seq)
        else:
            stored = seq
        sequence_dict.setdefault(found, []).append(stored)
    return sequence_dict, unidentified


def scar_search(sample, sequence_list, target, cutsite, n_limit, min_length, hr_donor):
    """Stages 2-4 for one sample library (ScarSearch.data_processing's loop)."""
    sw = build_ref.load()
    summary = [sample, 0, 0, 0, 0, 0, [0, 0], [0, 0], "junction data", "target", [0, 0]]
    freq = {}
    n_scar = 0
    if sw is not None:
        left, right = orc.window_mapping(target, cutsite)
        cutwindow = target[cutsite - 4:cutsite + 4]
        L = len(target)
        for seq in sequence_list:
            if seq.count("N") / len(seq) > n_limit:
                summary[7][0] += 1
                continue
            if len(seq) <= min_length:
                summary[7][0] += 1
                continue
            summary[1] += 1
            sub, summary = sw.sliding_window(seq, target, cutsite, L, 15, L - 15, summary, left, right,
                                             cutwindow, hr_donor)
            if not sub:
                continue
            n_scar += 1
            key = "{}|{}|{}|{}|{}".format(sub[0], sub[1], sub[2], sub[3], sub[9])
            if key in freq:
                freq[key][0] += 1
            else:
                freq[key] = [1, sub]
    else:
        prob = orc.Problem([target], [cutsite], [0], n_limit=n_limit, min_length=min_length, hr_donor=hr_donor)
        rec = orc.classify_batch(prob, sequence_list, sample_in=[0] * len(sequence_list))
        tab = orc.aggregate(prob, rec, sequence_list)[0]
        freq = {k: v for k, v in tab.items()}
        n_scar = int((rec["status"] == orc.ST_SCAR).sum())
        summary[1] = int((rec["status"] >= orc.ST_NO_JUNCTION).sum())
    return sample, summary[1], n_scar, len(freq)


def _star(args):
    return scar_search(*args)


class CpuReference:
    """Holds the worker pool so that repeated timed steps do not pay process start-up."""

    def __init__(self, cores: int | None = None):
        self.cores = cores or max(1, (os.cpu_count() or 2) - 1)   # Spawn <= CPUs - 1 (run_ScarMapper_IndelProcessing.sh:21)
        self.pool = mp.get_context("spawn").Pool(self.cores) if self.cores > 1 else None
        if self.pool:
            self.pool.map(_warm, range(self.cores))

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def run(self, names, seqs, sample_ids, left_index, right_index, sample_target, targets, cutsites,
            platform=0, mismatch=1, n_limit=0.01, min_length=100, hr_donor=""):
        """One pass over the sample; returns timings and a few result totals."""
        t0 = time.perf_counter()
        sequence_dict, unidentified = demultiplex(names, seqs, sample_ids, left_index, right_index, platform, mismatch)
        t1 = time.perf_counter()
        jobs = [(s, sequence_dict[s], targets[sample_target[s]], int(cutsites[sample_target[s]]), n_limit,
                 min_length, hr_donor)
                for s in sorted(sequence_dict, key=lambda k: len(sequence_dict[k]), reverse=True)]
        if self.pool:
            res = self.pool.map(_star, jobs, chunksize=1)
        else:
            res = [scar_search(*j) for j in jobs]
        t2 = time.perf_counter()
        return {"n_reads": len(seqs), "demux_s": t1 - t0, "search_s": t2 - t1, "total_s": t2 - t0,
                "unidentified": unidentified, "passing": sum(r[1] for r in res), "scar": sum(r[2] for r in res),
                "unique_keys": sum(r[3] for r in res)}


def _warm(_):
    build_ref.load()
    orc.lib()
    return 0
