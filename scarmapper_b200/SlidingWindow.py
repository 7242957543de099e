"""
SlidingWindow.py — drop-in for the reference's compiled Cython module scarmapper/SlidingWindow.pyx.

Same name, same `sliding_window` signature, same return value and the same in-place mutation of
summary_data (SlidingWindow.pyx:20-22,59,97,109-112,145-171), computed by the CUDA path through
the C ABI: a one-target, one-sample context per (target_region, cutsite, hr_donor, cutwindow) is
cached and the read goes through pack -> K3.  It exists so that code written against the per-read
interface (and the parity tests) keeps working; throughput comes from the batched entry points.
There is no CPU fallback: without the CUDA library / a GPU this raises.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .engine import ScarEngine, ScarProblem

__version__ = "0.5.0-b200"

_cache = {}
_MAX_CACHE = 8


def _engine_for(target_region, cutsite, cutwindow, hr_donor, max_len, left_target_windows, right_target_windows):
    key = (target_region, cutsite, cutwindow if len(cutwindow) == 10 else "", hr_donor)
    eng = _cache.get(key)
    if eng is not None and eng.prob.max_read_len >= max_len:
        return eng
    # the window lists must be the ones window_mapping builds (INDEL_Processing.py:57-80): the CUDA
    # tables are derived from (target_region, cutsite)
    L = len(target_region)
    left = [target_region[cutsite - 10 - i:cutsite - i] for i in range(max(0, cutsite - 15))]
    right = [target_region[cutsite + i:cutsite + i + 10] for i in range(max(0, L - 15 - cutsite))]
    if list(left_target_windows) != left or list(right_target_windows) != right:
        raise ValueError("left/right_target_windows differ from window_mapping(target_region, cutsite); "
                         "non-standard window lists are not supported by the CUDA path")
    if eng is not None:
        eng.close()
    if len(_cache) >= _MAX_CACHE:
        _cache.pop(next(iter(_cache))).close()
    prob = ScarProblem(targets=[target_region], cutsites=[cutsite], sample_target=[0], left_index=[""],
                       right_index=[""], platform="TruSeq",      # TruSeq + empty indices: no header fields needed
                       mismatch=1 << 20, hr_donor=hr_donor, n_limit=2.0, min_length=-1,
                       max_read_len=max(160, ((max_len + 12 + 63) // 64) * 64), table_capacity=1 << 10,
                       cutwindows=[cutwindow] if len(cutwindow) == 10 else None)
    eng = ScarEngine(prob)
    _cache[key] = eng
    return eng


def sliding_window(consensus, target_region, cutsite, target_length, lower_limit, upper_limit, summary_data,
                   left_target_windows, right_target_windows, cutwindow, hr_donor):
    if lower_limit != 15:
        raise ValueError("lower_limit {} is not supported (the driver always passes 15, INDEL_Processing.py:65)"
                         .format(lower_limit))
    eng = _engine_for(target_region, cutsite, cutwindow, hr_donor, len(consensus), left_target_windows,
                      right_target_windows)
    # TruSeq stores seq[6:-6] (INDEL_Processing.py:1181-1182): pad so that the stored sequence is `consensus`
    raw = "A" * 6 + consensus + "A" * 6
    eng.reset()
    rec = eng.process_host([raw])[0]
    st, fl, n = int(rec["status"]), int(rec["flags"]), int(rec["hr_n"])
    if n:                                        # :108-112 (counted before the N-in-insertion exit)
        summary_data[10][0] += 1
        summary_data[10][1] += n - 1
    if st == N.ST_NO_JUNCTION or st == N.ST_FILTERED:   # FILTERED here can only mean a zero-length read
        summary_data[6][0] += 1
        return [], summary_data
    if st == N.ST_NO_CUT:
        summary_data[6][1] += 1
        return [], summary_data
    if st != N.ST_SCAR:                          # N in the insertion (:140-141), or a zero-length read
        return [], summary_data
    summary_data[4] += bool(fl & N.FL_INSERTION)
    summary_data[2] += bool(fl & N.FL_LEFT_DEL)
    summary_data[3] += bool(fl & N.FL_RIGHT_DEL)
    summary_data[5] += bool(fl & N.FL_MICROHOM)
    clj, crj, tlj, trj = int(rec["clj"]), int(rec["crj"]), int(rec["tlj"]), int(rec["trj"])
    ins = consensus[clj:crj] if fl & N.FL_INSERTION else ""
    mh = consensus[crj:clj] if fl & N.FL_MICROHOM else ""
    return [target_region[tlj:cutsite], target_region[cutsite:trj], ins, mh, consensus, clj, crj, tlj, trj,
            "HR" if fl & N.FL_HR else ""], summary_data
