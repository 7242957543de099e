"""
The drop-in under the reference's REAL process model: DataProcessing.main_loop (reference
scarmapper/INDEL_Processing.py:1038-1061) runs consensus_demultiplex in the parent, THEN creates
pathos.multiprocessing.Pool(Spawn) - forked workers, dill pickling - and starmaps ScarSearch over the samples.
Here pathos.multiprocessing.Pool is multiprocess.Pool (the class pathos re-exports) with Spawn > 1, the driver is
started through scarmapper_b200.launcher.main from a real options file, and every Frequency / SNV_Frequency /
Summary file must equal the reference's own golden file line for line, every Raw_Data file by digest.

  CPU  (not gpu): the engine is the oracle-backed test double - checks the pickling boundary and that the workers
                  need nothing but their SampleResult;
  GPU  (-m gpu) : the real ScarEngine: CUDA is initialised in the parent before the pool forks, all GPU work
                  (classification, pattern searches) happens there, the forked workers never touch CUDA.
The reference tree is /root/reference in the build container and its byte-for-byte copy baseline/_ref/ScarMapper
(oracle/build_ref.build_baseline, git-ignored, shipped by gpurun) on the GPU box.
"""
import os
import sys

import pytest

import golden_util as gu
from oracle import ref_shims

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def run_launcher(tmp_path, name, spawn, engine):
    """tests/run_launcher_case.py in a fresh process (the reference's Logger holds sys.stderr, which pytest replaces
    with an unpicklable capture object), then every output file against the golden."""
    import subprocess
    golden = gu.load(name)["golden"]
    wd = str(tmp_path) + os.sep
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_launcher_case.py")
    proc = subprocess.run([sys.executable, script, name, wd, str(spawn), engine], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=900)
    assert proc.returncode == 0, proc.stderr[-4000:]
    assert "launcher ok" in proc.stdout, proc.stdout[-2000:]
    n = gu.compare_output_files(wd, golden, ignore_version=True)
    assert n == len(golden["files"]) + len(golden["file_sha256"]) and n >= 3


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["c2_96_samples", "ragged_hr_donor", "demux_c2_96_samples", "demux_truseq"])
def test_launcher_with_forked_workers_and_dill_pickling(tmp_path, name):
    run_launcher(tmp_path, name, 4, "oracle")


def test_sample_result_pickles_small():
    """What crosses the pool boundary is the sample's own slice: a few hundred bytes per pattern, no run-sized array."""
    import dill
    import numpy as np
    from scarmapper_b200 import _native as N
    from scarmapper_b200.dropin import SampleResult
    res = SampleResult("D501+D701", 1000)
    k = 50
    res.counts = np.ones(k, np.uint32)
    res.first_records = np.zeros(k, N.RECORD_DTYPE)
    res.first_bytes = np.full(150 * k, 65, np.uint8)
    res.first_off = np.arange(k + 1, dtype=np.int64) * 150
    res.pattern_out = np.zeros((k, 16), np.int32)
    res.target_region, res.cutsite, res.counters = "ACGT" * 100, 200, np.zeros(N.SCAR_N_COUNTERS, np.int64)
    blob = dill.dumps(res)
    assert len(blob) < 400 * k + 4096
    back = dill.loads(blob)
    assert back.n_patterns() == k and back.first_read(3) == "A" * 150 and len(back) == 1000


def test_a_registered_sample_result_travels_as_a_token_to_forked_workers():
    """main_loop forks its pool after consensus_demultiplex: a worker already holds every SampleResult of the run, so
    what is pickled is a token (dropin._INHERITED), and a real forked multiprocess.Pool resolves it to the arrays."""
    import dill
    import multiprocess
    import numpy as np
    from scarmapper_b200 import dropin
    res = dropin.SampleResult("D502+D703", 77)
    res.counts = np.arange(5000, dtype=np.uint32)
    res.first_bytes = np.full(150 * 5000, 67, np.uint8)
    res.first_off = np.arange(5001, dtype=np.int64) * 150
    res._token = (os.getpid(), 1, res.index_name)
    dropin._INHERITED[res._token] = res
    try:
        assert len(dill.dumps(res)) < 400
        assert dill.loads(dill.dumps(res)) is res                      # (same process: the registry answers)
        with multiprocess.get_context("fork").Pool(2) as pool:          # created AFTER the registration, as in main_loop
            got = pool.starmap(_probe_result, [[res, 4999], [res, 7]])
        assert got == [(77, 4999, "C" * 150), (77, 7, "C" * 150)]
        dropin.INHERIT_RESULTS[0] = False
        assert len(dill.dumps(res)) > 150 * 5000                      # the arrays themselves
    finally:
        dropin.INHERIT_RESULTS[0] = True
        dropin._INHERITED.clear()
    assert len(dill.dumps(res)) > 150 * 5000                          # not registered any more: pickled whole


def _probe_result(res, k):
    return len(res), int(res.counts[k]), res.first_read(k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.launcher_golden_names())
def test_gpu_launcher_with_forked_workers_writes_the_reference_files(tmp_path, name):
    assert ref_shims.available(), ("no reference tree on this box: baseline/_ref/ScarMapper is made by "
                                   "__graft_entry__.build() in the build container and shipped by gpurun")
    run_launcher(tmp_path, name, 8, "cuda")
