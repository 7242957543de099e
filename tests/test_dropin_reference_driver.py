"""
The drop-in glue end to end, in the build container: the UNMODIFIED reference driver
(DataProcessing.main_loop, frequency_output, raw_data_output, data_output) runs with
scarmapper_b200.dropin installed, and the files it writes must equal the golden files the reference
wrote on its own (tests/golden/*.json.gz "files": Frequency and Summary outputs minus wall-clock lines).
The engine is the oracle-backed test double (no GPU here); tests/test_gpu_parity.py proves
engine == oracle on the GPU box, so the two together cover FASTQ -> files.
Skipped where /root/reference is absent.
"""
import argparse
import os
import shutil
import sys
import tempfile

import pytest

import golden_util as gu
from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.mark.parametrize("native_frequency", [True, False], ids=["native_frequency", "reference_frequency"])
@pytest.mark.parametrize("name", ["c1_single_target", "c3_12_targets_mh", "ragged_hr_donor", "truseq", "ramsden"])
def test_reference_driver_with_dropin_writes_identical_files(name, native_frequency):
    """native_frequency=True: the Frequency / SNV files come from scar_frequency_rows_host (the product default);
    False: the reference's own frequency_output runs, its two searches served from the batched results."""
    import make_golden
    from oracle_engine import OracleEngine, host_homeology_batch
    blob = gu.load(name)
    case, golden = blob["case"], blob["golden"]
    ref_shims.install()
    from scarmapper import INDEL_Processing, TargetMapper
    from Valkyries import Tool_Box
    from scarmapper_b200 import dropin

    saved = (INDEL_Processing.DataProcessing.consensus_demultiplex, INDEL_Processing.ScarSearch.data_processing)
    wd = tempfile.mkdtemp(prefix="scar_dropin_") + os.sep
    try:
        make_golden.write_inputs(wd, case)
        args = argparse.Namespace(
            WorkingFolder=wd, Job_Name="golden", Verbose="ERROR", Platform=case["platform"], PEAR=True,
            Demultiplex=False, HR_Donor=case["hr_donor"], N_Limit=case["n_limit"],
            Minimum_Length=case["min_length"], OutputRawData=True, PatternThreshold=0.0001,
            RefSeq=wd + "refseq.fa", TargetFile=wd + "targets.tsv", SampleManifest=wd + "manifest.tsv",
            Master_Index_File=wd + "master_index.tsv", Spawn=1, FASTQ1=wd + "reads.fastq.gz", FASTQ2=None,
            Search_KMER=10, FigureType="pdf", IndelProcessing=True)
        log = Tool_Box.Logger(args)
        dropin._ENGINE_FACTORY[0] = OracleEngine
        dropin._HOMEOLOGY_BATCH[0] = host_homeology_batch
        dropin.NATIVE_FREQUENCY[0] = native_frequency
        dropin.install(INDEL_Processing)

        class _Fastq:
            input_file = args.FASTQ1

        manifest = Tool_Box.FileParser.indices(log, args.SampleManifest)
        dp = INDEL_Processing.DataProcessing(log, args, "run-start", "golden",
                                             TargetMapper.TargetMapper(log, args, manifest), _Fastq(), None)
        dropin.PATTERN_SEARCH_STATS.update(served=0, missed=0)
        written = {}

        def hook(index_name, rows, records, reads, res):     # the parent wrote this sample's Raw_Data file
            written[index_name] = (res, [dropin.record_to_sublist(
                records[r], dropin.stored_sequence(case["platform"], reads.item(int(r))), res.target_region, res.cutsite)
                for r in rows])
        dropin._RAW_DATA_HOOK[0] = hook
        try:
            dp.main_loop()
        finally:
            dropin._RAW_DATA_HOOK[0] = None
        # every homeology / templated-insertion search of frequency_output was answered from the batch
        assert dropin.PATTERN_SEARCH_STATS["missed"] == 0, dropin.PATTERN_SEARCH_STATS
        assert (dropin.PATTERN_SEARCH_STATS["served"] > 0) == (not native_frequency)
        assert dp.read_count == golden["read_count"]
        assert dict(dp.read_count_dict) == golden["read_count_dict"]
        assert {k: dict(v) for k, v in dp.phase_count.items()} == golden["phase_count"]
        # Frequency / SNV / Summary files line for line, Raw_Data files by the digest of the reference's own file
        assert gu.compare_output_files(wd, golden) >= 3
        # Raw_Data files: written natively from the records (scar_raw_rows_host); the rows must equal what the
        # reference's own raw_data_output writes from the per-read lists
        import types
        ref_dir = tempfile.mkdtemp(prefix="scar_raw_ref_") + "/"
        checked = 0
        for iid, (res, read_results) in written.items():
            stub = types.SimpleNamespace(args=argparse.Namespace(WorkingFolder=ref_dir, Job_Name="golden"),
                                         target_region=res.target_region, index_dict=dp.index_dict,
                                         target_dict=dp.target_dict, common_page_header=lambda idx: "")
            INDEL_Processing.ScarSearch.raw_data_output(stub, iid, read_results)
            fn = "golden_{}_ScarMapper_Raw_Data.txt".format(iid)
            want_rows = open(ref_dir + fn).read().split(dropin.RAW_DATA_COLUMNS, 1)[1]
            got_rows = open(wd + fn).read().split(dropin.RAW_DATA_COLUMNS, 1)[1]
            assert got_rows == want_rows, iid
            checked += bool(want_rows)
        shutil.rmtree(ref_dir, ignore_errors=True)
        assert checked >= 1 and len(written) == len(golden["samples"])
    finally:
        dropin._ENGINE_FACTORY[0] = None
        dropin._HOMEOLOGY_BATCH[0] = None
        dropin.NATIVE_FREQUENCY[0] = True
        INDEL_Processing.DataProcessing.consensus_demultiplex, INDEL_Processing.ScarSearch.data_processing = saved
        shutil.rmtree(wd, ignore_errors=True)


def test_launcher_runs_the_reference_pipeline_from_an_options_file(tmp_path):
    """scarmapper_b200.launcher.main with a real tab-delimited options file (docs/run_ScarMapper_IndelProcessing.sh
    format), the reference tree, and the oracle-backed engine: the files written equal the golden files."""
    import make_golden
    from oracle_engine import OracleEngine, host_homeology_batch
    blob = gu.load("c2_96_samples")
    case, golden = blob["case"], blob["golden"]
    ref_shims.install()
    from scarmapper import INDEL_Processing
    from scarmapper_b200 import dropin, launcher
    saved = (INDEL_Processing.DataProcessing.consensus_demultiplex, INDEL_Processing.ScarSearch.data_processing)
    wd = str(tmp_path) + os.sep
    make_golden.write_inputs(wd, case)
    gu.write_options_file(wd, case)
    try:
        dropin._ENGINE_FACTORY[0] = OracleEngine
        dropin._HOMEOLOGY_BATCH[0] = host_homeology_batch
        launcher.main(["--options_file", wd + "run.sh", "--reference", ref_shims.REFERENCE])
        assert gu.compare_output_files(wd, golden, ignore_version=True) == len(golden["files"]) + 96
    finally:
        dropin._ENGINE_FACTORY[0] = None
        dropin._HOMEOLOGY_BATCH[0] = None
        INDEL_Processing.DataProcessing.consensus_demultiplex, INDEL_Processing.ScarSearch.data_processing = saved


def test_lower_limit_arithmetic_is_scipys_bit_for_bit():
    """consensus_demultiplex's plot threshold (scarmapper/INDEL_Processing.py:1014-1017): dropin.norm_interval_lower / sem
    against scipy.stats.norm.interval / scipy.stats.sem on per-sample scar-pattern counts, compared as hex floats."""
    stats = pytest.importorskip("scipy.stats")
    import statistics

    import numpy as np
    from scarmapper_b200 import dropin
    rng = np.random.default_rng(12)
    cases = [[1, 2], [5, 5, 5], [0, 10 ** 7], list(range(96))]
    for _ in range(3000):
        n = int(rng.integers(2, 200))
        hi = int(rng.choice([3, 50, 10 ** 4, 10 ** 7]))
        cases.append([int(v) for v in rng.integers(0, hi, n)])
    import warnings
    for key_counts in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)        # scipy's own: scale == 0 when every count is equal
            want_lower, _ = stats.norm.interval(0.9, loc=statistics.mean(key_counts), scale=stats.sem(key_counts))
        assert float(dropin.sem(key_counts)).hex() == float(stats.sem(key_counts)).hex()
        got = dropin.norm_interval_lower(statistics.mean(key_counts), dropin.sem(key_counts))
        assert float(got).hex() == float(want_lower).hex(), key_counts[:5]


def test_launcher_defers_scipy_stats_until_something_needs_it():
    import subprocess
    code = ("import sys; sys.path.insert(0, {root!r})\n"
            "from scarmapper_b200 import launcher\n"
            "launcher.defer_scipy_stats()\n"
            "from scipy import stats\n"
            "assert 'scipy.stats._stats_py' not in sys.modules\n"
            "assert abs(stats.sem([1, 2, 4]) - 0.8819171036881969) < 1e-15\n"
            "assert 'scipy.stats._stats_py' in sys.modules and sys.modules['scipy.stats'].__name__ == 'scipy.stats'\n"
            "import scipy.stats as again\n"
            "assert again.norm.interval(0.9)[1] > 1.6\n"
            "print('ok')\n").format(root=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]
