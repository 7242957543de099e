"""
launcher.py — run the UNMODIFIED reference pipeline with the CUDA hot path plugged in.

    python -m scarmapper_b200.launcher --options_file run.sh --reference /path/to/ScarMapper [--device 0]

This mirrors scarmapper.py:main for IndelProcessing (scarmapper.py:221-286) — options file ->
argparse namespace -> Logger -> FASTQ_Reader -> TargetMapper -> DataProcessing(...).main_loop() —
with the reference's own modules imported from --reference, and with three deviations that are
unavoidable on Python >= 3.12, each outside the hot path:
  1. scarmapper.py itself is not imported: its import-time hack looks for a cpython-37 build of
     SlidingWindow and shells out to python3.7 (scarmapper.py:28-69), and
     VersionDependencies.python_check rejects Python > 3.11.0 (Version_Dependencies.py:25-30);
  2. `scarmapper.SlidingWindow` resolves to scarmapper_b200.SlidingWindow (same signature);
  3. the FASTQ is read by scarmapper_b200.fastq (FASTQ_Reader's text branch uses open(..., 'rU'),
     invalid on >= 3.11, FASTQ_Tools.py:607); FASTQ_Reader is still constructed for its path checks
     when the file is gzip.
The two hot loops are swapped for their batched bodies by scarmapper_b200.dropin.install().
PEAR merging (scarmapper.py:78-145, an external binary) runs when --FASTQ2 is set.
"""
from __future__ import annotations

import argparse
import datetime
import os
import subprocess
import sys
import time


def _strtobool(v) -> bool:
    return str(v).strip().lower() in ("y", "yes", "t", "true", "on", "1")


def parse_options(tool_box, argv):
    """string_to_boolean + error_checking of scarmapper.py:443-535, without importing scarmapper.py."""
    parser = argparse.ArgumentParser(description="ScarMapper (reference driver) with the B200 hot path")
    parser.add_argument("--options_file", action="store", dest="options_file", required=True)
    known, _ = parser.parse_known_args(argv)
    saved = sys.argv
    sys.argv = [saved[0], "--options_file", known.options_file]
    try:
        options_parser = tool_box.options_file(parser)      # Valkyries/Tool_Box.py:402-431 (reads sys.argv)
        args = options_parser.parse_args()
    finally:
        sys.argv = saved
    args.IndelProcessing = _strtobool(args.IndelProcessing)
    if args.IndelProcessing:
        args.PEAR = True
        args.Demultiplex = _strtobool(getattr(args, "Demultiplex", "False"))
        args.OutputRawData = _strtobool(getattr(args, "OutputRawData", "False"))
        args.DeleteConsensusFASTQ = _strtobool(getattr(args, "DeleteConsensusFASTQ", "True"))
    args.Verbose = args.Verbose.upper()
    args.Search_KMER = int(args.Search_KMER)
    args.Spawn = int(args.Spawn)
    args.N_Limit = float(args.N_Limit)
    args.PatternThreshold = float(args.PatternThreshold)
    args.Minimum_Length = int(args.Minimum_Length)
    # scarmapper.py:457-458 crashes on a non-empty HR_Donor (set_defaults returns None); the intent is upper-casing
    args.HR_Donor = (getattr(args, "HR_Donor", "") or "").upper()
    for key in ("WorkingFolder", "FASTQ1", "RefSeq", "Master_Index_File", "SampleManifest", "TargetFile"):
        v = getattr(args, key, "")
        if not v or not os.path.exists(v):
            print("\033[1;31mERROR:\n\t--{}: {} Not Found.  Check Options File.".format(key, v))
            raise SystemExit(1)
    return args


def run_pear(reference, args, log):
    """PEAR consensus exactly as scarmapper.py:78-145 launches it."""
    prefix = "{}{}".format(args.WorkingFolder, args.Job_Name)
    opt = "-y {} -j {} ".format(args.Memory, args.Spawn)
    for flag, key in (("-n", "MinConsensusLength"), ("-p", "PValue"), ("-v", "MinOverlap"),
                      ("-q", "QualityThreshold"), ("-b", "PhredValue"), ("-g", "TestMethod")):
        v = getattr(args, key, "")
        if v:
            opt += "{} {} ".format(flag, v)
    cmd = "{}{}Pear{}bin{}./pear -f {} -r {} -o {} {}".format(reference, os.sep, os.sep, os.sep, args.FASTQ1,
                                                             args.FASTQ2, prefix, opt)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, shell=True)
    if proc.stderr:
        log.error("{}\n{}\n".format(proc.stderr.decode(), proc.stdout.decode()))
        raise SystemExit(1)
    return "{}.assembled.fastq".format(prefix)


def defer_scipy_stats():
    """`from scipy import stats` (scarmapper/INDEL_Processing.py:16) takes 0.8 s of a start-up of about one second, for
    two numbers of consensus_demultiplex that the batched path computes itself (dropin.norm_interval_lower / sem).
    A stand-in module takes the name; the real scipy.stats is imported the moment anything else is asked of it."""
    import importlib
    import types
    if "scipy.stats" in sys.modules:
        return

    class _Deferred(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__") and name.endswith("__"):
                raise AttributeError(name)
            real = self.__dict__.get("_real")
            if real is None:
                if sys.modules.get("scipy.stats") is self:
                    del sys.modules["scipy.stats"]
                real = importlib.import_module("scipy.stats")
                self.__dict__["_real"] = real
            return getattr(real, name)
    sys.modules["scipy.stats"] = _Deferred("scipy.stats")


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--reference", default=os.environ.get("SCARMAPPER_REFERENCE", ""))
    ap.add_argument("--device", type=int, default=0)
    own, rest = ap.parse_known_args(argv)
    if not own.reference or not os.path.isdir(os.path.join(own.reference, "scarmapper")):
        raise SystemExit("--reference must point at a ScarMapper checkout (the directory holding scarmapper.py)")
    start_time = time.time()
    sys.path.insert(0, own.reference)
    # CUDA context creation (~1 s on a fresh process) runs beside the imports of the reference's modules (scipy alone
    # takes as long); the thread has long finished when the reference forks its pool
    import threading
    from scarmapper_b200 import _native

    def _warm_device():
        try:
            _native.lib().scar_device_memory(own.device, None, None)
        except Exception:       # a missing library / device is reported by the first real call
            pass
    warm = threading.Thread(target=_warm_device, daemon=True)
    warm.start()

    defer_scipy_stats()
    # deviation 2: the reference imports `scarmapper.SlidingWindow` at module import
    import scarmapper                                   # the reference package
    from scarmapper_b200 import SlidingWindow as sw, dropin
    sys.modules["scarmapper.SlidingWindow"] = sw
    scarmapper.SlidingWindow = sw
    from scarmapper import INDEL_Processing, TargetMapper
    from Valkyries import Tool_Box

    args = parse_options(Tool_Box, rest)
    if not args.IndelProcessing:
        raise SystemExit("only --IndelProcessing True runs through the CUDA path (Combine mode is the reference's own)")
    run_start = datetime.datetime.today().strftime("%H:%M:%S %Y  %a %b %d")
    log = Tool_Box.Logger(args)
    log.info("ScarMapper reference driver with scarmapper_b200 hot path")
    if getattr(args, "FASTQ2", ""):
        consensus = run_pear(own.reference, args, log)
    else:
        consensus = args.FASTQ1

    class _Fastq:            # what DataProcessing needs from FASTQ_Reader: the path
        input_file = consensus

    dropin.install(INDEL_Processing, device=own.device)
    warm.join()
    sample_manifest = Tool_Box.FileParser.indices(log, args.SampleManifest)
    processing = INDEL_Processing.DataProcessing(log, args, run_start, INDEL_Processing.__version__,
                                                 TargetMapper.TargetMapper(log, args, sample_manifest), _Fastq(), None)
    processing.main_loop()
    log.info("Run time: {:.1f} s".format(time.time() - start_time))


if __name__ == "__main__":
    main()
