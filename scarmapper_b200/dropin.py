"""
dropin.py — batched replacements for the bodies of the reference's two hot loops.

install(INDEL_Processing) swaps two methods of the (otherwise unmodified) reference module:

  DataProcessing.consensus_demultiplex   (scarmapper/INDEL_Processing.py:891-1024, HOT LOOP A)
      reference: a serial Python loop, per read a loop over the manifest with two Levenshtein calls.
      here: the FASTQ text is indexed once, copied to the GPU, and ONE call (scar_process_host) runs
      demultiplex -> pack -> sliding window -> aggregation for every read of the run.  Everything that
      needs the GPU or the whole run happens HERE, in the parent process, before main_loop forks its pool
      (CUDA does not survive a fork): the homeology / templated-insertion searches of every scar pattern of
      every sample (one scar_homeology_batch call) and, with --OutputRawData, the Raw_Data files (one pass
      that buckets the scarred reads by sample, then scar_raw_rows_indexed_host per file).
  ScarSearch.data_processing             (:82-207, HOT LOOP B, one pathos process per sample)
      reference: per read filters + SlidingWindow.sliding_window + dict aggregation.
      here: nothing is computed and no CUDA call is made; the worker unpacks ITS sample's slice of the GPU
      results (counters, pattern counts, the first read of every pattern, the pattern-search results) into
      exactly the objects the reference's reporting code consumes and runs the UNCHANGED frequency_output.

What crosses the pool's pickling boundary is one SampleResult per sample, holding that sample's data only:
O(patterns of the sample), never the run's records or FASTQ text.

Everything else — options file, Logger, TargetMapper, main_loop, the pathos Pool, data_output,
frequency_output, plots — stays the reference's own code.
"""
from __future__ import annotations

import collections
import math
import os
import statistics
import sys
import time
import types

import numpy as np

from . import _native as N
from . import fastq
from .engine import HostRagged, ScarEngine, ScarProblem, cutsite_search

_ENGINE_FACTORY = [None]      # tests substitute a CPU stand-in here; None -> ScarEngine
_HOMEOLOGY_BATCH = [None]     # likewise for the per-pattern searches; None -> engine.homeology_batch (GPU)
_RAW_DATA_HOOK = [None]       # tests: called as hook(index_name, rows, records, reads, result) per Raw_Data file
PATTERN_SEARCH_STATS = {"served": 0, "missed": 0}   # calls answered from the batch / handed to the reference's method
FILE_INGEST = [True]          # the FASTQ file is streamed through the GPU (scar_process_fastq_file); False: read + indexed on the host
FILE_INGEST_THREADS = [0]     # threads of the producer (0: min(8, cores))
NATIVE_FREQUENCY = [True]     # False: the reference's own frequency_output runs over results_freq_dict (cross-check)
STAGE_SECONDS = {}           # wall time of the stages of the last batched_consensus_demultiplex (parent process)
READ_LIMIT_DEBUG = 1000002    # --Verbose DEBUG: the reference's loop stops after 1 000 002 reads (:906-914, see below)


def _make_engine(prob: ScarProblem, device: int):
    f = _ENGINE_FACTORY[0]
    return f(prob, device) if f else ScarEngine(prob, device)


INHERIT_RESULTS = [True]      # pool workers take their SampleResult from the memory they were forked with (below)
_INHERITED = {}               # token -> SampleResult of the current run, filled in the parent BEFORE main_loop forks its pool
_RUN_SERIAL = [0]


def _inherited_result(token, index_name):
    """Unpickling side of SampleResult.__reduce_ex__: in a worker forked after batched_consensus_demultiplex the
    registry still holds the parent's objects (copy-on-write pages, nothing travels through the pipe)."""
    res = _INHERITED.get(token)
    if res is None:
        raise RuntimeError("the SampleResult of {} was not inherited by this process (a spawn-type pool?): set "
                           "scarmapper_b200.dropin.INHERIT_RESULTS[0] = False to send it through the pipe".format(index_name))
    return res


class SampleResult:
    """What the GPU pass produced for ONE sample library; stands in for sequence_dict[index]
    (main_loop only takes its len() and hands it to ScarSearch; INDEL_Processing.py:1051-1053).
    Plain numpy arrays of this sample's own data.  main_loop creates its pool AFTER consensus_demultiplex (:1044-1047),
    so a forked worker already holds every SampleResult: pickling one sends a token, not the arrays (dill-pickling
    the ~250 MB of a 10 M-read run cost 0.5 s in the parent alone); INHERIT_RESULTS[0] = False pickles the arrays."""
    _token = None

    def __reduce_ex__(self, protocol):
        if self._token is not None and INHERIT_RESULTS[0] and _INHERITED.get(self._token) is self:
            return (_inherited_result, (self._token, self.index_name))
        return object.__reduce_ex__(self, protocol)

    def __init__(self, index_name, n_reads):
        self.index_name = index_name
        self.n_reads = n_reads
        self.platform = "Illumina"
        self.target_region = None
        self.cutsite = None
        self.reverse_comp = False
        self.counters = None          # int64[SCAR_N_COUNTERS]
        self.counts = np.zeros(0, np.uint32)            # per pattern, first-seen order
        self.first_records = np.zeros(0, N.RECORD_DTYPE)  # record of the first read of each pattern
        self.first_bytes = np.zeros(0, np.uint8)        # stored sequences of those reads, CSR
        self.first_off = np.zeros(1, np.int64)
        self.pattern_out = None       # int32[n_patterns, 16] (struct ScarHomeology) or None
        self.raw_written = False      # the Raw_Data file was written in the parent
        self.read_results = None      # per-read lists for the reference's raw_data_output (only if somebody sets it)

    def __len__(self):
        return self.n_reads

    def n_patterns(self) -> int:
        return int(self.counts.size)

    def first_read(self, k: int) -> str:
        return self.first_bytes[int(self.first_off[k]):int(self.first_off[k + 1])].tobytes().decode("latin-1")

    def table(self):
        """[(count, sub_list of the first read)] in first-seen order (results_freq_dict's values)."""
        return [(int(self.counts[k]), record_to_sublist(self.first_records[k], self.first_read(k), self.target_region,
                                                        self.cutsite)) for k in range(self.n_patterns())]


def record_to_sublist(rec, read, target, cutsite):
    """The 10-field list SlidingWindow.sliding_window returns (SlidingWindow.pyx:171)."""
    clj, crj, tlj, trj = int(rec["clj"]), int(rec["crj"]), int(rec["tlj"]), int(rec["trj"])
    ins = read[clj:crj] if rec["flags"] & N.FL_INSERTION else ""
    mh = read[crj:clj] if rec["flags"] & N.FL_MICROHOM else ""
    return [target[tlj:cutsite], target[cutsite:trj], ins, mh, read, clj, crj, tlj, trj,
            "HR" if rec["flags"] & N.FL_HR else ""]


_COMP = bytes.maketrans(b"ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", b"TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn")


def stored_sequence(platform: str, seq: bytes) -> str:
    """Sequence handed to ScarSearch per platform (INDEL_Processing.py:946,978,981)."""
    if platform == "TruSeq":
        seq = seq[6:-6]
    elif platform == "Ramsden":
        seq = seq.translate(_COMP)[::-1]
    return seq.decode("latin-1")


def fetch_target_region(module, args, target_dict, target_name, log):
    """Target region exactly as ScarSearch.data_processing fetches it (:101-126)."""
    refseq = module.pysam.FastaFile(args.RefSeq)
    start, stop, chrm = target_dict[target_name][2], target_dict[target_name][3], target_dict[target_name][1]
    try:
        refseq.fetch(chrm, start, stop)
    except KeyError:
        chrm = "chr{}".format(chrm)
    try:
        return str(refseq.fetch(chrm, start, stop)).upper()
    except KeyError:
        return str(module.pyfaidx.Fasta(args.RefSeq)[0]).upper()


# scipy.stats.norm.interval(0.9, loc, scale)[0] and scipy.stats.sem(values) as consensus_demultiplex uses them
# (scarmapper/INDEL_Processing.py:1016-1017), without importing scipy.stats (0.8 s of a ~1 s start-up; the launcher
# defers that import).  interval(): ppf((1 - 0.9) / 2) = ndtri(0.04999999999999999) * scale + loc, the constant below;
# sem(): std(ddof=1) / n ** 0.5 over the values as an int64 array.  Bit-identical to scipy 1.18 on random inputs
# (tests/test_dropin_reference_driver.py); the value is only compared with read counts (:540), never printed.
_NDTRI_LOWER_90 = -1.6448536269514729


def norm_interval_lower(loc, scale):
    if not (scale > 0) or loc != loc:          # rv_continuous.ppf: NaN unless scale > 0 (every sample with the same count)
        return np.float64("nan")
    return np.float64(_NDTRI_LOWER_90) * np.float64(scale) + np.float64(loc)


def sem(values):
    a = np.asarray(values)
    return np.std(a, ddof=1) / a.shape[0] ** 0.5


def _rcomp(s: str) -> str:
    return s.encode("latin-1").translate(_COMP)[::-1].decode("latin-1")


def pattern_search_batch(records, reads, platform, first, sample_of, sample_target, targets, target_rc, device, gathered=None,
                         gathered_flipped=None):
    """homeology_search + templated_insertion_search (INDEL_Processing.py:550-700) for every scar pattern of the run
    in ONE batched call.  Pattern k is represented by its first read `first[k]` (record + stored sequence); the
    arguments are derived per pattern exactly as frequency_output derives them (:299-328, incl. the swaps and
    reverse complements for 3'-strand guides), vectorised.  Returns int32[n, 16] (struct ScarHomeology)."""
    from . import engine as _engine
    t0 = time.perf_counter()
    n = int(first.size)
    if n == 0:
        return np.zeros((0, 16), np.int32)
    tix = np.asarray(sample_target, dtype=np.int32)[sample_of]
    rc = np.asarray(target_rc, dtype=bool)[tix]
    if gathered is not None and not rc.any():
        cons, off = gathered                # the first reads as already gathered by the caller (no 3'-strand guide: no flips)
    elif gathered_flipped is not None:
        cons, off = gathered_flipped        # ... gathered with the rows of 3'-strand guides reverse-complemented
    else:
        cons, off = _engine.stored_gather(reads, first, platform, flip=rc.astype(np.uint8) if rc.any() else None)
    L = np.diff(off).astype(np.int64)
    rec = records[first]
    clj, crj = rec["clj"].astype(np.int64), rec["crj"].astype(np.int64)
    tlj, trj = rec["tlj"].astype(np.int64), rec["trj"].astype(np.int64)
    Lt = np.asarray([len(t) for t in targets], dtype=np.int64)[tix]
    a = np.where(rc, L - crj, clj)
    b = np.where(rc, L - clj, crj)
    c = np.where(rc, Lt - trj, tlj)
    d = np.where(rc, Lt - tlj, trj)
    # the insertion frequency_output passes is consensus[a:b] in the orientation it works in
    has_ins = (rec["flags"] & N.FL_INSERTION) != 0
    ins = HostRagged.from_views(cons, off[:-1] + np.where(has_ins, a, 0), np.where(has_ins, b - a, 0).astype(np.int32))
    oriented = [_rcomp(t) if r else t for t, r in zip(targets, target_rc)]
    batch = _HOMEOLOGY_BATCH[0] or _engine.homeology_batch
    t1 = time.perf_counter()
    out = batch(HostRagged(cons, offsets=off), ins, a, b, c, d, tix, oriented, list(targets), device)
    if os.environ.get("SCAR_TRACE"):
        print("[scar dropin] pattern_search_batch: {} patterns, arguments {:.1f} ms, batched call {:.1f} ms".format(
            n, 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)), file=sys.stderr)
    return out


def batched_consensus_demultiplex(self, module, device=0):
    """Replacement body of DataProcessing.consensus_demultiplex."""
    args, log = self.args, self.log
    log.info("Consensus Index Search (batched, GPU)")
    if args.Platform not in N.PLATFORMS:
        log.error("--Platform {} not correctly defined.  Edit parameter file and try again".format(args.Platform))
        raise SystemExit(1)

    # ---- static problem: manifest order, targets, cutsites, phasing
    sample_ids = list(self.index_dict)
    target_names, target_pos = [], {}
    for iid in sample_ids:
        t = self.index_dict[iid][7]
        if t not in target_pos:
            target_pos[t] = len(target_names)
            target_names.append(t)
    targets, cutsites, target_rc = [], [], []
    for t in target_names:
        region = fetch_target_region(module, args, self.target_dict, t, log)
        sgrna = self.target_dict[t][4]
        try:
            cut = cutsite_search(region, sgrna, self.target_dict[t][5] == "YES")
        except N.ScarError:
            log.error("sgRNA {} does not map to locus {}; chr{}:{}-{}.  Check --TargetFile and try again."
                      .format(sgrna, t, self.target_dict[t][1], self.target_dict[t][2], self.target_dict[t][3]))
            raise SystemExit(1)
        targets.append(region)
        cutsites.append(cut)
        target_rc.append(self.target_dict[t][5] == "YES")
    phasing = None
    if args.Platform == "Illumina":
        phasing = [([p[0] for p in self.phase_dict[t]["R1"]], [p[0] for p in self.phase_dict[t]["R2"]])
                   for t in target_names]

    # ---- reads: one text buffer, views for the sequence and the two header index fields.
    # --Verbose DEBUG (:906-914): the reference tests `read_count > 1000000` at the top of an iteration that still
    # reads and processes one more record, so it handles 1 000 002 reads.
    limit = READ_LIMIT_DEBUG if args.Verbose == "DEBUG" else None
    STAGE_SECONDS.clear()
    t_stage = time.perf_counter()

    def stage(name):
        nonlocal t_stage
        now = time.perf_counter()
        STAGE_SECONDS[name] = STAGE_SECONDS.get(name, 0.0) + now - t_stage
        t_stage = now
    sample_target = [target_pos[self.index_dict[i][7]] for i in sample_ids]
    common = dict(
        targets=targets, cutsites=cutsites, sample_target=sample_target,
        left_index=[self.index_dict[i][0] for i in sample_ids], right_index=[self.index_dict[i][2] for i in sample_ids],
        platform=args.Platform, phasing=phasing, hr_donor=args.HR_Donor or "", n_limit=args.N_Limit,
        min_length=args.Minimum_Length)
    streaming = FILE_INGEST[0] and hasattr(_ENGINE_FACTORY[0] or ScarEngine, "process_fastq_file")
    device_gather = device_gather_flipped = None
    if streaming:
        # The file is streamed through the GPU (scar_process_fastq_file): read / inflated by a producer thread while
        # its pieces are copied, indexed and classified on the device.  Nothing about the run is known up front: the
        # engine re-plans itself for the longest read and grows its scar table as the patterns arrive.  The text stays
        # resident, so every scarred read's full key is compared with the first read of its table entry at the end
        # (exact-or-fail identity; a mismatch raises instead of merging two scar patterns under one 128-bit tag).
        if len(self.fastq1.input_file) < 3 or not os.path.isfile(self.fastq1.input_file):
            raise FileNotFoundError("FASTQ file {} not found".format(self.fastq1.input_file))
        trace = [] if os.environ.get("SCAR_TRACE") else None

        def mark(what):
            if trace is not None:
                trace.append((what, time.perf_counter()))
        mark("start")
        eng = _make_engine(ScarProblem(max_read_len=160, table_capacity=1 << 18, **common), device)
        mark("context created")
        try:
            eng.set_verify(True)
            STAGE_SECONDS["keys_verified_exactly"] = 1.0
            # the host needs the text only for the Raw_Data / demultiplexed-FASTQ writers: otherwise the first read of
            # every scar pattern is gathered on the device, where the text is (below), and no copy is kept
            keep_text = bool(args.OutputRawData or args.Demultiplex) or not hasattr(eng, "fastq_gather")
            batch = eng.process_fastq_file(self.fastq1.input_file, limit=limit, threads=FILE_INGEST_THREADS[0],
                                           keep_text=keep_text)
            mark("file through the GPU")
            records = batch.records
            self.read_count = batch.n
            STAGE_SECONDS["fastq_source_" + batch.source] = 1.0
            STAGE_SECONDS["fastq_producer_wait"] = batch.wait_seconds
            counters, unassigned = eng.counters()
            phase_hist = eng.phase_hist()
            table = eng.table()                 # sorted by (sample, first_idx): per-sample first-seen order
            mark("counters and table read")
            if not keep_text:
                firsts = table["first_idx"].astype(np.int64)
                device_gather = eng.fastq_gather(firsts, max_len=batch.max_seq_len)
                if any(target_rc):
                    flips = np.asarray(target_rc, dtype=bool)[np.asarray(sample_target, dtype=np.int64)[table["sample"].astype(np.int64)]]
                    device_gather_flipped = eng.fastq_gather(firsts, flip=flips.astype(np.uint8), max_len=batch.max_seq_len)
                mark("first reads gathered on the device")
        except N.ScarError as exc:
            if exc.code != N.ERR_NOMEM:
                raise
            # the text does not fit the device: the host loader below feeds the GPU in chunks instead
            log.warning("FASTQ text does not fit the device ({}); reading it on the host and processing in chunks".format(exc))
            streaming = False
        finally:
            eng.close()
            mark("context closed")
            if trace:
                print("[scar dropin] " + ", ".join("{} {:.0f} ms".format(w, 1e3 * (t - trace[i][1])) for i, (w, t) in
                                                   enumerate(trace[1:])), file=sys.stderr)
    if streaming:
        stage("fastq_stream_gpu_classify_aggregate")
    else:
        batch = fastq.load(self.fastq1.input_file, want_index_fields=args.Platform == "Illumina", limit=limit)
        self.read_count = batch.n
        stage("fastq_read_inflate_index")
        prob = ScarProblem(max_read_len=max(16, batch.seq.max_len()),
                           table_capacity=1 << max(16, int(math.ceil(math.log2(max(2 * batch.n, 2))))), **common)
        eng = _make_engine(prob, device)
        try:
            # Exact-or-fail key identity: when the run fits the device (text + ~60 bytes of views / records per read) it
            # is processed as one resident chunk and every scarred read's full key is compared with the first read of
            # its table entry; a mismatch raises instead of merging two scar patterns under one 128-bit tag.
            if hasattr(eng, "set_verify"):
                free, _total = eng.device_memory()
                need = int(batch.text.size) + 64 * batch.n + (256 << 20)
                if need < 0.8 * free:
                    eng.set_verify(True)
                    STAGE_SECONDS["keys_verified_exactly"] = 1.0
                else:
                    log.warning("Run too large for one resident chunk ({:.1f} GB needed, {:.1f} GB free): scar patterns are "
                                "identified by their 128-bit key tags without the exact comparison".format(need / 1e9, free / 1e9))
                    STAGE_SECONDS["keys_verified_exactly"] = 0.0
            records = eng.process_host(batch.seq, batch.f0, batch.f1)
            counters, unassigned = eng.counters()
            phase_hist = eng.phase_hist()
            table = eng.table()                     # sorted by (sample, first_idx): per-sample first-seen order
        finally:
            eng.close()
        stage("gpu_classify_aggregate")

    # ---- unpack into the reference's bookkeeping
    n_samples = len(sample_ids)
    n_assigned = int(counters[:, N.CT_ASSIGNED].sum())
    # read_count_dict: a key appears once the manifest loop has visited it (:1173-1174)
    seen = np.nonzero(counters[:, N.CT_ASSIGNED])[0]
    last_visited = n_samples - 1 if unassigned else (int(seen[-1]) if seen.size else -1)
    for s in range(last_visited + 1):
        self.read_count_dict[sample_ids[s]] = int(counters[s, N.CT_ASSIGNED])
    if unassigned:
        self.read_count_dict["unidentified"] = int(unassigned)

    # sequence_dict in order of each sample's first read (dict insertion order, :946).  One pass, no sort: a fancy
    # assignment keeps the LAST value written to a repeated index, so writing the read numbers in descending order
    # leaves every sample's first read.
    samples16 = np.ascontiguousarray(records["sample"])              # unidentified reads carry SCAR_SAMPLE_NONE
    first_read = np.full(N.SAMPLE_NONE + 1, -1, dtype=np.int64)
    first_read[samples16[::-1]] = np.arange(samples16.size - 1, -1, -1, dtype=np.int64)
    first_of = {int(s): int(first_read[s]) for s in np.nonzero(first_read[:n_samples] >= 0)[0]}
    samples = None                                                   # (int64 copy: only the optional writers need it)

    # ---- per-pattern work of the whole run, in the parent: first reads gathered once, searches in one GPU call
    tab_sample = table["sample"].astype(np.int64)
    tab_first = table["first_idx"].astype(np.int64)
    bounds = np.searchsorted(tab_sample, np.arange(n_samples + 1))
    first_bytes, first_off = device_gather if device_gather is not None else _gather(batch.seq, tab_first, args.Platform)
    stage("unpack")
    pattern_out = pattern_search_batch(records, batch.seq, args.Platform, tab_first, tab_sample, sample_target, targets,
                                       target_rc, device, gathered=(first_bytes, first_off),
                                       gathered_flipped=device_gather_flipped)
    stage("gpu_pattern_searches")
    # scarred reads bucketed by sample once (stable: read order inside a sample)
    scar_rows, scar_bounds = None, None
    if args.OutputRawData:
        scar_idx = np.nonzero(records["status"] == N.ST_SCAR)[0]
        samples = samples16.astype(np.int64)
        order = np.argsort(samples[scar_idx], kind="stable")
        scar_rows = scar_idx[order].astype(np.int64)
        scar_bounds = np.searchsorted(samples[scar_rows], np.arange(n_samples + 1))

    if args.Demultiplex:
        write_demultiplexed_fastq(self, batch, records, sample_ids)
        stage("demultiplexed_fastq")

    _INHERITED.clear()
    _RUN_SERIAL[0] += 1
    for s in sorted(first_of, key=first_of.get):
        iid = sample_ids[s]
        t = sample_target[s]
        res = SampleResult(iid, int(counters[s, N.CT_ASSIGNED]))
        res._token = (os.getpid(), _RUN_SERIAL[0], iid)
        _INHERITED[res._token] = res
        res.platform = args.Platform
        res.target_region, res.cutsite, res.reverse_comp = targets[t], cutsites[t], target_rc[t]
        res.counters = counters[s].copy()
        lo, hi = int(bounds[s]), int(bounds[s + 1])
        res.counts = table["count"][lo:hi].copy()
        res.first_records = records[tab_first[lo:hi]]
        b0, b1 = int(first_off[lo]), int(first_off[hi])
        res.first_bytes = first_bytes[b0:b1].copy()
        res.first_off = first_off[lo:hi + 1] - b0
        res.pattern_out = pattern_out[lo:hi].copy()
        if args.OutputRawData:
            rows = scar_rows[int(scar_bounds[s]):int(scar_bounds[s + 1])]
            write_raw_data(self, module, iid, res, rows, records, batch.seq)
        self.sequence_dict[iid] = res

        # phase_count (:948-975)
        if args.Platform == "Illumina":
            locus = self.index_dict[iid][7]
            key = "{}+{}".format(iid, locus)
            for k, (r2_phase, r1_phase) in enumerate(zip(self.phase_dict[locus]["R2"], self.phase_dict[locus]["R1"])):
                l1, l2 = "Phase " + r1_phase[1], "Phase " + r2_phase[1]
                if not r1_phase[0]:
                    self.phase_count[key][l1] = -1
                    self.phase_count[key][l2] = -1
                else:
                    self.phase_count[key][l1] += int(phase_hist[s, 0, 1 + k])
                    self.phase_count[key][l2] += int(phase_hist[s, 1, 1 + k])
            if phase_hist[s, 1, 0]:
                self.phase_count[key]["No Read 2 Phasing"] += int(phase_hist[s, 1, 0])
            if phase_hist[s, 0, 0]:
                self.phase_count[key]["No Read 1 Phasing"] += int(phase_hist[s, 0, 0])

    stage("per_sample_results_and_raw_data")
    log.info("Batched path: {} reads, {} scar patterns; ".format(self.read_count, int(tab_first.size)) +
             ", ".join("{} {:.2f} s".format(k, v) for k, v in STAGE_SECONDS.items()))

    # ---- lower limit for plots (:1003-1024, unchanged arithmetic)
    key_counts = [len(self.sequence_dict[k]) for k in self.sequence_dict]
    if len(key_counts) == 0:
        log.error("No Scar Patterns Found")
        raise SystemExit(1)
    if len(key_counts) > 1:
        lower_limit = statistics.mean(key_counts) - norm_interval_lower(statistics.mean(key_counts), sem(key_counts))
    else:
        lower_limit = float("NaN")
    if math.isnan(lower_limit):
        lower_limit = args.PatternThreshold * 0.00001
    return n_assigned, lower_limit


def _gather(reads, rows, platform):
    from . import engine as _engine
    return _engine.stored_gather(reads, rows, platform)


def write_demultiplexed_fastq(self, batch, records, sample_ids):
    """--Demultiplex True (scarmapper/INDEL_Processing.py:876-889,988-1002): the reference keeps [name, seq, qual] of
    every read per sample in Python lists and writes "@name\\nseq\\n+\\nqual\\n" per read into the R1 file that
    dictionary_build opened for the sample ({Job}_{index}_R1.fastq; unmatched reads into {Job}_Unknown_R1.fastq).
    Here the reads are bucketed by sample once and every file is formatted natively from the FASTQ text
    (scar_demux_fastq_rows_host).  The reference then asks gzip to compress files named *_Consensus.fastq, which it
    never wrote (:986,1000,1002 -> Tool_Box.compress_files: gzip reports 'No such file'); that step has no effect
    and is not repeated."""
    from . import engine as _engine
    if not self.args.PEAR:
        self.log.error("--Demultiplex with un-merged read pairs (PEAR False) is not part of the batched path")
        raise SystemExit(1)
    samples = records["sample"].astype(np.int64)
    samples[records["status"] == N.ST_UNASSIGNED] = len(sample_ids)         # 'Unknown' sorts last
    order = np.argsort(samples, kind="stable").astype(np.int64)
    bounds = np.searchsorted(samples[order], np.arange(len(sample_ids) + 2))
    for s, key in enumerate(list(sample_ids) + ["Unknown"]):
        rows = order[int(bounds[s]):int(bounds[s + 1])]
        if rows.size == 0:
            continue                                   # the reference writes (and closes) only the files that got reads
        r1 = self.fastq_outfile_dict[key][0]           # FASTQ_Tools.Writer opened by dictionary_build
        path = r1.file.name
        r1.close()
        with open(path, "wb") as f:
            f.write(_engine.demux_fastq_rows(batch.text, batch.seq, rows))


RAW_DATA_COLUMNS = ("Left Deletions\tRight Deletions\tDeletion Size\tMicrohomology\tInsertion\tInsertion Size\t"
                    "Consensus Left Junction\tConsensus Right Junction\tRef Left Junction\tRef Right Junction\t"
                    "Consensus\tTarget Region\n")


def write_raw_data(self, module, index_name, res, rows, records, reads):
    """ScarSearch.raw_data_output (scarmapper/INDEL_Processing.py:702-757) without the per-read Python lists, run
    in the PARENT for one sample: same file name, the reference's own page header (ScarSearch.common_page_header
    evaluated on a stand-in carrying the attributes it reads), and the rows written by scar_raw_rows_indexed_host
    straight from the records of the sample's scarred reads `rows` (one line per scarred read, swaps and reverse
    complements for 3'-strand guides, unaltered reads skipped)."""
    from . import engine as _engine
    stub = types.SimpleNamespace(version=self.version, run_start=self.run_start, index_dict=self.index_dict,
                                 target_dict=self.target_dict, args=self.args)
    header = module.ScarSearch.common_page_header(stub, index_name)
    body = _engine.raw_rows_indexed(records, reads, rows, res.platform, res.target_region, res.cutsite, res.reverse_comp)
    path = "{}{}_{}_ScarMapper_Raw_Data.txt".format(self.args.WorkingFolder, self.args.Job_Name, index_name)
    with open(path, "wb") as f:
        f.write("{}{}".format(header, RAW_DATA_COLUMNS).encode())
        f.write(body)
    res.raw_written = True
    if _RAW_DATA_HOOK[0]:
        _RAW_DATA_HOOK[0](index_name, rows, records, reads, res)


def batched_data_processing(self):
    """Replacement body of ScarSearch.data_processing: unpack, then the reference's own writers.  Runs inside a
    forked pool worker: no CUDA call, only this sample's SampleResult."""
    res = self.sequence_list
    if not isinstance(res, SampleResult):
        raise TypeError("batched ScarSearch.data_processing needs the SampleResult produced by the batched demultiplex")
    self.log.info("Begin Processing {}".format(self.index_name))
    target_name = self.index_dict[self.index_name][7]
    c = res.counters
    self.summary_data = [self.index_name, int(c[N.CT_PASSING]), int(c[N.CT_LEFT_DEL]), int(c[N.CT_RIGHT_DEL]),
                         int(c[N.CT_INSERTION]), int(c[N.CT_MICROHOM]),
                         [int(c[N.CT_NO_JUNCTION]), int(c[N.CT_NO_CUT])], [int(c[N.CT_FILTERED]), 0],
                         'junction data', target_name, [int(c[N.CT_HR_FIRST]), int(c[N.CT_HR_MORE])]]
    junction_type_data = [0, 0, 0, 0, 0, 0]
    self.target_region = res.target_region
    self.cutsite = res.cutsite
    self.window_mapping()
    self.log.info("Finished Processing {}".format(self.index_name))
    if NATIVE_FREQUENCY[0]:
        # the per-pattern loop of frequency_output as one native call (no Python object per pattern)
        native_frequency_output(self, self.index_name, res, junction_type_data)
    else:
        # the reference's own frequency_output over results_freq_dict, its two searches answered from the batch
        results_freq_dict = collections.defaultdict(list)
        for count, sub_list in res.table():
            key = "{}|{}|{}|{}|{}".format(sub_list[0], sub_list[1], sub_list[2], sub_list[3], sub_list[9])
            results_freq_dict[key] = [count, sub_list]
        serve_pattern_searches(self, results_freq_dict, res.pattern_out)
        self.frequency_output(self.index_name, results_freq_dict, junction_type_data)
    if self.args.OutputRawData and not res.raw_written:
        self.raw_data_output(self.index_name, res.read_results or [])
    # the returned ScarSearch objects are pickled back to the parent for data_output, which reads summary_data only
    self.sequence_list = None
    return self.summary_data


FREQUENCY_COLUMNS = (
    "# Total\tFrequency\tScar Type\tLeft Deletions\tRight Deletions\tDeletion Size\tMicrohomology\t"
    "Microhomology Size\tInsertion\tInsertion Size\tLeft Template\tRight Template\tConsensus Left Junction\t"
    "Consensus Right Junction\tTarget Left Junction\tTarget Right Junction\tLeft Homeology\t"
    "Left Homeology Position\tLeft Query Seq\tLeft Homeologous Seq\tRight Homeology\t"
    "Right Homeology Position\tRight Query Seq\tRight Homeologous Seq\tConsensus\tTarget Region\n")


def native_frequency_output(self, index_name, res, junction_type_data):
    """ScarSearch.frequency_output (scarmapper/INDEL_Processing.py:233-548) with its per-pattern Python loop replaced by
    ONE native call (scar_frequency_rows_host: classification, junction-type totals, SNV tallies, threshold, ordering,
    the rows of the file).  What stays here is the framing the reference does once per sample: the page header it
    generates itself, the SNV file (:480-530) and the bookkeeping handed to its own plot routine (:396-470,531-548).
    Files identical to the reference's on every golden case (tests/test_dropin_*.py run both this and the
    unchanged frequency_output)."""
    from . import engine as _engine
    module = __import__(type(self).__module__, fromlist=["ScarMapperPlot"])
    self.log.info("Working on Frequency Files for {}".format(index_name))
    target_name = self.index_dict[index_name][7]
    rc = self.target_dict[target_name][5] == "YES"
    out = _engine.frequency_rows(res.counts, res.first_records, res.first_bytes, res.first_off, res.pattern_out,
                                 self.target_region, self.cutsite, rc, self.summary_data[1], self.summary_data[6][1],
                                 self.args.PatternThreshold)
    for i, v in enumerate(out["junction_type_data"]):
        junction_type_data[i] += v
    path = "{}{}_{}_ScarMapper_Frequency.txt".format(self.args.WorkingFolder, self.args.Job_Name, index_name)
    with open(path, "wb") as f:
        f.write("{}{}".format(self.common_page_header(index_name), FREQUENCY_COLUMNS).encode())
        f.write(out["rows"])
    scar_count, total_scars = out["scar_count"], out["total_scars"]
    lft_kmer, rt_kmer = self.left_target_windows[0], self.right_target_windows[0]
    kmer_size = len(lft_kmer)
    if rc and scar_count:            # (:315-316: swapped and reverse-complemented inside the loop, i.e. once any pattern exists)
        lft_kmer, rt_kmer = _rcomp(self.right_target_windows[0]), _rcomp(self.left_target_windows[0])

    # SNV file (:480-530)
    if junction_type_data[5] > 0:
        snv = out["snv"]             # [left, left x count, right, right x count][position][G, A, T, C]
        lft_kmer_string, rt_kmer_string = "\t".join(lft_kmer), "\t".join(rt_kmer)
        lft_position_label = "".join("\t{}".format(-1 * (kmer_size - i)) for i in range(kmer_size))
        rt_position_label = "".join("\t{}".format(i + 1) for i in range(kmer_size))
        snv_outstring = "{}\n\n# Values normalized to number of SNV scars\n\t{}\t{}\n# NT{}{}" \
            .format(self.common_page_header(index_name), lft_kmer_string, rt_kmer_string, lft_position_label, rt_position_label)
        snv_all_outstring = "\n\n\n# Values normalized to total number of scars\n\t{}\t{}\n# NT{}{}" \
            .format(lft_kmer_string, rt_kmer_string, lft_position_label, rt_position_label)
        for q, nt in enumerate(["G", "A", "T", "C"]):
            snv_outstring += "\n# {}".format(nt)
            snv_all_outstring += "\n# {}".format(nt)
            data = [int(snv[0, i, q]) for i in range(kmer_size)] + [int(snv[2, i, q]) for i in range(kmer_size)]
            data_all = [int(snv[1, i, q]) for i in range(kmer_size)] + [int(snv[3, i, q]) for i in range(kmer_size)]
            for v in data:
                snv_outstring += "\t{}".format(round(v / scar_count, 4))
            for v in data_all:
                snv_all_outstring += "\t{}".format(round(v / total_scars, 4))
        with open("{}{}_{}_SNV_Frequency.txt".format(self.args.WorkingFolder, self.args.Job_Name, index_name), "w") as f:
            f.write(snv_outstring + snv_all_outstring)

    self.summary_data[8] = junction_type_data
    scar_fraction = (junction_type_data[5] + self.summary_data[1] - self.summary_data[6][1] - self.summary_data[6][0]) / \
        self.summary_data[1]
    if self.summary_data[1] >= self.lower_limit_count and scar_fraction >= 0.08:
        # the plot's inputs, built from the rows that were written exactly as :414-467 builds them
        plot_data_dict = collections.defaultdict(list)
        marker_list = []
        label_dict = collections.defaultdict(float)
        for t in out["label_order"]:
            label_dict[_engine.SCAR_TYPES[t]] = float(out["label_sums"][t])
        for t, freq, lft_del, rt_del, mh_size, ins_size in out["plot"].tolist():
            scar_type = _engine.SCAR_TYPES[int(t)]
            lft_del, rt_del, mh_size, ins_size = int(lft_del), int(rt_del), int(mh_size), int(ins_size)
            if freq < 0.0025:
                freq = 0.0025
            y_value = freq * 0.5
            lft_ins_width = rt_ins_width = freq
            marker_list.extend([lft_del + (mh_size * 0.5), rt_del + (mh_size * 0.5), ins_size])
            lft_del_plot_value = (lft_del + (mh_size * 0.5)) * -1
            rt_del_plot_value = (rt_del + (mh_size * 0.5))
            lft_ins_plot_value = (ins_size * 0.5) * -1
            rt_ins_plot_value = (ins_size * 0.5)
            if lft_del != 0:
                lft_ins_width = freq * 0.5
            if rt_del != 0:
                rt_ins_width = freq * 0.5
            if scar_type not in plot_data_dict:
                plot_data_dict[scar_type] = [[freq], [lft_del_plot_value], [rt_del_plot_value], [lft_ins_plot_value],
                                             [rt_ins_plot_value], [lft_ins_width], [rt_ins_width], [y_value]]
            else:
                count = len(plot_data_dict[scar_type][0])
                previous_freq = plot_data_dict[scar_type][0][count - 1]
                previous_y = plot_data_dict[scar_type][7][count - 1]
                plot_data_dict[scar_type][0].append(freq)
                plot_data_dict[scar_type][1].append(lft_del_plot_value)
                plot_data_dict[scar_type][2].append(rt_del_plot_value)
                plot_data_dict[scar_type][3].append(lft_ins_plot_value)
                plot_data_dict[scar_type][4].append(rt_ins_plot_value)
                plot_data_dict[scar_type][5].append(lft_ins_width)
                plot_data_dict[scar_type][6].append(rt_ins_width)
                plot_data_dict[scar_type][7].append(previous_y + 0.002 + (0.5 * previous_freq) + y_value)
        sample_name = "{}.{}".format(self.index_dict[index_name][5], self.index_dict[index_name][6])
        plot_data_dict['Marker'] = [(max(marker_list)) * -1, max(marker_list)]
        module.ScarMapperPlot.scarmapperplot(self.args, datafile=None, sample_name=sample_name,
                                             plot_data_dict=plot_data_dict, label_dict=label_dict)


def serve_pattern_searches(self, results_freq_dict, pattern_out):
    """Serve the results of the batched pattern searches (computed in the parent, pattern_search_batch) to the
    unchanged frequency_output through instance attributes named like the two methods it calls
    (homeology_search / templated_insertion_search, scarmapper/INDEL_Processing.py:550-700).  The lookup keys are
    derived per pattern exactly as frequency_output derives its arguments (:299-328); a call whose arguments are not
    in the batch simply runs the reference's own method, so correctness never rests on this derivation."""
    if pattern_out is None or len(results_freq_dict) == 0:
        return
    assert len(pattern_out) == len(results_freq_dict)
    target_name = self.index_dict[self.index_name][7]
    rc = self.target_dict[target_name][5] == "YES"
    region = self.target_region
    oriented = _rcomp(region) if rc else region
    hom, tmpl = {}, {}
    for (_count, sub), o in zip(results_freq_dict.values(), pattern_out.tolist()):
        insertion, consensus = sub[2], sub[4]
        a, b, c, d = sub[5], sub[6], sub[7], sub[8]
        if rc:
            insertion, consensus = _rcomp(insertion), _rcomp(consensus)
            a, b = len(consensus) - sub[6], len(consensus) - sub[5]
            c, d = len(region) - sub[8], len(region) - sub[7]
        left = ["", "", "", ""] if o[0] < 0 else (o[0], o[1], consensus[o[2]:o[2] + o[3]], oriented[o[4]:o[4] + o[5]])
        right = ["", "", "", ""] if o[6] < 0 else (o[6], o[7], consensus[o[8]:o[8] + o[9]], oriented[o[10]:o[10] + o[11]])
        hom[(consensus, a, b, insertion, c, d)] = (left, right)
        if len(insertion) >= 5:
            lt = "" if o[12] < 0 else region[o[12]:o[12] + 5]
            rt = "" if o[13] < 0 else region[o[13]:o[13] + 5]
            tmpl[(insertion, c, d)] = (_rcomp(lt), _rcomp(rt)) if rc else (lt, rt)
    cls = type(self)
    misses = [0]

    def homeology_search(consensus, consensus_lft_junction, consensus_rt_junction, target_sequence, insertion,
                         ref_lft_junction, ref_rt_junction):
        got = hom.get((consensus, consensus_lft_junction, consensus_rt_junction, insertion, ref_lft_junction,
                       ref_rt_junction)) if target_sequence == oriented else None
        PATTERN_SEARCH_STATS["served" if got is not None else "missed"] += 1
        if got is None:
            misses[0] += 1
            return cls.homeology_search(self, consensus, consensus_lft_junction, consensus_rt_junction, target_sequence,
                                        insertion, ref_lft_junction, ref_rt_junction)
        return got

    def templated_insertion_search(insertion, lft_target_junction, rt_target_junction, name):
        got = tmpl.get((insertion, lft_target_junction, rt_target_junction)) if name == target_name else None
        PATTERN_SEARCH_STATS["served" if got is not None else "missed"] += 1
        if got is None:
            misses[0] += 1
            return cls.templated_insertion_search(self, insertion, lft_target_junction, rt_target_junction, name)
        return got

    self.homeology_search = homeology_search
    self.templated_insertion_search = templated_insertion_search
    self.pattern_search_misses = misses


def _drop_served_searches(search):
    """The two closures installed by serve_pattern_searches are not picklable and are not needed once
    frequency_output has run; the pool pickles every ScarSearch back to the parent."""
    for name in ("homeology_search", "templated_insertion_search"):
        search.__dict__.pop(name, None)


def install(module, device: int = 0):
    """Swap the two hot-loop bodies of the reference's INDEL_Processing module for the batched ones."""
    def consensus_demultiplex(self):
        return batched_consensus_demultiplex(self, module, device)

    def data_processing(self):
        try:
            return batched_data_processing(self)
        finally:
            _drop_served_searches(self)
    module.DataProcessing.consensus_demultiplex = consensus_demultiplex
    module.ScarSearch.data_processing = data_processing
    return module
