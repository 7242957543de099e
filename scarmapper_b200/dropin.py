"""
dropin.py — batched replacements for the bodies of the reference's two hot loops.

install(INDEL_Processing) swaps two methods of the (otherwise unmodified) reference module:

  DataProcessing.consensus_demultiplex   (scarmapper/INDEL_Processing.py:891-1024, HOT LOOP A)
      reference: a serial Python loop, per read a loop over the manifest with two Levenshtein calls.
      here: the FASTQ text is indexed once, copied to the GPU, and ONE call (scar_process_host) runs
      demultiplex -> pack -> sliding window -> aggregation for every read of the run.
  ScarSearch.data_processing             (:82-207, HOT LOOP B, one pathos process per sample)
      reference: per read filters + SlidingWindow.sliding_window + dict aggregation.
      here: nothing is computed; the per-sample results of the GPU pass are unpacked into exactly
      the objects the reference's reporting code consumes (summary_data, results_freq_dict in
      first-seen order with the first read's sub_list, read_results_list) and frequency_output /
      raw_data_output (unchanged reference code) write the files.

Everything else — options file, Logger, TargetMapper, main_loop, the pathos Pool, data_output,
frequency_output, plots — stays the reference's own code.
"""
from __future__ import annotations

import collections
import math
import statistics

import numpy as np

from . import _native as N
from . import fastq
from .engine import ScarEngine, ScarProblem, cutsite_search

_ENGINE_FACTORY = [None]      # tests substitute a CPU stand-in here; None -> ScarEngine
_HOMEOLOGY_BATCH = [None]     # likewise for the per-pattern searches; None -> engine.homeology_batch (GPU)
PATTERN_SEARCH_STATS = {"served": 0, "missed": 0}   # calls answered from the batch / handed to the reference's method


def _make_engine(prob: ScarProblem, device: int):
    f = _ENGINE_FACTORY[0]
    return f(prob, device) if f else ScarEngine(prob, device)


class SampleResult:
    """What the GPU pass produced for one sample library; stands in for sequence_dict[index]
    (main_loop only takes its len() and hands it to ScarSearch; INDEL_Processing.py:1051-1053)."""

    def __init__(self, index_name, n_reads):
        self.index_name = index_name
        self.n_reads = n_reads
        self.target_region = None
        self.cutsite = None
        self.counters = None          # int64[SCAR_N_COUNTERS]
        self.table = []               # [(count, sub_list of the first read)] in first-seen order
        self._read_results = None     # every SCAR read's sub_list (OutputRawData only), built on first use
        self._read_results_fn = None
        self.raw_source = None        # what the native Raw_Data writer needs (records, reads, sample, ...)
        self.device = 0               # CUDA device the sample's GPU work runs on

    @property
    def read_results(self):
        if self._read_results is None and self._read_results_fn is not None:
            self._read_results = self._read_results_fn()
        return self._read_results

    @read_results.setter
    def read_results(self, value):
        self._read_results = value

    def __len__(self):
        return self.n_reads


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


def batched_consensus_demultiplex(self, module, device=0):
    """Replacement body of DataProcessing.consensus_demultiplex."""
    args, log = self.args, self.log
    log.info("Consensus Index Search (batched, GPU)")
    if args.Demultiplex:
        log.error("--Demultiplex True (writing demultiplexed FASTQ) is not part of the batched path; "
                  "set it to False or run the reference loop.")
        raise SystemExit(1)
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
    targets, cutsites = [], []
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
    phasing = None
    if args.Platform == "Illumina":
        phasing = [([p[0] for p in self.phase_dict[t]["R1"]], [p[0] for p in self.phase_dict[t]["R2"]])
                   for t in target_names]

    # ---- reads: one text buffer, views for the sequence and the two header index fields
    limit = 1000001 if args.Verbose == "DEBUG" else None       # :906-914
    batch = fastq.load(self.fastq1.input_file, want_index_fields=args.Platform == "Illumina", limit=limit)
    self.read_count = batch.n
    prob = ScarProblem(
        targets=targets, cutsites=cutsites, sample_target=[target_pos[self.index_dict[i][7]] for i in sample_ids],
        left_index=[self.index_dict[i][0] for i in sample_ids], right_index=[self.index_dict[i][2] for i in sample_ids],
        platform=args.Platform, phasing=phasing, hr_donor=args.HR_Donor or "", n_limit=args.N_Limit,
        min_length=args.Minimum_Length, max_read_len=max(16, batch.seq.max_len()),
        table_capacity=1 << max(16, int(math.ceil(math.log2(max(2 * batch.n, 2))))))
    eng = _make_engine(prob, device)
    try:
        records = eng.process_host(batch.seq, batch.f0, batch.f1)
        counters, unassigned = eng.counters()
        phase_hist = eng.phase_hist()
        table = eng.table()
    finally:
        eng.close()

    # ---- unpack into the reference's bookkeeping
    assigned = records["status"] != N.ST_UNASSIGNED
    samples = records["sample"].astype(np.int64)
    n_samples = len(sample_ids)
    # read_count_dict: a key appears once the manifest loop has visited it (:1173-1174)
    last_visited = n_samples - 1 if unassigned else (int(samples[assigned].max()) if assigned.any() else -1)
    for s in range(last_visited + 1):
        self.read_count_dict[sample_ids[s]] = int(counters[s, N.CT_ASSIGNED])
    if unassigned:
        self.read_count_dict["unidentified"] = int(unassigned)

    # sequence_dict in order of each sample's first read (dict insertion order, :946)
    first_of = {}
    idx = np.nonzero(assigned)[0]
    if idx.size:
        u, first = np.unique(samples[idx], return_index=True)
        first_of = {int(s): int(idx[f]) for s, f in zip(u, first)}
    by_sample = collections.defaultdict(list)
    for e in table:
        by_sample[int(e["sample"])].append((int(e["count"]), int(e["first_idx"])))
    scar_reads = collections.defaultdict(list)
    if args.OutputRawData:
        for r in np.nonzero(records["status"] == N.ST_SCAR)[0]:
            scar_reads[int(samples[r])].append(int(r))
    for s in sorted(first_of, key=first_of.get):
        iid = sample_ids[s]
        t = prob.sample_target[s]
        res = SampleResult(iid, int(counters[s, N.CT_ASSIGNED]))
        res.device = device
        res.target_region, res.cutsite, res.counters = targets[t], cutsites[t], counters[s].copy()

        def sub(r, t=t):
            return record_to_sublist(records[r], stored_sequence(args.Platform, batch.seq.item(r)), targets[t], cutsites[t])
        res.table = [(count, sub(first)) for count, first in by_sample.get(s, [])]
        if args.OutputRawData:
            # the per-read list of lists the reference's raw_data_output consumes is only built if somebody asks
            # for it; the Raw_Data file itself is written natively from the records (native_raw_data_output)
            res._read_results_fn = (lambda rows=scar_reads.get(s, []), sub=sub: [sub(r) for r in rows])
            res.raw_source = dict(records=records, reads=batch.seq, sample=s, platform=args.Platform,
                                  region=targets[t], cutsite=cutsites[t])
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

    # ---- lower limit for plots (:1003-1024, unchanged arithmetic)
    key_counts = [len(self.sequence_dict[k]) for k in self.sequence_dict]
    if len(key_counts) == 0:
        log.error("No Scar Patterns Found")
        raise SystemExit(1)
    if len(key_counts) > 1:
        lower, _ = module.stats.norm.interval(0.9, loc=statistics.mean(key_counts), scale=module.stats.sem(key_counts))
        lower_limit = statistics.mean(key_counts) - lower
    else:
        lower_limit = float("NaN")
    if math.isnan(lower_limit):
        lower_limit = args.PatternThreshold * 0.00001
    return int(assigned.sum()), lower_limit


def batched_data_processing(self):
    """Replacement body of ScarSearch.data_processing: unpack, then the reference's own writers."""
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
    results_freq_dict = collections.defaultdict(list)
    for count, sub_list in res.table:
        key = "{}|{}|{}|{}|{}".format(sub_list[0], sub_list[1], sub_list[2], sub_list[3], sub_list[9])
        results_freq_dict[key] = [count, sub_list]
    self.log.info("Finished Processing {}".format(self.index_name))
    batched_pattern_searches(self, results_freq_dict, getattr(res, "device", 0))
    self.frequency_output(self.index_name, results_freq_dict, junction_type_data)
    if self.args.OutputRawData:
        if res.raw_source is not None:
            native_raw_data_output(self, self.index_name, res)
        else:
            self.raw_data_output(self.index_name, res.read_results or [])
    return self.summary_data


RAW_DATA_COLUMNS = ("Left Deletions\tRight Deletions\tDeletion Size\tMicrohomology\tInsertion\tInsertion Size\t"
                    "Consensus Left Junction\tConsensus Right Junction\tRef Left Junction\tRef Right Junction\t"
                    "Consensus\tTarget Region\n")


def native_raw_data_output(self, index_name, res):
    """ScarSearch.raw_data_output (scarmapper/INDEL_Processing.py:702-757) without the per-read Python lists:
    same file name, the reference's own page header, and the rows written by scar_raw_rows_host straight from the
    records (one line per scarred read, swaps and reverse complements for 3'-strand guides, unaltered reads
    skipped)."""
    from . import engine as _engine
    src = res.raw_source
    target_name = self.index_dict[index_name][7]
    body = _engine.raw_rows(src["records"], src["reads"], src["sample"], src["platform"], src["region"], src["cutsite"],
                            self.target_dict[target_name][5] == "YES")
    path = "{}{}_{}_ScarMapper_Raw_Data.txt".format(self.args.WorkingFolder, self.args.Job_Name, index_name)
    with open(path, "wb") as f:
        f.write("{}{}".format(self.common_page_header(index_name), RAW_DATA_COLUMNS).encode())
        f.write(body)


def batched_pattern_searches(self, results_freq_dict, device: int = 0):
    """The two per-pattern searches frequency_output runs in its Python loop - homeology_search and
    templated_insertion_search (scarmapper/INDEL_Processing.py:550-700) - for EVERY pattern of the sample in
    one batched GPU call (scar_homeology_batch), served to the unchanged frequency_output through instance
    attributes of the same names.  The arguments are derived per key exactly as frequency_output derives them
    (:299-328, incl. the swaps for 3'-strand guides); a call whose arguments are not in the batch simply runs
    the reference's own method, so correctness never rests on this derivation."""
    from .Sequence_Magic import rcomp
    from . import engine as _engine
    n = len(results_freq_dict)
    if n == 0:
        return
    target_name = self.index_dict[self.index_name][7]
    rc = self.target_dict[target_name][5] == "YES"
    region = self.target_region
    oriented = rcomp(region) if rc else region
    keys, cons, inss, clj, crj, tlj, trj = [], [], [], [], [], [], []
    for _count, sub in results_freq_dict.values():
        insertion, consensus = sub[2], sub[4]
        a, b, c, d = sub[5], sub[6], sub[7], sub[8]
        if rc:
            insertion, consensus = rcomp(insertion), rcomp(consensus)
            a, b = len(consensus) - sub[6], len(consensus) - sub[5]
            c, d = len(region) - sub[8], len(region) - sub[7]
        keys.append((consensus, a, b, insertion, c, d))
        cons.append(consensus); inss.append(insertion)
        clj.append(a); crj.append(b); tlj.append(c); trj.append(d)
    batch = _HOMEOLOGY_BATCH[0] or _engine.homeology_batch
    out = batch(cons, inss, clj, crj, tlj, trj, np.zeros(n, dtype=np.int32), [oriented], [region], device)
    hom, tmpl = {}, {}
    for key, o in zip(keys, out.tolist()):
        consensus, _a, _b, insertion, c, d = key
        left = ["", "", "", ""] if o[0] < 0 else (o[0], o[1], consensus[o[2]:o[2] + o[3]], oriented[o[4]:o[4] + o[5]])
        right = ["", "", "", ""] if o[6] < 0 else (o[6], o[7], consensus[o[8]:o[8] + o[9]], oriented[o[10]:o[10] + o[11]])
        hom[key] = (left, right)
        if len(insertion) >= 5:
            lt = "" if o[12] < 0 else region[o[12]:o[12] + 5]
            rt = "" if o[13] < 0 else region[o[13]:o[13] + 5]
            tmpl[(insertion, c, d)] = (rcomp(lt), rcomp(rt)) if rc else (lt, rt)
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


def install(module, device: int = 0):
    """Swap the two hot-loop bodies of the reference's INDEL_Processing module for the batched ones."""
    def consensus_demultiplex(self):
        return batched_consensus_demultiplex(self, module, device)
    module.DataProcessing.consensus_demultiplex = consensus_demultiplex
    module.ScarSearch.data_processing = batched_data_processing
    return module
