"""
make_golden.py — generate golden vectors by running the REFERENCE ITSELF in this container.

For each case below the unmodified reference driver
    DataProcessing(log, args, run_start, version, TargetMapper(...), FASTQ_Reader(...)).main_loop()
(scarmapper.py:268-273; INDEL_Processing.py:1038-1061) is run on seeded synthetic inputs with the
third-party stand-ins of oracle/ref_shims.py and the reference's own compiled SlidingWindow.pyx.
Captured per sample: results_freq_dict (key, count, first sub_list) in first-seen order,
summary_data, read_results_list coordinates; per run: read_count_dict, phase_count, and the
Frequency / Summary files minus their wall-clock lines.  Inputs are stored beside the outputs,
so the fixtures are self-contained on the GPU box (where /root/reference does not exist).

Run here only:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from oracle import ref_shims  # noqa: E402


def write_inputs(wd, case):
    """Files in the formats the reference parses."""
    with open(os.path.join(wd, "targets.tsv"), "w") as f:
        f.write("# TargetName\tChr\tStart\tStop\tsgRNA_Seq\tReverseComp\tComments\n")
        for nm, t, sg, rc in zip(case["target_names"], case["targets"], case["sgrnas"], case["reverse_comp"]):
            f.write("{}\t{}\t0\t{}\t{}\t{}\t\n".format(nm, "ctg_" + nm, len(t), sg, "Yes" if rc else "No"))
    with open(os.path.join(wd, "refseq.fa"), "w") as f:
        for nm, t in zip(case["target_names"], case["targets"]):
            f.write(">ctg_{}\n{}\n".format(nm, t))
    with open(os.path.join(wd, "master_index.tsv"), "w") as f:
        f.write("# Index_ID\tForward\tReverse\n")
        for iid, a, b in zip(case["sample_ids"], case["index_col2"], case["index_col3"]):
            f.write("{}\t{}\t{}\n".format(iid, a, b))
    with open(os.path.join(wd, "manifest.tsv"), "w") as f:
        f.write("# Index\tSample Name\tReplicate\tSpecies\tForward Phasing\tReverse Phasing\tTarget Name\n")
        for s, iid in enumerate(case["sample_ids"]):
            t = case["sample_target"][s]
            fwd, rev = case["phasing_seqs"][t]
            f.write("{}\tS{}\t1\tsyn\t{}\t{}\t{}\n".format(iid, s, fwd, rev, case["target_names"][t]))
    with gzip.open(os.path.join(wd, "reads.fastq.gz"), "wt") as f:
        for i, (name, seq) in enumerate(zip(case["names"], case["reads"])):
            f.write("@{}\n{}\n+\n{}\n".format(name, seq, "I" * len(seq)))


def raw_data_digest(path):
    """sha256 of a Raw_Data file without its wall-clock and version header lines."""
    import hashlib
    lines = open(path, "rb").read().split(b"\n")
    keep = [ln for ln in lines if not ln.startswith((b"# Run Start", b"# Run End", b"# ScarMapper Search v"))]
    return hashlib.sha256(b"\n".join(keep)).hexdigest()


def run_reference(case, demultiplex=False):
    ref_shims.install()
    from scarmapper import INDEL_Processing, TargetMapper
    from Valkyries import FASTQ_Tools, Tool_Box

    wd = tempfile.mkdtemp(prefix="scar_golden_") + os.sep
    try:
        write_inputs(wd, case)
        args = argparse.Namespace(
            WorkingFolder=wd, Job_Name="golden", Verbose="ERROR", Platform=case["platform"], PEAR=True,
            Demultiplex=demultiplex, HR_Donor=case["hr_donor"], N_Limit=case["n_limit"],
            Minimum_Length=case["min_length"], OutputRawData=True, PatternThreshold=0.0001,
            RefSeq=wd + "refseq.fa", TargetFile=wd + "targets.tsv", SampleManifest=wd + "manifest.tsv",
            Master_Index_File=wd + "master_index.tsv", Spawn=1, FASTQ1=wd + "reads.fastq.gz", FASTQ2=None,
            Search_KMER=10, FigureType="pdf", IndelProcessing=True)
        log = Tool_Box.Logger(args)
        captured = {}
        orig_freq = INDEL_Processing.ScarSearch.frequency_output
        orig_raw = INDEL_Processing.ScarSearch.raw_data_output

        def cap_freq(self, index_name, results_freq_dict, junction_type_data):
            captured.setdefault(index_name, {})["table"] = [
                [k, v[0]] + [v[1][i] for i in (0, 1, 2, 3, 5, 6, 7, 8, 9)] + [v[1][4]]
                for k, v in results_freq_dict.items()]
            return orig_freq(self, index_name, results_freq_dict, junction_type_data)

        def cap_raw(self, index_name, read_results_list):
            captured.setdefault(index_name, {})["read_results"] = [
                [r[5], r[6], r[7], r[8], r[9]] for r in read_results_list]
            return orig_raw(self, index_name, read_results_list)

        INDEL_Processing.ScarSearch.frequency_output = cap_freq
        INDEL_Processing.ScarSearch.raw_data_output = cap_raw
        try:
            sample_manifest = Tool_Box.FileParser.indices(log, args.SampleManifest)
            fq1 = FASTQ_Tools.FASTQ_Reader(args.FASTQ1, log)
            targeting = TargetMapper.TargetMapper(log, args, sample_manifest)
            dp = INDEL_Processing.DataProcessing(log, args, "run-start", "golden", targeting, fq1=fq1)
            results = {}
            orig_out = INDEL_Processing.DataProcessing.data_output

            def cap_out(self, summary_data_list):
                for ss in summary_data_list:
                    sd = ss.summary_data
                    results[sd[0]] = {"summary": [sd[1], sd[2], sd[3], sd[4], sd[5], sd[6], sd[7], sd[10]],
                                      "junction_types": sd[8], "target": sd[9], "cutsite": ss.cutsite,
                                      "n_left_windows": len(ss.left_target_windows),
                                      "n_right_windows": len(ss.right_target_windows)}
                return orig_out(self, summary_data_list)

            INDEL_Processing.DataProcessing.data_output = cap_out
            try:
                dp.main_loop()
            finally:
                INDEL_Processing.DataProcessing.data_output = orig_out
        finally:
            INDEL_Processing.ScarSearch.frequency_output = orig_freq
            INDEL_Processing.ScarSearch.raw_data_output = orig_raw
        for k in results:
            results[k].update(captured.get(k, {}))
        files, file_sha256 = {}, {}
        for fn in sorted(os.listdir(wd)):
            if fn.endswith(("_ScarMapper_Frequency.txt", "_ScarMapper_Summary.txt", "_SNV_Frequency.txt")):
                lines = open(os.path.join(wd, fn)).read().split("\n")
                keep = [ln for ln in lines if not ln.startswith(("# Run Start", "# Run End", "Start:", "End:",
                                                                 "FASTQ1:", "FASTQ2:"))]
                files[fn] = "\n".join(keep)
            elif fn.endswith("_R1.fastq"):
                import hashlib                 # --Demultiplex True: one FASTQ per sample (digest; empty files included)
                file_sha256[fn] = hashlib.sha256(open(os.path.join(wd, fn), "rb").read()).hexdigest()
            elif fn.endswith("_ScarMapper_Raw_Data.txt"):
                # one line per scarred read: too large to store, so the fixture keeps a digest of the file minus
                # its wall-clock and version lines
                file_sha256[fn] = raw_data_digest(os.path.join(wd, fn))
        return {
            "samples": results,
            "read_count": dp.read_count,
            "read_count_dict": dict(dp.read_count_dict),
            "phase_count": {k: dict(v) for k, v in dp.phase_count.items()},
            "files": files,
            "file_sha256": file_sha256,
        }
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def case_from_synth(cfg, n, hr_donor="", mutate=None, first=0):
    batch = cfg.generate(first, n)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    if mutate:
        reads, f0, f1 = mutate(reads, f0, f1)
    names = ["SYN:1:FC:1:{}:{}:{} 1:N:0:{}+{}".format(i % 97, i, i * 7 % 1013, a, b)
             for i, (a, b) in enumerate(zip(f0, f1))]
    return {
        "platform": cfg.platform, "hr_donor": hr_donor, "n_limit": cfg.n_limit, "min_length": cfg.min_length,
        "target_names": cfg.target_names, "targets": cfg.targets, "sgrnas": cfg.sgrnas,
        "reverse_comp": cfg.reverse_comp, "sample_ids": cfg.sample_ids,
        "index_col2": cfg.index_d7, "index_col3": cfg.index_d5, "sample_target": cfg.sample_target,
        "phasing_seqs": cfg.phasing_seqs, "names": names, "reads": reads,
    }


def rc(s):
    return synth.rcomp(s)


def build_cases():
    cases = {}
    rng = np.random.Generator(np.random.PCG64(7))

    cases["c1_single_target"] = case_from_synth(synth.make_config("C1", seed=11), 3000)
    cases["c2_96_samples"] = case_from_synth(synth.make_config("C2", seed=12), 5000)
    cases["c3_12_targets_mh"] = case_from_synth(synth.make_config("C3", seed=13), 5000)
    cases["c4_long_300bp"] = case_from_synth(synth.make_config("C4", seed=14), 1500)
    # a locus with a chance 10-mer repeat across the cut: wild-type reads hit a far window first
    rep = synth.make_config("C4", seed=14, allow_repeats=True)
    kmers = [rep.targets[0][i:i + 10] for i in range(len(rep.targets[0]) - 9)]
    assert len(set(kmers)) < len(kmers), "seed 14 no longer yields a repeat; pick another"
    cases["c4_repeat_locus"] = case_from_synth(rep, 1500)

    # ragged lengths around the loop bounds and the Minimum_Length filter, N-rich reads, odd index
    # fields (short / long / N), HR donor
    cfg = synth.make_config("C3", seed=15, min_len=30, p_n=0.05, p_idx1=0.2, p_idx2=0.05)

    def mutate(reads, f0, f1):
        reads, f0, f1 = list(reads), list(f0), list(f1)
        for i in range(len(reads)):
            u = rng.random()
            if u < 0.03:
                f0[i] = f0[i][:-1]                      # 7-nt i7
            elif u < 0.06:
                f1[i] = f1[i] + "ACGT"[int(rng.integers(4))]  # 9-nt i5
            elif u < 0.09:
                p = int(rng.integers(8)); f0[i] = f0[i][:p] + "N" + f0[i][p + 1:]
            elif u < 0.11:
                f0[i] = f0[i][1:] + "A"                 # shifted
            elif u < 0.12:
                f1[i] = ""
            if rng.random() < 0.03 and len(reads[i]) > 20:
                k = int(rng.integers(2, 6))
                pos = rng.integers(0, len(reads[i]), k)
                r = list(reads[i])
                for p in pos:
                    r[int(p)] = "N"
                reads[i] = "".join(r)
            if rng.random() < 0.02 and len(reads[i]) > 80:
                p = int(rng.integers(30, len(reads[i]) - 40))
                reads[i] = reads[i][:p] + reads[i][p:p + 10].lower() + reads[i][p + 10:]
        return reads, f0, f1

    donor = next(r for r in cfg.reads_as_strings(cfg.generate(0, 200)) if len(r) >= 140)[40:52]
    cases["ragged_hr_donor"] = case_from_synth(cfg, 4000, hr_donor=donor, mutate=mutate)

    # TruSeq: 6-nt indices at both read ends, threshold 0, stored sequence seq[6:-6]
    base = synth.make_config("C1", seed=16, min_len=110)
    b = case_from_synth(base, 1500)
    six = ["ACGTAC", "TTGACA", "GGATCC", "CATCAT"]
    six2 = ["TGCATG", "AACCGG", "CTAGCT", "GTGTGT"]
    b["platform"] = "TruSeq"
    b["sample_ids"] = ["T{}".format(i) for i in range(4)]
    b["index_col2"], b["index_col3"] = six, six2
    b["sample_target"] = [0, 0, 0, 0]
    reads = []
    for i, r in enumerate(b["reads"]):
        s = int(rng.integers(4))
        pre, suf = six[s], six2[s]
        if rng.random() < 0.1:
            pre = pre[:2] + "T" + pre[3:]
        reads.append(pre + r + suf)
    b["reads"] = reads
    cases["truseq"] = b

    # Ramsden: 26-nt primers, threshold 3, stored sequence rcomp(seq); lower-case letters in the
    # master index are upper-cased by dictionary_build
    base = synth.make_config("C1", seed=17, min_len=120)
    b = case_from_synth(base, 1200)
    fwd = ["ACTTGATCAGTTGGGCTGTTTTGGAG", "ACTTGATCAGTTGGGCTGTTTTGGAG", "TTGATCAGTTGGGCTGTTTTGGAGCA"]
    rev = ["ATTATTGCTTGTGATCCGCCTCAAGT", "ATTGCTTGTGATCCGCtcgCACATCG", "CTTGTGATCCGCCgcacaggATTGGC"]
    b["platform"] = "Ramsden"
    b["sample_ids"] = ["F1+R1", "F1+R2", "F2+R3"]
    b["index_col2"], b["index_col3"] = fwd, rev
    b["sample_target"] = [0, 0, 0]
    reads = []
    for i, r in enumerate(b["reads"]):
        s = int(rng.integers(3))
        pre, suf = list(fwd[s].upper()), list(rev[s].upper())
        for _ in range(int(rng.integers(0, 5))):
            side = pre if rng.random() < 0.5 else suf
            p = int(rng.integers(len(side)))
            side[p] = "ACGT"[int(rng.integers(4))]
        if rng.random() < 0.05:
            del pre[int(rng.integers(len(pre)))]
        body = rc(r)       # stored sequence is rcomp(seq): feed the reverse complement
        reads.append("".join(pre) + body + "".join(suf))
    b["reads"] = reads
    cases["ramsden"] = b
    return cases


def main():
    if not ref_shims.available():
        raise SystemExit("needs the reference tree at " + ref_shims.REFERENCE)
    only = set(sys.argv[1:])
    cases = build_cases()
    # --Demultiplex True on two small cases: the same inputs, the per-sample FASTQ files digested as well
    cases["demux_c2_96_samples"] = dict(cases["c2_96_samples"], demultiplex=True)
    cases["demux_truseq"] = dict(cases["truseq"], demultiplex=True)
    for name, case in cases.items():
        if only and name not in only:
            continue
        out = run_reference(case, demultiplex=bool(case.get("demultiplex")))
        blob = {"case": case, "golden": out}
        path = os.path.join(HERE, name + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(blob, sort_keys=True).encode())
        ns = len(out["samples"])
        print("{:24s} reads {:6d} samples-with-reads {:3d} unique keys {:6d} -> {} ({} KB)".format(
            name, len(case["reads"]), ns, sum(len(v.get("table", [])) for v in out["samples"].values()),
            os.path.basename(path), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
