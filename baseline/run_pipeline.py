"""
run_pipeline.py — the whole ScarMapper IndelProcessing pipeline on synthetic inputs, two ways, for bench.py.

    python baseline/run_pipeline.py --mode reference|dropin --workdir W [--config C2 --reads N --seed S]
                                    [--spawn K] [--repeat R] [--device 0] [--keep-inputs]

  reference : the UNMODIFIED reference driver, exactly as scarmapper.py:261-273 starts it:
              FASTQ_Reader -> TargetMapper -> DataProcessing(...).main_loop() (INDEL_Processing.py:1038-1061:
              serial consensus_demultiplex, then pathos Pool(Spawn).starmap(ScarSearch), then data_output), with
              the reference's own compiled SlidingWindow.pyx.  This is bench.py's `--impl reference` arm and its
              cpu_baseline.
  dropin    : the same driver started by scarmapper_b200.launcher.main with the CUDA hot path installed
              (scarmapper_b200.dropin) - what a user of this repository runs.

Both read the same gzip FASTQ / manifest / target / index files (written here from the `synth` generator when
they are not already in the work directory) and write the reference's output files into W/out_<mode>/.
The reference tree is /root/reference or its byte-for-byte copy baseline/_ref/ScarMapper; its absent third-party
imports are the stand-ins of baseline/shims (README there).  One JSON line per repeat on stdout.
"""
from __future__ import annotations

import argparse
import datetime
import gzip
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def write_inputs(wd, cfg, n_reads, first=0):
    """targets.tsv, refseq.fa, master_index.tsv, manifest.tsv, reads.fastq.gz in the formats the reference parses
    (SURVEY.md Appendix A), from a synth.SynthConfig."""
    with open(os.path.join(wd, "targets.tsv"), "w") as f:
        f.write("# TargetName\tChr\tStart\tStop\tsgRNA_Seq\tReverseComp\tComments\n")
        for nm, t, sg, rc in zip(cfg.target_names, cfg.targets, cfg.sgrnas, cfg.reverse_comp):
            f.write("{}\t{}\t0\t{}\t{}\t{}\t\n".format(nm, "ctg_" + nm, len(t), sg, "Yes" if rc else "No"))
    with open(os.path.join(wd, "refseq.fa"), "w") as f:
        for nm, t in zip(cfg.target_names, cfg.targets):
            f.write(">ctg_{}\n{}\n".format(nm, t))
    with open(os.path.join(wd, "master_index.tsv"), "w") as f:
        f.write("# Index_ID\tForward\tReverse\n")
        for iid, a, b in zip(cfg.sample_ids, cfg.index_d7, cfg.index_d5):
            f.write("{}\t{}\t{}\n".format(iid, a, b))
    with open(os.path.join(wd, "manifest.tsv"), "w") as f:
        f.write("# Index\tSample Name\tReplicate\tSpecies\tForward Phasing\tReverse Phasing\tTarget Name\n")
        for s, iid in enumerate(cfg.sample_ids):
            t = cfg.sample_target[s]
            fwd, rev = cfg.phasing_seqs[t]
            f.write("{}\tS{}\t1\tsyn\t{}\t{}\t{}\n".format(iid, s, fwd, rev, cfg.target_names[t]))
    path = os.path.join(wd, "reads.fastq.gz")
    with gzip.open(path, "wb", compresslevel=1) as f:
        step = 200_000
        for a in range(0, n_reads, step):
            m = min(step, n_reads - a)
            batch = cfg.generate(first + a, m)
            reads = cfg.reads_as_strings(batch)
            f0, f1 = cfg.fields_as_strings(batch)
            out = []
            for i, (r, x, y) in enumerate(zip(reads, f0, f1), a):
                out.append("@SYN:1:FC:1:{}:{}:{} 1:N:0:{}+{}\n{}\n+\n{}\n".format(i % 97, i, i * 7 % 1013, x, y, r, "I" * len(r)))
            f.write("".join(out).encode())
    return path


def write_options(wd, out_dir, spawn, cfg, raw=False, verbose="INFO"):
    """Tab-delimited options file, docs/run_ScarMapper_IndelProcessing.sh:8-42 format."""
    opts = {
        "IndelProcessing": "True", "FASTQ1": os.path.join(wd, "reads.fastq.gz"), "FASTQ2": "",
        "RefSeq": os.path.join(wd, "refseq.fa"), "Master_Index_File": os.path.join(wd, "master_index.tsv"),
        "SampleManifest": os.path.join(wd, "manifest.tsv"), "TargetFile": os.path.join(wd, "targets.tsv"),
        "WorkingFolder": out_dir, "Verbose": verbose, "Job_Name": "bench", "Spawn": str(spawn), "Demultiplex": "False",
        "DeleteConsensusFASTQ": "True", "HR_Donor": "", "Platform": cfg.platform, "Search_KMER": "10",
        "N_Limit": str(cfg.n_limit), "Minimum_Length": str(cfg.min_length), "OutputRawData": "True" if raw else "False",
        "PatternThreshold": "0.0001", "FigureType": "pdf",
    }
    path = os.path.join(out_dir, "run.sh")
    with open(path, "w") as f:
        f.write("#!/bin/bash\n#Parameter file\n\npython3 scarmapper.py --options_file run.sh\nexit\n\n")
        for k, v in opts.items():
            f.write("--{}\t{}\n".format(k, v))
    return path


def install_shims():
    from oracle import ref_shims          # locates the reference tree, baseline/shims and the compiled SlidingWindow
    ref_shims.install(real_pool=True)
    return ref_shims.REFERENCE


def run_reference(options_file):
    """scarmapper.py:221-286 for IndelProcessing without scarmapper.py's Python-3.7 import hack and version gate."""
    reference = install_shims()
    from scarmapper import INDEL_Processing, TargetMapper
    from Valkyries import FASTQ_Tools, Tool_Box
    from scarmapper_b200 import launcher                      # only its options-file parser (no CUDA)
    args = launcher.parse_options(Tool_Box, ["--options_file", options_file])
    run_start = datetime.datetime.today().strftime("%H:%M:%S %Y  %a %b %d")
    log = Tool_Box.Logger(args)
    stamps = {}
    orig = INDEL_Processing.DataProcessing.consensus_demultiplex

    def timed_demux(self):                                    # a stopwatch around the reference's own method
        t = time.perf_counter()
        try:
            return orig(self)
        finally:
            stamps["demux_s"] = time.perf_counter() - t
    INDEL_Processing.DataProcessing.consensus_demultiplex = timed_demux
    try:
        t0 = time.perf_counter()
        fq1 = FASTQ_Tools.FASTQ_Reader(args.FASTQ1, log)
        sample_manifest = Tool_Box.FileParser.indices(log, args.SampleManifest)
        dp = INDEL_Processing.DataProcessing(log, args, run_start, INDEL_Processing.__version__,
                                             TargetMapper.TargetMapper(log, args, sample_manifest), fq1=fq1)
        dp.main_loop()
        total = time.perf_counter() - t0
    finally:
        INDEL_Processing.DataProcessing.consensus_demultiplex = orig
    return {"mode": "reference", "reads": int(dp.read_count), "total_s": total, "demux_s": stamps.get("demux_s"),
            "search_s": total - stamps.get("demux_s", 0.0), "spawn": int(args.Spawn), "reference": reference}


def run_dropin(options_file, device):
    reference = install_shims()
    from scarmapper_b200 import dropin, launcher
    t0 = time.perf_counter()
    launcher.main(["--options_file", options_file, "--reference", reference, "--device", str(device)])
    total = time.perf_counter() - t0
    return {"mode": "dropin", "total_s": total, "reference": reference, "stages": dict(dropin.STAGE_SECONDS),
            "pattern_searches": dict(dropin.PATTERN_SEARCH_STATS)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["reference", "dropin"])
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--spawn", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--raw", action="store_true", help="--OutputRawData True")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    wd = os.path.abspath(a.workdir)
    os.makedirs(wd, exist_ok=True)
    import synth
    cfg = synth.make_config(a.config, n_reads=a.reads, seed=a.seed)
    marker = os.path.join(wd, "inputs.json")
    want = {"config": a.config, "reads": a.reads, "seed": a.seed}
    if not os.path.exists(marker) or json.load(open(marker)) != want:
        write_inputs(wd, cfg, a.reads)
        json.dump(want, open(marker, "w"))
    out_dir = os.path.join(wd, "out_" + a.mode) + os.sep
    for k in range(a.repeat):
        shutil.rmtree(out_dir, ignore_errors=True)
        os.makedirs(out_dir)
        opt = write_options(wd, out_dir, a.spawn, cfg, raw=a.raw, verbose="ERROR" if a.quiet else "INFO")
        res = run_reference(opt) if a.mode == "reference" else run_dropin(opt, a.device)
        res.update(config=a.config, reads=a.reads, repeat=k, cores=os.cpu_count())
        print("PIPELINE " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
