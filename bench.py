"""
bench.py — merged reads classified per second on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] — synthetic PEAR-merged 150-bp amplicon reads
demultiplexed across the 96 Illumina dual-index samples, Rosa26a-shaped 472-bp locus, 10 M reads per
GPU (weak scaling: rank g owns global reads [g*10M, (g+1)*10M)).

A step is one pass of the hot path over the rank's batch:
  value : inputs (packed rows, raw bytes, index fields) already resident in HBM; the timed region is
          reset -> K2 demux -> K3 sliding window -> K4 aggregation (-> table merge across ranks when
          N > 1), CUDA events on the launching stream, barrier + synchronize on both sides, max over
          ranks.  K1 (pack) is timed separately (stages_ms.k1_pack), as SURVEY.md §8d defines;
          from_raw_bytes is the same step INCLUDING K1 (scar_run_dev on raw read bytes in HBM).
  e2e   : the same batch from pinned HOST buffers through the public API (ScarEngine.process_host ->
          scar_process_host): H2D of reads + index fields, K2, K1, K3, K4, D2H of the 16-byte records.
  roofline : K3 (scar_window_kernel): algorithmic bytes = 182 B/read (150 read + 16 index + 16 record,
          SURVEY.md §8d) x reads per launch / mean K3 launch time, against the measured HBM copy peak.
  cpu_baseline : the reference's CPU path (oracle/cpu_reference.py) on a bounded sample, N=1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_READ = 182          # 150 (1 B/base) + 16 (i7+i5) + 16 (result record)
METRIC = "merged_reads_classified_per_sec"
WORKLOAD = ("C2: synthetic PEAR-merged 150-bp amplicon reads, 96 Illumina dual-index samples "
            "(Levenshtein index matching + per-sample scar tables), Rosa26a-shaped 472-bp locus")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=60_000, help="reads of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-reads", type=int, default=1_000_000, help="reads of the launcher-only pipeline run")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4/C5-shape lines")
    ap.add_argument("--kernels-only", action="store_true",
                    help="tuning runs: skip the end-to-end arms (the line then carries e2e.value = null)")
    ap.add_argument("--seed", type=int, default=20260101)
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(gpu_index):
    """One rank per GPU: run on (and first-touch the pinned host buffers from) the CPUs NVML names as local to this GPU, so
    that the H2D copies of the end-to-end arm do not all cross one socket.  Returns the CPU list, or None when NVML cannot
    tell / the process may not change its affinity (a no-op on boxes whose GPUs all hang off one NUMA node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(cpus) & allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
        return cpus or None
    except Exception:
        return None


def build_case(args, rank, world):
    import numpy as np
    import synth
    cfg = synth.make_config("C2", n_reads=args.reads * world, seed=args.seed)
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target,
               left_index=cfg.index_d5, right_index=cfg.index_d7, platform="Illumina",
               phasing=cfg.phasing_lists(), n_limit=cfg.n_limit, min_length=cfg.min_length)
    return cfg, pin


def cpu_sample_inputs(cfg, n):
    batch = cfg.generate(0, n)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    names = ["SYN:1:FC:1:{}:{}:{} 1:N:0:{}+{}".format(i % 97, i, i * 7 % 1013, a, b)
             for i, (a, b) in enumerate(zip(f0, f1))]
    return names, reads


def run_cpu_reference(cfg, n_sample, cores=None, repeats=1, warm=0):
    from oracle import cpu_reference as cr
    names, reads = cpu_sample_inputs(cfg, n_sample)
    ref = cr.CpuReference(cores)
    try:
        kw = dict(sample_ids=cfg.sample_ids, left_index=cfg.index_d5, right_index=cfg.index_d7,
                  sample_target=cfg.sample_target, targets=cfg.targets, cutsites=cfg.cutsites,
                  platform=0, mismatch=1, n_limit=cfg.n_limit, min_length=cfg.min_length)
        for _ in range(warm):
            ref.run(names, reads, **kw)
        runs = [ref.run(names, reads, **kw) for _ in range(repeats)]
    finally:
        cores_used = ref.cores
        ref.close()
    return runs, cores_used, cr.baseline_kind()


def pipeline_available():
    from oracle import ref_shims
    return ref_shims.available()


def run_pipeline(mode, wd, config, reads, seed, spawn, repeat=1, device=0, timeout=1500):
    """baseline/run_pipeline.py in a fresh process: the whole IndelProcessing pipeline (gzip FASTQ -> output files),
    mode 'reference' = the UNMODIFIED reference driver (DataProcessing.main_loop), 'dropin' = the same driver started
    by scarmapper_b200.launcher with the CUDA hot path.  Returns one dict per repeat."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_pipeline.py"), "--mode", mode, "--workdir", wd,
           "--config", config, "--reads", str(reads), "--seed", str(seed), "--spawn", str(spawn), "--repeat", str(repeat),
           "--device", str(device), "--quiet"]
    t0 = time.perf_counter()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    lines = [json.loads(ln[len("PIPELINE "):]) for ln in proc.stdout.splitlines() if ln.startswith("PIPELINE ")]
    if proc.returncode != 0 or len(lines) != repeat:
        raise RuntimeError("run_pipeline {} failed ({}): {}".format(mode, proc.returncode, proc.stderr[-1500:]))
    for ln in lines:
        ln["process_s"] = wall
    return lines


def compare_pipeline_outputs(wd):
    """Frequency / SNV_Frequency / Summary files of out_reference and out_dropin, wall-clock lines excluded."""
    skip = ("# Run Start", "# Run End", "Start:", "End:", "FASTQ1:", "FASTQ2:")
    a, b = os.path.join(wd, "out_reference"), os.path.join(wd, "out_dropin")
    keep = lambda d: sorted(f for f in os.listdir(d) if f.endswith(("_Frequency.txt", "_Summary.txt")))
    fa, fb = keep(a), keep(b)
    same = 0
    for fn in fa:
        if fn not in fb:
            continue
        x = [ln for ln in open(os.path.join(a, fn)).read().split("\n") if not ln.startswith(skip)]
        y = [ln for ln in open(os.path.join(b, fn)).read().split("\n") if not ln.startswith(skip)]
        same += x == y
    return {"reference_files": len(fa), "dropin_files": len(fb), "identical": same,
            "all_identical": same == len(fa) == len(fb) and same > 0}


def reference_spawn(cfg):
    return max(1, min(cfg.n_samples, (os.cpu_count() or 2) - 1))   # one library per process, Spawn <= CPUs - 1


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores - the UNMODIFIED
    DataProcessing.main_loop() (scarmapper/INDEL_Processing.py:1038-1061: serial consensus_demultiplex, then one
    pathos process per sample library with the compiled SlidingWindow.pyx, then data_output) on a bounded gzip-FASTQ
    sample of the workload; restated driver loops (oracle/cpu_reference.py) only where no reference tree is present."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile
    import synth  # noqa: F401
    cfg, _ = build_case(args, 0, 1)
    n_runs = args.steps + args.warmup
    if pipeline_available():
        spawn = reference_spawn(cfg)
        wd = tempfile.mkdtemp(prefix="scar_ref_arm_")
        try:
            cal = run_pipeline("reference", os.path.join(wd, "cal"), "C2", 8000, args.seed, spawn)[0]
            rate = cal["reads"] / cal["total_s"]
            budget = min(25.0, 170.0 / max(1, n_runs))
            n_sample = int(max(8000, min(300_000, rate * budget)) // 1000 * 1000)
            runs = run_pipeline("reference", os.path.join(wd, "run"), "C2", n_sample, args.seed, spawn, repeat=n_runs)
        finally:
            import shutil
            shutil.rmtree(wd, ignore_errors=True)
        runs = runs[args.warmup:]
        total = sum(r["total_s"] for r in runs)
        value = n_sample * len(runs) / total
        demux = n_sample * len(runs) / sum(r["demux_s"] for r in runs)
        search = n_sample * len(runs) / sum(r["search_s"] for r in runs)
        cores, kind = spawn, "reference"
        sample = ("{} reads of the workload (gzip FASTQ) per step through the UNMODIFIED DataProcessing.main_loop(): "
                  "consensus_demultiplex serial in the main process ({:.0f} reads/s), then Pool({}).starmap(ScarSearch) "
                  "with the compiled reference SlidingWindow.pyx + frequency_output + data_output ({:.0f} reads/s); "
                  "python-Levenshtein / pathos / natsort / pysam stand-ins from baseline/shims").format(
                      n_sample, demux, spawn, search)
    else:
        cal, cores, kind = run_cpu_reference(cfg, 12_000)
        rate = cal[0]["n_reads"] / cal[0]["total_s"]
        budget = min(20.0, 150.0 / max(1, n_runs))
        n_sample = int(max(10_000, min(200_000, rate * budget)))
        runs, cores, kind = run_cpu_reference(cfg, n_sample, repeats=args.steps, warm=args.warmup)
        total = sum(r["total_s"] for r in runs)
        value = n_sample * len(runs) / total
        sample = "{} reads per step, restated driver loops (no reference tree on this box), {} workers".format(n_sample, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(runs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reads_per_step": n_sample, "read_len": 150, "n_samples": cfg.n_samples},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_config(name, n, seed, device, steps=5, chk=4000):
    """One BASELINE.json config on one GPU: parity of the first `chk` reads against the oracle, then the resident-batch
    step (reset, K2, K1, K3, K4 from raw bytes in HBM) timed with CUDA events."""
    import numpy as np
    import torch
    import synth
    from oracle import scar_oracle as orc
    from scarmapper_b200.engine import HostRagged, ScarEngine, ScarProblem
    cfg = synth.make_config(name, n_reads=n, seed=seed)
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target,
               left_index=cfg.index_d5, right_index=cfg.index_d7, platform="Illumina",
               phasing=cfg.phasing_lists(), n_limit=cfg.n_limit, min_length=cfg.min_length)
    log2 = 22
    while (1 << log2) < n // 2:
        log2 += 1
    eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << log2, **pin), device=device)
    try:
        batch = cfg.generate(0, n)
        sub = {k: (v[:chk] if v is not None else None) for k, v in batch.items()}
        want = orc.classify_batch(orc.Problem(**pin), cfg.reads_as_strings(sub), *cfg.fields_as_strings(sub))
        got = eng.process_host(HostRagged.from_matrix(batch["seq"][:chk], lengths=batch["len"][:chk]), batch["f0"][:chk], batch["f1"][:chk])
        parity = got.tobytes() == want.tobytes()
        eng.reset()
        res = eng.make_resident(HostRagged.from_matrix(batch["seq"], lengths=batch["len"]), HostRagged.from_matrix(batch["f0"]),
                                HostRagged.from_matrix(batch["f1"]))
        stream = torch.cuda.current_stream()

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e
        for _ in range(2):
            eng.reset_async()
            eng.run_resident(res, first_index=0)
        torch.cuda.synchronize()
        info = eng.info()
        marks = []
        t0 = ev()
        for _ in range(steps):
            eng.reset_async()
            m = [ev()]
            eng.demux_resident(res); m.append(ev())
            if info["fused"]:
                eng.classify_raw_resident(res, first_index=0); m.append(ev())
            else:
                eng.pack_resident(res); m.append(ev())
                eng.classify_resident(res); m.append(ev())
                eng.aggregate_resident(res, first_index=0); m.append(ev())
            marks.append(m)
        t1 = ev()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
        st = [sum(m[i].elapsed_time(m[i + 1]) for m in marks) / steps for i in range(len(marks[0]) - 1)]
        n_entries = eng.table_size()
        cnt, un = eng.counters()
        algo = cfg.read_len + 16 + 16
        peak, _ = measured_peak()
        return {"workload": "{}: {} reads x {} nt, {} samples, {} targets".format(name, n, cfg.read_len, cfg.n_samples, len(cfg.targets)),
                "parity_first_{}_reads_vs_oracle".format(chk): bool(parity),
                "value": n / (ms * 1e-3), "unit": "reads/s", "ms_per_step": ms, "steps": steps,
                "read_bases_per_s": n * cfg.read_len / (ms * 1e-3),
                "timed_region": "from raw read bytes resident in HBM: reset, K2 demux, " +
                                ("fused kernel (K1 + K3 + K4)" if info["fused"] else "K1 pack, K3, K4"),
                "stages_ms": ({"k2_demux": st[0], "fused_pack_window_aggregate": st[1]} if info["fused"] else
                              {"k2_demux": st[0], "k1_pack": st[1], "k3_window": st[2], "k4_aggregate": st[3]}),
                "fused": info["fused"], "fused_ctas_per_sm": info["fused_ctas_per_sm"], "fused_smem_bytes": info["fused_smem_bytes"],
                "algorithmic_bytes_per_read": algo, "hbm_frac_of_step": algo * n / (ms * 1e-3) / 1e9 / peak,
                "cells_per_read": eng.cells_per_read, "k3_cells_staged_in_smem": bool(info.get("k3_stage", False)),
                "k3_ctas_per_sm": info.get("k3_ctas_per_sm"), "tables_in_smem": info["tables_in_smem"],
                "static_blob_bytes": info["static_bytes"], "scar_patterns": int(n_entries), "unassigned_reads": int(un),
                "scar_reads": int(cnt[:, 12].sum())}
    finally:
        eng.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from scarmapper_b200.engine import HostRagged, ScarEngine, ScarProblem
    from scarmapper_b200.distributed import DistributedScar
    from scarmapper_b200._native import RECORD_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE {} != --gpus {}".format(world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # SCAR_BENCH_MERGE_PRIORITY=high: the merge's side stream and NCCL's stream get the high CUDA priority, so their
        # (small, latency-bound) kernels are dispatched beside the next step's K2 instead of queueing behind it
        pg_opts = None
        if os.environ.get("SCAR_BENCH_MERGE_PRIORITY", "normal") == "high":
            pg_opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=pg_opts)

    # the CPU baseline runs first (rank 0, N=1 only), before this process touches the GPU
    cpu_baseline = None
    cfg, pin = build_case(args, rank, world)
    pipeline, pipe_wd = None, None
    if world == 1 and not args.no_cpu_baseline:
        if pipeline_available():
            import tempfile
            pipe_wd = tempfile.mkdtemp(prefix="scar_pipeline_")
            spawn = reference_spawn(cfg)
            r = run_pipeline("reference", pipe_wd, "C2", args.cpu_sample, args.seed, spawn)[0]
            cpu_baseline = {
                "value": r["reads"] / r["total_s"], "unit": "reads/s", "cores": spawn, "kind": "reference",
                "sample": "first {} reads of the workload as gzip FASTQ through the UNMODIFIED DataProcessing.main_loop(): "
                          "consensus_demultiplex (serial) {:.1f} s, Pool({}).starmap(ScarSearch) + data_output {:.1f} s"
                          .format(r["reads"], r["demux_s"], spawn, r["search_s"]),
            }
            pipeline = {"reads": r["reads"], "fastq": "gzip", "spawn": spawn, "reference_s": r["total_s"],
                        "reference_demux_s": r["demux_s"], "reference_search_s": r["search_s"]}
        else:
            runs, cores, kind = run_cpu_reference(cfg, args.cpu_sample)
            r = runs[0]
            cpu_baseline = {
                "value": r["n_reads"] / r["total_s"], "unit": "reads/s", "cores": cores, "kind": kind,
                "sample": "first {} reads of the workload: stage 1 serial {:.1f} s, stages 2-4 on {} processes {:.1f} s "
                          "(restated driver loops: no reference tree on this box)"
                          .format(r["n_reads"], r["demux_s"], cores, r["search_s"]),
            }

    n = args.reads
    first = rank * n
    batch = cfg.generate(first, n)
    # scar-table slots: ~1 M patterns per 10 M reads of this spectrum; after the owner-partitioned merge a rank
    # holds ~1/world of the merged patterns, so the load factor stays near 0.25 for every world size
    table_capacity = 1 << int(os.environ.get("SCAR_BENCH_TABLE_LOG2", "22"))
    eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=table_capacity, **pin), device=local)

    # pinned host buffers for the end-to-end arm
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = a
        return t
    p_seq, p_f0, p_f1 = pinned(batch["seq"]), pinned(batch["f0"]), pinned(batch["f1"])
    h_seq = HostRagged.from_matrix(p_seq.numpy())
    h_f0 = HostRagged.from_matrix(p_f0.numpy())
    h_f1 = HostRagged.from_matrix(p_f1.numpy())
    p_rec = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    rec_out = p_rec.numpy().view(RECORD_DTYPE).reshape(-1)

    # small parity check against the oracle before any timing (outside every timed region)
    from oracle import scar_oracle as orc
    chk = 4000
    sub = {k: (v[:chk] if v is not None else None) for k, v in batch.items()}
    want = orc.classify_batch(orc.Problem(**pin), cfg.reads_as_strings(sub), *cfg.fields_as_strings(sub))
    got = eng.process_host(HostRagged.from_matrix(batch["seq"][:chk]), batch["f0"][:chk], batch["f1"][:chk])
    if got.tobytes() != want.tobytes():
        raise SystemExit("bench: CUDA records differ from the oracle on the first {} reads".format(chk))
    eng.reset()

    res = eng.make_resident(h_seq, h_f0, h_f1)
    dscar = DistributedScar(eng, entry_capacity=1 << 21) if world > 1 else None
    stream = torch.cuda.current_stream()
    # N > 1: the table merge of step k (export, all_gather over NVLink, fold) runs on a side stream while
    # step k+1 computes into a second context, so two contexts (tables + accumulators) alternate.
    engines = [eng]
    dscars = [dscar]
    if world > 1:
        eng_b = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=table_capacity, **pin), device=local)
        engines.append(eng_b)
        dscars.append(DistributedScar(eng_b, entry_capacity=1 << 21))
    merge_priority = os.environ.get("SCAR_BENCH_MERGE_PRIORITY", "normal")
    merge_order = os.environ.get("SCAR_BENCH_MERGE_ORDER", "under_next_step")      # or "beside_k2"
    side = torch.cuda.Stream(priority=-1 if merge_priority == "high" else 0) if world > 1 else None
    # owner-partitioned merge: a rank sends each peer the entries that peer owns.  The exchange moves whole
    # regions, so they are sized from the table this batch actually produces (one untimed pass): the largest
    # per-rank entry count / world (owners are hash-uniform) + 25 % + 4096, in multiples of 4096 entries.
    region_capacity = 4096
    if world > 1:
        eng.reset_async()
        eng.run_resident(res, first_index=first)
        torch.cuda.synchronize()
        t = torch.tensor([eng.table_size()], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_capacity = int(-(-(int(t.item()) * 1.25 / world + 4096) // 4096) * 4096)
        eng.reset()
    merge_done = [None, None]
    step_no = [0]
    # N > 1: the persistent grids leave a few SMs to the merge kernels and NCCL, which then really run beside the
    # next step instead of waiting for a CTA slot (SCAR_BENCH_SM_RESERVE; measured at N=2: 0, 4 and 8 reserved SMs give
    # 1.138e10, 1.135e10 and 1.115e10 reads/s, so the default is 0)
    sm_reserve = 0
    if world > 1:
        sm_reserve = int(os.environ.get("SCAR_BENCH_SM_RESERVE", "0"))
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        for e in engines:
            e.set_sm_limit(sms - sm_reserve)

    # N > 1, untimed: the union of the ranks' merged partitions of a small sharded batch must be the table one GPU
    # builds from the whole batch (entries, counts, first-seen indices, counters)
    merge_check = None
    if world > 1:
        per = 50_000
        chk_batch = cfg.generate(rank * per, per)
        e_chk = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 20, **pin), device=local)
        e_chk.process_host(HostRagged.from_matrix(chk_batch["seq"]), chk_batch["f0"], chk_batch["f1"], first_index=rank * per,
                           want_records=False)
        DistributedScar(e_chk).merge_partitioned(region_capacity=1 << 16)
        torch.cuda.synchronize()
        mine = e_chk.table()
        cnt_m, un_m = e_chk.counters()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        e_chk.close()
        if rank == 0:
            whole = cfg.generate(0, per * world)
            e_one = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 20, **pin), device=local)
            e_one.process_host(HostRagged.from_matrix(whole["seq"]), whole["f0"], whole["f1"], want_records=False)
            want_tab = e_one.table()
            cnt_1, un_1 = e_one.counters()
            e_one.close()
            union = np.concatenate(parts)
            union = union[np.lexsort((union["first_idx"], union["sample"]))]
            merge_check = {"reads": per * world, "entries": int(len(want_tab)),
                           "union_of_partitions_equals_single_gpu_table": bool(union.tobytes() == want_tab.tobytes()),
                           "counters_equal": bool((cnt_m == cnt_1).all()) and int(un_m) == int(un_1)}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    fused = eng.info()["fused"]

    # the three-kernel path (K1 pack -> K3 -> K4 through HBM), each kernel timed on its own: what the fused kernel replaces
    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a = ev()
        for _ in range(reps):
            fn()
        b = ev()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    eng.demux_resident(res)
    k1_ms = timed(lambda: eng.pack_resident(res))
    k3_alone_ms = timed(lambda: eng.classify_resident(res))
    k4_alone_ms = timed(lambda: eng.aggregate_resident(res, first_index=first))
    eng.reset()

    marks = []

    def step(record_marks):
        k = step_no[0] % len(engines)
        step_no[0] += 1
        e = engines[k]
        if merge_done[k] is not None:
            stream.wait_event(merge_done[k])          # this context's previous merge has drained
        e.reset_async()
        m0 = ev() if record_marks else None
        e.demux_resident(res)
        m1 = ev() if record_marks else None
        if world > 1 and merge_order == "beside_k2":
            # the previous step's merge goes to the side stream NOW and this step's fused kernel waits for it: the merge's
            # small kernels and the all-to-all run beside reset + K2 (which leave SM resources free) instead of queueing
            # behind a persistent kernel that holds every SM for 1.4 ms
            done = run_pending_merge(record_marks)
            if done is not None:
                stream.wait_event(done)
        if fused:
            e.classify_raw_resident(res, first_index=first)       # K1 + K3 + K4 in one pass over the raw bytes
            m2 = m3 = ev() if record_marks else None
        else:
            e.pack_resident(res)
            e.classify_resident(res)
            m2 = ev() if record_marks else None
            e.aggregate_resident(res, first_index=first)
            m3 = ev() if record_marks else None
        if world > 1:
            computed = torch.cuda.Event()
            computed.record(stream)
            # software pipeline on the host as well: this step's kernels are queued, NOW run the previous
            # step's merge (its small D2H syncs block the host while the GPU is busy with this step)
            if merge_order != "beside_k2":
                run_pending_merge(record_marks)
            pending.append((k, computed))
        if record_marks:
            marks.append((m0, m1, m2, m3))

    pending = []

    def run_pending_merge(record_marks):
        if not pending:
            return None
        k, computed = pending.pop(0)
        with torch.cuda.stream(side):
            side.wait_event(computed)
            ms = torch.cuda.Event(enable_timing=True)
            ms.record(side)
            dscars[k].merge_partitioned(region_capacity=region_capacity)
            me = torch.cuda.Event(enable_timing=True)
            me.record(side)
        merge_done[k] = me
        if record_marks:
            merge_marks.append((ms, me))
        return me

    merge_marks = []

    def drain(record_marks=False):
        if side is not None:
            run_pending_merge(record_marks)
            stream.wait_stream(side)                  # the last merge is part of the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)
    drain()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sum(e.launch_count() for e in engines)
    t_start = ev()
    for _ in range(args.steps):
        step(True)
    drain(True)
    t_end = ev()
    barrier()
    launches = sum(e.launch_count() for e in engines) - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    k2 = sum(m[0].elapsed_time(m[1]) for m in marks) / len(marks)
    k3 = sum(m[1].elapsed_time(m[2]) for m in marks) / len(marks)      # fused: K1 + K3 + K4 in one kernel
    k4 = sum(m[2].elapsed_time(m[3]) for m in marks) / len(marks)
    merge_ms = sum(a.elapsed_time(b) for a, b in merge_marks) / len(merge_marks) if merge_marks else 0.0
    last = engines[(step_no[0] - 1) % len(engines)]
    n_entries = last.table_size()                     # N > 1: the entries this rank owns after the merge
    cnt, un = last.counters()
    if world > 1:
        t = torch.tensor([n_entries], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_entries = int(t.item())

    # the same step from RAW read bytes resident in HBM (scar_run_dev: K2, K1 pack, K3, K4), single stream
    raw_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        eng.reset_async()
        eng.run_resident(res, first_index=first)
    barrier()
    r0 = ev()
    for _ in range(raw_steps):
        eng.reset_async()
        eng.run_resident(res, first_index=first)
    r1 = ev()
    barrier()
    raw_ms = r0.elapsed_time(r1) / raw_steps

    # N > 1: what a real run does - every rank streams its batches into ONE table and the ranks merge ONCE at the end
    # (`value` above merges after every batch of 10 M reads, the worst case for the merge)
    job_ms = 0.0
    if world > 1:
        barrier()
        j0 = ev()
        eng.reset_async()
        for k in range(raw_steps):
            eng.run_resident(res, first_index=first + k * world * n)
        dscar.merge_partitioned(region_capacity=region_capacity)
        j1 = ev()
        barrier()
        job_ms = j0.elapsed_time(j1)
        eng.reset()

    # end-to-end arm: host buffers -> records on the host, through the public API
    e2e_steps = 0 if args.kernels_only else max(3, min(args.steps, 10))
    for _ in range(0 if args.kernels_only else 2):
        eng.reset_async()
        eng.process_host(h_seq, h_f0, h_f1, first_index=first, out=rec_out)
        if dscar:
            dscar.merge_partitioned(region_capacity=region_capacity)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.reset_async()
        eng.process_host(h_seq, h_f0, h_f1, first_index=first, out=rec_out)
        if dscar:
            dscar.merge_partitioned(region_capacity=region_capacity)
    barrier()
    e2e_s = time.perf_counter() - t0
    # the same from FASTQ TEXT (uncompressed, pinned): one H2D copy of the text, record indexing on the GPU
    fastq_e2e = None
    if world == 1 and not args.kernels_only:
        head = b"@SYN:1:FC:1:0:0:0 1:N:0:"
        rl = cfg.read_len
        rec_bytes = len(head) + 8 + 1 + 8 + 1 + rl + 3 + rl + 1
        p_txt = torch.empty((n, rec_bytes), dtype=torch.uint8, pin_memory=True)
        tn = p_txt.numpy()
        o = 0
        tn[:, o:o + len(head)] = np.frombuffer(head, dtype=np.uint8); o += len(head)
        tn[:, o:o + 8] = batch["f0"]; o += 8
        tn[:, o] = ord("+"); o += 1
        tn[:, o:o + 8] = batch["f1"]; o += 8
        tn[:, o] = 10; o += 1
        tn[:, o:o + rl] = batch["seq"]; o += rl
        tn[:, o:o + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8); o += 3
        tn[:, o:o + rl] = ord("I"); o += rl
        tn[:, o] = 10
        flat = tn.reshape(-1)
        eng.reset_async()
        got = eng.process_fastq_text(flat, first_index=first, out=rec_out)
        if len(got) != n or got[:chk].tobytes() != want.tobytes():
            raise SystemExit("bench: FASTQ-text path disagrees with the oracle")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.reset_async()
            eng.process_fastq_text(flat, first_index=first, out=rec_out)
        torch.cuda.synchronize()
        fq_s = time.perf_counter() - t0
        fastq_e2e = {"value": n * 3 / fq_s, "unit": "reads/s", "h2d_bytes_per_step": int(flat.size),
                     "d2h_bytes_per_step": int(n * 16), "steps": 3,
                     "path": "ScarEngine.process_fastq_text -> scar_process_fastq_host (uncompressed FASTQ text, pinned; "
                             "records indexed on the GPU)"}
        # and from GZIP-compressed FASTQ (what a sequencer delivers) held in memory: the host inflates the one stream
        # with all its threads (scarmapper_b200.fastq.gunzip -> scar_gunzip_host, csrc/scar_pgzip.h), the GPU does the
        # rest; 1 M reads, bounded so the bench stays short.  zlib on one core is timed beside it.
        try:
            import zlib
            from scarmapper_b200 import fastq as _fq
            n_gz = min(n, 1_000_000)
            gz_text = flat[:n_gz * rec_bytes]
            co = zlib.compressobj(1, zlib.DEFLATED, 31)
            gz = co.compress(gz_text.tobytes()) + co.flush()
            t0 = time.perf_counter()
            zlib.decompress(gz, 31)
            t_zlib = time.perf_counter() - t0
            gz_threads = min(16, os.cpu_count() or 1)
            _fq.gunzip(gz, threads=gz_threads, capacity=int(gz_text.size))      # (threads started once, pages touched)
            eng.reset_async()
            t0 = time.perf_counter()
            raw = _fq.gunzip(gz, threads=gz_threads, capacity=int(gz_text.size))
            t_inflate = time.perf_counter() - t0
            got_gz = eng.process_fastq_text(raw, first_index=first, out=rec_out)
            gz_s = time.perf_counter() - t0
            if len(got_gz) != n_gz or got_gz[:chk].tobytes() != want.tobytes():
                raise RuntimeError("gz path disagrees with the oracle")
            fastq_e2e["from_gzip"] = {"value": n_gz / gz_s, "unit": "reads/s", "reads": int(n_gz), "gz_bytes": len(gz),
                                      "inflate_s": t_inflate, "total_s": gz_s, "inflate_threads": gz_threads,
                                      "zlib_one_core_inflate_s": t_zlib,
                                      "path": "fastq.gunzip (one stream, speculative parallel inflate, scar_gunzip_host) -> "
                                              "process_fastq_text"}
            del raw, gz, gz_text
        except Exception as exc:   # diagnostic only
            fastq_e2e["from_gzip"] = {"error": str(exc)[:200]}
        # and from FASTQ FILES through scar_process_fastq_file (what the drop-in calls): the file is read / inflated by a
        # producer thread while the GPU indexes and classifies the pieces that have arrived; bounded sizes
        try:
            import shutil
            import struct
            import tempfile
            import zlib
            tmp = tempfile.mkdtemp(prefix="scar_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            files = {}
            try:
                n_txt, n_z = min(n, 2_000_000), min(n, 1_000_000)
                flat[:n_txt * rec_bytes].tofile(os.path.join(tmp, "r.fastq"))
                z_text = flat[:n_z * rec_bytes].tobytes()
                co = zlib.compressobj(1, zlib.DEFLATED, 31)
                with open(os.path.join(tmp, "r.fastq.gz"), "wb") as f:
                    f.write(co.compress(z_text) + co.flush())
                with open(os.path.join(tmp, "r.bgzf.gz"), "wb") as f:       # BGZF as bgzip writes it (64 KB blocks)
                    for a in list(range(0, len(z_text), 65280)) + [None]:
                        chunk = b"" if a is None else z_text[a:a + 65280]
                        co = zlib.compressobj(1, zlib.DEFLATED, -15)
                        payload = co.compress(chunk) + co.flush()
                        f.write(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" +
                                struct.pack("<HH", 2, 12 + 6 + len(payload) + 8 - 1) + payload +
                                struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
                del z_text
                for key, fn, n_f in (("text", "r.fastq", n_txt), ("gzip", "r.fastq.gz", n_z), ("bgzf", "r.bgzf.gz", n_z)):
                    best, run = None, None
                    for _ in range(3):
                        eng.reset_async()
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        run = eng.process_fastq_file(os.path.join(tmp, fn), first_index=first)
                        dt = time.perf_counter() - t0
                        best = dt if best is None else min(best, dt)
                        if run.n != n_f or run.records[:chk].tobytes() != want.tobytes():
                            raise RuntimeError("FASTQ file path ({}) disagrees with the oracle".format(key))
                    files[key] = {"value": n_f / best, "unit": "reads/s", "reads": int(n_f), "seconds": best,
                                  "file_bytes": os.path.getsize(os.path.join(tmp, fn)), "text_bytes": run.text_bytes,
                                  "source": run.source, "threads": run.threads, "producer_wait_s": run.wait_seconds}
                    run.close()
                files["path"] = ("ScarEngine.process_fastq_file -> scar_process_fastq_file: producer thread(s) read / inflate "
                                 "into pinned pieces, the GPU indexes and classifies them as they arrive; best of 3, file in "
                                 + ("/dev/shm" if tmp.startswith("/dev/shm") else "the temp directory"))
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
            fastq_e2e["from_file"] = files
        except Exception as exc:   # diagnostic only
            fastq_e2e["from_file"] = {"error": str(exc)[:300]}
        del p_txt, tn, flat
    # SURVEY 8f F1 (outside the metric): the per-pattern homeology / templated-insertion searches of
    # frequency_output as one batched kernel, timed through the host-buffer C ABI on synthetic patterns
    pattern_searches = None
    if world == 1 and not args.kernels_only:
        try:
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            import homeology_bench as hb
            from scarmapper_b200.engine import homeology_batch
            nk = 200_000
            region, pc, pi, a4, b4, c4, d4 = hb.patterns(nk)
            tix = np.zeros(nk, dtype=np.int32)
            homeology_batch(pc[:1000], pi[:1000], a4[:1000], b4[:1000], c4[:1000], d4[:1000], tix[:1000], [region], [region])
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                homeology_batch(pc, pi, a4, b4, c4, d4, tix, [region], [region])
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            pattern_searches = {"value": nk / best, "unit": "scar patterns/s", "patterns": nk, "best_of": 3,
                                "path": "engine.homeology_batch -> scar_homeology_batch (host buffers; Python string "
                                        "packing, H2D, kernel, D2H)"}
        except Exception as exc:   # diagnostic only: never lose the bench line over it
            pattern_searches = {"error": str(exc)[:200]}
    clocks = sampler.stop() if rank == 0 else None

    # The whole pipeline a user runs (gzip FASTQ -> Frequency / Summary files): the reference's unmodified driver under
    # scarmapper_b200.launcher with the CUDA hot path and a real forking pool, on the SAME files the reference arm
    # above processed on this box's CPUs; the output files of the two runs are diffed.  Then the launcher alone on a
    # larger file.  Wall times of launcher.main (context creation, inflate, GPU, pool, reporting all included).
    if pipeline is not None and not args.kernels_only:
        try:
            # two fresh processes, the faster one reported (both kept): CUDA context creation alone varies between 0.7
            # and 1.4 s from process to process on one box, which is as much as everything else the launcher does
            runs = [run_pipeline("dropin", pipe_wd, "C2", args.cpu_sample, args.seed, pipeline["spawn"], device=local)[0]
                    for _ in range(2)]
            d = min(runs, key=lambda x: x["total_s"])
            pipeline.update(launcher_s=d["total_s"], launcher_runs_s=[x["total_s"] for x in runs], launcher_stages_s=d["stages"],
                            speedup=pipeline["reference_s"] / d["total_s"], outputs=compare_pipeline_outputs(pipe_wd))
            big = run_pipeline("dropin", os.path.join(pipe_wd, "big"), "C2", args.pipeline_reads, args.seed,
                               pipeline["spawn"], device=local)[0]
            pipeline["launcher_large"] = {"reads": args.pipeline_reads, "launcher_s": big["total_s"],
                                          "reads_per_s": args.pipeline_reads / big["total_s"], "stages_s": big["stages"]}
        except Exception as exc:   # diagnostic: never lose the bench line over it
            pipeline["error"] = str(exc)[-600:]
    if pipe_wd:
        import shutil
        shutil.rmtree(pipe_wd, ignore_errors=True)

    # the other BASELINE.json configs (parity-checked against the oracle on a prefix, timed over resident batches)
    configs = None
    if world == 1 and not args.no_configs and not args.kernels_only:
        for e in engines:
            e.close()
        engines = []
        del res
        torch.cuda.empty_cache()
        configs = {}
        for key, name, n_cfg in (("C1", "C1", 1_000_000), ("C3", "C3", 10_000_000), ("C4", "C4", 10_000_000),
                                 ("C5shape", "C2", 25_000_000)):
            try:
                configs[key] = measure_config(name, n_cfg, args.seed, local, steps=max(3, min(args.steps, 5)))
            except Exception as exc:
                configs[key] = {"error": str(exc)[-400:]}

    # N > 1: BASELINE.json's config 5 as it is defined - ONE job of 25 M reads per GPU (200 M reads on 8 GPUs), the tables
    # merged once at its end
    c5_ms, c5_entries = 0.0, 0
    if world > 1 and not args.kernels_only and not args.no_configs:
        n5 = 25_000_000
        for e in engines:
            e.close()
        engines = []
        del res
        torch.cuda.empty_cache()
        eng5 = res5 = None
        try:                               # (host or device memory: every rank must get through, or all skip together)
            b5 = cfg.generate(rank * n5, n5)
            eng5 = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 23, **pin), device=local)
            res5 = eng5.make_resident(HostRagged.from_matrix(b5["seq"], lengths=b5["len"]), HostRagged.from_matrix(b5["f0"]),
                                      HostRagged.from_matrix(b5["f1"]))
            del b5
            ok5 = 1
        except Exception:                  # noqa: BLE001
            ok5 = 0
        t = torch.tensor([ok5], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok5 = int(t.item())
    if world > 1 and not args.kernels_only and not args.no_configs and ok5:
        d5 = DistributedScar(eng5, entry_capacity=1 << 22)
        first5 = rank * n5
        eng5.reset_async()
        eng5.run_resident(res5, first_index=first5)
        torch.cuda.synchronize()
        t = torch.tensor([eng5.table_size()], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cap5 = int(-(-(int(t.item()) * 1.25 / world + 4096) // 4096) * 4096)
        d5.merge_partitioned(region_capacity=cap5)
        barrier()
        j0 = ev()
        eng5.reset_async()
        eng5.run_resident(res5, first_index=first5)
        d5.merge_partitioned(region_capacity=cap5)
        j1 = ev()
        barrier()
        c5_ms = j0.elapsed_time(j1)
        t = torch.tensor([eng5.table_size()], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        c5_entries = int(t.item())
        eng5.close()

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_s, k3, raw_ms, job_ms, c5_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_s, k3, raw_ms, job_ms, c5_ms = (float(x) for x in t.tolist())

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = ALGO_BYTES_PER_READ * n / (k3 * 1e-3) / 1e9
        traffic, secondary = None, None
        tp = os.path.join(ROOT, "profiles", "k3_traffic.json")
        if os.path.exists(tp):
            try:
                prof = json.load(open(tp))
                traffic = prof.get("dram_bytes_per_launch")
                # SURVEY.md §8d secondary bound: the kernel is integer-ALU / issue bound, not DRAM bound
                secondary = {"bound": "int_alu_pipe", "alu_pipe_pct_of_peak": prof.get("alu_pipe_pct_of_peak"),
                             "issue_active_pct_of_peak": prof.get("issue_active_pct_of_peak"),
                             "dram_throughput_pct_of_peak": prof.get("dram_throughput_pct_of_peak"),
                             "warp_instructions_per_read": prof.get("warp_instructions_per_read"),
                             "warp_instructions_per_cycle_per_smsp": prof.get("warp_instructions_per_cycle_per_smsp"),
                             "best_mixed_integer_issue_rate_measured": prof.get("best_mixed_integer_issue_rate_measured"),
                             "issue_rate_source": prof.get("issue_rate_source"),
                             "source": prof.get("source", "ncu --set full under profiles/ (not measured in this run)")}
            except Exception:
                traffic = None
        h2d = int(p_seq.numel() + p_f0.numel() + p_f1.numel())
        line = {
            "metric": METRIC, "value": world * n * args.steps / (elapsed_ms * 1e-3), "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reads_per_gpu": n, "read_len": cfg.read_len, "n_samples": cfg.n_samples,
                       "n_targets": len(cfg.targets), "global_reads_per_step": world * n,
                       "parallelism": "reads sharded in contiguous blocks, dp{}".format(world),
                       "l2": "inputs_exceed_l2 ({} MB packed rows + {} MB raw per step vs 126 MB L2)".format(
                           n * eng.cells_per_read * 8 // 2 ** 20, n * cfg.read_len // 2 ** 20),
                       "timed_region": "from raw read bytes resident in HBM: reset, K2 demux, fused kernel (K1 pack + K3 sliding "
                                       "window + K4 aggregation in one pass)" +
                                       (", owner-partitioned table merge over NCCL all-to-all + accumulator all-reduce every "
                                        "step (side stream, overlapped with the next step's kernels; the last merge "
                                        "drains inside the timed region)"
                                        if world > 1 else ""),
                       "scar_patterns": int(n_entries), "unassigned_reads": int(un),
                       "table_slots": table_capacity},
            "from_raw_bytes": {"value": world * n / (raw_ms * 1e-3), "unit": "reads/s", "ms_per_step": raw_ms,
                               "steps": raw_steps,
                               "timed_region": "reset, K2 demux, K1 pack, K3, K4 (scar_run_dev) from raw read bytes "
                                               "resident in HBM; no cross-rank merge"},
            "stages_ms": ({"k2_demux": k2, "fused_pack_window_aggregate": k3, "merge": merge_ms if world > 1 else 0.0}
                          if fused else {"k2_demux": k2, "k1_pack_k3_window": k3, "k4_aggregate": k4,
                                         "merge": merge_ms if world > 1 else 0.0}),
            "three_kernel_path_ms": {"k1_pack": k1_ms, "k3_window": k3_alone_ms, "k4_aggregate": k4_alone_ms,
                                     "note": "each kernel timed alone over the same resident batch (packed rows and "
                                             "records travel through HBM); the fused kernel replaces all three"},
            "roofline": {"bound": "hbm", "kernel": "scar_window_kernel<fused> (K1 pack + K3 sliding window + K4 aggregation, "
                                                   "one pass over the raw bytes)" if fused else "scar_window_kernel (K3)",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ, "secondary": secondary},
            "e2e": {"value": world * n * e2e_steps / e2e_s if e2e_steps else None, "unit": "reads/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(n * 16), "steps": e2e_steps,
                    "path": "ScarEngine.process_host -> scar_process_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if fastq_e2e:
            line["e2e_fastq_text"] = fastq_e2e
        if pattern_searches:
            line["pattern_searches"] = pattern_searches
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if affinity:
            line["config"]["rank0_cpu_affinity"] = "{} CPUs local to the GPU (NVML)".format(len(affinity))
        if world > 1 and job_ms > 0:
            line["job_with_one_merge"] = {"value": world * n * raw_steps / (job_ms * 1e-3), "unit": "reads/s", "batches": raw_steps,
                                          "ms": job_ms,
                                          "timed_region": "reset, {} batches of {} reads per GPU into one table per rank (K2 + fused "
                                                          "kernel each), then ONE owner-partitioned merge + accumulator "
                                                          "all-reduce; max over ranks".format(raw_steps, n)}
        if world > 1 and c5_ms > 0:
            line["config5_job"] = {"value": world * 25_000_000 / (c5_ms * 1e-3), "unit": "reads/s", "reads_per_gpu": 25_000_000,
                                   "global_reads": world * 25_000_000, "ms": c5_ms, "merged_scar_patterns": c5_entries,
                                   "timed_region": "BASELINE.json config 5 (200 M reads on 8 GPUs) as one job: reset, K2 + fused "
                                                   "kernel over 25 M resident reads per GPU, ONE owner-partitioned merge + "
                                                   "accumulator all-reduce; max over ranks"}
        if merge_check:
            line["merge_check"] = merge_check
            line["config"]["sm_reserved_for_merge"] = sm_reserve
            line["config"]["merge_stream_priority"] = merge_priority
            line["config"]["merge_order"] = merge_order
        if pipeline:
            line["pipeline"] = pipeline
        if configs:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    for e in engines:
        e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
