"""profiles/probe.py — a short resident-batch run of one config for ncu captures and quick A/B timing.

    python profiles/probe.py [--config C2] [--reads 10000000] [--steps 3] [--unfused]
Prints the per-step time (CUDA events) of reset + K2 + fused kernel (or K1/K3/K4 with --unfused)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--seed", type=int, default=20260101)
    a = ap.parse_args()
    if a.unfused:
        os.environ["SCAR_NO_FUSE"] = "1"
    import torch
    import synth
    from scarmapper_b200.engine import HostRagged, ScarEngine, ScarProblem
    cfg = synth.make_config(a.config, n_reads=a.reads, seed=a.seed)
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target, left_index=cfg.index_d5,
               right_index=cfg.index_d7, platform="Illumina", phasing=cfg.phasing_lists(), n_limit=cfg.n_limit,
               min_length=cfg.min_length)
    eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 22, **pin), device=0)
    batch = cfg.generate(0, a.reads)
    res = eng.make_resident(HostRagged.from_matrix(batch["seq"], lengths=batch["len"]), HostRagged.from_matrix(batch["f0"]),
                            HostRagged.from_matrix(batch["f1"]))
    stream = torch.cuda.current_stream()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e
    for _ in range(2):
        eng.reset_async()
        eng.run_resident(res)
    torch.cuda.synchronize()
    marks = []
    for _ in range(a.steps):
        eng.reset_async()
        m0 = ev()
        eng.demux_resident(res)
        m1 = ev()
        if eng.info()["fused"]:
            eng.classify_raw_resident(res)
        else:
            eng.pack_resident(res); eng.classify_resident(res); eng.aggregate_resident(res)
        m2 = ev()
        marks.append((m0, m1, m2))
    torch.cuda.synchronize()
    # the whole path in one call (scar_run_dev: with the direct index tables the fused kernel assigns the samples itself)
    runs = []
    for _ in range(a.steps):
        eng.reset_async()
        m0 = ev()
        eng.run_resident(res)
        runs.append((m0, ev()))
    torch.cuda.synchronize()
    run_ms = sum(x.elapsed_time(y) for x, y in runs) / len(runs)
    k2 = sum(m[0].elapsed_time(m[1]) for m in marks) / len(marks)
    rest = sum(m[1].elapsed_time(m[2]) for m in marks) / len(marks)
    print(json.dumps({"config": a.config, "reads": a.reads, "fused": eng.info()["fused"], "k2_ms": k2, "rest_ms": rest, "run_dev_ms": run_ms,
                      "info": eng.info(), "patterns": eng.table_size()}))
    eng.close()


if __name__ == "__main__":
    main()
