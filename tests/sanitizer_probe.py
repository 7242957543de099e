"""A small end-to-end pass for `compute-sanitizer` (memcheck / racecheck): every kernel of the path on a
few thousand reads, both K3 strategies, the staged and the global-memory cell paths, the FASTQ-text path in
several pieces, and the table export.  Not collected by pytest; run as a script on a GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from oracle import scar_oracle as orc  # noqa: E402
from scarmapper_b200.engine import ScarEngine, ScarProblem  # noqa: E402


def problem(cfg):
    return dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target, left_index=cfg.index_d5,
                right_index=cfg.index_d7, platform="Illumina", phasing=cfg.phasing_lists(), n_limit=cfg.n_limit,
                min_length=cfg.min_length)


def main():
    n = int(os.environ.get("PROBE_READS", "3000"))
    for name, kw in (("C2", {}), ("C4", {})):
        cfg = synth.make_config(name, n_reads=n, seed=5, p_n=0.02, **kw)
        batch = cfg.generate(0, n)
        reads = cfg.reads_as_strings(batch)
        f0, f1 = cfg.fields_as_strings(batch)
        pin = problem(cfg)
        want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
        for env in ({}, {"SCAR_NO_ANCHOR": "1"}, {"SCAR_K3_NO_STAGE": "1"}):
            for k in ("SCAR_NO_ANCHOR", "SCAR_K3_NO_STAGE"):
                os.environ.pop(k, None)
            os.environ.update(env)
            eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 14, **pin), device=0)
            rec = eng.process_host(reads, f0, f1)
            assert rec.tobytes() == want.tobytes(), (name, env)
            eng.table()
            eng.close()
        os.environ.pop("SCAR_K3_NO_STAGE", None)
        # FASTQ text in 4 KB pieces
        os.environ["SCAR_FASTQ_PIECE_BYTES"] = "4096"
        text = "".join("@x:1:{}+{}\n{}\n+\n{}\n".format(a, b, r, "I" * len(r)) for r, a, b in zip(reads, f0, f1))
        eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 14, **pin), device=0)
        rec = eng.process_fastq_text(np.frombuffer(text.encode(), dtype=np.uint8))
        assert rec.tobytes() == want.tobytes(), (name, "fastq")
        eng.close()
        os.environ.pop("SCAR_FASTQ_PIECE_BYTES", None)
        # the FILE path (gzip stream inflated by several threads, pieces of 128 KB), then the first reads of the table's
        # entries gathered on the device and the per-pattern searches
        import tempfile
        import zlib
        from scarmapper_b200 import engine as E
        os.environ["SCAR_FASTQ_FILE_PIECE_BYTES"] = str(1 << 17)
        with tempfile.NamedTemporaryFile(suffix=".fastq.gz", delete=False) as f:
            co = zlib.compressobj(4, zlib.DEFLATED, 31)
            f.write(co.compress(text.encode() * 3) + co.flush())
        eng = ScarEngine(ScarProblem(max_read_len=48, table_capacity=1 << 10, **pin), device=0)
        run = eng.process_fastq_file(f.name, threads=4, keep_text=False)
        assert run.n == 3 * n and run.records[:n].tobytes() == want.tobytes(), (name, "file")
        tab = eng.table()
        rows = tab["first_idx"].astype(np.int64)
        flip = (np.arange(rows.size) % 2).astype(np.uint8)
        got, off = eng.fastq_gather(rows, flip=flip, max_len=run.max_seq_len)
        ref, ref_off = E.stored_gather(reads, rows % n, "Illumina", flip=flip)
        assert (off == ref_off).all() and got.tobytes() == ref.tobytes(), (name, "gather")
        run.close()
        eng.close()
        os.remove(f.name)
        os.environ.pop("SCAR_FASTQ_FILE_PIECE_BYTES", None)
        k = min(500, rows.size)
        cons = [r for r in reads[:k]]
        out = E.homeology_batch(cons, [c[60:66] for c in cons], np.full(k, 60), np.full(k, 66), np.full(k, cfg.cutsites[0] - 5),
                                np.full(k, cfg.cutsites[0] + 5), np.zeros(k, np.int32), cfg.targets[:1], cfg.targets[:1], 0)
        assert out.shape == (k, 16)
        print("sanitizer_probe", name, "ok", flush=True)


if __name__ == "__main__":
    main()
