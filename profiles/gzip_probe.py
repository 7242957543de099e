"""profiles/gzip_probe.py - the single-stream gzip FASTQ path: zlib (1 thread) against the parallel inflate (scar_pgzip.h)
through scar_process_fastq_file, and the inflate alone through scar_gunzip_host.  PROBE_READS (default 4 M)."""
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import synth
from scarmapper_b200 import fastq
from test_gpu_ingest import synth_problem, engine_for

n = int(os.environ.get("PROBE_READS", "4000000"))
cfg = synth.make_config("C2", n_reads=n, seed=1)
b = cfg.generate(0, n)
rl = cfg.read_len
rng = np.random.default_rng(9)


def build_text(kind):
    if kind == "bench":          # what bench.py writes: constant header and quality line
        head = b"@SYN:1:FC:1:0:0:0 1:N:0:"
        qual = np.full((n, rl), ord("I"), dtype=np.uint8)
    else:                        # closer to an instrument's file: coordinates in the header, binned qualities in runs
        head = b"@M01234:56:000000000-ABCDE:1:"
        q = np.frombuffer(b"#,:FFFFFFFFF", dtype=np.uint8)
        qual = q[rng.integers(0, len(q), (n, rl // 5 + 1))].repeat(5, axis=1)[:, :rl]
    coord = np.char.encode(np.char.add(np.char.add((1101 + np.arange(n) // 4000).astype(str), ":"),
                                       np.char.add(np.char.add(rng.integers(1000, 30000, n).astype(str), ":"),
                                                   rng.integers(10000, 30000, n).astype(str)))) if kind != "bench" else None
    cw = 0 if coord is None else 16
    rec = len(head) + cw + 8 + 1 + 8 + 1 + rl + 3 + rl + 1 + (9 if coord is not None else 0)
    t = np.empty((n, rec), dtype=np.uint8)
    o = 0
    t[:, o:o + len(head)] = np.frombuffer(head, dtype=np.uint8); o += len(head)
    if coord is not None:
        cc = np.frombuffer(np.char.ljust(coord, cw, b"0").tobytes(), dtype=np.uint8).reshape(n, cw)
        t[:, o:o + cw] = cc; o += cw
        t[:, o:o + 9] = np.frombuffer(b" 1:N:0:AA", dtype=np.uint8)[:9]; o += 9
    t[:, o:o + 8] = b["f0"]; o += 8
    t[:, o] = ord("+"); o += 1
    t[:, o:o + 8] = b["f1"]; o += 8
    t[:, o] = 10; o += 1
    t[:, o:o + rl] = b["seq"]; o += rl
    t[:, o:o + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8); o += 3
    t[:, o:o + rl] = qual; o += rl
    t[:, o] = 10
    return t.reshape(-1)


pin = synth_problem(cfg)
out = {}
for kind, level in (("bench", 1), ("instrument", 6)):
    text = build_text(kind)
    t0 = time.perf_counter()
    co = zlib.compressobj(level, zlib.DEFLATED, 31)
    gz = co.compress(text.tobytes()) + co.flush()
    path = "/dev/shm/probe_{}.fastq.gz".format(kind)
    with open(path, "wb") as f:
        f.write(gz)
    res = {"reads": n, "text_bytes": int(text.size), "gz_bytes": len(gz), "level": level, "compress_s": round(time.perf_counter() - t0, 2)}
    # inflate alone
    for th in (1, 4, 8, 16):
        best = None
        for _ in range(2):
            st = {}
            t0 = time.perf_counter()
            got = fastq.gunzip(gz, threads=th, capacity=text.size, stats=st)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        assert got.size == text.size and (got[:1 << 20] == text[:1 << 20]).all() and (got[-(1 << 20):] == text[-(1 << 20):]).all()
        res["gunzip_host_{}t".format(th)] = {"s": round(best, 4), "text_GBps": round(text.size / best / 1e9, 2), **st}
    t0 = time.perf_counter(); zlib.decompress(gz, 31); res["python_zlib_s"] = round(time.perf_counter() - t0, 3)
    # the whole path
    eng = engine_for(pin, 160, table_capacity=1 << 22)
    ref_tab = None
    for th in (1, 8, 16):
        best, run = None, None
        for _ in range(2):
            eng.reset()
            t0 = time.perf_counter()
            run = eng.process_fastq_file(path, threads=th, keep_text=False)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert run.n == n
            wait = run.wait_seconds
            run.close()
        tab = eng.table().tobytes()
        ref_tab = ref_tab or tab
        assert tab == ref_tab
        res["file_{}t".format(th)] = {"s": round(best, 4), "reads_per_s": round(n / best), "producer_wait_s": round(wait, 4)}
    eng.close()
    os.remove(path)
    out[kind] = res
    print(json.dumps({kind: res}), flush=True)
