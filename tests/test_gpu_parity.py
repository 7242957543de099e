"""
GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI
(libscar_b200.so via scarmapper_b200.engine) and is compared bit for bit with
  (a) the golden vectors captured from the unmodified reference driver, and
  (b) the CPU oracle on seeded synthetic inputs of the BASELINE.json shapes.
"""
import numpy as np
import pytest

import golden_util as gu
import synth
from oracle import scar_oracle as orc

pytestmark = pytest.mark.gpu


def engine_for(pin, max_read_len, **kw):
    from scarmapper_b200.engine import ScarEngine, ScarProblem
    return ScarEngine(ScarProblem(max_read_len=max_read_len, **pin, **kw), device=0)


def assert_records_equal(got, want, reads=None):
    if got.tobytes() == want.tobytes():
        return
    bad = np.nonzero(got != want)[0]
    i = int(bad[0])
    raise AssertionError("{} of {} records differ; first at {}: got {} want {} read {}".format(
        len(bad), len(got), i, got[i], want[i], None if reads is None else reads[i]))


def table_of(eng, s):
    tab = eng.table()
    sel = tab[tab["sample"] == s]
    return [(int(e["count"]), int(e["first_idx"])) for e in sel]


@pytest.mark.parametrize("tables", ["bucketed", "compact"])
@pytest.mark.parametrize("anchor", ["anchor", "scan"])
@pytest.mark.parametrize("name", gu.golden_names())
def test_golden_vectors_through_c_abi(name, anchor, tables, monkeypatch):
    # both K3 junction searches: anchor-and-extend (unique-10-mer loci) and the exhaustive cell scan; both layouts of
    # the window tables: bucketed perfect hash (one probe) and compact hash-and-displace (many loci)
    if anchor == "scan":
        monkeypatch.setenv("SCAR_NO_ANCHOR", "1")
    monkeypatch.setenv("SCAR_COMPACT_TABLES", "1" if tables == "compact" else "0")
    blob = gu.load(name)
    case, golden = blob["case"], blob["golden"]
    pin = gu.problem_inputs(case)
    f0, f1 = gu.index_fields(case)
    reads = case["reads"]
    eng = engine_for(pin, max(16, max(len(r) for r in reads)))
    rec = eng.process_host(reads, f0, f1)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    assert_records_equal(rec, want, reads)
    tab = eng.table()
    by_sample = {}
    for e in tab:
        by_sample.setdefault(int(e["sample"]), []).append((int(e["count"]), int(e["first_idx"])))
    cnt = gu.compare_with_golden(case, golden, rec, table_fn=lambda s: by_sample.get(s, []))
    mine, un = eng.counters()
    assert (mine == cnt).all()
    assert un == golden["read_count_dict"].get("unidentified", 0)
    eng.close()


def synth_problem(cfg, **over):
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target,
               left_index=cfg.index_d5, right_index=cfg.index_d7, platform="Illumina",
               phasing=cfg.phasing_lists(), n_limit=cfg.n_limit, min_length=cfg.min_length, hr_donor=cfg.hr_donor)
    pin.update(over)
    return pin


@pytest.mark.parametrize("name,n,kw", [
    ("C1", 60000, {}), ("C2", 60000, {}), ("C3", 60000, {}), ("C4", 20000, {}),
    ("C3", 40000, {"min_len": 20, "p_n": 0.05}), ("C1", 40000, {"p_sub": 0.5, "p_n": 0.2, "min_len": 100}),
    ("C4", 20001, {"p_sub": 0.3}),
])
def test_synthetic_configs_match_oracle(name, n, kw, monkeypatch):
    if name == "C4" and n == 20000:
        monkeypatch.setenv("SCAR_NO_ANCHOR", "1")     # exhaustive scan on the long-amplicon shape
    cfg = synth.make_config(name, n_reads=n, seed=101, **kw)
    batch = cfg.generate(0, n)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    pin = synth_problem(cfg)
    prob = orc.Problem(**pin)
    want = orc.classify_batch(prob, reads, f0, f1)
    eng = engine_for(pin, cfg.read_len)
    # fixed-stride host input (what the synthetic generator and the batched loader produce)
    from scarmapper_b200.engine import HostRagged
    seq = HostRagged.from_matrix(batch["seq"], lengths=batch["len"])
    rec = eng.process_host(seq, batch["f0"], batch["f1"])
    assert_records_equal(rec, want, reads)
    # aggregated table: (sample, first_idx, count) in the reference's first-seen order
    tables = orc.aggregate(prob, want, reads)
    ref = [(s, v[1], v[0]) for s, d in enumerate(tables) for v in d.values()]
    mine = [(int(e["sample"]), int(e["first_idx"]), int(e["count"])) for e in eng.table()]
    assert mine == ref
    cnt, un = eng.counters()
    assert (cnt == orc.counters(prob, want)).all()
    assert un == int((want["status"] == orc.ST_UNASSIGNED).sum())
    # phase histograms
    ph = eng.phase_hist()
    ok = want["status"] != orc.ST_UNASSIGNED
    for s in range(0, prob.n_samples, 7):
        sel = ok & (want["sample"] == s)
        for side, col in enumerate(("phase_r1", "phase_r2")):
            v = want[col][sel].astype(int)
            assert ph[s, side, 0] == int((v == -1).sum())
            for k in range(6):
                assert ph[s, side, 1 + k] == int((v == k).sum())
    eng.close()


def test_resident_path_equals_host_path_and_verifies():
    cfg = synth.make_config("C3", n_reads=50000, seed=77)
    batch = cfg.generate(0, 50000)
    pin = synth_problem(cfg)
    eng = engine_for(pin, cfg.read_len)
    from scarmapper_b200.engine import HostRagged
    seq = HostRagged.from_matrix(batch["seq"], lengths=batch["len"])
    rec_host = eng.process_host(seq, batch["f0"], batch["f1"]).copy()
    tab_host = eng.table().copy()
    cnt_host = eng.counters()[0].copy()
    eng.reset()
    import torch
    res = eng.make_resident(seq, batch["f0"], batch["f1"])
    eng.run_resident(res)
    torch.cuda.synchronize()
    assert_records_equal(eng.records_of(res), rec_host)
    assert eng.table().tobytes() == tab_host.tobytes()
    assert (eng.counters()[0] == cnt_host).all()
    assert eng.verify_resident(res) == 0
    # idempotence of the accumulators: a second pass doubles every count and keeps first_idx
    eng.run_resident(res)
    torch.cuda.synchronize()
    t2 = eng.table()
    assert (t2["count"] == 2 * tab_host["count"]).all() and (t2["first_idx"] == tab_host["first_idx"]).all()
    eng.close()


def test_multi_chunk_pipeline_keeps_global_first_index(monkeypatch):
    monkeypatch.setenv("SCAR_CHUNK_READS", "4096")
    cfg = synth.make_config("C2", n_reads=30000, seed=55)
    batch = cfg.generate(0, 30000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    pin = synth_problem(cfg)
    prob = orc.Problem(**pin)
    want = orc.classify_batch(prob, reads, f0, f1)
    eng = engine_for(pin, cfg.read_len)
    rec = eng.process_host(reads, f0, f1, first_index=1000)     # CSR input, 8 chunks
    assert_records_equal(rec, want, reads)
    tables = orc.aggregate(prob, want, reads)
    ref = [(s, v[1] + 1000, v[0]) for s, d in enumerate(tables) for v in d.values()]
    assert [(int(e["sample"]), int(e["first_idx"]), int(e["count"])) for e in eng.table()] == ref
    eng.close()


def pack_reference(reads, W):
    """numpy restatement of the packed layout (scar_internal.cuh): cell-major [W][n_reads]."""
    out = np.zeros((W, len(reads)), dtype=np.uint64)
    code = {ord("A"): 0, ord("C"): 1, ord("T"): 2, ord("G"): 3}
    for r, s in enumerate(reads):
        b = s.encode("latin-1") if isinstance(s, str) else s
        for k, ch in enumerate(b):
            j, o = divmod(k, 16)
            if ch in code:
                out[j, r] |= np.uint64(code[ch] << (2 * o)) | np.uint64(1 << (32 + o))
            elif ch == ord("N"):
                out[j, r] |= np.uint64(1 << (48 + o))
    return out


@pytest.mark.parametrize("platform", ["Illumina", "TruSeq", "Ramsden"])
def test_pack_kernel_layout_and_transforms(platform):
    import random
    import torch
    rng = random.Random(4)
    reads = []
    for i in range(3000):
        n = rng.choice([0, 1, 5, 6, 11, 12, 13, 15, 16, 17, 31, 32, 33, 100, 149, 150, 151, 160])
        reads.append("".join(rng.choice("ACGTACGTACGTNacgtRY.") for _ in range(n)))
    pin = dict(targets=["ACGT" * 40], cutsites=[80], sample_target=[0], left_index=["ACGTAC"], right_index=["TTGACA"],
               platform=platform)
    eng = engine_for(pin, 160)
    res = eng.make_resident(reads)
    eng.pack_resident(res)
    torch.cuda.synchronize()
    pf = orc.PLATFORMS[platform]
    stored = [orc.platform_seq(pf, r.encode("latin-1")).decode("latin-1") for r in reads]
    want = pack_reference(stored, eng.cells_per_read)
    got = res["cells"].cpu().numpy().view(np.uint64)
    assert (got == want).all()
    lens = res["lens"].cpu().numpy().view(np.uint32)
    assert list(lens & 0xFFFF) == [len(s) for s in stored]
    assert list(lens >> 16) == [s.count("N") for s in stored]
    eng.close()


def test_demux_kernel_all_platform_rules():
    import random
    import torch
    rng = random.Random(8)
    # Illumina with odd index fields: short, long, empty, N, shifted
    idx = synth.illumina_dual_indices()
    right = [d7 for _, d7, _ in idx]
    left = [d5 for _, _, d5 in idx]
    f0, f1 = [], []
    for i in range(20000):
        s = rng.randrange(96)
        a, b = right[s], left[s]
        u = rng.random()
        if u < 0.2:
            p = rng.randrange(8); a = a[:p] + rng.choice("ACGTN") + a[p + 1:]
        elif u < 0.3:
            a = a[1:]
        elif u < 0.4:
            b = b + rng.choice("ACGT")
        elif u < 0.45:
            b = rng.choice("ACGT") + b
        elif u < 0.5:
            a = ""
        elif u < 0.55:
            a = a[:3] + a[4:] + "G"
        f0.append(a); f1.append(b)
    pin = dict(targets=["ACGT" * 40], cutsites=[80], sample_target=[0] * 96, left_index=left, right_index=right,
               platform="Illumina")
    eng = engine_for(pin, 160)
    res = eng.make_resident(["ACGT"] * len(f0), f0, f1)
    eng.demux_resident(res)
    torch.cuda.synchronize()
    got = res["samples"].cpu().numpy().view(np.uint16).astype(np.int64)
    L = orc.lib()
    lb, lo = orc.csr(left); rb, ro = orc.csr(right)
    for i in range(len(f0)):
        a, b = f0[i].encode(), f1[i].encode()
        w = L.orc_index_matching(0, 1, 96, lb.ctypes.data, lo.ctypes.data, rb.ctypes.data, ro.ctypes.data,
                                 b"", 0, a, len(a), b, len(b))
        assert (got[i] if got[i] != 0xFFFF else -1) == w, (i, f0[i], f1[i], got[i], w)
    eng.close()


def test_match_maker_batch_matches_oracle():
    import random
    from scarmapper_b200.engine import match_maker_batch
    rng = random.Random(12)
    q = ["".join(rng.choice("ACGTN") for _ in range(rng.randint(0, 40))) for _ in range(5000)]
    u = []
    for s in q:
        t = list(s)
        for _ in range(rng.randint(0, 4)):
            op = rng.random()
            if op < 0.4 and t:
                t[rng.randrange(len(t))] = rng.choice("ACGT")
            elif op < 0.7 and t:
                del t[rng.randrange(len(t))]
            else:
                t.insert(rng.randint(0, len(t)), rng.choice("ACGT"))
        u.append("".join(t))
    got = match_maker_batch(q, u)
    want = [orc.match_maker(a, b) for a, b in zip(q, u)]
    assert list(got) == want


def test_edge_cases_empty_and_degenerate_batches():
    cfg = synth.make_config("C1", n_reads=10, seed=3)
    pin = synth_problem(cfg)
    eng = engine_for(pin, 300)
    # empty batch
    rec = eng.process_host([], [], [])
    assert rec.size == 0 and eng.table_size() == 0
    # reads around every loop bound, all-N, lowercase, and one wild-type
    t, cut = cfg.targets[0], cfg.cutsites[0]
    reads = ["", "A", "ACGT" * 8 + "AC", "N" * 150, t[cut - 75:cut + 75], t[cut - 75:cut + 75].lower(),
             t[cut - 18:cut + 18], t[cut - 17:cut + 18], t[cut - 150:cut + 150], t[:101], t[:100]]
    f0 = [cfg.index_d7[0]] * len(reads)
    f1 = [cfg.index_d5[0]] * len(reads)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    rec = eng.process_host(reads, f0, f1)
    assert_records_equal(rec, want, reads)
    eng.close()


def test_read_longer_than_context_limit_is_an_error():
    from scarmapper_b200._native import ScarError
    cfg = synth.make_config("C1", n_reads=10, seed=3)
    eng = engine_for(synth_problem(cfg), 150)
    with pytest.raises(ScarError) as e:
        eng.process_host(["A" * 200], [cfg.index_d7[0]], [cfg.index_d5[0]])
    assert e.value.code == -3
    eng.close()


def test_unsupported_inputs_fail_loudly():
    from scarmapper_b200._native import ScarError
    with pytest.raises(ScarError):
        engine_for(dict(targets=["ACGT" * 20 + "N" + "ACGT" * 20], cutsites=[80], sample_target=[0],
                        left_index=["ACGTACGT"], right_index=["ACGTACGT"]), 150)
    with pytest.raises(ScarError):
        engine_for(dict(targets=["ACGT" * 40], cutsites=[80], sample_target=[0], left_index=["ACGTACGT"],
                        right_index=["ACGTACGT"], hr_donor="ACGTNACGTAC"), 150)


@pytest.mark.slow
def test_full_size_properties_config1():
    """BASELINE config 1 at full size (1M reads): size-independent properties instead of the oracle:
    counters add up, table counts add up to the SCAR counter, shuffling the batch order leaves every
    (sample, key) count unchanged, and an oracle spot-check on a stride sample."""
    n = 1_000_000
    cfg = synth.make_config("C1", n_reads=n, seed=2024)
    batch = cfg.generate(0, n)
    pin = synth_problem(cfg)
    eng = engine_for(pin, cfg.read_len, table_capacity=1 << 22)
    from scarmapper_b200.engine import HostRagged
    rec = eng.process_host(HostRagged.from_matrix(batch["seq"], lengths=batch["len"]), batch["f0"], batch["f1"])
    cnt, un = eng.counters()
    tab = eng.table()
    assert cnt[:, 0].sum() + un == n
    assert int(tab["count"].sum()) == int(cnt[:, 12].sum()) == int((rec["status"] == orc.ST_SCAR).sum())
    assert (cnt[:, 1] == cnt[:, 6] + cnt[:, 7] + cnt[:, 11] + cnt[:, 12]).all()
    # first_idx really is the first read carrying that key
    firsts = tab["first_idx"].astype(np.int64)
    assert (rec["status"][firsts] == orc.ST_SCAR).all()
    # permutation invariance of the counts
    perm = np.random.default_rng(1).permutation(n)
    eng.reset()
    eng.process_host(HostRagged.from_matrix(batch["seq"][perm], lengths=batch["len"][perm]),
                     batch["f0"][perm], batch["f1"][perm], want_records=False)
    tab2 = eng.table()
    a = np.sort(np.stack([tab["h1"], tab["h2"], tab["count"].astype(np.uint64)], 1).view("u8,u8,u8").ravel())
    b = np.sort(np.stack([tab2["h1"], tab2["h2"], tab2["count"].astype(np.uint64)], 1).view("u8,u8,u8").ravel())
    assert (a == b).all()
    # oracle spot check on every 97th read
    sel = np.arange(0, n, 97)
    sub = {k: (v[sel] if v is not None else None) for k, v in batch.items()}
    reads = cfg.reads_as_strings(sub)
    f0, f1 = cfg.fields_as_strings(sub)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    assert_records_equal(rec[sel], want, reads)
    eng.close()


def test_per_read_dropin_matches_compiled_reference(ref_sliding_window):
    """scarmapper_b200.SlidingWindow.sliding_window (same signature as the Cython module) against the
    reference's own compiled SlidingWindow.pyx: return value and summary_data mutation."""
    import random
    from scarmapper_b200 import SlidingWindow as mine
    import test_oracle_vs_reference as tov
    rng = random.Random(21)
    targets = []
    for _ in range(6):
        L = rng.choice([120, 235, 251, 472])
        t = tov.rand_seq(rng, L, "ACGT")
        targets.append((t, rng.randint(L // 3, 2 * L // 3)))
    n_scar = 0
    for k in range(600):
        target, cut = targets[k % len(targets)]
        rl = rng.choice([20, 35, 36, 37, 60, 100, 150, 150, 151, 200])
        dl, dr = rng.choice([0, 0, 1, 5, 20]), rng.choice([0, 0, 2, 7, 30])
        ins = tov.rand_seq(rng, rng.choice([0, 0, 1, 3, 12]), rng.choice(["ACGT", "ACGTN", "acgt"]))
        a = rl // 2 - rng.randint(0, 5)
        lo = max(0, cut - dl - a)
        read = (target[lo:max(lo, cut - dl)] + ins + target[min(len(target), cut + dr):])[:rl]
        read += tov.rand_seq(rng, rl - len(read), "ACGT")
        hr = read[40:52] if (k % 5 == 0 and len(read) > 100 and set(read[40:52]) <= set("ACGT")) else ""
        left, right = orc.window_mapping(target, cut)
        cw = target[cut - 4:cut + 4]
        s_ref, s_mine = tov.fresh_summary(), tov.fresh_summary()
        want, s1 = ref_sliding_window.sliding_window(read, target, cut, len(target), 15, len(target) - 15, s_ref,
                                                     left, right, cw, hr)
        got, s2 = mine.sliding_window(read, target, cut, len(target), 15, len(target) - 15, s_mine, left, right, cw, hr)
        assert got == want, (read, target, cut, hr)
        assert s2 is s_mine and s_mine == s_ref
        n_scar += bool(want)
    assert n_scar > 100


def test_sequence_magic_dropin():
    from scarmapper_b200 import Sequence_Magic as sm
    assert sm.match_maker("ATTACTCG", "ATTACTCGA") == orc.match_maker("ATTACTCG", "ATTACTCGA") == 0
    assert sm.match_maker("ATTACTCG", "TTACTCG") == orc.match_maker("ATTACTCG", "TTACTCG")
    assert sm.rcomp("ACGTNacgtRYX.") == orc.rcomp("ACGTNacgtRYX.")


def test_fastq_text_views_path_matches_oracle():
    """FASTQ text -> scar_fastq_index views -> one shared H2D copy -> whole path, vs the oracle on the
    parsed strings (CRLF line ends, a record without a final newline, odd headers)."""
    from scarmapper_b200 import fastq
    cfg = synth.make_config("C3", n_reads=20000, seed=31, min_len=60, p_n=0.02)
    batch = cfg.generate(0, 20000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    eol = ["\n", "\r\n"]
    parts = []
    for i, (r, a, b) in enumerate(zip(reads, f0, f1)):
        e = eol[i % 2]
        parts.append("@SYN:1:FC:1:{}:{}:{} 1:N:0:{}+{}{}{}{}+{}{}{}".format(i % 97, i, i * 7 % 1013, a, b, e, r, e, e,
                                                                             "I" * len(r), e))
    text = "".join(parts).rstrip("\r\n")          # last record without a newline
    fb = fastq.index_text(np.frombuffer(text.encode(), dtype=np.uint8))
    assert fb.n == len(reads)
    assert [fb.read(i) for i in range(0, fb.n, 501)] == reads[::501]
    pin = synth_problem(cfg)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    eng = engine_for(pin, cfg.read_len)
    rec = eng.process_host(fb.seq, fb.f0, fb.f1)
    assert_records_equal(rec, want, reads)
    tables = orc.aggregate(orc.Problem(**pin), want, reads)
    ref = [(s, v[1], v[0]) for s, d in enumerate(tables) for v in d.values()]
    assert [(int(e["sample"]), int(e["first_idx"]), int(e["count"])) for e in eng.table()] == ref
    eng.close()


@pytest.mark.slow
def test_full_size_properties_config2():
    """BASELINE config 2 at full size (10 M reads, 96 samples, device-resident path): counter identities,
    table counts add up, first_idx really first, two-shard processing with global indices + table merge
    equals one pass, verification finds no tag collision, and an oracle spot check on a stride sample."""
    import torch
    n = 10_000_000
    cfg = synth.make_config("C2", n_reads=n, seed=2025)
    batch = cfg.generate(0, n)
    pin = synth_problem(cfg)
    from scarmapper_b200.engine import HostRagged
    eng = engine_for(pin, cfg.read_len, table_capacity=1 << 22)
    res = eng.make_resident(HostRagged.from_matrix(batch["seq"]), batch["f0"], batch["f1"])
    eng.run_resident(res)
    torch.cuda.synchronize()
    rec = eng.records_of(res)
    cnt, un = eng.counters()
    tab = eng.table()
    assert eng.verify_resident(res) == 0
    assert cnt[:, 0].sum() + un == n
    assert int(tab["count"].sum()) == int(cnt[:, 12].sum()) == int((rec["status"] == orc.ST_SCAR).sum())
    assert (cnt[:, 1] == cnt[:, 6] + cnt[:, 7] + cnt[:, 11] + cnt[:, 12]).all()
    firsts = tab["first_idx"].astype(np.int64)
    assert (rec["status"][firsts] == orc.ST_SCAR).all() and (rec["sample"][firsts] == tab["sample"]).all()
    assert (np.diff(tab["sample"].astype(np.int64)) >= 0).all()          # sorted by sample, then first_idx
    # two shards with global indices, exported and merged into one table == one pass (the multi-GPU fold)
    half = n // 2
    a = engine_for(pin, cfg.read_len, table_capacity=1 << 22)
    b = engine_for(pin, cfg.read_len, table_capacity=1 << 22)
    a.process_host(HostRagged.from_matrix(batch["seq"][:half]), batch["f0"][:half], batch["f1"][:half],
                   first_index=0, want_records=False)
    b.process_host(HostRagged.from_matrix(batch["seq"][half:]), batch["f0"][half:], batch["f1"][half:],
                   first_index=half, want_records=False)
    buf = torch.empty((b.table_size() + 8, 32), dtype=torch.uint8, device="cuda")
    nb = b.export_table_dev(buf)
    a.merge_table_dev(buf, nb)
    torch.cuda.synchronize()
    merged = a.table()
    assert merged.tobytes() == tab.tobytes()
    a.close(); b.close()
    # oracle spot check on every 997th read
    sel = np.arange(0, n, 997)
    sub = {k: (v[sel] if v is not None else None) for k, v in batch.items()}
    reads = cfg.reads_as_strings(sub)
    f0, f1 = cfg.fields_as_strings(sub)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    assert_records_equal(rec[sel], want, reads)
    eng.close()


def test_multi_gpu_merges_match_single_gpu():
    """Both table merges under real NCCL (needs >= 2 GPUs; skipped on a one-GPU box)."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("anchor", ["anchor", "scan"])
def test_tables_in_global_memory_path(anchor, monkeypatch):
    """Blobs too large for shared memory are read through the global read-only path: same results."""
    monkeypatch.setenv("SCAR_FORCE_GLOBAL_TABLES", "1")
    if anchor == "scan":
        monkeypatch.setenv("SCAR_NO_ANCHOR", "1")
    blob = gu.load("ragged_hr_donor")
    case = blob["case"]
    pin = gu.problem_inputs(case)
    f0, f1 = gu.index_fields(case)
    eng = engine_for(pin, max(len(r) for r in case["reads"]))
    rec = eng.process_host(case["reads"], f0, f1)
    assert_records_equal(rec, orc.classify_batch(orc.Problem(**pin), case["reads"], f0, f1), case["reads"])
    eng.close()


@pytest.mark.parametrize("anchor", ["anchor", "scan"])
def test_cells_read_from_global_memory_path(anchor, monkeypatch):
    """Short reads normally have their tile's cells staged in shared memory; SCAR_K3_NO_STAGE=1 runs the same
    batch through the global-memory cell path (what reads longer than 192 nt take): same records and table."""
    if anchor == "scan":
        monkeypatch.setenv("SCAR_NO_ANCHOR", "1")
    cfg = synth.make_config("C2", n_reads=40000, seed=77, min_len=60, p_n=0.02)
    batch = cfg.generate(0, 40000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    pin = synth_problem(cfg)
    eng = engine_for(pin, cfg.read_len)
    staged = eng.process_host(reads, f0, f1).copy()
    tab = eng.table().copy()
    eng.close()
    monkeypatch.setenv("SCAR_K3_NO_STAGE", "1")
    eng = engine_for(pin, cfg.read_len)
    rec = eng.process_host(reads, f0, f1)
    assert rec.tobytes() == staged.tobytes() and eng.table().tobytes() == tab.tobytes()
    assert_records_equal(rec, orc.classify_batch(orc.Problem(**pin), reads, f0, f1), reads)
    eng.close()


@pytest.mark.parametrize("anchor", ["anchor", "scan"])
def test_long_reads_up_to_1024(anchor, monkeypatch):
    """Reads of 300-1000 nt on a 1.2-kb locus (W up to 64 cells), large deletions and insertions."""
    import random
    if anchor == "scan":
        monkeypatch.setenv("SCAR_NO_ANCHOR", "1")
    rng = random.Random(77)
    while True:
        target = "".join(rng.choice("ACGT") for _ in range(1200))
        if len({target[i:i + 10] for i in range(1191)}) == 1191:
            break
    cut = 600
    reads = []
    for i in range(3000):
        rl = rng.choice([300, 512, 513, 700, 1000, 1024])
        dl, dr = rng.choice([0, 0, 3, 40, 200]), rng.choice([0, 0, 5, 60, 250])
        ins = "".join(rng.choice("ACGTN" if i % 9 == 0 else "ACGT") for _ in range(rng.choice([0, 0, 2, 30])))
        a = rl // 2
        lo = max(0, cut - dl - a)
        r = (target[lo:cut - dl] + ins + target[cut + dr:])[:rl]
        r += "".join(rng.choice("ACGT") for _ in range(rl - len(r)))
        if i % 11 == 0:
            p = rng.randrange(len(r)); r = r[:p] + rng.choice("ACGTN") + r[p + 1:]
        reads.append(r)
    pin = dict(targets=[target], cutsites=[cut], sample_target=[0], left_index=["TATAGCCT"], right_index=["ATTACTCG"],
               platform="Illumina", n_limit=0.01, min_length=100)
    f0, f1 = ["ATTACTCG"] * len(reads), ["TATAGCCT"] * len(reads)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    assert (want["status"] == orc.ST_SCAR).sum() > 1000
    eng = engine_for(pin, 1024)
    rec = eng.process_host(reads, f0, f1)
    assert_records_equal(rec, want, reads)
    eng.close()


def test_many_samples_global_counter_path():
    """600 samples: the per-CTA shared-memory histogram does not fit, counters go through global atomics."""
    import random
    rng = random.Random(5)
    pool7 = ["".join(rng.choice("ACGT") for _ in range(8)) for _ in range(25)]
    pool5 = ["".join(rng.choice("ACGT") for _ in range(8)) for _ in range(24)]
    right = [a for a in pool7 for _ in pool5]
    left = [b for _ in pool7 for b in pool5]
    cfg = synth.make_config("C1", n_reads=30000, seed=41)
    batch = cfg.generate(0, 30000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = [], []
    for i in range(len(reads)):
        s = rng.randrange(len(right))
        a, b = right[s], left[s]
        if rng.random() < 0.1:
            p = rng.randrange(8); a = a[:p] + rng.choice("ACGT") + a[p + 1:]
        f0.append(a); f1.append(b)
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=[0] * len(right), left_index=left,
               right_index=right, platform="Illumina", phasing=cfg.phasing_lists(), n_limit=0.01, min_length=100)
    prob = orc.Problem(**pin)
    want = orc.classify_batch(prob, reads, f0, f1)
    eng = engine_for(pin, cfg.read_len)
    rec = eng.process_host(reads, f0, f1)
    assert_records_equal(rec, want, reads)
    cnt, un = eng.counters()
    assert (cnt == orc.counters(prob, want)).all()
    eng.close()


def test_index_sequences_with_non_acgt_symbols():
    """Index strings that are not pure ACGT disable the 2-bit shortcut: class-mapped bit-vector distance."""
    import random
    rng = random.Random(6)
    right = ["ATTANTCG", "TCCGGAGA", "CGCTCATT", "NNNNNNNN"]
    left = ["TATAGCCT", "ATAGAGGC", "CCTRYCCT", "GGCTCTGA"]
    f0, f1 = [], []
    for i in range(4000):
        s = rng.randrange(4)
        a, b = list(right[s]), list(left[s])
        for _ in range(rng.choice([0, 0, 1, 2])):
            side = a if rng.random() < 0.5 else b
            side[rng.randrange(8)] = rng.choice("ACGTNRY")
        if rng.random() < 0.1:
            del a[rng.randrange(len(a))]
        f0.append("".join(a)); f1.append("".join(b))
    cfg = synth.make_config("C1", n_reads=4000, seed=42)
    reads = cfg.reads_as_strings(cfg.generate(0, 4000))
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=[0] * 4, left_index=left, right_index=right,
               platform="Illumina", n_limit=0.01, min_length=100)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    assert len(set(want["sample"].tolist())) >= 4
    eng = engine_for(pin, cfg.read_len)
    assert_records_equal(eng.process_host(reads, f0, f1), want, reads)
    eng.close()


def run_dropin_demultiplex(name, tmp_path):
    """dropin.batched_consensus_demultiplex with the REAL engine on one golden case, driven through stand-ins for the
    reference objects it touches; returns the DataProcessing stand-in (+ .indexed, .golden)."""
    import argparse
    import collections
    import sys
    import types
    sys.path.insert(0, str(gu.GOLDEN_DIR))
    import make_golden
    from oracle import ref_shims
    from scarmapper_b200 import dropin

    blob = gu.load(name)
    case, golden = blob["case"], blob["golden"]
    wd = str(tmp_path) + "/"
    make_golden.write_inputs(wd, case)
    args = argparse.Namespace(
        WorkingFolder=wd, Job_Name="g", Verbose="ERROR", Platform=case["platform"], PEAR=True, Demultiplex=False,
        HR_Donor=case["hr_donor"], N_Limit=case["n_limit"], Minimum_Length=case["min_length"], OutputRawData=False,
        PatternThreshold=0.0001, RefSeq=wd + "refseq.fa", FASTQ1=wd + "reads.fastq.gz")

    # stand-ins shaped like the reference's objects (DataProcessing.dictionary_build / TargetMapper)
    pin = gu.problem_inputs(case)
    index_dict = collections.OrderedDict()
    for s, iid in enumerate(case["sample_ids"]):
        index_dict[iid] = [pin["left_index"][s], 0, pin["right_index"][s], 0, iid, "S%d" % s, "1",
                           case["target_names"][case["sample_target"][s]]]
    target_dict = {nm: (nm, "ctg_" + nm, 0, len(t), sg.upper(), "YES" if rc else "NO")
                   for nm, t, sg, rc in zip(case["target_names"], case["targets"], case["sgrnas"], case["reverse_comp"])}
    phase_dict = collections.defaultdict(lambda: collections.defaultdict(list))
    if case["platform"] != "TruSeq":
        for t, (r1, r2) in zip(case["target_names"], gu.phasing_lists(case)):
            phase_dict[t]["R1"] = [[p, "F%d" % k] for k, p in enumerate(r1)]
            phase_dict[t]["R2"] = [[p, "R%d" % k] for k, p in enumerate(r2)]

    sys.path.insert(0, ref_shims.SHIMS)
    import pysam as pysam_shim

    class Fasta:                       # pysam.FastaFile stand-in (same semantics as baseline/shims/pysam)
        def __init__(self, path):
            self.recs = dict(pysam_shim.read_fasta(path))

        def fetch(self, chrom, start=None, stop=None):
            if chrom not in self.recs:
                raise KeyError(chrom)
            return self.recs[chrom][start:stop]

    from scipy import stats
    module = types.SimpleNamespace(pysam=types.SimpleNamespace(FastaFile=Fasta), pyfaidx=None, stats=stats)
    log = types.SimpleNamespace(info=lambda *a: None, error=lambda *a: None, warning=lambda *a: None)
    dp = types.SimpleNamespace(
        args=args, log=log, index_dict=index_dict, target_dict=target_dict, phase_dict=phase_dict,
        phase_count=collections.defaultdict(lambda: collections.defaultdict(int)),
        read_count_dict=collections.defaultdict(), sequence_dict=collections.defaultdict(list), read_count=0,
        fastq1=types.SimpleNamespace(input_file=args.FASTQ1))
    dp.indexed, dp.lower_limit = dropin.batched_consensus_demultiplex(dp, module)
    dp.golden = golden
    return dp


@pytest.mark.parametrize("name", ["c3_12_targets_mh", "ragged_hr_donor", "truseq", "ramsden"])
def test_dropin_glue_on_gpu_from_fastq_file(name, tmp_path):
    """scarmapper_b200.dropin.batched_consensus_demultiplex with the REAL engine, fed from a gz FASTQ file
    and the reference's input file formats, driven through stand-ins for the reference objects it touches
    (index_dict, target_dict, phase_dict, args, a FastaFile reader).  What it hands back - read_count_dict,
    phase_count, per-sample SampleResult (counters, table with the first read's sub_list) - must equal what the
    unmodified reference driver produced (golden).  (The files, Raw_Data included, are compared by
    tests/test_dropin_forking_pool.py through the launcher.)"""
    dp = run_dropin_demultiplex(name, tmp_path)
    golden, indexed = dp.golden, dp.indexed
    assert dp.read_count == golden["read_count"]
    assert dict(dp.read_count_dict) == golden["read_count_dict"]
    # (data_output later touches phase_count[key] for every sample, leaving empty dicts on non-Illumina runs)
    assert {k: dict(v) for k, v in dp.phase_count.items()} == {k: v for k, v in golden["phase_count"].items() if v}
    assert indexed == sum(v for k, v in golden["read_count_dict"].items() if k != "unidentified")
    assert set(dp.sequence_dict) == set(golden["samples"])
    for iid, g in golden["samples"].items():
        res = dp.sequence_dict[iid]
        c = res.counters
        assert [int(c[1]), int(c[2]), int(c[3]), int(c[4]), int(c[5]), [int(c[6]), int(c[7])], [int(c[8]), 0],
                [int(c[9]), int(c[10])]] == g["summary"]
        assert res.cutsite == g["cutsite"]
        got = [["{}|{}|{}|{}|{}".format(s[0], s[1], s[2], s[3], s[9]), n] + [s[i] for i in (0, 1, 2, 3, 5, 6, 7, 8, 9)] + [s[4]]
               for n, s in res.table()]
        assert got == g.get("table", []), iid
        assert res.pattern_out is not None and len(res.pattern_out) == res.n_patterns()     # searched in the parent


@pytest.mark.parametrize("piece_bytes", [None, 4096, 65536])
def test_gpu_fastq_indexer_matches_host_indexer_and_oracle(piece_bytes, monkeypatch):
    """FASTQ text -> GPU indexer (newline count / scan / positions, one thread per record) -> whole path.
    Records must equal the host-indexed views path and the oracle; malformed text must raise where the
    reference raises.  Small pieces force the pipelined path through many pieces (records straddling a piece
    boundary, pieces that complete no record, the unterminated last record in the last piece)."""
    import random
    if piece_bytes:
        monkeypatch.setenv("SCAR_FASTQ_PIECE_BYTES", str(piece_bytes))
    from scarmapper_b200 import fastq
    from scarmapper_b200._native import ScarError
    cfg = synth.make_config("C3", n_reads=30000, seed=61, min_len=40, p_n=0.02)
    batch = cfg.generate(0, 30000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    rng = random.Random(3)
    parts = []
    for i, (r, a, b) in enumerate(zip(reads, f0, f1)):
        e = "\r\n" if i % 3 == 0 else "\n"
        pad = " " if i % 17 == 0 else ""
        extra = "+X" if i % 29 == 0 else ""
        parts.append("@{}SYN:{}:x y:N:0:{}+{}{}{}{}{}{}+{}{}{}".format("@" * (i % 2), i, a, b, extra, e, r, pad, e, e,
                                                                    "I" * len(r), e))
    for final_newline in (True, False):
        text = "".join(parts)
        if not final_newline:
            text = text.rstrip("\r\n")
        arr = np.frombuffer(text.encode(), dtype=np.uint8)
        pin = synth_problem(cfg)
        eng = engine_for(pin, cfg.read_len)
        rec = eng.process_fastq_text(arr)
        assert len(rec) == len(reads)
        want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
        assert_records_equal(rec, want, reads)
        tab = eng.table().copy()
        eng.reset()
        fb = fastq.index_text(arr)
        rec2 = eng.process_host(fb.seq, fb.f0, fb.f1)
        assert rec2.tobytes() == rec.tobytes() and eng.table().tobytes() == tab.tobytes()
        eng.close()
    eng = engine_for(synth_problem(cfg), cfg.read_len)
    with pytest.raises(ScarError):
        eng.process_fastq_text(np.frombuffer(b"@a:AC+GT\nACGT\n+\nII\n", dtype=np.uint8))
    with pytest.raises(ScarError):
        eng.process_fastq_text(np.frombuffer(b"@a:ACGT\nACGT\n+\nIIII\n", dtype=np.uint8))
    assert len(eng.process_fastq_text(np.frombuffer(b"@a:AC+GT\nACGT\n+\nIIII\n@trunc\nAC\n", dtype=np.uint8))) == 1
    eng.close()


def test_context_info_config3_stays_in_shared_memory():
    cfg = synth.make_config("C3", n_reads=10, seed=101)
    eng = engine_for(synth_problem(cfg), cfg.read_len)
    info = eng.info()
    assert info["anchor_targets"] == 12 and info["tables_in_smem"] and info["static_bytes"] < 200 * 1024
    eng.close()
    blob = gu.load("c4_repeat_locus")
    eng = engine_for(gu.problem_inputs(blob["case"]), 300)
    assert eng.info()["anchor_targets"] == 0          # a 10-mer repeat forces the exhaustive scan
    eng.close()


@pytest.mark.gpu
def test_homeology_batch_kernel_equals_the_host_build_of_the_same_header():
    """scar_homeology_batch (one thread per scar pattern) against scar_homeology_core.h compiled for the host,
    which tests/test_homeology_host.py pins to the reference's homeology_search / templated_insertion_search.
    Includes junctions at the string ends (Python slice wrap-around), empty strings, non-ACGT insertions and
    two targets with separate oriented / stored sequences."""
    import random
    from oracle_engine import host_homeology_batch
    from scarmapper_b200 import Sequence_Magic
    from scarmapper_b200.engine import homeology_batch
    rng = random.Random(23)
    regions = ["".join(rng.choice("ACGT") for _ in range(L)) for L in (472, 97)]
    targets = [regions[0], Sequence_Magic.rcomp(regions[1])]
    cons, inss, clj, crj, tlj, trj, tix = [], [], [], [], [], [], []
    for case in range(20000):
        t = case % 2
        L = len(regions[t])
        cut = rng.randint(5, L - 5)
        dl, dr = rng.randint(0, min(30, cut)), rng.randint(0, min(30, L - cut))
        ins = "".join(rng.choice("ACGT" if case % 6 else "ACGTNacgr") for _ in range(rng.choice([0, 0, 1, 3, 5, 8, 20])))
        if len(ins) >= 5 and rng.random() < 0.5:
            a = rng.randint(max(0, cut - 45), max(0, cut - 6))
            ins = regions[t][a:a + len(ins)]
            if rng.random() < 0.5:
                ins = Sequence_Magic.rcomp(ins)
        lf = targets[t][max(0, cut - dl - rng.randint(0, 60)):cut - dl]
        rf = targets[t][cut + dr:cut + dr + rng.randint(0, 60)]
        c = lf + ins + rf
        a, b, x, y = len(lf), len(lf) + len(ins), cut - dl, cut + dr
        if case % 9 == 0:
            a, b = rng.randint(0, len(c)), rng.randint(0, len(c))
            x = rng.randint(0, L)
            y = rng.randint(x, L)
        cons.append(c); inss.append(ins); clj.append(a); crj.append(b); tlj.append(x); trj.append(y); tix.append(t)
    got = homeology_batch(cons, inss, clj, crj, tlj, trj, tix, targets, regions)
    want = host_homeology_batch(cons, inss, clj, crj, tlj, trj, tix, targets, regions)
    assert got.shape == want.shape == (20000, 16)
    assert np.array_equal(got, want)
    assert (want[:, 0] >= 0).sum() > 1000 and (want[:, 12] >= 0).sum() > 200      # the searches really find things
    assert homeology_batch([], [], [], [], [], [], [], targets, regions).shape == (0, 16)


# ---- fused kernel (K1 + K3 + K4 in one pass over the raw bytes) vs the three-kernel path ---------------------
def _run_both_paths(cfg, n, pin, monkeypatch, mutate=None, max_read_len=None):
    """(records, table, counters, unassigned, phase hist) of the same batch through scar_run_dev twice: fused
    (default) and K1 -> K3 -> K4 (SCAR_NO_FUSE=1); also returns the oracle's records."""
    import torch
    from scarmapper_b200.engine import HostRagged
    batch = cfg.generate(0, n)
    if mutate:
        mutate(batch)
    seq = HostRagged.from_matrix(batch["seq"], lengths=batch["len"])
    out = []
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("SCAR_NO_FUSE", raising=False)
        else:
            monkeypatch.setenv("SCAR_NO_FUSE", "1")
        eng = engine_for(pin, max_read_len or cfg.read_len)
        assert eng.info()["fused"] == fused
        res = eng.make_resident(seq, batch["f0"], batch["f1"])
        eng.run_resident(res, first_index=7)
        torch.cuda.synchronize()
        cnt, un = eng.counters()
        out.append((eng.records_of(res).copy(), eng.table().copy(), cnt.copy(), un, eng.phase_hist().copy(),
                    eng.verify_resident(res, first_index=7)))
        eng.close()
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    return out, want, reads


@pytest.mark.parametrize("anchor", ["anchor", "scan"])
@pytest.mark.parametrize("name,n,kw", [
    ("C1", 40000, {}), ("C2", 60000, {}), ("C3", 60000, {}), ("C4", 30000, {}),
    ("C3", 30000, dict(min_len=30, p_n=0.05, p_idx1=0.2, p_idx2=0.05)),
])
def test_fused_kernel_equals_three_kernel_path_and_oracle(name, n, kw, anchor, monkeypatch):
    if anchor == "scan":
        monkeypatch.setenv("SCAR_NO_ANCHOR", "1")
    cfg = synth.make_config(name, n_reads=n, seed=321, **kw)
    pin = synth_problem(cfg)

    def mutate(batch):            # lower-case runs and N bytes inside reads: insertion segments with non-ACGT bytes
        rng = np.random.Generator(np.random.PCG64(9))
        seq, ln = batch["seq"], batch["len"]
        for i in rng.integers(0, len(ln), max(50, len(ln) // 100)):
            if ln[i] > 90:
                p = int(rng.integers(60, ln[i] - 25))
                seq[i, p:p + 4] = np.frombuffer(b"acgt", dtype=np.uint8) if rng.random() < 0.5 else ord("R")
    (fu, un), want, reads = _run_both_paths(cfg, n, pin, monkeypatch, mutate=mutate)
    assert_records_equal(fu[0], want, reads)
    assert_records_equal(un[0], want, reads)
    assert fu[1].tobytes() == un[1].tobytes(), "scar tables differ (entries, counts, first_idx, hashes)"
    assert (fu[2] == un[2]).all() and fu[3] == un[3] and (fu[4] == un[4]).all()
    assert fu[5] == 0 and un[5] == 0
    prob = orc.Problem(**pin)
    tables = orc.aggregate(prob, want, reads)
    ref = [(s, v[1] + 7, v[0]) for s, d in enumerate(tables) for v in d.values()]
    assert [(int(e["sample"]), int(e["first_idx"]), int(e["count"])) for e in fu[1]] == ref
    assert (fu[2] == orc.counters(prob, want)).all()


def test_fused_kernel_on_fastq_views_takes_several_passes():
    """FASTQ text: the reads are a third of the bytes, so a tile's raw span is wider than the staging buffer and is
    taken in passes (and, with a tiny buffer, read from global memory); same records either way."""
    import os
    cfg = synth.make_config("C2", n_reads=20000, seed=99)
    batch = cfg.generate(0, 20000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    text = "".join("@SYN:1:FC:1:0:{}:0 1:N:0:{}+{}\n{}\n+\n{}\n".format(i, a, b, r, "I" * len(r))
                   for i, (r, a, b) in enumerate(zip(reads, f0, f1))).encode()
    pin = synth_problem(cfg)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    for raw_bytes in (None, "16384", "2048"):
        if raw_bytes is None:
            os.environ.pop("SCAR_FUSED_RAW_BYTES", None)
        else:
            os.environ["SCAR_FUSED_RAW_BYTES"] = raw_bytes
        try:
            eng = engine_for(pin, cfg.read_len)
            rec = eng.process_fastq_text(np.frombuffer(text, dtype=np.uint8))
            assert_records_equal(rec, want, reads)
            tables = orc.aggregate(orc.Problem(**pin), want, reads)
            ref = [(s, v[1], v[0]) for s, d in enumerate(tables) for v in d.values()]
            assert [(int(e["sample"]), int(e["first_idx"]), int(e["count"])) for e in eng.table()] == ref
            eng.close()
        finally:
            os.environ.pop("SCAR_FUSED_RAW_BYTES", None)


def test_fused_kernel_truseq_golden_and_many_samples(monkeypatch):
    """TruSeq (stored sequence = seq[6:-6]) goes through the fused kernel; so does a manifest too large for the
    shared-memory accumulators (global atomics)."""
    blob = gu.load("truseq")
    case, golden = blob["case"], blob["golden"]
    pin = gu.problem_inputs(case)
    reads = case["reads"]
    eng = engine_for(pin, max(len(r) for r in reads))
    assert eng.info()["fused"]
    rec = eng.process_host(reads, None, None)
    assert_records_equal(rec, orc.classify_batch(orc.Problem(**pin), reads, None, None), reads)
    by_sample = {}
    for e in eng.table():
        by_sample.setdefault(int(e["sample"]), []).append((int(e["count"]), int(e["first_idx"])))
    gu.compare_with_golden(case, golden, rec, table_fn=lambda s: by_sample.get(s, []))
    eng.close()


def test_process_host_verify_mode_is_exact_or_fail():
    """scar_ctx_set_verify: the batch is one resident chunk and every scarred read's full key is compared with the
    first read of its table entry.  Same results as the chunked pipeline; and when the tags are made to collide
    (test hook: tags that leave the insertion / microhomology bytes out) the call FAILS instead of merging keys."""
    from scarmapper_b200._native import ScarError
    cfg = synth.make_config("C3", n_reads=40000, seed=404)
    batch = cfg.generate(0, 40000)
    pin = synth_problem(cfg)
    from scarmapper_b200.engine import HostRagged
    seq = HostRagged.from_matrix(batch["seq"], lengths=batch["len"])
    eng = engine_for(pin, cfg.read_len)
    rec0 = eng.process_host(seq, batch["f0"], batch["f1"]).copy()
    tab0 = eng.table().copy()
    eng.close()
    for fused in (True, False):
        import os
        if not fused:
            os.environ["SCAR_NO_FUSE"] = "1"
        try:
            eng = engine_for(pin, cfg.read_len)
            eng.set_verify(True)
            rec1 = eng.process_host(seq, batch["f0"], batch["f1"])
            assert_records_equal(rec1, rec0)
            assert eng.table().tobytes() == tab0.tobytes()
            eng.close()
            eng = engine_for(pin, cfg.read_len)
            eng._L.scar_ctx_set_verify(eng._ctx, 2)            # verification on, tags weakened
            with pytest.raises(ScarError) as err:
                eng.process_host(seq, batch["f0"], batch["f1"])
            assert err.value.code == -6                          # SCAR_ERR_HASH_COLLISION
            eng.close()
        finally:
            os.environ.pop("SCAR_NO_FUSE", None)
