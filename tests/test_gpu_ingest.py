"""
GPU tests of the FASTQ FILE path (scar_process_fastq_file): a file - plain text, gzip (one stream
or several members), BGZF - streamed through the GPU must give exactly what the host-indexed path gives for the same
text: records against the oracle, table / counters against an engine fed from host buffers, the sequence views
against the host indexer (which tests/test_fastq_host.py pins to the reference's FASTQ_Reader).  The engine starts
with a read length and a scar table that are too small on purpose: the run must re-plan and grow them.
"""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import synth
from oracle import scar_oracle as orc

pytestmark = pytest.mark.gpu


def synth_problem(cfg, **over):
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target,
               left_index=cfg.index_d5, right_index=cfg.index_d7, platform="Illumina",
               phasing=cfg.phasing_lists(), n_limit=cfg.n_limit, min_length=cfg.min_length, hr_donor=cfg.hr_donor)
    pin.update(over)
    return pin


def engine_for(pin, max_read_len, **kw):
    from scarmapper_b200.engine import ScarEngine, ScarProblem
    return ScarEngine(ScarProblem(max_read_len=max_read_len, **pin, **kw), device=0)


def bgzf_bytes(data: bytes, block: int = 60000) -> bytes:
    """BGZF as bgzip writes it: gzip members with the 'BC' extra field, an empty member at the end."""
    out = []
    for a in list(range(0, len(data), block)) + [None]:
        chunk = b"" if a is None else data[a:a + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(payload) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) +
                   payload + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def fastq_text(reads, f0, f1, odd=True):
    parts = []
    for i, (r, a, b) in enumerate(zip(reads, f0, f1)):
        e = "\r\n" if odd and i % 3 == 0 else "\n"
        pad = " " if odd and i % 17 == 0 else ""
        parts.append("@SYN:{}:x y:N:0:{}+{}{}{}{}{}+{}{}{}".format(i, a, b, e, r, pad, e, e, "I" * len(r), e))
    return "".join(parts).encode()


@pytest.fixture(scope="module")
def case():
    # ragged reads (40 .. 150 nt) of the 12-locus shape: short reads first, so that the engine meets longer ones later
    cfg = synth.make_config("C3", n_reads=30000, seed=77, min_len=40, p_n=0.02)
    batch = cfg.generate(0, 30000)
    reads = cfg.reads_as_strings(batch)
    f0, f1 = cfg.fields_as_strings(batch)
    order = sorted(range(len(reads)), key=lambda i: (len(reads[i]) > 60, i))[:30000]
    reads, f0, f1 = [reads[i] for i in order], [f0[i] for i in order], [f1[i] for i in order]
    pin = synth_problem(cfg)
    want = orc.classify_batch(orc.Problem(**pin), reads, f0, f1)
    text = fastq_text(reads, f0, f1)
    # the host-buffer engine: the table / counters the file path must reproduce
    from scarmapper_b200 import fastq
    eng = engine_for(pin, cfg.read_len)
    fb = fastq.index_text(np.frombuffer(text, dtype=np.uint8))
    rec = eng.process_host(fb.seq, fb.f0, fb.f1)
    assert rec.tobytes() == want.tobytes()
    ref = dict(table=eng.table().copy(), counters=eng.counters()[0].copy(), unassigned=eng.counters()[1],
               phase=eng.phase_hist().copy(), seq_off=fb.seq.offsets.copy(), seq_len=fb.seq.lengths.copy())
    eng.close()
    return dict(cfg=cfg, pin=pin, reads=reads, text=text, want=want, ref=ref)


def write_source(kind, text, path):
    if kind == "text":
        data = text
    elif kind == "gzip":
        data = gzip.compress(text, 4)
    elif kind == "gzip_members":      # three members back to back, zero padding after the last (as tape / cat leave it)
        k = len(text) // 3
        data = gzip.compress(text[:k], 1) + gzip.compress(text[k:2 * k + 7], 9) + gzip.compress(text[2 * k + 7:], 5) + b"\0" * 64
    elif kind == "bgzf":
        data = bgzf_bytes(text)
    else:
        raise ValueError(kind)
    with open(path, "wb") as f:
        f.write(data)


@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("kind", ["text", "gzip", "gzip_members", "bgzf"])
def test_fastq_file_matches_the_host_indexed_path(kind, threads, case, tmp_path, monkeypatch):
    monkeypatch.setenv("SCAR_FASTQ_FILE_PIECE_BYTES", str(1 << 17))      # ~90 pieces: records straddle every boundary
    path = str(tmp_path / ("reads.fastq" + ("" if kind == "text" else ".gz")))
    write_source(kind, case["text"], path)
    # too small on purpose: reads of up to 150 nt against max_read_len 48, ~10 k scar patterns against 1 Ki slots
    eng = engine_for(case["pin"], 48, table_capacity=1 << 10)
    eng.set_verify(True)
    run = eng.process_fastq_file(path, threads=threads)
    assert run.source == {"text": "text", "gzip": "gzip", "gzip_members": "gzip", "bgzf": "bgzf"}[kind]
    assert run.n == len(case["reads"]) and run.text_bytes == len(case["text"])
    if kind.startswith("gzip"):      # one stream inflated by all the threads (csrc/scar_pgzip.h); zlib with one
        assert run.threads == threads and os.path.getsize(path) > 256 << 10
    assert run.max_seq_len == max(len(r) for r in case["reads"])       # (stripped of white space, as the views are)
    assert run.records.tobytes() == case["want"].tobytes()
    assert run.text.tobytes() == case["text"]
    ref = case["ref"]
    assert (run.seq_off == ref["seq_off"]).all() and (run.seq_len == ref["seq_len"]).all()
    assert eng.table().tobytes() == ref["table"].tobytes()
    cnt, un = eng.counters()
    assert (cnt == ref["counters"]).all() and un == ref["unassigned"]
    assert (eng.phase_hist() == ref["phase"]).all()
    # a second file through the same context accumulates (global read indices continue)
    run2 = eng.process_fastq_file(path, first_index=run.n, threads=threads, keep_text=False)
    assert run2.text is None and run2.records.tobytes() == case["want"].tobytes()
    tab2 = eng.table()
    assert (tab2["count"] == 2 * ref["table"]["count"]).all() and (tab2["first_idx"] == ref["table"]["first_idx"]).all()
    run.close(); run2.close()
    eng.close()


@pytest.mark.parametrize("platform", ["Illumina", "TruSeq", "Ramsden"])
def test_device_gather_of_stored_sequences_matches_the_host_gather(platform, case, tmp_path):
    """scar_fastq_gather_host: the stored sequences of chosen reads, gathered on the device from the resident text, are
    what scar_stored_gather_host builds from the host text (platform transform, flipped rows reverse-complemented)."""
    from scarmapper_b200 import engine as E, fastq
    path = str(tmp_path / "reads.fastq.gz")
    write_source("gzip", case["text"], path)
    pin = dict(case["pin"])
    pin["platform"] = platform
    if platform != "Illumina":
        pin["left_index"] = pin["right_index"] = None
        pin["phasing"] = None
    try:
        eng = engine_for(pin, 160, table_capacity=1 << 16)
    except Exception:
        pytest.skip("problem not expressible for " + platform)
    run = eng.process_fastq_file(path, keep_text=False)
    assert run.text is None and run.n == len(case["reads"])
    fb = fastq.index_text(np.frombuffer(case["text"], dtype=np.uint8))
    rng = np.random.default_rng(3)
    for n_rows in (0, 1, 40000):
        rows = rng.integers(0, run.n, n_rows).astype(np.int64)
        for flip in (None, rng.integers(0, 2, n_rows).astype(np.uint8)):
            got, got_off = eng.fastq_gather(rows, flip=flip)
            want, want_off = E.stored_gather(fb.seq, rows, platform, flip=flip)
            assert (got_off == want_off).all() and got.tobytes() == want.tobytes()
    with pytest.raises(Exception):
        eng.fastq_gather(np.asarray([run.n], dtype=np.int64))
    run.close()
    eng.close()


def test_fastq_file_three_kernel_path_grows_its_packed_batch(case, tmp_path, monkeypatch):
    """SCAR_NO_FUSE=1: K1 -> K3 -> K4 with the packed batch in HBM; the context starts at 48 nt, so the packed buffers
    are re-planned for 150-nt reads in mid-file."""
    monkeypatch.setenv("SCAR_NO_FUSE", "1")
    monkeypatch.setenv("SCAR_FASTQ_FILE_PIECE_BYTES", str(1 << 18))
    path = str(tmp_path / "reads.fastq")
    write_source("text", case["text"], path)
    eng = engine_for(case["pin"], 48, table_capacity=1 << 12)
    assert not eng.info()["fused"]
    run = eng.process_fastq_file(path)
    assert run.records.tobytes() == case["want"].tobytes()
    assert eng.table().tobytes() == case["ref"]["table"].tobytes()
    eng.close()


def test_fastq_file_record_limit_empty_file_and_last_record_without_newline(case, tmp_path):
    text = case["text"].rstrip(b"\r\n")
    path = str(tmp_path / "r.fastq.gz")
    with open(path, "wb") as f:
        f.write(gzip.compress(text, 1))
    eng = engine_for(case["pin"], 160)
    run = eng.process_fastq_file(path)
    assert run.n == len(case["reads"]) and run.records.tobytes() == case["want"].tobytes()
    eng.reset()
    run = eng.process_fastq_file(path, limit=1002)                       # --Verbose DEBUG stops after a fixed number
    assert run.n == 1002 and run.records.tobytes() == case["want"][:1002].tobytes()
    assert int(eng.counters()[0][:, 0].sum()) + eng.counters()[1] == 1002
    eng.reset()
    empty = str(tmp_path / "empty.fastq")
    open(empty, "wb").close()
    run = eng.process_fastq_file(empty)
    assert run.n == 0 and len(run.records) == 0
    eng.close()


def test_fastq_file_errors_are_loud(case, tmp_path):
    from scarmapper_b200._native import ScarError
    eng = engine_for(case["pin"], 160)
    with pytest.raises(ScarError):
        eng.process_fastq_file(str(tmp_path / "missing.fastq"))
    gz = gzip.compress(case["text"][:400000], 6)
    for name, data in (("truncated.gz", gz[:len(gz) // 2]), ("corrupt.gz", gz[:2000] + b"\xff" * 64 + gz[2064:]),
                       ("badcrc.gz", gz[:-8] + b"\0\0\0\0" + gz[-4:])):
        p = str(tmp_path / name)
        with open(p, "wb") as f:
            f.write(data)
        with pytest.raises(ScarError):
            eng.process_fastq_file(p)
        eng.reset()
    bad_block = bytearray(bgzf_bytes(case["text"][:300000]))
    bad_block[len(bad_block) // 2] ^= 0x55
    p = str(tmp_path / "bad_bgzf.gz")
    with open(p, "wb") as f:
        f.write(bytes(bad_block))
    with pytest.raises(ScarError):
        eng.process_fastq_file(p)
    eng.reset()
    p = str(tmp_path / "mismatch.fastq")
    with open(p, "wb") as f:
        f.write(b"@a:AC+GT\nACGT\n+\nII\n")
    with pytest.raises(ScarError):                                       # sequence / quality lengths differ
        eng.process_fastq_file(p)
    eng.reset()
    # the context is still usable after every failure
    p = str(tmp_path / "ok.fastq")
    with open(p, "wb") as f:
        f.write(case["text"])
    assert eng.process_fastq_file(p).records.tobytes() == case["want"].tobytes()
    eng.close()


def test_fastq_file_verify_mode_is_exact_or_fail(case, tmp_path):
    from scarmapper_b200._native import ScarError
    p = str(tmp_path / "r.fastq")
    with open(p, "wb") as f:
        f.write(case["text"])
    eng = engine_for(case["pin"], 160)
    eng._L.scar_ctx_set_verify(eng._ctx, 2)                              # verification on, tags weakened (test hook)
    with pytest.raises(ScarError) as err:
        eng.process_fastq_file(p)
    assert err.value.code == -6                                          # SCAR_ERR_HASH_COLLISION
    eng.close()


def test_table_reserve_keeps_every_entry(case):
    from scarmapper_b200 import fastq
    eng = engine_for(case["pin"], 160, table_capacity=1 << 16)
    fb = fastq.index_text(np.frombuffer(case["text"], dtype=np.uint8))
    eng.process_host(fb.seq, fb.f0, fb.f1)
    before = eng.table().copy()
    eng.table_reserve(1 << 20)                                           # forces the rebuild at a larger size
    assert eng.table().tobytes() == before.tobytes()
    eng.process_host(fb.seq, fb.f0, fb.f1, first_index=len(case["reads"]))
    after = eng.table()
    assert (after["count"] == 2 * before["count"]).all() and (after["first_idx"] == before["first_idx"]).all()
    eng.close()


@pytest.mark.parametrize("name", ["c3_12_targets_mh", "truseq"])
def test_dropin_streams_the_file_and_matches_the_host_loader(name, tmp_path):
    """batched_consensus_demultiplex with FILE_INGEST on (the default) and off: identical bookkeeping."""
    import sys
    import golden_util as gu
    sys.path.insert(0, str(gu.GOLDEN_DIR))
    import test_gpu_parity as tgp
    from scarmapper_b200 import dropin
    results = []
    for ingest in (True, False, "too large"):
        dropin.FILE_INGEST[0] = bool(ingest)
        if ingest == "too large":          # the text does not fit the device: the drop-in falls back to the host loader
            os.environ["SCAR_FASTQ_FILE_MAX_TEXT_BYTES"] = "100000"
        try:
            sub = tmp_path / str(ingest).replace(" ", "_")
            sub.mkdir()
            dp = tgp.run_dropin_demultiplex(name, sub)
        finally:
            dropin.FILE_INGEST[0] = True
            os.environ.pop("SCAR_FASTQ_FILE_MAX_TEXT_BYTES", None)
        assert ("fastq_stream_gpu_classify_aggregate" in dropin.STAGE_SECONDS) == (ingest is True)
        results.append((dp.read_count, dict(dp.read_count_dict), {k: dict(v) for k, v in dp.phase_count.items()},
                        {iid: (r.counters.tolist(), r.counts.tolist(), r.first_records.tobytes(), r.first_bytes.tobytes(),
                               r.pattern_out.tolist()) for iid, r in dp.sequence_dict.items()}))
    assert results[0] == results[1] == results[2]
