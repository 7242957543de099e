"""
engine.py — Python face of the C ABI: one ScarEngine per GPU.

ScarEngine holds what the reference computes once per run (window tables, index tables, phasing
tables: INDEL_Processing.py:57-80,1063-1124; TargetMapper.py:30-71) and runs the batched hot path
K2 demux -> K1 pack -> K3 sliding window -> K4 aggregation on its device.  Two ways in:

  process_host(reads, f0, f1)   host buffers; H2D / kernels / D2H are pipelined inside the library
  run_resident(batch)           inputs already in HBM (torch uint8 tensors), kernels only

Both accumulate into the engine's per-sample counters, phase histograms and scar table, read back
with counters() / phase_hist() / table().
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _native as N


def _b(x) -> bytes:
    return x.encode("latin-1") if isinstance(x, str) else bytes(x)


def _csr(strings):
    bs = [_b(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    return data, off


class HostRagged:
    """A list of byte strings in host memory: CSR (bytes + int64 offsets) or fixed stride
    (2-D uint8 array, optional int32 lengths).  Keeps the numpy arrays alive for the C call."""

    def __init__(self, bytes_, offsets=None, lengths=None, stride=0, n=None):
        self.bytes = np.ascontiguousarray(bytes_, dtype=np.uint8)
        self.offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        self.lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        self.stride = int(stride)
        if n is None:
            if self.offsets is not None:
                n = len(self.lengths) if self.lengths is not None else len(self.offsets) - 1
            else:
                n = self.bytes.size // max(self.stride, 1)
        self.n = int(n)

    @classmethod
    def from_strings(cls, strings):
        data, off = _csr(strings)
        return cls(data, offsets=off)

    @classmethod
    def from_matrix(cls, mat, lengths=None):
        mat = np.ascontiguousarray(mat, dtype=np.uint8)
        return cls(mat.reshape(-1), lengths=lengths, stride=mat.shape[1], n=mat.shape[0])

    @classmethod
    def from_views(cls, text, offsets, lengths):
        """Items are (offset, length) views into one text buffer (scar_ragged views mode)."""
        return cls(text, offsets=offsets, lengths=lengths, n=len(lengths))

    def max_len(self) -> int:
        if self.lengths is not None:
            return int(self.lengths.max()) if self.n else 0
        if self.offsets is not None:
            return int(np.max(np.diff(self.offsets))) if self.n else 0
        return self.stride

    def item(self, i: int) -> bytes:
        if self.offsets is not None:
            a = int(self.offsets[i])
            b = a + int(self.lengths[i]) if self.lengths is not None else int(self.offsets[i + 1])
        else:
            a = i * self.stride
            b = a + (int(self.lengths[i]) if self.lengths is not None else self.stride)
        return self.bytes[a:b].tobytes()

    def c_struct(self) -> N.Ragged:
        return N.Ragged(self.bytes.ctypes.data,
                        None if self.offsets is None else self.offsets.ctypes.data,
                        None if self.lengths is None else self.lengths.ctypes.data, self.stride)


def _view(ptr, count, dtype, writeable=True):
    """numpy view of `count` items of library-owned memory at `ptr` (no copy)."""
    dtype = np.dtype(dtype)
    if not count or not ptr:
        return np.zeros(0, dtype)
    a = np.ctypeslib.as_array((C.c_uint8 * (int(count) * dtype.itemsize)).from_address(ptr)).view(dtype)
    a.flags.writeable = writeable
    return a


class FastqFileRun:
    """What ScarEngine.process_fastq_file returns.  The arrays are views of memory the library owns (for a plain-text
    file the text IS the mapped file); they stay valid until close() or until this object is collected - copy what
    must live longer."""
    KINDS = ("text", "gzip", "bgzf")

    def __init__(self, lib, raw):
        import weakref
        self._raw = raw
        self.n = int(raw.n_records)
        self.text_bytes = int(raw.text_bytes)
        self.records = _view(raw.records, self.n, N.RECORD_DTYPE)
        self.seq_off = _view(raw.seq_off, self.n, np.int64)
        self.seq_len = _view(raw.seq_len, self.n, np.int32)
        self.text = _view(raw.text, self.text_bytes, np.uint8, writeable=False) if raw.text else None
        self.seq = HostRagged.from_views(self.text, self.seq_off, self.seq_len) if self.text is not None else None
        self.max_seq_len = int(raw.max_seq_len)
        self.source = self.KINDS[raw.source_kind]
        self.threads = int(raw.threads)
        self.wait_seconds, self.total_seconds = float(raw.wait_seconds), float(raw.total_seconds)
        self._final = weakref.finalize(self, lib.scar_fastq_run_release, C.byref(raw))

    def close(self):
        self.records = self.seq_off = self.seq_len = self.text = self.seq = None
        self._final()


def as_ragged(x):
    if x is None or isinstance(x, HostRagged):
        return x
    if isinstance(x, np.ndarray) and x.ndim == 2:
        return HostRagged.from_matrix(x)
    return HostRagged.from_strings(x)


@dataclass
class ScarProblem:
    """Static inputs of a run, in the reference's own terms."""
    targets: list                      # upper-cased target regions (ScarSearch.target_region)
    cutsites: list                     # ScarSearch.cutsite per target
    sample_target: list                # target id per manifest row
    left_index: list = None            # index_dict[s][0] (Master_Index_File column 3, upper-cased)
    right_index: list = None           # index_dict[s][2] (Master_Index_File column 2, upper-cased)
    platform: str = "Illumina"
    mismatch: int = -1                 # -1: platform default (INDEL_Processing.py:1142-1147)
    phasing: list = None               # per target ([R1 strings], [R2 strings]) (TargetMapper.phasing)
    hr_donor: str = ""
    n_limit: float = 0.01
    min_length: int = 100
    max_read_len: int = 160
    table_capacity: int = 0
    cutwindows: list = None            # per-read drop-in only: 10-nt cutwindow per target
    extra: dict = field(default_factory=dict)


def cutsite_search(target, sgrna, reverse_comp: bool) -> int:
    """ScarSearch.cutsite_search (INDEL_Processing.py:759-795); raises ScarError when the sgRNA does not map."""
    t, s = _b(target), _b(sgrna)
    out = C.c_int32(-1)
    N.check(N.lib().scar_cutsite_search(t, len(t), s, len(s), int(bool(reverse_comp)), C.byref(out)))
    return out.value


class ScarEngine:
    def __init__(self, prob: ScarProblem, device: int = 0):
        L = N.lib()
        self.prob = prob
        self.device = device
        self.n_samples = len(prob.sample_target)
        self.n_targets = len(prob.targets)
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data

        tb, to = _csr(prob.targets)
        lb, lo = _csr(prob.left_index if prob.left_index is not None else [""] * self.n_samples)
        rb, ro = _csr(prob.right_index if prob.right_index is not None else [""] * self.n_samples)
        cfg = N.Config()
        cfg.device = device
        cfg.platform = N.PLATFORMS[prob.platform] if isinstance(prob.platform, str) else int(prob.platform)
        cfg.mismatch = prob.mismatch
        cfg.n_targets = self.n_targets
        cfg.target_bytes, cfg.target_off = arr(tb, np.uint8), arr(to, np.int64)
        cfg.target_cutsite = arr(prob.cutsites, np.int32)
        cfg.n_samples = self.n_samples
        cfg.sample_target = arr(prob.sample_target, np.int32)
        cfg.left_index_bytes, cfg.left_index_off = arr(lb, np.uint8), arr(lo, np.int64)
        cfg.right_index_bytes, cfg.right_index_off = arr(rb, np.uint8), arr(ro, np.int64)
        if prob.phasing is not None and cfg.platform == 0:
            r1, r2, poff = [], [], [0]
            for p1, p2 in prob.phasing:
                r1 += list(p1); r2 += list(p2); poff.append(len(r1))
            b1, o1 = _csr(r1)
            b2, o2 = _csr(r2)
            cfg.phase_off = arr(poff, np.int32)
            cfg.phase_r1_bytes, cfg.phase_r1_off = arr(b1, np.uint8), arr(o1, np.int64)
            cfg.phase_r2_bytes, cfg.phase_r2_off = arr(b2, np.uint8), arr(o2, np.int64)
        hr = _b(prob.hr_donor or "")
        if hr:
            cfg.hr_donor = arr(np.frombuffer(hr, dtype=np.uint8), np.uint8)
            cfg.hr_len = len(hr)
        cfg.n_limit = float(prob.n_limit)
        cfg.min_length = int(prob.min_length)
        cfg.max_read_len = int(prob.max_read_len)
        cfg.table_capacity = int(prob.table_capacity)
        if prob.cutwindows is not None:
            cw = b"".join((_b(w) + b"\0" * 10)[:10] if len(_b(w)) == 10 else b"\0" * 10 for w in prob.cutwindows)
            cfg.cutwindow_bytes = arr(np.frombuffer(cw, dtype=np.uint8), np.uint8)
        ctx = C.c_void_p()
        N.check(L.scar_ctx_create(C.byref(cfg), C.byref(ctx)))
        self._ctx = ctx
        self._L = L
        self.cells_per_read = L.scar_cells_per_read(ctx)

    # ---- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.scar_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        N.check(self._L.scar_reset(self._ctx))

    def reset_async(self):
        """Clear table and accumulators on the current torch stream (no host synchronisation)."""
        N.check(self._L.scar_reset_dev(self._ctx, self._stream()))

    def set_verify(self, on: bool = True):
        """process_host keeps the batch resident as one chunk and verifies every scarred read's full key against the
        first read of its table entry (ScarError SCAR_ERR_HASH_COLLISION on a mismatch)."""
        N.check(self._L.scar_ctx_set_verify(self._ctx, 1 if on else 0))

    def device_memory(self):
        f, t = C.c_int64(), C.c_int64()
        N.check(self._L.scar_device_memory(self.device, C.byref(f), C.byref(t)))
        return f.value, t.value

    def set_sm_limit(self, n_sms: int = 0):
        """Size this engine's persistent grids for n_sms SMs (0 = all): leaves the rest to a merge running beside."""
        N.check(self._L.scar_ctx_set_sm_limit(self._ctx, int(n_sms)))

    def launch_count(self) -> int:
        return int(self._L.scar_launch_count(self._ctx))

    def info(self) -> dict:
        b, s, k = C.c_int64(), C.c_int32(), C.c_int32()
        N.check(self._L.scar_ctx_info(self._ctx, C.byref(b), C.byref(s), C.byref(k)))
        st, per, sm = C.c_int32(), C.c_int32(), C.c_int32()
        N.check(self._L.scar_k3_plan(self._ctx, C.byref(st), C.byref(per), C.byref(sm)))
        fe, fper, fsm = C.c_int32(), C.c_int32(), C.c_int32()
        N.check(self._L.scar_fused_info(self._ctx, C.byref(fe), C.byref(fper), C.byref(fsm)))
        return {"static_bytes": b.value, "tables_in_smem": bool(s.value), "anchor_targets": k.value,
                "k3_stage": bool(st.value), "k3_ctas_per_sm": per.value, "k3_smem_bytes": sm.value,
                "fused": bool(fe.value), "fused_ctas_per_sm": fper.value, "fused_smem_bytes": fsm.value}

    def window_counts(self, target: int):
        a, b = C.c_int32(), C.c_int32()
        N.check(self._L.scar_window_counts(self._ctx, target, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- host path ------------------------------------------------------------------------------
    def process_host(self, reads, f0=None, f1=None, first_index: int = 0, want_records: bool = True,
                     out: np.ndarray | None = None):
        """Whole path from host buffers.  Returns the per-read records (or None)."""
        seq, f0, f1 = as_ragged(reads), as_ragged(f0), as_ragged(f1)
        n = seq.n
        rec = None
        if want_records:
            rec = out if out is not None else np.empty(n, dtype=N.RECORD_DTYPE)
            assert rec.dtype == N.RECORD_DTYPE and rec.size >= n
        s, a, b = seq.c_struct(), (f0.c_struct() if f0 else None), (f1.c_struct() if f1 else None)
        N.check(self._L.scar_process_host(self._ctx, C.byref(s), C.byref(a) if a else None,
                                          C.byref(b) if b else None, n, first_index,
                                          rec.ctypes.data if rec is not None else None))
        return rec

    def process_fastq_text(self, text, first_index: int = 0, want_records: bool = True, out=None):
        """Whole path from FASTQ TEXT (uint8 array, whole records): one H2D copy of the text, record indexing
        on the GPU, then K2/K1/K3/K4 on views into the text.  Returns the records in file order (or their
        number when want_records is False)."""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        n = C.c_int64(0)
        N.check(self._L.scar_process_fastq_host(self._ctx, text.ctypes.data, text.size, first_index, None, 0,
                                                C.byref(n)))
        if not want_records:
            return n.value
        rec = out[:n.value] if out is not None else np.empty(n.value, dtype=N.RECORD_DTYPE)
        N.check(self._L.scar_fastq_records_fetch(self._ctx, rec.ctypes.data, n.value))
        return rec

    def process_fastq_file(self, path: str, limit: int | None = None, first_index: int = 0, threads: int = 0,
                           keep_text: bool = True) -> "FastqFileRun":
        """Whole path from a FASTQ FILE (plain text, gzip or BGZF; scar_process_fastq_file): the file is read /
        inflated by a producer thread while its pieces are copied to the GPU, indexed and classified there.
        Returns the run: records, sequence views and (keep_text) the text, as numpy views of library-owned memory."""
        raw = N.FastqRun()
        N.check(self._L.scar_process_fastq_file(self._ctx, path.encode(), int(limit or 0), first_index, int(threads),
                                                1 if keep_text else 0, C.byref(raw)))
        return FastqFileRun(self._L, raw)

    def fastq_gather(self, rows, flip=None, max_len=None):
        """Stored sequences of the reads `rows` of the last process_fastq_file run, gathered on the device from the
        resident text (scar_fastq_gather_host): (uint8 bytes, int64 offsets[len(rows) + 1]), as stored_gather.
        max_len: an upper bound of the stored length (FastqFileRun.max_seq_len) - one call instead of size + fill."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        fl = None if flip is None else np.ascontiguousarray(flip, dtype=np.uint8)
        assert fl is None or fl.size == rows.size
        off = np.zeros(rows.size + 1, dtype=np.int64)
        fp = None if fl is None else fl.ctypes.data
        if max_len is not None:
            out = np.empty(max(int(rows.size) * int(max_len), 1), dtype=np.uint8)
            N.check(self._L.scar_fastq_gather_host(self._ctx, rows.ctypes.data, rows.size, fp, out.ctypes.data, out.size,
                                                   off.ctypes.data))
            return out[:int(off[-1])], off
        N.check(self._L.scar_fastq_gather_host(self._ctx, rows.ctypes.data, rows.size, fp, None, 0, off.ctypes.data))
        out = np.empty(max(int(off[-1]), 1), dtype=np.uint8)
        N.check(self._L.scar_fastq_gather_host(self._ctx, rows.ctypes.data, rows.size, fp, out.ctypes.data, int(off[-1]),
                                               off.ctypes.data))
        return out[:int(off[-1])], off

    def table_reserve(self, n_more: int):
        """Room for n_more further scar patterns (the table is rebuilt at a larger size when needed)."""
        N.check(self._L.scar_table_reserve(self._ctx, int(n_more)))

    # ---- device-resident path -------------------------------------------------------------------
    def make_resident(self, reads, f0=None, f1=None):
        """Copy a host batch into HBM (torch tensors) and allocate its scratch/outputs."""
        import torch
        dev = torch.device("cuda", self.device)
        seq, f0, f1 = as_ragged(reads), as_ragged(f0), as_ragged(f1)

        def up(r):
            if r is None:
                return None
            d = {"bytes": torch.from_numpy(np.concatenate([r.bytes, np.zeros(16, np.uint8)])).to(dev),
                 "offsets": None if r.offsets is None else torch.from_numpy(r.offsets).to(dev),
                 "lengths": None if r.lengths is None else torch.from_numpy(r.lengths).to(dev),
                 "stride": r.stride}
            return d

        n = seq.n
        batch = {"n": n, "seq": up(seq), "f0": up(f0), "f1": up(f1),
                 "cells": torch.empty((self.cells_per_read, max(n, 1)), dtype=torch.int64, device=dev),  # cell-major
                 "lens": torch.empty(n, dtype=torch.int32, device=dev),        # stored length | N count << 16
                 "samples": torch.empty(n, dtype=torch.int16, device=dev),
                 "records": torch.empty((n, 16), dtype=torch.uint8, device=dev)}
        return batch

    @staticmethod
    def _dev_ragged(d):
        if d is None:
            return None
        return N.Ragged(d["bytes"].data_ptr(), None if d["offsets"] is None else d["offsets"].data_ptr(),
                        None if d["lengths"] is None else d["lengths"].data_ptr(), d["stride"])

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def demux_resident(self, batch):
        s, a, b = (self._dev_ragged(batch[k]) for k in ("seq", "f0", "f1"))
        N.check(self._L.scar_demux_dev(self._ctx, C.byref(s), C.byref(a) if a else None, C.byref(b) if b else None,
                                       batch["n"], batch["samples"].data_ptr(), self._stream()))

    def pack_resident(self, batch):
        s = self._dev_ragged(batch["seq"])
        N.check(self._L.scar_pack_dev(self._ctx, C.byref(s), batch["n"], batch["cells"].data_ptr(),
                                      batch["lens"].data_ptr(), self._stream()))

    def classify_resident(self, batch):
        N.check(self._L.scar_classify_dev(self._ctx, batch["cells"].data_ptr(), batch["lens"].data_ptr(),
                                          batch["samples"].data_ptr(), batch["n"], batch["records"].data_ptr(),
                                          self._stream()))

    def classify_raw_resident(self, batch, first_index: int = 0, aggregate: bool = True):
        """The fused kernel: raw bytes -> records (+ counters, phase histograms, scar table) in one pass."""
        s = self._dev_ragged(batch["seq"])
        N.check(self._L.scar_classify_raw_dev(self._ctx, C.byref(s), batch["samples"].data_ptr(), batch["n"], first_index,
                                              batch["records"].data_ptr(), 1 if aggregate else 0, self._stream()))

    def aggregate_resident(self, batch, first_index: int = 0):
        s = self._dev_ragged(batch["seq"])
        N.check(self._L.scar_aggregate_dev(self._ctx, C.byref(s), batch["records"].data_ptr(), batch["n"],
                                           first_index, self._stream()))

    def run_resident(self, batch, first_index: int = 0):
        """K2 -> K1 -> K3 -> K4 on the current torch stream; asynchronous."""
        s, a, b = (self._dev_ragged(batch[k]) for k in ("seq", "f0", "f1"))
        N.check(self._L.scar_run_dev(self._ctx, C.byref(s), C.byref(a) if a else None, C.byref(b) if b else None,
                                     batch["n"], first_index, batch["cells"].data_ptr(), batch["lens"].data_ptr(),
                                     batch["samples"].data_ptr(), batch["records"].data_ptr(), self._stream()))

    def verify_resident(self, batch, first_index: int = 0) -> int:
        s = self._dev_ragged(batch["seq"])
        ncol = C.c_int64(0)
        N.check(self._L.scar_verify_dev(self._ctx, C.byref(s), batch["records"].data_ptr(), batch["n"], first_index,
                                        C.byref(ncol), self._stream()))
        return ncol.value

    # ---- multi-GPU building blocks (see distributed.py) -------------------------------------------
    def export_table_dev(self, buf) -> int:
        """Compact this rank's table into `buf` (torch uint8 [cap, 32] on this device); returns the entry count."""
        n = C.c_int64(0)
        N.check(self._L.scar_table_export_dev(self._ctx, buf.data_ptr(), buf.shape[0], C.byref(n), self._stream()))
        return n.value

    def merge_table_dev(self, buf, n: int):
        """Merge `n` foreign entries (count SUM, first_idx MIN) into this rank's table."""
        N.check(self._L.scar_table_merge_dev(self._ctx, buf.data_ptr(), int(n), self._stream()))

    def export_parts_dev(self, regions, n_parts: int):
        """Owner-partitioned export into `regions` (torch uint8 [n_parts * region_capacity, 32])."""
        N.check(self._L.scar_table_export_parts_dev(self._ctx, regions.data_ptr(), n_parts, regions.shape[0] // n_parts,
                                                    self._stream()))

    def export_p2p_dev(self, peer_ptrs, my_rank: int, region_capacity: int):
        """Owner-partitioned export straight into the peers' receive buffers (device pointers mapped into this
        process, one per rank): the exchange is part of the kernel, over NVLink peer memory."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        N.check(self._L.scar_table_export_p2p_dev(self._ctx, arr, len(peer_ptrs), int(my_rank), int(region_capacity),
                                                  self._stream()))

    def clear_table_dev(self):
        N.check(self._L.scar_table_clear_dev(self._ctx, self._stream()))

    def merge_parts_dev(self, regions, n_parts: int):
        N.check(self._L.scar_table_merge_parts_dev(self._ctx, regions.data_ptr(), n_parts, regions.shape[0] // n_parts,
                                                   self._stream()))

    def accumulator_words(self) -> int:
        return int(self._L.scar_accumulator_words(self._ctx))

    def export_accumulators_dev(self, out):
        N.check(self._L.scar_accumulators_export_dev(self._ctx, out.data_ptr(), self._stream()))

    def import_accumulators_dev(self, src):
        N.check(self._L.scar_accumulators_import_dev(self._ctx, src.data_ptr(), self._stream()))

    @staticmethod
    def records_of(batch) -> np.ndarray:
        return batch["records"].cpu().numpy().view(N.RECORD_DTYPE).reshape(-1)

    # ---- results --------------------------------------------------------------------------------
    def counters(self):
        out = np.zeros((self.n_samples, N.SCAR_N_COUNTERS), dtype=np.int64)
        un = C.c_int64(0)
        N.check(self._L.scar_counters_read(self._ctx, out.ctypes.data, C.byref(un)))
        return out, un.value

    def phase_hist(self):
        out = np.zeros((self.n_samples, 2, N.SCAR_PHASE_BINS), dtype=np.int64)
        N.check(self._L.scar_phase_hist_read(self._ctx, out.ctypes.data))
        return out

    def table_size(self) -> int:
        n = C.c_int64(0)
        N.check(self._L.scar_table_size(self._ctx, C.byref(n)))
        return n.value

    def table(self) -> np.ndarray:
        """Entries sorted by (sample, first_idx): the reference's per-sample first-seen dict order."""
        n = self.table_size()
        out = np.zeros(max(n, 1), dtype=N.ENTRY_DTYPE)
        got = C.c_int64(0)
        N.check(self._L.scar_table_read(self._ctx, out.ctypes.data, out.size, C.byref(got)))
        return out[:got.value]


def raw_rows(records, reads, sample: int, platform, target_region: str, cutsite: int, reverse_comp: bool) -> bytes:
    """The body of ScarSearch.raw_data_output (scarmapper/INDEL_Processing.py:702-757) for one sample, built
    natively from the per-read records and the raw reads (host buffers)."""
    r = as_ragged(reads)
    rec = np.ascontiguousarray(records)
    assert rec.dtype == N.RECORD_DTYPE and rec.shape[0] == r.n
    plat = N.PLATFORMS[platform] if isinstance(platform, str) else int(platform)
    reg = target_region.encode("latin-1")
    rs = r.c_struct()
    need = C.c_int64(0)
    N.check(N.lib().scar_raw_rows_host(rec.ctypes.data, C.byref(rs), r.n, int(sample), plat, reg, len(reg), int(cutsite),
                                       1 if reverse_comp else 0, None, 0, C.byref(need)))
    buf = np.empty(max(need.value, 1), dtype=np.uint8)
    N.check(N.lib().scar_raw_rows_host(rec.ctypes.data, C.byref(rs), r.n, int(sample), plat, reg, len(reg), int(cutsite),
                                       1 if reverse_comp else 0, buf.ctypes.data, need.value, C.byref(need)))
    return buf[:need.value].tobytes()


def raw_rows_indexed(records, reads, rows, platform, target_region: str, cutsite: int, reverse_comp: bool) -> bytes:
    """raw_rows for a list of read indices (the scarred reads of one sample, in read order): O(len(rows))."""
    r = as_ragged(reads)
    rec = np.ascontiguousarray(records)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    assert rec.dtype == N.RECORD_DTYPE and rec.shape[0] == r.n
    assert rows.size == 0 or (0 <= int(rows.min()) and int(rows.max()) < r.n)
    plat = N.PLATFORMS[platform] if isinstance(platform, str) else int(platform)
    reg = target_region.encode("latin-1")
    rs = r.c_struct()
    need = C.c_int64(0)
    args = (rec.ctypes.data, C.byref(rs), rows.ctypes.data, rows.size, plat, reg, len(reg), int(cutsite),
            1 if reverse_comp else 0)
    N.check(N.lib().scar_raw_rows_indexed_host(*args, None, 0, C.byref(need)))
    buf = np.empty(max(need.value, 1), dtype=np.uint8)
    N.check(N.lib().scar_raw_rows_indexed_host(*args, buf.ctypes.data, need.value, C.byref(need)))
    return buf[:need.value].tobytes()


def demux_fastq_rows(text, seq_views, rows) -> bytes:
    """"@name\\nseq\\n+\\nqual\\n" of the reads `rows` (file order), formatted from the FASTQ text; seq_views is the
    HostRagged of sequence views over `text` (scarmapper_b200.fastq)."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    assert seq_views.offsets is not None and seq_views.lengths is not None
    need = C.c_int64(0)
    args = (text.ctypes.data, text.size, seq_views.offsets.ctypes.data, seq_views.lengths.ctypes.data, rows.ctypes.data, rows.size)
    N.check(N.lib().scar_demux_fastq_rows_host(*args, None, 0, C.byref(need)))
    buf = np.empty(max(need.value, 1), dtype=np.uint8)
    N.check(N.lib().scar_demux_fastq_rows_host(*args, buf.ctypes.data, need.value, C.byref(need)))
    return buf[:need.value].tobytes()


def stored_gather(reads, rows, platform, flip=None):
    """Stored sequences (platform transform applied; row k reverse-complemented when flip[k]) of the reads `rows`,
    as one CSR buffer: (uint8 bytes, int64 offsets[len(rows) + 1])."""
    r = as_ragged(reads)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    assert rows.size == 0 or (0 <= int(rows.min()) and int(rows.max()) < r.n)
    plat = N.PLATFORMS[platform] if isinstance(platform, str) else int(platform)
    fl = None if flip is None else np.ascontiguousarray(flip, dtype=np.uint8)
    assert fl is None or fl.size == rows.size
    off = np.zeros(rows.size + 1, dtype=np.int64)
    rs = r.c_struct()
    fp = None if fl is None else fl.ctypes.data
    N.check(N.lib().scar_stored_gather_host(C.byref(rs), rows.ctypes.data, rows.size, plat, fp, None, 0, off.ctypes.data))
    out = np.empty(max(int(off[-1]), 1), dtype=np.uint8)
    N.check(N.lib().scar_stored_gather_host(C.byref(rs), rows.ctypes.data, rows.size, plat, fp, out.ctypes.data,
                                            int(off[-1]), off.ctypes.data))
    return out[:int(off[-1])], off


SCAR_TYPES = ("HR", "SNV", "TMEJ", "NHEJ", "Non-MH Deletion", "Insertion", "Other")


def frequency_rows(counts, first_records, first_bytes, first_off, pattern_out, region: str, cutsite: int, reverse_comp: bool,
                   passing: int, no_cut: int, pattern_threshold: float) -> dict:
    """The per-pattern body of ScarSearch.frequency_output for one sample (scar_frequency_rows_host): the rows of the
    Frequency file (bytes), junction_type_data, label sums / order, SNV tallies, plot tuples, totals."""
    n = int(len(counts))
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    first = np.ascontiguousarray(first_records)
    assert first.dtype == N.RECORD_DTYPE and first.shape[0] == n
    rb = np.ascontiguousarray(first_bytes, dtype=np.uint8)
    if rb.size == 0:
        rb = np.zeros(1, np.uint8)
    ro = np.ascontiguousarray(first_off, dtype=np.int64)
    pat = np.ascontiguousarray(pattern_out if pattern_out is not None and n else np.zeros((max(n, 1), 16)), dtype=np.int32)
    reg = np.frombuffer(region.encode("latin-1"), dtype=np.uint8)
    fin = N.FrequencyIn(n, counts.ctypes.data, first.ctypes.data, rb.ctypes.data, ro.ctypes.data, pat.ctypes.data,
                        reg.ctypes.data, len(region), int(cutsite), 1 if reverse_comp else 0, int(passing), int(no_cut),
                        float(pattern_threshold))
    jtd = np.zeros(6, np.int64)
    label_sums, label_order = np.zeros(7, np.float64), np.zeros(7, np.int32)
    snv = np.zeros((4, 10, 4), np.int64)
    plot = np.zeros((max(n, 1), 6), np.float64)
    n_plot, nbytes = C.c_int64(0), C.c_int64(0)
    totals = np.zeros(2, np.int64)
    N.check(N.lib().scar_frequency_rows_host(C.byref(fin), C.byref(nbytes), jtd.ctypes.data, label_sums.ctypes.data,
                                             label_order.ctypes.data, snv.ctypes.data, plot.ctypes.data, plot.shape[0],
                                             C.byref(n_plot), totals.ctypes.data))
    rows = np.empty(max(nbytes.value, 1), dtype=np.uint8)
    N.check(N.lib().scar_frequency_rows_fetch(rows.ctypes.data, rows.size))
    return {"rows": rows[:nbytes.value].tobytes(), "junction_type_data": [int(v) for v in jtd], "label_sums": label_sums,
            "label_order": [int(t) for t in label_order if t >= 0], "snv": snv, "plot": plot[:n_plot.value],
            "scar_count": int(totals[0]), "total_scars": int(totals[1])}


def homeology_batch(consensus, insertions, clj, crj, tlj, trj, target_idx, targets, regions, device: int = 0) -> np.ndarray:
    """ScarSearch.homeology_search + templated_insertion_search (scarmapper/INDEL_Processing.py:550-700) for a
    batch of scar patterns on the GPU: [n, 16] int32 (struct ScarHomeology, include/scar_b200.h)."""
    c, i, t, r = as_ragged(consensus), as_ragged(insertions), as_ragged(targets), as_ragged(regions)
    n = c.n
    assert i.n == n and t.n == r.n
    ints = [np.ascontiguousarray(v, dtype=np.int32) for v in (clj, crj, tlj, trj, target_idx)]
    assert all(v.size == n for v in ints)
    out = np.zeros((n, 16), dtype=np.int32)
    cs, is_, ts, rs = c.c_struct(), i.c_struct(), t.c_struct(), r.c_struct()
    N.check(N.lib().scar_homeology_batch(device, C.byref(cs), C.byref(is_), *[v.ctypes.data for v in ints],
                                         C.byref(ts), C.byref(rs), t.n, n, out.ctypes.data))
    return out


def match_maker_batch(queries, unknowns, device: int = 0) -> np.ndarray:
    """Sequence_Magic.match_maker over a batch on the GPU (Valkyries/Sequence_Magic.py:123-136)."""
    q, u = as_ragged(queries), as_ragged(unknowns)
    assert q.n == u.n
    out = np.zeros(q.n, dtype=np.int32)
    a, b = q.c_struct(), u.c_struct()
    N.check(N.lib().scar_match_maker_batch(device, C.byref(a), C.byref(b), q.n, out.ctypes.data))
    return out
