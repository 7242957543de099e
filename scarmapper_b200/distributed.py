"""
distributed.py — multi-GPU plumbing: one process per GPU, reads sharded, one exchange at the end.

The reference's only parallelism is one pathos process per sample library
(INDEL_Processing.py:1047-1059).  Here every read is independent through demux / pack / classify,
so the global read stream is cut into contiguous blocks, rank g owning reads
[shard_bounds(n, G)[g], shard_bounds(n, G)[g+1]) with GLOBAL read indices (the reference keeps the
first-seen record per scar key and orders ties by first-seen rank, so first_idx must be global).
No data-path collective runs until the tables are complete; then

  1. dense accumulators (unassigned, per-sample counters, phase histograms): all_reduce(SUM, int64)
  2. scar tables (sparse): all_gather of entry counts, all_gather of the compacted 32-byte entries
     padded to the largest rank, and every rank folds the other ranks' entries into its own table
     on its GPU (count SUM, first_idx MIN) — so all ranks end with the identical merged table.

Payload is unique-keys x 32 B (KB..MB): latency-bound over NVLink 5 / NVSwitch.  torch.distributed
is the transport (NCCL on GPUs, gloo in the CPU tests); the fold itself is the library's merge kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_reads: int, world: int):
    """Contiguous blocks: [b[g], b[g+1]) for rank g; sizes differ by at most one."""
    base, rem = divmod(int(n_reads), int(world))
    b = [0]
    for g in range(world):
        b.append(b[-1] + base + (1 if g < rem else 0))
    return b


def merge_tables(export_fn, merge_fn, entry_buf: torch.Tensor, group=None):
    """All-gather merge of sparse tables.

    export_fn(buf) -> n      compacts the local table into buf ([cap, 32] uint8) and returns the count
    merge_fn(buf, n)         folds n foreign entries into the local table
    entry_buf                scratch on the rank's device, at least as large as the largest local table

    Returns the list of per-rank entry counts.  Every rank folds every other rank's entries, so all
    ranks hold the same merged table afterwards.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = int(export_fn(entry_buf))
    counts = torch.zeros(world, dtype=torch.int64, device=entry_buf.device)
    mine = torch.tensor([n_local], dtype=torch.int64, device=entry_buf.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_h = [int(c) for c in counts.tolist()]
    n_max = max(counts_h)
    if n_max > entry_buf.shape[0]:
        raise RuntimeError("entry buffer too small: {} < {}".format(entry_buf.shape[0], n_max))
    if world == 1 or n_max == 0:
        return counts_h
    flat = torch.empty((world * n_max, entry_buf.shape[1]), dtype=entry_buf.dtype, device=entry_buf.device)
    dist.all_gather_into_tensor(flat, entry_buf[:n_max].contiguous(), group=group)
    gathered = flat.view(world, n_max, entry_buf.shape[1])
    for g in range(world):
        if g != rank and counts_h[g]:
            merge_fn(gathered[g], counts_h[g])
    return counts_h


def reduce_accumulators(export_fn, import_fn, words: int, device, group=None):
    """all_reduce(SUM) of the dense int64 accumulator block."""
    buf = torch.empty(words, dtype=torch.int64, device=device)
    export_fn(buf)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    import_fn(buf)
    return buf


def exchange_regions(send: torch.Tensor, recv: torch.Tensor, group=None):
    """Region o of `send` goes to rank o; region g of `recv` comes from rank g.  One all-to-all on NCCL; on
    backends without it (gloo, CPU tests) the same exchange is spelled with an all-gather."""
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, group=group)
        return
    rank = dist.get_rank(group)
    per = send.shape[0] // world
    everything = torch.empty((world * send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(everything, send.contiguous(), group=group)
    for g in range(world):
        src = everything[g * send.shape[0] + rank * per: g * send.shape[0] + (rank + 1) * per]
        recv[g * per:(g + 1) * per] = src


def merge_tables_partitioned(export_parts_fn, clear_fn, merge_parts_fn, send: torch.Tensor, recv: torch.Tensor, group=None):
    """Owner-partitioned merge: every rank ends up with the merged entries it owns.  Stream-ordered, no host
    synchronisation: the per-region counts travel inside the regions."""
    world = dist.get_world_size(group)
    export_parts_fn(send, world)
    if dist.get_backend(group) == "nccl":
        # the table is cleared (it only needs the export to be done) while the regions travel
        work = dist.all_to_all_single(recv, send, group=group, async_op=True)
        clear_fn()
        work.wait()                      # stream-ordered: the current stream waits for NCCL's, the host does not
    else:
        exchange_regions(send, recv, group)
        clear_fn()
    merge_parts_fn(recv, world)


def peer_receive_buffers(rows: int, device, group=None):
    """A receive buffer of `rows` 32-byte entries on every rank, each mapped into every other rank's address
    space (torch symmetric memory: CUDA IPC over NVLink).  Returns (local tensor, handle) or None when the
    backend cannot do it (CPU tests, no peer access)."""
    try:
        import torch.distributed._symmetric_memory as symm_mem
        if dist.get_backend(group) != "nccl":
            return None
        pg = group if group is not None else dist.group.WORLD
        buf = symm_mem.empty((rows, 32), dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, pg)
        if len(hdl.buffer_ptrs) != dist.get_world_size(group):
            return None
        return buf, hdl
    except Exception as exc:  # noqa: BLE001 - any failure here means "use the all-to-all instead"
        import warnings
        warnings.warn("peer-memory receive buffers unavailable ({}); using NCCL all-to-all".format(exc))
        return None


class DistributedScar:
    """Glue between a ScarEngine and the process group."""

    def __init__(self, engine, entry_capacity: int = 1 << 20, group=None):
        self.engine = engine
        self.group = group
        self.device = torch.device("cuda", engine.device)
        self.entry_buf = torch.empty((entry_capacity, 32), dtype=torch.uint8, device=self.device)
        self._send = self._recv = None
        self._peer = None          # (recv tensor, symmetric-memory handle, region_capacity) of the fused export
        # Export fused with the exchange over NVLink peer memory (scar_table_export_p2p_dev).  Off by default: in
        # isolation it costs the same as export + NCCL all-to-all (0.23 ms per 10 M reads at N=2), but its remote
        # stores hold SMs longer than NCCL's copy does and bench.py runs the merge under the next step's kernels
        # (1.25e10 vs 1.32e10 reads/s at N=2).
        self.use_peer_memory = False

    def _peer_buffers(self, world: int, region_capacity: int):
        if self._peer is None or self._peer[2] != region_capacity:
            got = peer_receive_buffers(world * region_capacity, self.device, self.group) if self.use_peer_memory else None
            self._peer = (got[0], got[1], region_capacity) if got else (None, None, region_capacity)
        return self._peer

    def merge_partitioned(self, region_capacity: int = 1 << 18):
        """Owner-partitioned merge (cost independent of the GPU count): afterwards this rank's table holds the
        merged entries whose key it owns; the dense accumulators are all-reduced.  Fully asynchronous.
        With use_peer_memory (NCCL on NVLink) the export kernel stores every entry directly into its owner's
        receive buffer - export and exchange are one kernel, fenced by two device-side barriers; otherwise the
        regions are exchanged with one all-to-all."""
        e = self.engine
        world = dist.get_world_size(self.group)
        recv, hdl, _ = self._peer_buffers(world, region_capacity)
        if hdl is not None:
            hdl.barrier(channel=0)                     # every peer has folded what it received last time
            e.export_p2p_dev(hdl.buffer_ptrs, dist.get_rank(self.group), region_capacity)
            hdl.barrier(channel=1)                     # every peer's stores into this rank's buffer have landed
            e.clear_table_dev()
            e.merge_parts_dev(recv, world)
            reduce_accumulators(e.export_accumulators_dev, e.import_accumulators_dev, e.accumulator_words(),
                                self.device, self.group)
            return
        if self._send is None or self._send.shape[0] != world * region_capacity:
            self._send = torch.empty((world * region_capacity, 32), dtype=torch.uint8, device=self.device)
            self._recv = torch.empty_like(self._send)
        merge_tables_partitioned(e.export_parts_dev, e.clear_table_dev, e.merge_parts_dev, self._send, self._recv,
                                 self.group)
        reduce_accumulators(e.export_accumulators_dev, e.import_accumulators_dev, e.accumulator_words(),
                            self.device, self.group)

    def merge(self):
        e = self.engine
        reduce_accumulators(e.export_accumulators_dev, e.import_accumulators_dev, e.accumulator_words(),
                            self.device, self.group)
        return merge_tables(e.export_table_dev, e.merge_table_dev, self.entry_buf, self.group)
