"""
World-size-2 gloo test (CPU) of the multi-GPU plumbing in scarmapper_b200/distributed.py:
contiguous sharding with global read indices, the all-gather merge of sparse scar tables
(count SUM, first_idx MIN) and the all-reduce of the dense accumulators.  The device-side export /
merge kernels are replaced by numpy stand-ins with the same contract; what is under test is the
collective plumbing (sizes, padding, which rank folds what) and that every rank ends with the table a
single process would have built.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRY = np.dtype([("h1", "<u8"), ("h2", "<u8"), ("first_idx", "<u8"), ("count", "<u4"), ("sample", "<u2"), ("pad", "<u2")])


def test_shard_bounds():
    from scarmapper_b200.distributed import shard_bounds
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(200_000_000, 8)[-1] == 200_000_000
    b = shard_bounds(7, 8)
    assert b[-1] == 7 and all(0 <= b[i + 1] - b[i] <= 1 for i in range(8))


def synthetic_keys(n, seed):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 500, n)            # 500 possible scar patterns, heavy repeats
    samples = keys % 7
    return keys.astype(np.uint64), samples.astype(np.uint16)


class NumpyTable:
    """Stand-in for the device table: key -> [count, first_idx, sample]."""

    def __init__(self):
        self.d = {}

    def insert(self, h1, h2, count, first, sample):
        e = self.d.get((h1, h2))
        if e is None:
            self.d[(h1, h2)] = [count, first, sample]
        else:
            e[0] += count
            e[1] = min(e[1], first)

    def export(self, buf):
        arr = np.zeros(len(self.d), dtype=ENTRY)
        for i, ((h1, h2), (c, f, s)) in enumerate(sorted(self.d.items())):
            arr[i] = (h1, h2, f, c, s, 0)
        raw = torch.from_numpy(arr.view(np.uint8).reshape(-1, 32).copy())
        buf[:raw.shape[0]] = raw
        return raw.shape[0]

    def export_parts(self, regions, n_parts):
        per = regions.shape[0] // n_parts
        regions.zero_()
        parts = [[] for _ in range(n_parts)]
        for (h1, h2), (c, f, s) in sorted(self.d.items()):
            parts[(h2 >> 40) % n_parts].append((h1, h2, f, c, s, 0))
        for o, rows in enumerate(parts):
            arr = np.zeros(per, dtype=ENTRY)
            arr[0]["h1"] = len(rows)
            if rows:
                arr[1:1 + len(rows)] = np.array(rows, dtype=ENTRY)
            regions[o * per:(o + 1) * per] = torch.from_numpy(arr.view(np.uint8).reshape(-1, 32).copy())

    def clear(self):
        self.d = {}

    def merge_parts(self, regions, n_parts):
        per = regions.shape[0] // n_parts
        for g in range(n_parts):
            arr = regions[g * per:(g + 1) * per].contiguous().numpy().reshape(-1).view(ENTRY)
            for e in arr[1:1 + int(arr[0]["h1"])]:
                self.insert(int(e["h1"]), int(e["h2"]), int(e["count"]), int(e["first_idx"]), int(e["sample"]))

    def merge(self, buf, n):
        arr = buf[:n].contiguous().numpy().reshape(-1).view(ENTRY)
        for e in arr:
            self.insert(int(e["h1"]), int(e["h2"]), int(e["count"]), int(e["first_idx"]), int(e["sample"]))


def worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    from scarmapper_b200.distributed import merge_tables, merge_tables_partitioned, reduce_accumulators, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys, samples = synthetic_keys(n_total, 5)
    b = shard_bounds(n_total, world)
    tab, tab2 = NumpyTable(), NumpyTable()
    acc = np.zeros(1 + 7 * 16, dtype=np.int64)
    for r in range(b[rank], b[rank + 1]):                 # GLOBAL read indices
        tab.insert(int(keys[r]), (int(keys[r]) * 0x9E3779B97F4A7C15) & (2 ** 64 - 1), 1, r, int(samples[r]))
        tab2.insert(int(keys[r]), (int(keys[r]) * 0x9E3779B97F4A7C15) & (2 ** 64 - 1), 1, r, int(samples[r]))
        acc[1 + int(samples[r]) * 16] += 1
    acc[0] = rank + 1
    buf = torch.zeros((1024, 32), dtype=torch.uint8)
    counts = merge_tables(tab.export, tab.merge, buf)
    holder = {}

    def exp(t):
        t.copy_(torch.from_numpy(acc))

    def imp(t):
        holder["acc"] = t.numpy().copy()

    reduce_accumulators(exp, imp, acc.size, torch.device("cpu"))
    # owner-partitioned merge on a second copy: the union over ranks must be the same table, disjointly
    send = torch.zeros((world * 512, 32), dtype=torch.uint8)
    recv = torch.zeros_like(send)
    merge_tables_partitioned(tab2.export_parts, tab2.clear, tab2.merge_parts, send, recv)
    part = sorted((k[0], v[0], v[1], v[2]) for k, v in tab2.d.items())
    np.save(os.path.join(out_dir, "part{}.npy".format(rank)), np.array(part, dtype=np.int64).reshape(-1, 4))
    final = sorted((k[0], v[0], v[1], v[2]) for k, v in tab.d.items())
    np.save(os.path.join(out_dir, "rank{}.npy".format(rank)), np.array(final, dtype=np.int64))
    np.save(os.path.join(out_dir, "acc{}.npy".format(rank)), holder["acc"])
    np.save(os.path.join(out_dir, "counts{}.npy".format(rank)), np.array(counts))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_merge_over_gloo_equals_single_process(tmp_path, world):
    n_total = 5000
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    keys, samples = synthetic_keys(n_total, 5)
    ref = {}
    for r in range(n_total):
        e = ref.setdefault(int(keys[r]), [0, r, int(samples[r])])
        e[0] += 1
    want = np.array(sorted((k, v[0], v[1], v[2]) for k, v in ref.items()), dtype=np.int64)
    for rank in range(world):
        got = np.load(tmp_path / "rank{}.npy".format(rank))
        assert (got == want).all(), "rank {} table differs from the single-process table".format(rank)
        acc = np.load(tmp_path / "acc{}.npy".format(rank))
        assert acc[0] == world * (world + 1) // 2
        assert acc[1:].reshape(7, 16)[:, 0].sum() == n_total
        counts = np.load(tmp_path / "counts{}.npy".format(rank))
        assert len(counts) == world and counts.sum() >= len(ref)
    parts = np.concatenate([np.load(tmp_path / "part{}.npy".format(r)) for r in range(world)])
    assert len(parts) == len(want)                       # ownership is a partition: no key on two ranks
    assert (parts[np.argsort(parts[:, 0])] == want).all()
