"""
Multi-GPU parity check, launched by torchrun (one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank classifies its contiguous shard with GLOBAL read indices, then the tables are merged both
ways (all-gather fold: every rank holds the full table; owner-partitioned all-to-all: every rank holds
the entries it owns).  Rank 0 compares both with the single-GPU table of the whole batch (which the
single-GPU tests pin against the oracle and the goldens): entries, counts, first_idx, counters.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from scarmapper_b200.distributed import DistributedScar, shard_bounds  # noqa: E402
from scarmapper_b200.engine import HostRagged, ScarEngine, ScarProblem  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 400_000
    cfg = synth.make_config("C3", n_reads=n, seed=99)
    batch = cfg.generate(0, n)
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target, left_index=cfg.index_d5,
               right_index=cfg.index_d7, platform="Illumina", phasing=cfg.phasing_lists(), n_limit=cfg.n_limit,
               min_length=cfg.min_length)
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]

    def shard_engine():
        e = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 20, **pin), device=local)
        e.process_host(HostRagged.from_matrix(batch["seq"][lo:hi]), batch["f0"][lo:hi], batch["f1"][lo:hi],
                       first_index=lo, want_records=False)
        return e

    # (a) all-gather fold
    ea = shard_engine()
    DistributedScar(ea, entry_capacity=1 << 19).merge()
    torch.cuda.synchronize()
    tab_a = ea.table()
    cnt_a, un_a = ea.counters()
    # (b) owner-partitioned, export fused with the exchange over NVLink peer memory (twice: the second merge
    #     reuses the peer buffers and must see the barriers)
    eb = shard_engine()
    db = DistributedScar(eb)
    db.use_peer_memory = True
    db.merge_partitioned(region_capacity=1 << 17)
    torch.cuda.synchronize()
    peer_path = db._peer is not None and db._peer[1] is not None
    tab_b = eb.table()
    cnt_b, un_b = eb.counters()
    # (c) owner-partitioned with the NCCL all-to-all
    ec = shard_engine()
    dc = DistributedScar(ec)
    dc.merge_partitioned(region_capacity=1 << 17)
    torch.cuda.synchronize()
    tab_c = ec.table()
    cnt_c, un_c = ec.counters()
    same_bc = tab_b.tobytes() == tab_c.tobytes() and bool((cnt_b == cnt_c).all()) and un_b == un_c
    flags = [None] * world
    dist.all_gather_object(flags, (same_bc, peer_path))
    sizes = [None] * world
    dist.all_gather_object(sizes, tab_b)
    ok = True
    if rank == 0:
        ref = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 20, **pin), device=local)
        ref.process_host(HostRagged.from_matrix(batch["seq"]), batch["f0"], batch["f1"], want_records=False)
        want = ref.table()
        cnt, un = ref.counters()
        full_b = np.concatenate(sizes)
        full_b = full_b[np.lexsort((full_b["first_idx"], full_b["sample"]))]
        checks = {
            "allgather table": tab_a.tobytes() == want.tobytes(),
            "allgather counters": bool((cnt_a == cnt).all()) and un_a == un,
            "partitioned table": full_b.tobytes() == want.tobytes(),
            "partitioned counters": bool((cnt_b == cnt).all()) and un_b == un,
            "partitions disjoint": len(full_b) == len(want),
            "peer-memory path == all-to-all path": all(f[0] for f in flags),
            "peer-memory path taken": all(f[1] for f in flags),
        }
        ok = all(checks.values())
        print("multi_gpu_check world={} entries={} {}".format(world, len(want), checks), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
