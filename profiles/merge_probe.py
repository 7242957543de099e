"""Time the pieces of the owner-partitioned merge in isolation (torchrun, N ranks): export, all-to-all, clear,
fold, accumulator all-reduce.  Diagnostic only (not a bench value)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from scarmapper_b200.distributed import exchange_regions, peer_receive_buffers, reduce_accumulators  # noqa: E402
from scarmapper_b200.engine import HostRagged, ScarEngine, ScarProblem  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 10_000_000
    cfg = synth.make_config("C2", n_reads=n * world, seed=20260101)
    batch = cfg.generate(rank * n, n)
    pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target, left_index=cfg.index_d5,
               right_index=cfg.index_d7, platform="Illumina", phasing=cfg.phasing_lists(), n_limit=cfg.n_limit,
               min_length=cfg.min_length)
    eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 22, **pin), device=local)
    res = eng.make_resident(HostRagged.from_matrix(batch["seq"]), batch["f0"], batch["f1"])
    region = 1 << 20 if world <= 2 else 1 << 18
    send = torch.empty((world * region, 32), dtype=torch.uint8, device="cuda")
    recv = torch.empty_like(send)
    names = ["reset", "compute", "export_parts", "all_to_all", "clear", "merge_parts", "accumulators"]
    tot = {k: 0.0 for k in names}
    reps = 6
    for it in range(reps + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record(); eng.reset_async()
        ev[1].record(); eng.run_resident(res, first_index=rank * n)
        ev[2].record(); eng.export_parts_dev(send, world)
        ev[3].record(); exchange_regions(send, recv)
        ev[4].record(); eng.clear_table_dev()
        ev[5].record(); eng.merge_parts_dev(recv, world)
        ev[6].record(); reduce_accumulators(eng.export_accumulators_dev, eng.import_accumulators_dev,
                                            eng.accumulator_words(), torch.device("cuda", local))
        ev[7].record(); torch.cuda.synchronize()
        if it >= 2:
            for i, k in enumerate(names):
                tot[k] += ev[i].elapsed_time(ev[i + 1])
            if rank == 0:
                print("merge_probe iter", it, " ".join("{:.3f}".format(ev[i].elapsed_time(ev[i + 1])) for i in range(len(names))),
                      flush=True)
    # the same with the export fused with the exchange over peer memory
    got = peer_receive_buffers(world * region, torch.device("cuda", local))
    if got is not None:
        precv, hdl = got
        names2 = ["barrier0", "export_p2p", "barrier1", "clear", "merge_parts"]
        tot2 = {k: 0.0 for k in names2}
        for it in range(reps + 2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names2) + 1)]
            eng.reset_async(); eng.run_resident(res, first_index=rank * n)
            dist.barrier(); torch.cuda.synchronize()
            ev[0].record(); hdl.barrier(channel=0)
            ev[1].record(); eng.export_p2p_dev(hdl.buffer_ptrs, rank, region)
            ev[2].record(); hdl.barrier(channel=1)
            ev[3].record(); eng.clear_table_dev()
            ev[4].record(); eng.merge_parts_dev(precv, world)
            ev[5].record(); torch.cuda.synchronize()
            if it >= 2:
                for i, k in enumerate(names2):
                    tot2[k] += ev[i].elapsed_time(ev[i + 1])
        if rank == 0:
            print("merge_probe p2p world={} ".format(world) + " ".join("{}={:.3f}ms".format(k, v / reps) for k, v in tot2.items()),
                  flush=True)
    if rank == 0:
        print("merge_probe world={} ".format(world) + " ".join("{}={:.3f}ms".format(k, v / reps) for k, v in tot.items()),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
