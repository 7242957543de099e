"""Single-GPU timing of the kernels of the owner-partitioned merge (export into n_parts regions, clear, fold of the same
regions back) on the table a 10 M-read C2 batch leaves: python profiles/merge_local_probe.py [n_parts ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from scarmapper_b200.engine import HostRagged, ScarEngine, ScarProblem  # noqa: E402

n = 10_000_000
cfg = synth.make_config("C2", n_reads=n, seed=20260101)
batch = cfg.generate(0, n)
pin = dict(targets=cfg.targets, cutsites=cfg.cutsites, sample_target=cfg.sample_target, left_index=cfg.index_d5,
           right_index=cfg.index_d7, platform="Illumina", phasing=cfg.phasing_lists(), n_limit=cfg.n_limit,
           min_length=cfg.min_length)
eng = ScarEngine(ScarProblem(max_read_len=cfg.read_len, table_capacity=1 << 22, **pin), device=0)
res = eng.make_resident(HostRagged.from_matrix(batch["seq"]), batch["f0"], batch["f1"])
for parts in [int(a) for a in sys.argv[1:]] or [2, 8]:
    region = (1 << 21) // parts + 1024
    send = torch.empty((parts * region, 32), dtype=torch.uint8, device="cuda")
    names = ["export_parts", "clear", "merge_parts"]
    tot = {k: 0.0 for k in names}
    reps = 8
    for it in range(reps + 2):
        eng.reset_async(); eng.run_resident(res)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); eng.export_parts_dev(send, parts)
        ev[1].record(); eng.clear_table_dev()
        ev[2].record(); eng.merge_parts_dev(send, parts)
        ev[3].record(); torch.cuda.synchronize()
        if it >= 2:
            for i, k in enumerate(names):
                tot[k] += ev[i].elapsed_time(ev[i + 1])
    print("merge_local_probe parts={} entries={} ".format(parts, len(eng.table())) +
          " ".join("{}={:.3f}ms".format(k, v / reps) for k, v in tot.items()), flush=True)
