"""cProfile of the launcher (drop-in mode of baseline/run_pipeline.py) on a synthetic gz FASTQ: where the wall time of
the pipeline goes once the hot path runs on the GPU.  python profiles/launcher_profile.py [reads]"""
import cProfile
import io
import json
import os
import pstats
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import run_pipeline as rp
import synth

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
wd = tempfile.mkdtemp(prefix="scar_prof_")
cfg = synth.make_config("C2", n_reads=reads, seed=20260101)
rp.write_inputs(wd, cfg, reads)
out_dir = os.path.join(wd, "out_dropin") + os.sep
os.makedirs(out_dir)
opt = rp.write_options(wd, out_dir, max(1, min(cfg.n_samples, (os.cpu_count() or 2) - 1)), cfg, verbose="ERROR")
pr = cProfile.Profile()
pr.enable()
res = rp.run_dropin(opt, 0)
pr.disable()
print(json.dumps(res))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
shutil.rmtree(wd, ignore_errors=True)
