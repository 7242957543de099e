"""
Run one golden case through scarmapper_b200.launcher.main in a FRESH process, with the reference's real process
model: pathos.multiprocessing.Pool = multiprocess.Pool (forked workers, dill pickling), Spawn > 1.

    python tests/run_launcher_case.py <golden name> <work dir> <spawn> <cuda|oracle>

cuda   : the real ScarEngine (GPU box);  oracle : the CPU test double (build container).
A separate process because the reference's Logger holds sys.stderr, which pytest replaces with an unpicklable
capture object, and because it is how the pipeline is actually started.  Used by tests/test_dropin_forking_pool.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    name, wd, spawn, engine = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    import golden_util as gu
    import make_golden
    from oracle import ref_shims
    case = gu.load(name)["case"]
    ref_shims.install(real_pool=True)
    from scarmapper_b200 import dropin, launcher
    if engine == "oracle":
        from oracle_engine import OracleEngine, host_homeology_batch
        dropin._ENGINE_FACTORY[0] = OracleEngine
        dropin._HOMEOLOGY_BATCH[0] = host_homeology_batch
    wd = wd.rstrip(os.sep) + os.sep
    make_golden.write_inputs(wd, case)
    gu.write_options_file(wd, case, spawn=spawn)
    launcher.main(["--options_file", wd + "run.sh", "--reference", ref_shims.REFERENCE])
    print("launcher ok: {} spawn={} engine={} reference={} served={} missed={}".format(
        name, spawn, engine, ref_shims.REFERENCE, dropin.PATTERN_SEARCH_STATS["served"], dropin.PATTERN_SEARCH_STATS["missed"]))


if __name__ == "__main__":
    main()
