"""Time scar_homeology_batch (host buffers in, results out) on synthetic scar patterns; diagnostic, prints one line."""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scarmapper_b200.engine import homeology_batch  # noqa: E402


def patterns(n, seed=1):
    rng = random.Random(seed)
    region = "".join(rng.choice("ACGT") for _ in range(472))
    cons, inss, clj, crj, tlj, trj = [], [], [], [], [], []
    for _ in range(n):
        cut = 236
        dl, dr = rng.randint(0, 40), rng.randint(0, 40)
        ins = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 0, 0, 1, 2, 6, 12])))
        lf, rf = region[cut - dl - 70:cut - dl], region[cut + dr:cut + dr + 70]
        cons.append(lf + ins + rf); inss.append(ins)
        clj.append(len(lf)); crj.append(len(lf) + len(ins)); tlj.append(cut - dl); trj.append(cut + dr)
    return region, cons, inss, clj, crj, tlj, trj


def main():
    n = int(os.environ.get("HOM_KEYS", "1000000"))
    region, cons, inss, clj, crj, tlj, trj = patterns(n)
    tix = np.zeros(n, dtype=np.int32)
    homeology_batch(cons[:1000], inss[:1000], clj[:1000], crj[:1000], tlj[:1000], trj[:1000], tix[:1000], [region], [region])
    t0 = time.perf_counter()
    out = homeology_batch(cons, inss, clj, crj, tlj, trj, tix, [region], [region])
    dt = time.perf_counter() - t0
    from oracle_engine import host_homeology_batch
    m = 20000
    t0 = time.perf_counter()
    want = host_homeology_batch(cons[:m], inss[:m], clj[:m], crj[:m], tlj[:m], trj[:m], tix[:m], [region], [region])
    dh = time.perf_counter() - t0
    assert np.array_equal(out[:m], want)
    print("homeology_bench keys={} gpu_batch_s={:.3f} ({:.3g} keys/s incl. host packing and copies) "
          "host_core_ctypes_per_key_us={:.2f}".format(n, dt, n / dt, 1e6 * dh / m), flush=True)


if __name__ == "__main__":
    main()
