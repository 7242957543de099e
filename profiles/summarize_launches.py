"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.

bench.py launches the kernels on several batch shapes (the resident 10 M-read batch of the timed `value`
region, the same batch from raw bytes incl. pack, the 1 M-read chunks of the host pipeline, the FASTQ-text
path).  Per kernel, launches are clustered by duration (a new cluster where the sorted durations jump by
more than 1.5x).  The `value` steps are recognised in launch order as demux -> window -> aggregate with no
pack in between, so the kernels' SHARES of a resident step can be compared with bench.py's stages_ms
(per-launch times under ncu are serialised and cold-cache: the share must agree, not the absolute)."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[start]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    seq = [(r[kn].split("(")[0][:64], float(r[mv].replace(",", "")) / 1e3) for r in rows[start + 1:] if len(r) > mv]
    per = {}
    for k, us in seq:
        per.setdefault(k, []).append(us)
    print("# per kernel, clustered by duration")
    print("kernel,cluster,launches,avg_us,min_us,max_us")
    for k, v in sorted(per.items(), key=lambda x: -sum(x[1])):
        v = sorted(v, reverse=True)
        clusters, cur = [], [v[0]]
        for x in v[1:]:
            if cur[-1] > 1.5 * x:
                clusters.append(cur)
                cur = []
            cur.append(x)
        clusters.append(cur)
        for i, c in enumerate(clusters):
            print("{},{},{},{:.1f},{:.1f},{:.1f}".format(k, i, len(c), sum(c) / len(c), min(c), max(c)))
    # value steps: demux, window, aggregate back to back on the full batch
    wmax = max(us for k, us in seq if "window" in k)
    steps = []
    for i in range(len(seq) - 2):
        a, b, c = seq[i], seq[i + 1], seq[i + 2]
        if "demux" in a[0] and "window" in b[0] and "aggregate" in c[0] and b[1] > wmax / 1.5:
            steps.append((a[1], b[1], c[1]))
    if steps:
        n = len(steps)
        k2, k3, k4 = (sum(s[j] for s in steps) / n for j in range(3))
        tot = k2 + k3 + k4
        print("# resident `value` steps found: {}".format(n))
        print("kernel,avg_us,share_of_step_pct")
        for name, x in (("scar_demux_kernel (K2)", k2), ("scar_window_kernel (K3)", k3), ("scar_aggregate_kernel (K4)", k4)):
            print("{},{:.1f},{:.1f}".format(name, x, 100 * x / tot))


if __name__ == "__main__":
    main(sys.argv[1])
