"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

bench.py launches every kernel on two batch sizes: the resident 10 M-read batch (the timed `value`
region: demux, window, aggregate; pack is timed on its own) and the 1 M-read chunks of the host pipeline
(the `e2e` arm).  Launches are split by duration (>= half of the kernel's longest launch = full batch) so
that the kernel's SHARE of a resident step can be compared with bench.py's stages_ms."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[start]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    per = {}
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        per.setdefault(r[kn].split("(")[0][:64], []).append(float(r[mv].replace(",", "")))
    for title, pick in (("full-batch launches (10 M reads; pack is outside the timed value region)", True),
                        ("host-pipeline chunk launches (1 M reads, e2e arm)", False)):
        agg = {}
        for k, v in per.items():
            big = max(v)
            sel = [x for x in v if (x >= 0.5 * big) == pick]
            if sel:
                agg[k] = (len(sel), sum(sel))
        tot = sum(a[1] for a in agg.values())
        step = sum(a[1] for k, a in agg.items() if "pack" not in k)
        print("# " + title)
        print("kernel,launches,total_us,avg_us,share_of_all_pct,share_of_step_without_pack_pct")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            print("{},{},{:.1f},{:.1f},{:.1f},{}".format(k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot,
                                                       "" if "pack" in k else "{:.1f}".format(100 * a[1] / step)))


if __name__ == "__main__":
    main(sys.argv[1])
