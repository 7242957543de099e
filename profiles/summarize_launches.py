"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[start]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        k = r[kn].split("(")[0][:64]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,avg_us,share_pct")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("{},{},{:.1f},{:.1f},{:.1f}".format(k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
