"""Basic-block execution summary of an `ncu --page source --csv` dump (per-warp execution counts)."""
import csv
import sys


def main(path, n_warps):
    rows = list(csv.reader(open(path)))
    sec, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            sec.append(cur)
            continue
        if r and r[0] == "Address":
            cur["hdr"] = r
            continue
        if cur is not None and len(r) > 5:
            cur["rows"].append(r)
    s = sec[0]
    h = s["hdr"]
    ia, isrc, ismp, ith = (h.index(k) for k in ("Instructions Executed", "Source", "# Samples", "Thread Instructions Executed"))
    blocks = []
    for i, r in enumerate(s["rows"]):
        c, th = int(r[ia]), int(r[ith])
        if blocks and abs(blocks[-1]["c"] - c) <= 0.03 * max(c, 1):
            b = blocks[-1]
            b["n"] += 1; b["tot"] += c; b["smp"] += int(r[ismp]); b["th"] += th
        else:
            blocks.append({"i": i, "c": c, "n": 1, "tot": c, "smp": int(r[ismp]), "first": r[isrc][:60], "th": th})
    tot = sum(b["tot"] for b in blocks)
    print(s["name"], "total warp instr", tot, "per warp", tot / n_warps)
    for b in blocks:
        if b["tot"] > 0.01 * tot:
            print("idx %4d n=%3d exec/warp=%6.2f tot%%=%5.1f avgthreads=%4.1f smp=%6d  %s" % (
                b["i"], b["n"], b["c"] / n_warps, 100 * b["tot"] / tot, b["th"] / max(b["tot"], 1), b["smp"], b["first"]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
