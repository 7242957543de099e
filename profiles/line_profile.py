"""profiles/line_profile.py — where a kernel's warp instructions and stall samples go, by SOURCE line.

    python profiles/line_profile.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING N_READS [--top 40] [--inner]

`ncu --page source --csv` lists per-SASS-instruction execution counts and stall samples but no source lines; the
lines come from `nvdisasm -gi` of the cubin inside OBJECT.o (compiled with -lineinfo): every instruction is tagged with
its innermost line and the chain of inlining sites, the last of which lies in the kernel body.  The two listings are
joined by instruction order (opcodes are cross-checked).  Output: warp instructions per read and stall samples
aggregated by kernel-body line (default) or by innermost line (--inner), largest first.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel_sub):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    out, name, hdr, take = [], None, None, False
    for r in rows:
        if r and r[0] == "Kernel Name":
            take = kernel_sub in r[1] and name is None
            if take:
                name = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if take and hdr and len(r) > 5:
            out.append(r)
    return name, hdr, out


def line_info(obj, mangled_sub):
    d = tempfile.mkdtemp(prefix="lineprof_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL, check=True)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
    ins, on, tags = [], False, []
    for ln in txt:
        if ln.startswith("\t.section\t.text."):
            on = mangled_sub in ln
            tags = []
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            tags.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if tags:
                cur = list(tags)
                tags = []
            ins.append((m.group(2).strip(), cur if "cur" in dir() else []))
            last = ins[-1][1]
            # instructions without a fresh tag inherit the previous one
            if not last and len(ins) > 1:
                ins[-1] = (ins[-1][0], ins[-2][1])
    return ins


def mangle(kernel_name):
    # "void scar_window_kernel<(bool)1, (bool)0, (int)2, (int)10>(K3Params)" -> ILb1ELb0ELi2ELi10EE
    base = re.match(r"(?:void )?(\w+)", kernel_name).group(1)
    args = re.findall(r"\((bool|int)\)(-?\d+)", kernel_name)
    tpl = "I" + "".join("L{}{}E".format("b" if t == "bool" else "i", v) for t, v in args) + "E" if args else ""
    return "{}{}{}".format(len(base), base, tpl)


def main():
    rep, obj, ksub, n_reads = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    inner = "--inner" in sys.argv
    name, hdr, rows = sass_rows(rep, ksub)
    if not rows:
        raise SystemExit("kernel not found in the report")
    ins = line_info(obj, mangle(name))
    if len(ins) != len(rows):
        print("warning: {} SASS rows in the report, {} instructions in the cubin (joined by order)".format(len(rows), len(ins)))
    ie, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    tot_i = tot_s = 0
    for k, r in enumerate(rows):
        tags = ins[k][1] if k < len(ins) else []
        if not tags:
            key = ("?", 0)
        else:
            key = tags[0] if inner else tags[-1]
        a = agg.setdefault(key, {"i": 0, "s": 0, "st": {}})
        c, s = int(r[ie]), int(r[ismp])
        a["i"] += c; a["s"] += s
        tot_i += c; tot_s += s
        for i, h in stall_cols:
            v = int(r[i]) if r[i] else 0
            if v:
                a["st"][h] = a["st"].get(h, 0) + v
    print("{}\nwarp instructions {:.0f} = {:.2f} per read; stall samples {}".format(name, tot_i, tot_i / n_reads, tot_s))
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
        f, ln = key
        if f not in src_cache:
            try:
                src_cache[f] = open(f).read().split("\n")
            except OSError:
                src_cache[f] = []
        text = src_cache[f][ln - 1].strip()[:70] if 0 < ln <= len(src_cache[f]) else ""
        st = ", ".join("{} {:.0f}%".format(h[6:], 100.0 * v / max(a["s"], 1)) for h, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
        print("{:>22s}:{:<5d} instr/read {:6.2f} ({:4.1f}%)  samples {:5.1f}%  [{}]  {}".format(
            os.path.basename(f), ln, a["i"] / n_reads, 100.0 * a["i"] / tot_i, 100.0 * a["s"] / max(tot_s, 1), st, text))


if __name__ == "__main__":
    main()
