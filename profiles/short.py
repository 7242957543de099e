"""Print the key fields of bench.py JSON lines: `short.py FILE [LABEL]` (or stdin when FILE is '-')."""
import json
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "-"
label = sys.argv[2] if len(sys.argv) > 2 else ""
src = sys.stdin if path == "-" else open(path)
for line in src:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    st = {k: round(v, 3) for k, v in d.get("stages_ms", {}).items()}
    raw = d.get("from_raw_bytes", {})
    print(label, "n_gpus", d.get("n_gpus"), "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], st,
          "frac %.3f" % d.get("roofline", {}).get("frac", 0), "e2e %.4g" % (d["e2e"]["value"] or 0),
          "raw %.4g (%.3f ms)" % (raw.get("value", 0), raw.get("ms_per_step", 0)), "launches", d.get("gpu_launches"))
