"""Print the key fields of a bench.py JSON line read from stdin (helper for gpurun one-liners)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    st = {k: round(v, 3) for k, v in d.get("stages_ms", {}).items()}
    print(sys.argv[1] if len(sys.argv) > 1 else "", "n_gpus", d.get("n_gpus"), "value %.4g" % d["value"],
          "ms/step %.3f" % d["ms_per_step"], st, "frac %.3f" % d.get("roofline", {}).get("frac", 0),
          "e2e %.4g" % d["e2e"]["value"], "launches", d.get("gpu_launches"))
