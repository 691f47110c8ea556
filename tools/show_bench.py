"""print the essentials of bench.py JSON lines: python tools/show_bench.py FILE [FILE...]"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        pr = d.get("pruned") or {}
        print(path, {k: d.get(k) for k in ["value", "ms_per_step", "parity", "gpu_launches"]}, "e2e", d.get("e2e") and round(d["e2e"]["value"]),
              "frac", round(d["roofline"]["frac"], 4), "kms", round(d["roofline"]["kernel_ms"], 2), "pruned", pr.get("value") and round(pr["value"]))
