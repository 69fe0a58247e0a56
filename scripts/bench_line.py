"""Prints the headline fields of a bench.py JSON line read from stdin (helper for A/B runs)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(" ".join(sys.argv[1:]), "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
      "ms", round(d["e2e"]["ms_per_step"], 3), "launches", d.get("gpu_launches"), d["config"].get("cuda_graph", "")[:58])
