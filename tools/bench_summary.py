#!/usr/bin/env python3
"""Print the headline numbers and the per-kernel roofline table of a bench.py JSON line.  usage: bench_summary.py file.json"""
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches", "ranks_identical")}, "e2e", d["e2e"]["value"])
r = d.get("roofline") or {}
print("dominant:", r.get("kernel"), "frac %.3f" % r.get("frac", 0), r.get("bound"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
for k in r.get("per_kernel", []):
    print("  %-62s n=%2d %7.3f ms  frac %.2f  %s" % (k["kernel"], k["launches_per_step"], k["ms_per_step"], k["frac"], k["bound"]))
for key in ("sr2", "inference"):
    if key in d:
        print(key, json.dumps(d[key])[:400])
