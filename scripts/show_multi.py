#!/usr/bin/env python
"""print the interesting numbers of a bench.py line (N > 1): python scripts/show_multi.py gpurun_out/r02_bench_n8.json"""
import json, sys
for path in sys.argv[1:]:
    d = [json.loads(l) for l in open(path) if l.startswith('{')][-1]
    print(path, "N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "shots/step", d["run"]["shots_per_step"], d["scaling"])
    print("  multichip_check", d.get("multichip_check"))
    for k in ("strong_config3", "strong_config4"):
        if k in d: print(" ", k, {a: (float("%.4g" % b) if isinstance(b, float) else b) for a, b in d[k].items() if a != "workload"})
    print("  kernels", {k.split(" ")[0]: round(v["ms_per_batch"], 4) for k, v in d["kernels"].items()})
    for k in ("k1", "reference_schedule_k64", "config3"):
        if k in d: print(" ", k, round(d[k]["value"]), "e2e", round(d[k]["e2e"]))
