#!/usr/bin/env python
"""Runs a few batches of a bench workload — the short command wrapped by ncu (see profiles/README.md).
    python scripts/prof_batches.py [--workload config2] [--batches 4] [--process-reps 0]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from radiosity_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config2")
ap.add_argument("--batches", type=int, default=4)
ap.add_argument("--process-reps", type=int, default=0)
a = ap.parse_args()
area, N, k, _, desc = WORKLOADS[a.workload]
scene = api.Scene(area)
ctx = api.context_for_scene(scene, N, k, select_mode=api.SELECT_TOPK if k > 1 else api.SELECT_REFERENCE)
ap_warm = ctx.shoot(a.batches) if a.batches >= 16 else None      # graph capture + warm-up outside the reported time
st = ctx.shoot(a.batches)
print(desc, "|", st.batches_done, "batches", st.gpu_ms, "ms", st.kernel_launches, "launches, big triangles in last batch:", st.big_triangles)
if a.process_reps:
    ctx.select(); ctx.render()
    print("process ms/launch", ctx.bench_process(a.process_reps))
ctx.close()
