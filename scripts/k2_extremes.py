#!/usr/bin/env python
"""ProcessHemicube (K2, item-buffer form) on the two synthetic extremes SURVEY.md 8d asks for, next to a real batch:
constant-ID atlas (maximum contention: every pixel of a hemicube adds to ONE F entry) and id = hash(px) mod P (no
coherence: every pixel is its own run).  Prints one JSON line.
    python scripts/k2_extremes.py [--workload config2] [--reps 20]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, measured_peak  # noqa: E402
from radiosity_b200 import api  # noqa: E402


def k2_extremes(ctx, P, reps=20, seed=0):
    """ctx: emitters selected and rendered.  Returns {name: ms per launch} for real / constant / hashed item buffers."""
    RES, k = ctx.RES, ctx.k
    out = {"real": ctx.bench_process(reps)}
    const = np.full(RES, (P // 2) + 1, np.uint32)
    for h in range(k):
        ctx.write_itembuffer(h, const)
    out["constant_id"] = ctx.bench_process(reps)
    px = np.arange(RES, dtype=np.uint64)
    for h in range(k):
        hsh = ((px + np.uint64(seed + h * 7919)) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(33)
        ctx.write_itembuffer(h, (hsh % np.uint64(P)).astype(np.uint32) + np.uint32(1))
    out["hashed_id"] = ctx.bench_process(reps)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    area, N, k, _, desc = WORKLOADS[a.workload]
    scene = api.Scene(area)
    ctx = api.context_for_scene(scene, N, k, select_mode=api.SELECT_TOPK if k > 1 else api.SELECT_REFERENCE)
    _, valid = ctx.select(); ctx.render()
    nv = int(np.count_nonzero(valid))
    ms = k2_extremes(ctx, scene.P, a.reps)
    peak, src = measured_peak()
    res = {"workload": desc, "slots": nv, "pixels_per_launch": nv * ctx.RES}
    for name, t in ms.items():
        res[name] = {"ms_per_launch": t, "gpix_per_s": nv * ctx.RES / (t * 1e-3) / 1e9, "frac_of_hbm_peak_at_8B_per_px": nv * ctx.RES * 8 / (t * 1e-3) / 1e9 / peak,
                     "dram_frac_at_4B_per_px": nv * ctx.RES * 4 / (t * 1e-3) / 1e9 / peak}
    print(json.dumps(res))
    ctx.close()
