#!/usr/bin/env python
"""BASELINE config 5: hemicube resolution sweep 256-2048 at fixed 64 659 patches (built-in scene, area 0.0035), k = 64:
K1 (raster set-up + chunks) and K2 (ProcessHemicube, item-buffer form) reported separately per resolution.
    python scripts/sweep_config5.py [--sides 256 512 1024 2048] > gpurun_out/config5.json"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosity_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sides", type=int, nargs="+", default=[256, 512, 1024, 2048])
ap.add_argument("--area", type=float, default=0.0035)
ap.add_argument("--k", type=int, default=64)
ap.add_argument("--raster", default="keys", help="keys (global key buffers, default path) or tiles (tile-binned rasteriser)")
a = ap.parse_args()
os.environ["RAD_RASTER"] = a.raster
peak = 6545.3
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
scene = api.Scene(a.area)
rows = []
for N in a.sides:
    ctx = api.context_for_scene(scene, N, a.k, select_mode=api.SELECT_TOPK)
    ctx.save_state()
    ctx.shoot(2)                                   # warm-up
    ctx.restore_state()
    prof = np.zeros(6)
    nb = 3
    for _ in range(nb):
        prof += ctx.profile_batch()
    prof /= nb
    ctx.restore_state()
    st = ctx.shoot(16)                             # the real thing: CUDA graph, raster lanes
    graph_ms = st.gpu_ms / 16
    ctx.restore_state(); ctx.select(); ctx.render()
    k2_ms = ctx.bench_process(10)
    RES = 3 * N * N
    px = a.k * RES
    rows.append({"hemicube": N, "patches": scene.P, "k": a.k, "raster": a.raster, "pixels_per_batch": px,
                 "graph_batch_ms": graph_ms, "graph_shots_per_s": a.k / (graph_ms * 1e-3),
                 "raster_setup_ms": float(prof[1]), "raster_chunks_ms": float(prof[2]), "fused_process_ms": float(prof[4]),
                 "select_ms": float(prof[0]), "apply_ms": float(prof[5]),
                 "batch_ms": float(prof.sum()), "shots_per_s": a.k / (float(prof.sum()) * 1e-3),
                 "K1_us_per_hemicube": float((prof[1] + prof[2]) * 1e3 / a.k),
                 "K1_Mpatch_faces_per_s": 5.0 * scene.P * a.k / ((prof[1] + prof[2]) * 1e-3) / 1e6,
                 "K2_itembuffer_ms": k2_ms, "K2_Gpix_per_s": px / (k2_ms * 1e-3) / 1e9,
                 "K2_frac_of_hbm_peak_at_8B_per_px": 8.0 * px / (k2_ms * 1e-3) / 1e9 / peak})
    ctx.close()
print(json.dumps({"workload": f"config5: built-in scene area {a.area} (P={scene.P}), k={a.k}", "hbm_peak_gbs": peak, "rows": rows}, indent=1))
