#!/usr/bin/env python
"""Time-boxed parity sweep on the GPU: many random shooters per scene through the staged render, every item buffer
compared with the oracle's bit for bit.  Scenes: quad soups with small shooters, the built-in box at 16 k / 250 k / 1 M
patches.  Prints one JSON object (mismatching shooters listed with their first differing pixels).
    python scripts/fuzz_parity.py [--seconds 30] [--raster keys|tiles] > gpurun_out/fuzz_parity.json"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--raster", default="keys")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--out", default="", help="also write the (partial) result here after every scene")
a = ap.parse_args()
os.environ["RAD_RASTER"] = a.raster

from radiosity_b200 import api  # noqa: E402
from oracle import orc  # noqa: E402
from test_gpu_parity import random_soup  # noqa: E402

threads = a.threads or orc.max_threads()
t_start = time.time()
scenarios = [("soup seed 5, 30000 quads of 0.03", lambda: random_soup(5, 30000, 0.03), 256),
             ("box, 250 063 patches", lambda: orc.scene_cornell(0.0009)[0], 512),
             ("box, 16 469 patches", lambda: orc.scene_cornell(0.014)[0], 512),
             ("soup seed 6, 8000 quads of 0.02", lambda: random_soup(6, 8000, 0.02), 128),
             ("box, 1 021 554 patches", lambda: orc.scene_cornell(0.00022)[0], 256)]
share = a.seconds / len(scenarios)
rows = []
for si, (name, make, N) in enumerate(scenarios):
    deadline = t_start + share * (si + 1)
    if time.time() > deadline:
        continue
    v = make()
    P = v.shape[0]
    c = np.full((P, 3), 0.5, np.float32); z = np.zeros((P, 3), np.float32)
    k = 16
    ctx = api.Context(N, k, P)
    ctx.set_formfactors(api.formfactors(N))
    ctx.upload_scene(v, c, z, z)
    rng = np.random.default_rng(1000 + si)
    row = {"scene": name, "patches": int(P), "hemicube": N, "shooters": 0, "pixels": 0, "mismatched_pixels": 0, "bad": []}
    while time.time() < deadline:
        shooters = [int(x) for x in rng.choice(P, k, replace=False)]
        ctx.set_emitters(shooters)
        ctx.render()
        for h, sh in enumerate(shooters):
            if time.time() > deadline:
                break
            got = ctx.read_itembuffer(h)
            exp = orc.render_hemicube(v, sh, N, threads=threads)
            bad = np.argwhere(got != exp)
            row["shooters"] += 1; row["pixels"] += int(exp.size); row["mismatched_pixels"] += int(len(bad))
            if len(bad) and len(row["bad"]) < 8:
                row["bad"].append({"shooter": sh, "n": int(len(bad)),
                                   "first": [[int(x), int(y), int(got[y, x]), int(exp[y, x])] for y, x in bad[:4]]})
    ctx.close()
    rows.append(row)
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"raster": a.raster, "oracle_threads": threads, "seconds": time.time() - t_start, "rows": rows}, f, indent=1)
print(json.dumps({"raster": a.raster, "oracle_threads": threads, "seconds": time.time() - t_start, "rows": rows}, indent=1))
