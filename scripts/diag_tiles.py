"""Diagnostic for the tile-binned rasteriser: quad soup (seed 4 of tests/test_gpu_tiles.py) through the fused path in both
raster modes and through the staged render, every mismatch against the oracle printed with the patches involved."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import random_soup
from radiosity_b200 import api
from oracle import orc

seed, n, size, N = 4, 20000, 0.05, 256
v = random_soup(seed, n, size)
P = v.shape[0]
rng = np.random.default_rng(100 + seed)
shooters = [int(x) for x in rng.choice(P, 8, replace=False)]
c = np.full((P, 3), 0.5, np.float32); r = np.zeros((P, 3), np.float32); il = np.zeros((P, 3), np.float32)
for j, s in enumerate(shooters):
    r[s] = 10.0 * (8 - j)
exp = [orc.render_hemicube(v, sh, N) for sh in shooters]
for mode in ("keys", "tiles"):
    os.environ["RAD_RASTER"] = mode
    ctx = api.Context(N, 8, P, select_mode=api.SELECT_TOPK, flags=api.FLAG_KEEP_ITEMBUFFER)
    ctx.set_formfactors(api.formfactors(N))
    ctx.upload_scene(v, c, r, il)
    ctx.shoot(1)
    fused = [ctx.read_itembuffer(h) for h in range(8)]
    ctx.upload_state(r, il)
    ctx.set_emitters(shooters)
    ctx.render()
    staged = [ctx.read_itembuffer(h) for h in range(8)]
    for h, sh in enumerate(shooters):
        for name, got in (("fused", fused[h]), ("staged", staged[h])):
            bad = np.argwhere(got != exp[h])
            if len(bad):
                print(mode, name, "shooter", sh, "mismatches", len(bad))
                for (y, x) in bad[:5]:
                    g, e = int(got[y, x]), int(exp[h][y, x])
                    print("  pixel", int(x), int(y), "got", g, "exp", e)
                    for pid in (g, e):
                        if pid:
                            print("    patch", pid - 1, v[pid - 1].reshape(4, 3).tolist())
                print("  shooter quad", v[sh].reshape(4, 3).tolist())
    ctx.close()
print("diag done")
