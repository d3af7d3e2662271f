"""RED.MIN.64 rate of the rasteriser's access pattern as a function of the key-buffer footprint (k x 3N^2 x 8 B)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosity_b200 import api
s = api.Scene(0.014)
for k, N in ((1, 512), (2, 512), (4, 512), (8, 512), (12, 512), (16, 512), (32, 512), (64, 512)):
    ctx = api.context_for_scene(s, N, k, select_mode=api.SELECT_TOPK if k > 1 else 0)
    print('k', k, 'keys MB', round(k * 3 * N * N * 8 / 1e6, 1), {pat: round(max(ctx.bench_atomics(pat, 1 << 27) for _ in range(2)), 1) for pat in (0, 1, 2)}, 'G RED.MIN.64/s (0 raster-like, 1 coalesced, 2 scattered)', flush=True)
    ctx.close()
