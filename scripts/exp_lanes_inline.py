#!/usr/bin/env python
"""lanes vs micro-triangle scenes (inline tier): graph shots/s of config 4 and of config 5 at hemicube 256 for RAD_LANES=1/2/4/8"""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from radiosity_b200 import api
    area, N, k, nb = float(sys.argv[2]), int(sys.argv[3]), 64, int(sys.argv[4])
    scene = api.Scene(area)
    ctx = api.context_for_scene(scene, N, k, select_mode=api.SELECT_TOPK)
    ctx.save_state(); ctx.shoot(nb); ctx.restore_state()
    st = ctx.shoot(nb)
    print(json.dumps({"area": area, "N": N, "P": scene.P, "lanes": os.environ.get("RAD_LANES"), "batch_ms": round(st.gpu_ms / nb, 4), "shots_per_s": round(st.shots_done / st.gpu_ms * 1e3, 1)}), flush=True)
    ctx.close()
else:
    for area, N, nb in ((0.0035, 256, 16), (0.00022, 1024, 16), (0.0009, 1024, 16)):
        for lanes in ("1", "2", "4", "8"):
            subprocess.run([sys.executable, __file__, "child", str(area), str(N), str(nb)], env=dict(os.environ, RAD_LANES=lanes))
