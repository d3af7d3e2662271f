#!/usr/bin/env python
"""Experiment: does the raster queue kernel get faster when the key buffers it updates stay L2-resident?
profile_batch runs the stages of one batch back to back on one stream (lanes off); RAD_L2_GROUP_MB caps the key
footprint of a launch group (groups recycle the same buffers).  Prints ms per batch per stage."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from radiosity_b200 import api
    k = int(sys.argv[2])
    scene = api.Scene(0.014)
    ctx = api.context_for_scene(scene, 512, k, select_mode=api.SELECT_TOPK)
    ctx.save_state()
    for _ in range(4): ctx.profile_batch()
    ctx.restore_state()
    st = np.zeros(6)
    n = 16
    for _ in range(n): st += ctx.profile_batch()
    st /= n
    print(json.dumps({"k": k, "group_mb": os.environ.get("RAD_L2_GROUP_MB"), "per_batch_ms": [round(float(x), 4) for x in st], "per_hemicube_us": [round(float(x) * 1e3 / k, 3) for x in st]}))
    ctx.close()
else:
    for k, mb in ((64, None), (64, 101), (64, 51), (64, 26), (64, 13), (16, None), (8, None), (4, None)):
        env = dict(os.environ, RAD_LANES="1")
        if mb: env["RAD_L2_GROUP_MB"] = str(mb)
        subprocess.run([sys.executable, __file__, "child", str(k)], env=env)
