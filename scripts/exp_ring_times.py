#!/usr/bin/env python
"""ring kernel role timings (RAD_RING_DEBUG=16 [+1]): a few profiled batches of config 2, role stamps on stderr"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosity_b200 import api
scene = api.Scene(0.014)
ctx = api.context_for_scene(scene, 512, 64, select_mode=api.SELECT_TOPK)
for i in range(6):
    print([round(float(x), 4) for x in ctx.profile_batch()], flush=True)
ctx.close()
