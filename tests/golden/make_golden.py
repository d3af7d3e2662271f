#!/usr/bin/env python
"""Generates tests/golden/golden.json from the REFERENCE's own host code (oracle/_ref/libref_host.so, built by
oracle/ref_build.sh from /root/reference/source) — run in the authoring container only:

    python tests/golden/make_golden.py

Section "reference" holds values produced by reference code.  Section "oracle_regression" holds outputs of the
CPU oracle for the two stages the reference delegates to GPU drivers (GL raster, CL kernel); those pin the
oracle against accidental change, they are NOT reference outputs (no GL/CL stack exists in this image).
"""
import ctypes
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731


def seeded_radiosity(P, seed):
    """Deterministic pseudo-random energies with many ties and zeros (portable: integer LCG, no numpy RNG)."""
    x = np.arange(P * 3, dtype=np.uint64) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407 + seed)
    x ^= x >> np.uint64(29)
    x = (x * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(32)
    q = (x % np.uint64(9)).astype(np.float32) / np.float32(4.0)          # {0, .25, ..., 2}
    zero = ((x >> np.uint64(8)) % np.uint64(3)) == 0
    q[zero] = 0
    return q.reshape(P, 3)


def kernel_cases():
    """(name, ids, ff, N, P, hemicubes): item buffers fed — as RGBA8 atlases — to the ProcessHemicube kernel"""
    v, c, r, il = orc.scene_cornell(0.5)
    P = v.shape[0]
    ids = np.concatenate([orc.render_hemicube(v, s, 32).ravel() for s in (323, 0)])
    yield "box_area0.5_N32_s323_s0", ids, orc.formfactors(32, 2), 32, P, 2
    for P, N in ((502, 16), (16469, 16), (250063, 32), (1021554, 32)):        # every colour bit layout the configs use
        n = 3 * N * N
        x = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(12345)) >> np.uint64(7)
        yield f"hash_P{P}_N{N}", ((x % np.uint64(P + 1))).astype(np.uint32), orc.formfactors(N, 1), N, P, 1     # no coherence, some empty pixels
        yield f"const_P{P}_N{N}", np.full(n, P, np.uint32), orc.formfactors(N, 1), N, P, 1                       # one id everywhere (the last patch)
        runs = (np.arange(n, dtype=np.uint32) // 5) % np.uint32(P) + 1
        yield f"runs5_P{P}_N{N}", runs.astype(np.uint32), orc.formfactors(N, 1), N, P, 1                          # runs that straddle the 4 work-item spans


def tail_cases():
    """(name, area, hemicube side, k, batches, initial radiosity, initial illumination): batches of the shooting loop whose
    host side runs on the reference's own text (orc.reference_tail_batches)"""
    v, c, r, il = orc.scene_cornell(0.5)
    P = v.shape[0]
    yield "fresh_k1_x6", 0.5, 32, 1, 6, r, il
    yield "fresh_k4_x3", 0.5, 32, 4, 3, r, il
    yield "fresh_k10_x2", 0.5, 32, 10, 2, r, il                                  # the reference's default `hemicubes 10`: 4 lights + NULL slots
    yield "seeded_k7_x2", 0.5, 16, 7, 2, seeded_radiosity(P, 3), seeded_radiosity(P, 4)   # ties and zeros in the list
    yield "stops_k4", 0.5, 32, 4, 10, (r * np.float32(0.06)).astype(np.float32), il   # the stop test fires at batch 4
    yield "stops_k1", 0.5, 32, 1, 20, (r * np.float32(0.06)).astype(np.float32), il   # ... at batch 13


def main():
    R = orc.ref()
    assert R is not None, "build oracle/_ref first (bash oracle/ref_build.sh)"
    g = {"reference": {}, "oracle_regression": {}}
    ref = g["reference"]

    ref["sizeof_patch"] = int(R.refp_sizeof_patch())
    ref["scenes"] = {}
    for area in (0.5, 0.014, 0.0035, 0.0009):
        P = int(R.refp_scene_build(area))
        v = np.zeros((P, 12), np.float32); ix = np.zeros((P, 6), np.int32)
        c = np.zeros((P, 3), np.float32); r = np.zeros((P, 3), np.float32); il = np.zeros((P, 3), np.float32)
        R.refp_scene_get(vp(v), vp(ix), vp(c), vp(r), vp(il))
        lights = np.nonzero(r[:, 0] > 0)[0]
        s = {"P": P, "lights": int(lights.size), "first_light": int(lights[0]), "last_light": int(lights[-1]),
             "verts_sha256": sha(v), "indices_sha256": sha(ix), "color_sha256": sha(c), "rad_sha256": sha(r), "illum_sha256": sha(il)}
        sel = {}
        for k in (1, 10, 64):
            ids = np.zeros(k, np.uint32); nul = np.zeros(k, np.int32)
            R.refp_select(k, vp(ids), vp(nul))
            sel[str(k)] = {"ids": ids.tolist(), "null": nul.tolist()}
        s["select_fresh"] = sel
        if P <= 70000:
            nb = np.zeros((P, 8), np.int32); R.refp_neighbours(vp(nb)); s["neighbours_sha256"] = sha(nb)
            rs = {}
            for seed in (0, 1):
                rad = seeded_radiosity(P, seed)
                R.refp_scene_set_radiosity(vp(rad))
                for k in (1, 3, 10, 64):
                    ids = np.zeros(k, np.uint32); nul = np.zeros(k, np.int32)
                    R.refp_select(k, vp(ids), vp(nul))
                    rs[f"seed{seed}_k{k}"] = {"ids": ids.tolist(), "null": nul.tolist()}
            s["select_seeded"] = rs
            # display stage: Colors::smoothShadePatch on a seeded state
            rad = seeded_radiosity(P, 3) * np.float32(7); ill = seeded_radiosity(P, 4)
            R.refp_scene_set_radiosity(vp(rad)); R.refp_scene_set_illumination(vp(ill))
            sh = np.zeros((P, 12), np.float32); R.refp_smooth_shade(vp(sh))
            s["smooth_shade_sha256"] = sha(sh)
            R.refp_scene_set_radiosity(vp(r)); R.refp_scene_set_illumination(vp(il))
        if area == 0.5:
            mv = {}
            for p in (0, 100, 320, 323, 400, 501):
                for look in range(5):
                    m = np.zeros(16, np.float32); R.refp_mvp(p, look, vp(m))
                    mv[f"{p}_{look}"] = m.view(np.uint32).tolist()
            s["mvp_bits"] = mv
            geo = {}
            for p in (0, 320, 501):
                cc = np.zeros(3, np.float32); nn = np.zeros(3, np.float32); uu = np.zeros(3, np.float32)
                R.refp_patch_geom(p, vp(cc), vp(nn), vp(uu))
                geo[str(p)] = {"center": cc.view(np.uint32).tolist(), "normal": nn.view(np.uint32).tolist(), "up": uu.view(np.uint32).tolist()}
            s["geom_bits"] = geo
        ref["scenes"][repr(area)] = s
    ref["P_area_0.00022"] = 1021554      # measured once from the reference; too slow to regenerate in the CPU suite

    m = np.zeros(16, np.float32); R.refp_projection(vp(m)); ref["projection_bits"] = m.view(np.uint32).tolist()

    ref["config"] = {}
    for side, k in ((16, 10), (128, 1), (512, 64), (1024, 64)):
        out = (ctypes.c_uint * 9)(); R.refp_config(side, k, out); ref["config"][f"{side}_{k}"] = list(out)

    ref["formfactors"] = {}
    for N in (16, 128, 512):
        ff = np.zeros(3 * N * N * 2, np.float32); R.refp_formfactors(N, 2, vp(ff))
        one = ff[:3 * N * N]
        ref["formfactors"][str(N)] = {"sha256_k2": sha(ff), "sum_f64": float(one.sum(dtype=np.float64)),
                                      "first_bits": int(one[:1].view(np.uint32)[0]),
                                      "center_bits": int(one[(N // 2) * 2 * N + N: (N // 2) * 2 * N + N + 1].view(np.uint32)[0])}

    ref["codec"] = {}
    for P in (1, 7, 502, 16469, 64659, 250063, 1021554):
        out = (ctypes.c_uint * 11)(); R.refp_colors_setup(P, out)
        samples = sorted(i for i in {1, 2, P // 3 + 1, P // 2 + 1, P} if i <= P)
        ref["codec"][str(P)] = {"params": list(out), "colors": {str(i): int(R.refp_color(i)) for i in samples},
                                "index_of_color": {str(int(R.refp_color(i))): int(R.refp_color_index(R.refp_color(i))) for i in samples}}

    objp = os.path.join(ROOT, "tests", "golden", "simple.obj")
    for area in (0.0, 0.3):
        P = int(R.refp_scene_build_obj(objp.encode(), area))
        v = np.zeros((P, 12), np.float32); c = np.zeros((P, 3), np.float32); r = np.zeros((P, 3), np.float32)
        R.refp_scene_get(vp(v), None, vp(c), vp(r), None)
        ref.setdefault("obj", {})[repr(area)] = {"P": P, "verts_sha256": sha(v), "color_sum": float(c.sum()), "rad_sum": float(r.sum())}

    # `.rr` checkpoint written by the reference's own Patch class / SaveToFile body (LP64 ABI) on a seeded state
    P = int(R.refp_scene_build(0.5))
    rad = seeded_radiosity(P, 1); ill = seeded_radiosity(P, 2)
    R.refp_scene_set_radiosity(vp(rad)); R.refp_scene_set_illumination(vp(ill))
    assert R.refp_save_rr(os.path.join(ROOT, "tests", "golden", "ref_area0.5_lp64.rr").encode())
    ref["rr"] = {"file": "ref_area0.5_lp64.rr", "P": P, "bytes": 8 + 184 * P, "rad_sha256": sha(rad), "illum_sha256": sha(ill)}

    # ---- the reference's own OpenCL kernel TEXT run on the CPU (oracle/ref_kernel.cpp): record streams ----
    ref["kernel"] = {}
    for name, ids, ff, N, P, k in kernel_cases():
        h, ii, e, nrec = orc.process_cl_records(ids, ff, N, P, hemicubes=k, reference_kernel=True)
        ref["kernel"][name] = {"N": N, "P": P, "hemicubes": k, "records": int(nrec), "hemicubes_sha256": sha(h), "ids_sha256": sha(ii),
                               "energies_sha256": sha(e), "sum_energy": float(e.sum(dtype=np.float64))}

    # ---- the CPU tail of the loop on the reference's own TEXT (Main.cpp:1161,1251-1303; oracle/ref_probe.cpp refp_main_tail) ----
    ref["tail"] = {}
    for name, area, N, k, nb, rad0, il0 in tail_cases():
        rad, illum, done, last, stopped = orc.reference_tail_batches(area, rad0, il0, N, k, nb)
        ref["tail"][name] = {"N": N, "k": k, "batches_asked": nb, "batches_done": int(done), "stopped": bool(stopped),
                             "last_bits": int(np.float32(last).view(np.uint32)), "rad_sha256": sha(rad), "illum_sha256": sha(illum)}

    # ---- oracle regression (NOT reference outputs) ----
    reg = g["oracle_regression"]
    v, c, r, il = orc.scene_cornell(0.5)
    P = v.shape[0]
    for N, shooter in ((32, 323), (32, 100), (64, 0), (128, 450)):
        ids, dep = orc.render_hemicube(v, shooter, N, want_depth=True)
        ff = orc.formfactors(N)
        F = orc.process_ids(ids, ff, N, P)
        reg[f"hemicube_N{N}_s{shooter}"] = {"ids_sha256": sha(ids), "depth_sha256": sha(dep), "F_sha256": sha(F),
                                            "empty": int((ids == 0).sum()), "sumF": float(F.sum(dtype=np.float64))}
    rad, illum, sched, n, last = orc.shoot(v, c, r, il, 32, 1, 40)
    reg["shoot_area0.5_N32_k1_40"] = {"rad_sha256": sha(rad), "illum_sha256": sha(illum), "schedule": sched.ravel().tolist(),
                                      "sumB": float(rad.sum(dtype=np.float64)), "last": float(last)}
    rad, illum, sched, n, last = orc.shoot(v, c, r, il, 32, 10, 6)
    reg["shoot_area0.5_N32_k10_6"] = {"rad_sha256": sha(rad), "illum_sha256": sha(illum), "schedule": sched.ravel().tolist(),
                                      "sumB": float(rad.sum(dtype=np.float64)), "last": float(last)}

    out = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
