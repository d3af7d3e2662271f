"""TEST INFRASTRUCTURE — ctypes view of oracle/liboracle.so (the CPU restatement, oracle.cpp) and of
oracle/_ref/libref_host.so (the reference's own host sources, built by ref_build.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp, _u32 = ctypes.c_void_p, ctypes.c_uint32
_orc = None
_ref = None


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def lib():
    global _orc
    if _orc is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle liboracle.so`")
        L = ctypes.CDLL(path)
        L.orc_scene_cornell.argtypes = [ctypes.c_double]; L.orc_scene_cornell.restype = _u32
        L.orc_scene_obj.argtypes = [ctypes.c_char_p, ctypes.c_double]
        L.orc_scene_get.argtypes = [_vp] * 4
        L.orc_config.argtypes = [_u32, _u32, _vp]
        L.orc_formfactors.argtypes = [_u32, _u32, _vp]
        L.orc_colors_setup.argtypes = [_u32, _vp]
        L.orc_color.argtypes = [_u32]; L.orc_color.restype = _u32
        L.orc_color_index.argtypes = [_u32]; L.orc_color_index.restype = _u32
        L.orc_select.argtypes = [_u32, _vp, _u32, ctypes.c_int, _vp, _vp]
        L.orc_patch_geom.argtypes = [_vp] * 4
        L.orc_mvp.argtypes = [_vp, ctypes.c_int, _vp]
        L.orc_projection.argtypes = [_vp]
        L.orc_render_hemicube.argtypes = [_u32, _vp, _u32, _u32, _vp, _vp, ctypes.c_int]
        L.orc_encode_atlas.argtypes = [_u32, _vp, ctypes.c_size_t, _vp]
        L.orc_process_hemicube_cl.argtypes = [_u32, _vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp]
        L.orc_process_hemicube_cl.restype = _u32
        L.orc_gather_records.argtypes = [_u32, _u32, _vp, _vp, _vp, _u32, _vp]; L.orc_gather_records.restype = _u32
        L.orc_process_hemicube_ids.argtypes = [_vp, _vp, _u32, _u32, _u32, _vp]
        L.orc_shoot.argtypes = [_u32, _vp, _vp, _vp, _vp, _u32, _u32, _u32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_float)]
        L.orc_shoot.restype = _u32
        L.orc_smooth_shade.argtypes = [_u32, _vp, _vp, _vp, _vp, _vp]
        _orc = L
    return _orc


def ref():
    """The reference's own host code (None when oracle/_ref was not built / did not travel)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libref_host.so")
        if not os.path.exists(path):
            return None
        L = ctypes.CDLL(path)
        L.refp_scene_build.argtypes = [ctypes.c_double]; L.refp_scene_build.restype = _u32
        L.refp_scene_build_obj.argtypes = [ctypes.c_char_p, ctypes.c_double]; L.refp_scene_build_obj.restype = _u32
        L.refp_scene_get.argtypes = [_vp] * 5
        L.refp_scene_set_radiosity.argtypes = [_vp]
        L.refp_neighbours.argtypes = [_vp]
        L.refp_select.argtypes = [_u32, _vp, _vp]
        L.refp_select_single.restype = _u32
        L.refp_patch_geom.argtypes = [_u32, _vp, _vp, _vp]
        L.refp_mvp.argtypes = [_u32, ctypes.c_int, _vp]
        L.refp_projection.argtypes = [_vp]
        L.refp_config.argtypes = [_u32, _u32, _vp]
        L.refp_formfactors.argtypes = [_u32, _u32, _vp]
        L.refp_colors_setup.argtypes = [_u32, _vp]
        L.refp_color.argtypes = [_u32]; L.refp_color.restype = _u32
        L.refp_color_index.argtypes = [_u32]; L.refp_color_index.restype = _u32
        L.refp_sizeof_patch.restype = _u32
        L.refp_save_rr.argtypes = [ctypes.c_char_p]
        L.refp_load_rr.argtypes = [ctypes.c_char_p]; L.refp_load_rr.restype = _u32
        L.refp_smooth_shade.argtypes = [_vp]
        L.refp_scene_set_illumination.argtypes = [_vp]
        if hasattr(L, "refp_process_hemicube_kernel"):
            L.refp_process_hemicube_kernel.argtypes = [_u32, _vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp]
            L.refp_process_hemicube_kernel.restype = _u32
        _ref = L
    return _ref


# ---- convenience wrappers around the oracle -------------------------------------------------------
def scene_cornell(area):
    L = lib()
    P = L.orc_scene_cornell(float(area))
    v = np.zeros((P, 12), np.float32); c = np.zeros((P, 3), np.float32)
    r = np.zeros((P, 3), np.float32); i = np.zeros((P, 3), np.float32)
    L.orc_scene_get(_ptr(v), _ptr(c), _ptr(r), _ptr(i))
    return v, c, r, i


def scene_obj(path, area):
    L = lib()
    P = L.orc_scene_obj(os.fsencode(path), float(area))
    if P < 0:
        raise IOError(path)
    v = np.zeros((P, 12), np.float32); c = np.zeros((P, 3), np.float32)
    r = np.zeros((P, 3), np.float32); i = np.zeros((P, 3), np.float32)
    L.orc_scene_get(_ptr(v), _ptr(c), _ptr(r), _ptr(i))
    return v, c, r, i


def formfactors(side, hemicubes=1):
    out = np.zeros(3 * side * side * hemicubes, np.float32)
    lib().orc_formfactors(side, hemicubes, _ptr(out))
    return out


def select(rad, count, mode=0):
    rad = np.ascontiguousarray(rad, np.float32)
    ids = np.zeros(count, np.uint32); nul = np.zeros(count, np.int32)
    lib().orc_select(rad.size // 3, _ptr(rad), count, mode, _ptr(ids), _ptr(nul))
    return ids, nul


def mvp(quad12, look):
    q = np.ascontiguousarray(quad12, np.float32); out = np.zeros(16, np.float32)
    lib().orc_mvp(_ptr(q), look, _ptr(out))
    return out


def render_hemicube(verts, shooter, side, threads=1, want_depth=False):
    verts = np.ascontiguousarray(verts, np.float32)
    W, H = 2 * side, side + side // 2
    ids = np.zeros((H, W), np.uint32)
    dep = np.zeros((H, W), np.uint32) if want_depth else None
    lib().orc_render_hemicube(verts.size // 12, _ptr(verts), int(shooter), side, _ptr(ids), _ptr(dep), threads)
    return (ids, dep) if want_depth else ids


def process_ids(ids, ff, side, P, workitems_x=4):
    ids = np.ascontiguousarray(ids, np.uint32); ff = np.ascontiguousarray(ff, np.float32)
    F = np.zeros(P, np.float32)
    lib().orc_process_hemicube_ids(_ptr(ids), _ptr(ff), side, workitems_x, P, _ptr(F))
    return F


def process_cl(ids, ff, side, P, hemicubes=1, workitems_x=4):
    """Reference-format path: RGBA8 colour atlas -> literal kernel restatement -> record gather."""
    L = lib()
    ids = np.ascontiguousarray(ids, np.uint32); ff = np.ascontiguousarray(ff, np.float32)
    n = ids.size
    rgba = np.zeros(n * 4, np.uint8)
    L.orc_encode_atlas(P, _ptr(ids), n, _ptr(rgba))
    h = np.zeros(n + 8, np.uint32); ii = np.zeros(n + 8, np.uint32); e = np.zeros(n + 8, np.float32)
    W, H = 2 * side, side + side // 2
    nrec = L.orc_process_hemicube_cl(P, _ptr(rgba), _ptr(ff), W, H, hemicubes, workitems_x, _ptr(h), _ptr(ii), _ptr(e))
    out = []
    bad = 0
    for hi in range(hemicubes):
        F = np.zeros(P, np.float32)
        bad += L.orc_gather_records(P, nrec, _ptr(h), _ptr(ii), _ptr(e), hi, _ptr(F))
        out.append(F)
    return out, nrec, bad


def process_cl_records(ids, ff, side, P, hemicubes=1, workitems_x=4, reference_kernel=False):
    """The raw record stream (hemicubes[], ids[], energies[], write index) of one kernel launch over an RGBA8 atlas encoded
    from `ids`: the oracle's restatement, or — reference_kernel=True — the reference's OWN kernel text compiled by
    oracle/ref_build.sh and run on the CPU (oracle/ref_kernel.cpp)."""
    L = lib()
    ids = np.ascontiguousarray(ids, np.uint32); ff = np.ascontiguousarray(ff, np.float32)
    n = ids.size
    rgba = np.zeros(n * 4, np.uint8)
    L.orc_encode_atlas(P, _ptr(ids), n, _ptr(rgba))
    h = np.zeros(n + 8, np.uint32); ii = np.zeros(n + 8, np.uint32); e = np.zeros(n + 8, np.float32)
    W, H = 2 * side, side + side // 2
    if reference_kernel:
        nrec = ref().refp_process_hemicube_kernel(P, _ptr(rgba), _ptr(ff), W, H, hemicubes, workitems_x, _ptr(h), _ptr(ii), _ptr(e))
    else:
        nrec = L.orc_process_hemicube_cl(P, _ptr(rgba), _ptr(ff), W, H, hemicubes, workitems_x, _ptr(h), _ptr(ii), _ptr(e))
    return h[:nrec], ii[:nrec], e[:nrec], nrec


def shoot(verts, color, rad, illum, side, k, n_batches, select_mode=0, via_codec=False, stop_test=False, threads=1):
    """Main.cpp:1137-1309 on the CPU.  Returns (rad, illum, schedule[n,k], batches_done, last_energy_len)."""
    verts = np.ascontiguousarray(verts, np.float32); color = np.ascontiguousarray(color, np.float32)
    rad = np.array(rad, np.float32, copy=True, order="C"); illum = np.array(illum, np.float32, copy=True, order="C")
    sched = np.zeros((n_batches, k), np.uint32)
    last = ctypes.c_float()
    n = lib().orc_shoot(verts.size // 12, _ptr(verts), _ptr(color), _ptr(rad), _ptr(illum), side, k, n_batches, select_mode,
                        1 if via_codec else 0, 1 if stop_test else 0, threads, _ptr(sched), ctypes.byref(last))
    return rad, illum, sched, n, last.value


def reference_tail_batches(area, rad, illum, side, k, n_batches):
    """`n_batches` batches of the shooting loop with every host-side step on the reference's OWN code (oracle/_ref):
    ModelContainer::getHighestRadiosityPatchesId for the emitters, the reference's kernel text for the record stream, and
    the text of Main.cpp:1161,1251-1303 (snapshot, record gather, energy transfer, emitter update, stop test; refp_main_tail)
    for the rest.  Only the item buffers come from the oracle's raster restatement (there is no GL here).
    Returns (rad, illum, batches_done, last_energy_len, stopped); stops like Main.cpp:1137 when the stop test fires."""
    R = ref()
    P = int(R.refp_scene_build(ctypes.c_double(area)))
    R.refp_main_tail.argtypes = [_u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]; R.refp_main_tail.restype = ctypes.c_int
    R.refp_scene_get_state.argtypes = [_vp, _vp]
    rad = np.ascontiguousarray(rad, np.float32); illum = np.ascontiguousarray(illum, np.float32)
    R.refp_scene_set_radiosity(_ptr(rad)); R.refp_scene_set_illumination(_ptr(illum))
    v = np.zeros((P, 12), np.float32)
    R.refp_scene_get(_ptr(v), None, None, None, None)
    ff = formfactors(side, k)
    RES = 3 * side * side
    last = ctypes.c_float(); done = 0; stopped = False
    for _ in range(n_batches):
        ids = np.zeros(k, np.uint32); nul = np.zeros(k, np.int32)
        R.refp_select(k, _ptr(ids), _ptr(nul))
        atlas = np.zeros(k * RES, np.uint32)
        for h in range(k):
            if not nul[h]:
                atlas[h * RES:(h + 1) * RES] = render_hemicube(v, int(ids[h]), side).ravel()
        rh, ri, re, nrec = process_cl_records(atlas, ff, side, P, hemicubes=k, reference_kernel=True)
        rh = np.ascontiguousarray(rh); ri = np.ascontiguousarray(ri); re = np.ascontiguousarray(re)
        go = R.refp_main_tail(k, _ptr(ids), _ptr(nul), int(nrec), _ptr(rh), _ptr(ri), _ptr(re), ctypes.byref(last))
        done += 1
        if not go:
            stopped = True
            break
    out_r = np.zeros((P, 3), np.float32); out_i = np.zeros((P, 3), np.float32)
    R.refp_scene_get_state(_ptr(out_r), _ptr(out_i))
    return out_r, out_i, done, last.value, stopped


def smooth_shade(color, rad, illum, nb8):
    color = np.ascontiguousarray(color, np.float32); rad = np.ascontiguousarray(rad, np.float32)
    illum = np.ascontiguousarray(illum, np.float32); nb8 = np.ascontiguousarray(nb8, np.int32)
    P = color.size // 3
    out = np.zeros((P, 12), np.float32)
    lib().orc_smooth_shade(P, _ptr(color), _ptr(rad), _ptr(illum), _ptr(nb8), _ptr(out))
    return out


def max_threads():
    return int(lib().orc_max_threads())
