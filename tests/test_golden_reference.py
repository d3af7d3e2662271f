"""CPU: the oracle restatement AND the product host library against tests/golden/golden.json (values produced by the
reference's own host code, see tests/golden/make_golden.py), and — when oracle/_ref is present — against the
reference's code live.  Everything here is bit-exact."""
import ctypes
import os

import numpy as np
import pytest

from util import sha, bits, seeded_radiosity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AREAS = (0.5, 0.014, 0.0035)
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


@pytest.fixture(scope="module", params=AREAS)
def area(request):
    return request.param


def _scene_checks(g, v, c, r, il):
    assert v.shape[0] == g["P"]
    assert sha(v) == g["verts_sha256"] and sha(c) == g["color_sha256"]
    assert sha(r) == g["rad_sha256"] and sha(il) == g["illum_sha256"]
    lights = np.nonzero(r[:, 0] > 0)[0]
    assert lights.size == g["lights"] and lights[0] == g["first_light"] and lights[-1] == g["last_light"]


def test_oracle_scene(orc, golden, area):
    _scene_checks(golden["reference"]["scenes"][repr(area)], *orc.scene_cornell(area))


def test_host_scene(api, golden, area):
    s = api.Scene(area)
    g = golden["reference"]["scenes"][repr(area)]
    v, ix, c, r, il = s.arrays()
    _scene_checks(g, v, c, r, il)
    assert sha(ix) == g["indices_sha256"]
    assert sha(s.neighbours()) == g["neighbours_sha256"]


def test_patch_counts_large(orc, api, golden):
    g = golden["reference"]["scenes"]["0.0009"]
    v, c, r, il = orc.scene_cornell(0.0009)
    assert v.shape[0] == g["P"] == 250063 and sha(v) == g["verts_sha256"]
    s = api.Scene(0.0009)
    assert s.P == g["P"] and sha(s.arrays()[0]) == g["verts_sha256"]


def test_select_fresh_and_seeded(orc, api, golden, area):
    g = golden["reference"]["scenes"][repr(area)]
    v, c, r, il = orc.scene_cornell(area)
    s = api.Scene(area)
    for k, exp in g["select_fresh"].items():
        ids, nul = orc.select(r, int(k), 0)
        assert ids.tolist() == exp["ids"] and nul.tolist() == exp["null"]
        ids, nul = s.select(int(k))
        assert ids.tolist() == exp["ids"] and nul.tolist() == exp["null"]
    for key, exp in g["select_seeded"].items():
        seed, k = int(key.split("_")[0][4:]), int(key.split("_k")[1])
        rad = seeded_radiosity(v.shape[0], seed)
        ids, nul = orc.select(rad, k, 0)
        assert ids.tolist() == exp["ids"] and nul.tolist() == exp["null"], key
        s.set_state(rad=rad)
        ids, nul = s.select(k)
        assert ids.tolist() == exp["ids"] and nul.tolist() == exp["null"], key


def test_display_stage_smooth_shade(orc, api, golden, area):
    """Colors::smoothShadePatch — oracle restatement and host library against the reference's own code."""
    g = golden["reference"]["scenes"][repr(area)]
    s = api.Scene(area)
    v, _, c, r, il = s.arrays()
    rad = seeded_radiosity(s.P, 3) * np.float32(7); ill = seeded_radiosity(s.P, 4)
    assert sha(orc.smooth_shade(c, rad, ill, s.neighbours())) == g["smooth_shade_sha256"]
    s.set_state(rad, ill)
    assert sha(s.smooth_shade()) == g["smooth_shade_sha256"]


def test_known_answers_from_survey(golden):
    sc = golden["reference"]["scenes"]
    assert sc["0.5"]["select_fresh"]["10"]["ids"] == [323, 321, 320, 322, 0, 0, 0, 0, 0, 0]
    assert sc["0.5"]["select_fresh"]["10"]["null"] == [0, 0, 0, 0, 0, 1, 1, 1, 1, 1]
    assert sc["0.014"]["select_fresh"]["1"]["ids"] == [11331] and sc["0.0035"]["select_fresh"]["1"]["ids"] == [44574]
    assert [sc[a]["P"] for a in ("0.5", "0.014", "0.0035", "0.0009")] == [502, 16469, 64659, 250063]
    assert [sc[a]["lights"] for a in ("0.5", "0.014", "0.0035", "0.0009")] == [4, 99, 396, 1540]
    assert golden["reference"]["sizeof_patch"] == 184


def test_mvp_and_geometry(orc, api, golden):
    g = golden["reference"]["scenes"]["0.5"]
    v, c, r, il = orc.scene_cornell(0.5)
    s = api.Scene(0.5)
    for key, exp in g["mvp_bits"].items():
        p, look = (int(x) for x in key.split("_"))
        assert bits(orc.mvp(v[p], look)).tolist() == exp, key
        assert bits(s.mvp(p, look)).tolist() == exp, key
    proj = np.zeros(16, np.float32); orc.lib().orc_projection(vp(proj))
    assert bits(proj).tolist() == golden["reference"]["projection_bits"]
    assert bits(api.projection()).tolist() == golden["reference"]["projection_bits"]
    for p, exp in g["geom_bits"].items():
        cc = np.zeros(3, np.float32); nn = np.zeros(3, np.float32); uu = np.zeros(3, np.float32)
        orc.lib().orc_patch_geom(vp(np.ascontiguousarray(v[int(p)])), vp(cc), vp(nn), vp(uu))
        assert bits(cc).tolist() == exp["center"] and bits(nn).tolist() == exp["normal"] and bits(uu).tolist() == exp["up"]
        api.host_lib().radhost_patch_geom(s.h, int(p), vp(cc), vp(nn), vp(uu))
        assert bits(cc).tolist() == exp["center"] and bits(nn).tolist() == exp["normal"] and bits(uu).tolist() == exp["up"]


def test_mvp_light_front_known_answer(orc):
    # MVP of light patch 320, FRONT face, printed row-major
    v, *_ = orc.scene_cornell(0.5)
    m = orc.mvp(v[320], 0).reshape(4, 4).T
    exp = np.array([[1, 0, 0, -3.105], [0, 0, 1, -2.5325], [0, -1.00002, 0, 5.46511], [0, -1, 0, 5.485]])
    assert np.allclose(m, exp, atol=2e-5)


def test_formfactors(orc, api, golden):
    for N, g in golden["reference"]["formfactors"].items():
        N = int(N)
        for ff in (orc.formfactors(N, 2), api.formfactors(N, 2)):
            assert sha(ff) == g["sha256_k2"]
            one = ff[:3 * N * N]
            assert abs(float(one.sum(dtype=np.float64)) - g["sum_f64"]) < 1e-12
            assert int(one[:1].view(np.uint32)[0]) == g["first_bits"]
            assert (ff[3 * N * N:] == one).all()                      # k copies of the same table
    # SURVEY.md §6 known answers
    f = golden["reference"]["formfactors"]
    assert abs(f["16"]["sum_f64"] - 1.0714602) < 1e-6 and abs(f["128"]["sum_f64"] - 1.0086914) < 1e-6 and abs(f["512"]["sum_f64"] - 1.0021666) < 1e-6


def test_config(orc, api, golden):
    for key, exp in golden["reference"]["config"].items():
        side, k = (int(x) for x in key.split("_"))
        out = (ctypes.c_uint * 9)()
        orc.lib().orc_config(side, k, out)
        assert list(out) == exp
        api.host_lib().radhost_config(side, k, 500, 0.0, out)
        assert list(out) == exp


def test_config_refuses_after_freeze(api, capfd):
    api.host_lib().radhost_config(64, 2, 500, 0.0, None)
    assert api.host_lib().radhost_config_set_when_frozen() == 1
    assert "frozen" in capfd.readouterr().err            # same message channel as the reference (cerr)


def test_codec(orc, golden):
    for P, g in golden["reference"]["codec"].items():
        out = (ctypes.c_uint * 11)()
        orc.lib().orc_colors_setup(int(P), out)
        assert list(out) == g["params"]
        for i, col in g["colors"].items():
            assert orc.lib().orc_color(int(i)) == col
            assert orc.lib().orc_color_index(col) == g["index_of_color"][str(col)] == int(i)
    assert golden["reference"]["codec"]["502"]["params"][9] == 66124863
    assert golden["reference"]["codec"]["502"]["colors"]["1"] == 0x80 and golden["reference"]["codec"]["502"]["colors"]["502"] == 0x380C0300


def test_obj_loader(orc, api, golden):
    path = os.path.join(ROOT, "tests", "golden", "simple.obj")
    for area, g in golden["reference"]["obj"].items():
        v, c, r, il = orc.scene_obj(path, float(area))
        assert v.shape[0] == g["P"] and sha(v) == g["verts_sha256"]
        s = api.Scene(float(area), obj=path)
        hv, _, hc, hr, hi = s.arrays()
        assert s.P == g["P"] and sha(hv) == g["verts_sha256"]
        # the #@color / #@emit directives are an extension: the reference (and the oracle) leave OBJ patches black and unlit
        assert g["color_sum"] == 0.0 and g["rad_sum"] == 0.0 and c.sum() == 0 and r.sum() == 0
        assert hc.sum() > 0 and hr.sum() > 0


def test_live_against_reference_build(orc, api, ref):
    """Same checks against the reference's code compiled here (skipped where oracle/_ref did not travel)."""
    for area in (0.5, 0.02):
        P = ref.refp_scene_build(area)
        rv = np.zeros((P, 12), np.float32); rc = np.zeros((P, 3), np.float32); rr = np.zeros((P, 3), np.float32); ri = np.zeros((P, 3), np.float32)
        ref.refp_scene_get(vp(rv), None, vp(rc), vp(rr), vp(ri))
        ov, oc, orr, oi = orc.scene_cornell(area)
        hv, _, hc, hr, hi = api.Scene(area).arrays()
        for a, b, c in ((rv, ov, hv), (rc, oc, hc), (rr, orr, hr), (ri, oi, hi)):
            assert (bits(a) == bits(b)).all() and (bits(a) == bits(c)).all()
        for p in range(0, P, max(1, P // 64)):
            for look in range(5):
                m = np.zeros(16, np.float32); ref.refp_mvp(p, look, vp(m))
                assert (bits(m) == bits(orc.mvp(ov[p], look))).all()
        for seed in (3, 4):
            rad = seeded_radiosity(P, seed)
            ref.refp_scene_set_radiosity(vp(rad))
            for k in (1, 2, 7, 33):
                ids = np.zeros(k, np.uint32); nul = np.zeros(k, np.int32)
                ref.refp_select(k, vp(ids), vp(nul))
                oi_, on_ = orc.select(rad, k, 0)
                assert ids.tolist() == oi_.tolist() and nul.tolist() == on_.tolist()
    for N in (32, 64):
        a = np.zeros(3 * N * N, np.float32); ref.refp_formfactors(N, 1, vp(a))
        assert (bits(a) == bits(orc.formfactors(N))).all() and (bits(a) == bits(api.formfactors(N))).all()


def _kernel_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m.kernel_cases()


def test_process_hemicube_kernel_restatement_equals_reference_kernel_text(orc, golden):
    """the record stream of Kernel_ProcessHemicube.h.  The golden digests were produced by the reference's
    OWN kernel text, compiled from the reference header by oracle/ref_build.sh and run on the CPU (oracle/ref_kernel.cpp);
    the oracle's restatement must emit the same records — hemicube, id and float energy, in the same order — for real item
    buffers and for the synthetic extremes (no coherence, one id, runs across the work-item spans) in every colour layout."""
    g = golden["reference"]["kernel"]
    n = 0
    for name, ids, ff, N, P, k in _kernel_cases():
        h, ii, e, nrec = orc.process_cl_records(ids, ff, N, P, hemicubes=k)
        assert int(nrec) == g[name]["records"], name
        assert sha(h) == g[name]["hemicubes_sha256"] and sha(ii) == g[name]["ids_sha256"] and sha(e) == g[name]["energies_sha256"], name
        # ... and gathering the records (Main.cpp:1257-1269) gives the per-patch sums of the decoded-id path
        F_cl, _, bad = orc.process_cl(ids, ff, N, P, hemicubes=k)
        assert bad == 0
        per = ids.size // k
        for hi in range(k):
            F_ids = orc.process_ids(ids[hi * per:(hi + 1) * per], ff[:per], N, P)
            assert np.abs(F_cl[hi] - F_ids).max() <= 1e-6 * max(1.0, float(np.abs(F_ids).max())), name
        n += 1
    assert n == len(g)


def test_process_hemicube_kernel_live(orc, ref):
    """The same comparison against the reference kernel compiled here (skipped where oracle/_ref did not travel)."""
    if not hasattr(ref, "refp_process_hemicube_kernel"):
        pytest.skip("libref_host.so built before the kernel was added")
    for name, ids, ff, N, P, k in _kernel_cases():
        a = orc.process_cl_records(ids, ff, N, P, hemicubes=k)
        b = orc.process_cl_records(ids, ff, N, P, hemicubes=k, reference_kernel=True)
        assert a[3] == b[3] and (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2].view(np.uint32) == b[2].view(np.uint32)).all(), name
