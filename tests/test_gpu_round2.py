"""GPU (`-m gpu`), round 2: the stop test inside the device loop, BASELINE configs 4 and 5 at full size, and the epoch tags of
the key buffers over many hemicube groups."""
import numpy as np
import pytest

from util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def box(orc):
    return orc.scene_cornell(0.5)


def _ctx(api, scene, N, k, **kw):
    v, c, r, il = scene
    ctx = api.Context(N, k, v.shape[0], **kw)
    ctx.set_formfactors(api.formfactors(N))
    ctx.upload_scene(v, c, r, il)
    return ctx


# (k, select mode, scale of the initial light energy, batches asked for): the oracle stops at batch 31 / 13 / 18 / 21 / 4 —
# inside the second CUDA-graph replay, inside the first, and (asked for 10 < 16) on the direct-launch path
@pytest.mark.parametrize("k,mode,scale,ask", [(1, 0, 0.08, 64), (1, 0, 0.06, 64), (4, 0, 0.1, 48), (8, 1, 0.15, 64), (4, 0, 0.06, 10)])
def test_stop_test_ends_the_run_at_the_oracles_batch(api, orc, box, k, mode, scale, ask):
    """Main.cpp:1137,1297-1300: the loop ends with the batch whose last emitter had |B| < 0.1 (energy read before the
    subtraction).  rad_shoot(stop_test=1) replays CUDA graphs of 16 batches: the batches a replay still holds after the stop
    must do nothing — same batch count, same state as the oracle's loop."""
    v, c, r, il = box
    N = 32
    r0 = (r * np.float32(scale)).astype(np.float32)
    orad, oillum, sched, done, last = orc.shoot(v, c, r0, il, N, k, ask, select_mode=mode, stop_test=True)
    assert 0 < done < ask
    ctx = _ctx(api, (v, c, r0, il), N, k, select_mode=mode)
    st = ctx.shoot(ask, stop_test=True)
    assert st.stopped == 1 and st.batches_done == done, (st.batches_done, done)
    assert st.shots_done == int((sched[:done] != 0xFFFFFFFF).sum())
    assert abs(st.last_energy_len - last) <= 1e-4 * last and st.last_energy_len < 0.1
    rad, illum = ctx.download_state()
    if k == 1:
        assert rel_l2(rad, orad) < 1e-3 and rel_l2(illum, oillum) < 1e-3
    else:       # k > 1: a near-tie may swap batch membership (see test_shoot_config1_vs_oracle); the energy bookkeeping must agree
        tot, otot = float(rad.sum(dtype=np.float64) + illum.sum(dtype=np.float64)), float(orad.sum(dtype=np.float64) + oillum.sum(dtype=np.float64))
        assert abs(tot - otot) < 1e-3 * otot
    # without the test the same call runs all the batches
    ctx.upload_state(r0, il)
    st2 = ctx.shoot(ask, stop_test=False)
    assert st2.batches_done == ask
    # ... and a stopped context shoots again when asked to (the test is re-armed per call and fires again at once or later)
    ctx.upload_state(r0, il)
    st3 = ctx.shoot(ask, stop_test=True)
    assert st3.batches_done == done and st3.stopped == 1
    ctx.close()


def _hemicube_invariants(api, orc, ctx, v, ff, N, P, shooters, oracle_for, max_empty):
    ctx.set_emitters(shooters)
    ctx.render()
    items = [ctx.read_itembuffer(h) for h in range(len(shooters))]
    for it in items:
        assert it.max() <= P and int((it == 0).sum()) <= max_empty
    ctx.process()
    ff64 = ff.astype(np.float64)
    for h, s in enumerate(shooters):
        F = ctx.read_formfactors(h)
        assert F[s] == 0                                                    # a patch does not see itself
        covered = float(ff64[items[h].ravel() > 0].sum())
        assert abs(float(F.sum(dtype=np.float64)) - covered) < 3e-5         # sum F == sum of dFF over the covered pixels
    ctx.render()                                                            # bit-exact run to run
    for h in range(len(shooters)):
        assert (ctx.read_itembuffer(h) == items[h]).all()
    for h in oracle_for:
        exp = orc.render_hemicube(v, shooters[h], N, threads=8)
        agree = float((items[h] == exp).mean())
        assert agree >= 0.999, agree                                        # north_star's bar
        assert agree == 1.0, int((items[h] != exp).sum())                   # what the implementation achieves
        assert rel_l2(ctx.read_formfactors(h), orc.process_ids(exp, ff, N, P)) < 1e-5


def test_config4_one_million_patches(api, orc):
    """BASELINE config 4: built-in scene at area 0.00022 -> P = 1 021 554, hemicube 1024.  Two hemicubes with the
    size-independent invariants, one of them against the oracle, then one k = 64 batch through rad_shoot."""
    N = 1024
    v, c, r, il = orc.scene_cornell(0.00022)
    P = v.shape[0]
    assert P == 1021554
    ff = api.formfactors(N)
    ctx = _ctx(api, (v, c, r, il), N, 2, select_mode=api.SELECT_TOPK)
    lights = np.nonzero(r[:, 0] > 0)[0]
    _hemicube_invariants(api, orc, ctx, v, ff, N, P, [int(lights[len(lights) // 2]), P // 3], oracle_for=[0], max_empty=4096)
    ctx.close()
    ctx = _ctx(api, (v, c, r, il), N, 64, select_mode=api.SELECT_TOPK)
    st = ctx.shoot(1)
    assert st.batches_done == 1 and st.shots_done == 64 and st.queue_overflow == 0
    rad, illum = ctx.download_state()
    shot = np.nonzero(illum[:, 0] > 1.5)[0]
    assert (shot == lights[:64]).all()                                      # equal energies: id order
    assert np.allclose(illum[shot], 101.0, rtol=1e-3) and np.isfinite(rad).all() and rad.min() > -1e-3
    # every shot moved S = 100 from B to I; what the scene received is S * F * rho (.) colour > 0
    assert float(rad.sum(dtype=np.float64)) > 0
    ctx.close()


@pytest.mark.parametrize("N", [256, 2048])
def test_config5_resolution_sweep_ends(api, orc, N):
    """BASELINE config 5: built-in scene at area 0.0035 -> P = 64 659, hemicube 256 and 2048 (the ends of the sweep)."""
    v, c, r, il = orc.scene_cornell(0.0035)
    P = v.shape[0]
    assert P == 64659
    ff = api.formfactors(N)
    ctx = _ctx(api, (v, c, r, il), N, 2, select_mode=api.SELECT_TOPK)
    lights = np.nonzero(r[:, 0] > 0)[0]
    _hemicube_invariants(api, orc, ctx, v, ff, N, P, [int(lights[0]), 40000], oracle_for=[0, 1] if N == 256 else [1], max_empty=64 if N == 256 else 4096)
    ctx.close()
    ctx = _ctx(api, (v, c, r, il), N, 8, select_mode=api.SELECT_TOPK)
    st = ctx.shoot(2)
    assert st.batches_done == 2 and st.shots_done == 16 and st.queue_overflow == 0
    ctx.close()


def test_epoch_tags_over_many_hemicube_groups(api, orc, box, monkeypatch):
    """Key buffers are recycled under decreasing epoch tags (254 .. 1).  With several hemicube groups per batch the tags of
    one batch must stay strictly decreasing when the counter wraps in the middle of a batch: 7 groups per
    batch (RAD_LANES=1, 1 MB of keys per group), 80 direct-launch batches — the wrap falls inside batch 37 — against the
    same run with one group per batch, and the first batches against the oracle."""
    v, c, r, il = box
    N, k, nb = 64, 64, 80
    out = []
    for env in ({"RAD_LANES": "1", "RAD_L2_GROUP_MB": "1"}, {"RAD_LANES": "8"}):
        for kk, vv in env.items():
            monkeypatch.setenv(kk, vv)
        ctx = _ctx(api, box, N, k, select_mode=api.SELECT_REFERENCE, flags=api.FLAG_KEEP_ITEMBUFFER)
        for _ in range(nb // 8):
            st = ctx.shoot(8)                                               # < 16: direct launches, no graph (no clear per replay)
            assert st.batches_done == 8 and st.queue_overflow == 0
        out.append((ctx.download_state(), [ctx.read_itembuffer(h) for h in range(k)]))
        ctx.close()
        for kk in env:
            monkeypatch.delenv(kk)
    (r0, i0), items0 = out[0]
    (r1, i1), items1 = out[1]
    assert rel_l2(r0, r1) < 1e-5 and rel_l2(i0, i1) < 1e-5
    for h in range(k):
        assert (items0[h] == items1[h]).all(), h
    orad, oillum, *_ = orc.shoot(v, c, r, il, N, k, 6, select_mode=0, threads=8)
    monkeypatch.setenv("RAD_LANES", "1"); monkeypatch.setenv("RAD_L2_GROUP_MB", "1")
    ctx = _ctx(api, box, N, k, select_mode=api.SELECT_REFERENCE)
    ctx.shoot(6)
    rad, illum = ctx.download_state()
    assert rel_l2(rad, orad) < 1e-3 and rel_l2(illum, oillum) < 1e-3
    ctx.close()


def test_partition_without_exchange_is_refused(api, box):
    """rad_set_partition(world > 1) without rad_comm_init / rad_peer_init: rad_shoot must not apply a partial dB silently."""
    ctx = _ctx(api, box, 32, 8, select_mode=api.SELECT_TOPK)
    ctx.set_partition(0, 2)
    with pytest.raises(api.RadError):
        ctx.shoot(1)
    ctx.set_partition(0, 1)
    assert ctx.shoot(1).batches_done == 1
    ctx.close()


def test_neighbours_dropped_when_the_scene_changes_size(api, orc, box):
    """rad_upload_scene with a different P invalidates the neighbour planes of the display stage (they are strided by P)."""
    big = api.Scene(0.1)
    ctx = api.context_for_scene(big, 32, 1)
    ctx.upload_neighbours(big.neighbours())
    ctx.shade_vertices()
    v, c, r, il = box
    ctx.upload_scene(v, c, r, il)                                           # 502 patches now
    with pytest.raises(api.RadError):
        ctx.shade_vertices()
    ctx.close()


def test_ring_path_is_bit_identical_to_the_lane_path(api, orc):
    """RAD_RING=1 (opt-in): per-slot work lists + raster_ring_kernel (walk and ProcessHemicube in one cooperative kernel through
    an L2-resident ring of key buffers, re-used under decreasing epoch tags).  Same item buffers as the oracle, and the same
    state as the default path within float-atomic rounding — also when the ring is as small as it gets (2 stages of 1 slot)
    and when a batch is larger than one launch group (k > 64).  The knob is read at context creation: child processes."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from radiosity_b200 import api
        from oracle import orc
        v, c, r, il = orc.scene_cornell(0.05)
        N, k = 128, int(sys.argv[2])
        ctx = api.Context(N, k, v.shape[0], select_mode=api.SELECT_TOPK, flags=api.FLAG_KEEP_ITEMBUFFER)
        ctx.set_formfactors(api.formfactors(N)); ctx.upload_scene(v, c, r, il)
        ids, valid = ctx.select()
        st = ctx.shoot(1)
        for h in range(0, k, max(1, k // 6)):
            if valid[h]:
                assert (ctx.read_itembuffer(h) == orc.render_hemicube(v, int(ids[h]), N)).all(), (h, ids[h])
        ctx.upload_state(r, il)
        st = ctx.shoot(20)
        assert st.batches_done == 20 and st.queue_overflow == 0
        rad, illum = ctx.download_state()
        np.save(sys.argv[1], np.concatenate([rad.ravel(), illum.ravel()]))
        print("ok")
    """ % (root, os.path.join(root, "tests")))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    for k in (8, 96):
        out = []
        for env_extra in ({"RAD_RING": "0"}, {"RAD_RING": "1"}, {"RAD_RING": "1", "RAD_RING_SG": "1", "RAD_RING_RS": "2"}):
            env = dict(os.environ, **env_extra)
            path = os.path.join(root, "gpurun_out", "ring_%d_%s.npy" % (k, "_".join(env_extra.values())))
            p = subprocess.run([sys.executable, "-c", code, path, str(k)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
            assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-2000:]
            out.append(np.load(path))
        assert rel_l2(out[1], out[0]) < 1e-5 and rel_l2(out[2], out[0]) < 1e-5


@pytest.mark.parametrize("area", [0.5, 0.014])
def test_reference_list_fast_path_and_fallback(api, orc, area):
    """RAD_SELECT_REFERENCE, k > 1: when the top-(k + 1) energies of {0} + {i : |B_i|^2 > 0 and >= |B_0|^2} are pairwise different
    the list of ModelContainer.cpp:259-299 is their top-k in energy order and is written by the top-k kernels (ref_mode);
    any tie inside the list or across its end falls back to the exact emulation.  Both must equal the oracle's list: random
    tie-free states with a dark / dim / bright patch 0, fewer candidates than slots, and ties planted inside and at the end."""
    v, c, r, il = orc.scene_cornell(area)
    P = v.shape[0]
    rng = np.random.default_rng(11)

    def state(e0, n_pos=None):
        rad = (rng.random((P, 3), dtype=np.float32) + np.float32(0.05)).astype(np.float32)
        if n_pos is not None:                       # only n_pos patches (besides patch 0) carry energy
            keep = rng.choice(np.arange(1, P), n_pos, replace=False)
            m = np.zeros(P, bool); m[keep] = True
            rad[~m] = 0
        rad[0] = np.float32(e0)
        return rad

    for k in (3, 10, 64):
        ctx = api.Context(32, k, P, select_mode=api.SELECT_REFERENCE)
        ctx.set_formfactors(api.formfactors(32)); ctx.upload_scene(v, c, r, il)
        cases = [state(0.0), state(0.2), state(0.9), state(5.0), state(0.0, n_pos=5), state(0.3, n_pos=k - 1), state(0.0, n_pos=k), state(0.0, n_pos=k + 1)]
        # ties: inside the list (two of the strongest patches equal), across its end (k-th and (k+1)-th equal), and with patch 0
        t1 = state(0.0); order = np.argsort(-(t1.astype(np.float32) ** 2).sum(1, dtype=np.float32)); t1[order[1]] = t1[order[0]]; cases.append(t1)
        t2 = state(0.0); order = np.argsort(-(t2.astype(np.float32) ** 2).sum(1, dtype=np.float32)); t2[order[k]] = t2[order[k - 1]]; cases.append(t2)
        t3 = state(0.4); order = np.argsort(-(t3.astype(np.float32) ** 2).sum(1, dtype=np.float32)); t3[order[2]] = t3[0]; cases.append(t3)
        for j, rad in enumerate(cases):
            ctx.upload_state(rad, il)
            ids, valid = ctx.select()
            exp, nul = orc.select(rad, k, 0)
            assert ids.tolist() == exp.tolist() and valid.tolist() == (1 - nul).tolist(), (k, j)
        ctx.close()


def test_speculative_strict_progressive_equals_the_one_shot_loop(api, orc):
    """k = 1 (default path): the hemicubes of the 64 strongest patches are rendered as one batch and spec_apply_kernel replays
    the reference's one-shot-at-a-time loop over them (argmax of the current B, S of that moment, emitter update, stop test).
    Same shots, same state as the oracle's strict loop and as the one-hemicube-per-launch path (RAD_SPEC=0) — on the
    direct-launch path (short runs, odd counts), on the CUDA-graph path (>= 256 shots) and across calls."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    v, c, r, il = orc.scene_cornell(0.5)
    N = 64
    # against the oracle, in-process (default = speculative)
    ctx = api.Context(N, 1, v.shape[0])
    ctx.set_formfactors(api.formfactors(N)); ctx.upload_scene(v, c, r, il)
    total = 0
    for n in (1, 37, 300, 2):
        st = ctx.shoot(n)
        assert st.batches_done == n and st.shots_done == n and st.queue_overflow == 0
        total += n
    rad, illum = ctx.download_state()
    orad, oillum, sched, *_ = orc.shoot(v, c, r, il, N, 1, total)
    assert rel_l2(rad, orad) < 1e-3 and rel_l2(illum, oillum) < 1e-3
    # a converged scene: every further shot is the reference's no-op shot of patch 0 and still counts
    ctx.upload_state(np.zeros_like(r), il)
    st = ctx.shoot(500)
    assert st.batches_done == 500 and st.stopped == 1
    ctx.close()
    # against the one-hemicube-per-launch path (the knob is read at context creation: child processes)
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r)
        from radiosity_b200 import api
        from oracle import orc
        v, c, r, il = orc.scene_cornell(0.05)
        ctx = api.Context(128, 1, v.shape[0])
        ctx.set_formfactors(api.formfactors(128)); ctx.upload_scene(v, c, r, il)
        st = ctx.shoot(700)
        assert st.batches_done == 700 and st.shots_done == 700
        rad, illum = ctx.download_state()
        np.save(sys.argv[1], np.concatenate([rad.ravel(), illum.ravel()]))
        print("ok")
    """ % root)
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    out = []
    for spec in ("1", "0"):
        path = os.path.join(root, "gpurun_out", "spec_%s.npy" % spec)
        p = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, RAD_SPEC=spec), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-2000:]
        out.append(np.load(path))
    assert rel_l2(out[0], out[1]) < 1e-5
