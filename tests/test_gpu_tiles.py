"""GPU (`-m gpu`): the tile-binned rasteriser (RAD_RASTER=tiles, radiosity_b200/csrc/raster_tiles.cu) against the
oracle and against the default global-key path.  Same bars: item buffers bit for bit, radiosity within 1e-5 rel-L2 of
the default path on the same schedule.

The arithmetic of the tile walks is also checked without a GPU: tests/test_tile_walk_cpu.py."""
import numpy as np
import pytest

from util import rel_l2
from test_gpu_parity import make_ctx, random_soup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def box(orc):
    return orc.scene_cornell(0.5)


@pytest.mark.parametrize("N,lanes", [(32, "1"), (48, "8"), (128, "3")])
def test_tiles_itembuffers_bit_exact(api, orc, box, monkeypatch, N, lanes):
    """First batch of the fresh box (k = 12, clean top-k): the item buffers the tile CTAs write equal the staged
    render of the same emitters (global keys) and the oracle's raster, bit for bit.  N = 32 / 48: atlases with partial
    tiles (64 x 48, 96 x 72)."""
    v, c, r, il = box
    k = 12
    monkeypatch.setenv("RAD_RASTER", "tiles")
    monkeypatch.setenv("RAD_LANES", lanes)
    ctx = make_ctx(api, orc, box, N, k=k, select_mode=api.SELECT_TOPK, flags=api.FLAG_KEEP_ITEMBUFFER)
    st = ctx.shoot(1)
    assert st.batches_done == 1 and st.queue_overflow == 0
    fused = [ctx.read_itembuffer(h) for h in range(k)]
    ctx.upload_state(r, il)
    ids, valid = ctx.select()
    ctx.render()
    assert valid.sum() >= 4
    for h in range(k):
        if valid[h]:
            staged = ctx.read_itembuffer(h)
            assert (staged == fused[h]).all(), (h, int((staged != fused[h]).sum()))
            assert (orc.render_hemicube(v, int(ids[h]), N) == fused[h]).all(), h
    ctx.close()


def test_tiles_shoot_equals_default_path(api, orc, box, monkeypatch):
    """20 batches of 8 (16 through the CUDA graph): same radiosity as the default path up to the order of the float
    atomics inside F, item buffers of the last batch identical."""
    N = 64; k = 8
    out = []
    for mode in ("keys", "tiles"):
        monkeypatch.setenv("RAD_RASTER", mode)
        ctx = make_ctx(api, orc, box, N, k=k, select_mode=api.SELECT_REFERENCE, flags=api.FLAG_KEEP_ITEMBUFFER)
        st = ctx.shoot(20)
        assert st.batches_done == 20 and st.queue_overflow == 0
        out.append((ctx.download_state(), [ctx.read_itembuffer(h) for h in range(k)]))
        ctx.close()
    (r0, i0), items0 = out[0]
    (r1, i1), items1 = out[1]
    assert rel_l2(r1, r0) < 1e-5 and rel_l2(i1, i0) < 1e-5
    for h in range(k):
        assert (items0[h] == items1[h]).all(), h


@pytest.mark.parametrize("seed,n,size,N", [(1, 1500, 0.35, 64), (2, 4000, 0.12, 128), (3, 600, 1.2, 128), (4, 20000, 0.05, 256)])
def test_tiles_random_quad_soup(api, orc, monkeypatch, seed, n, size, N):
    """Quad soups (non-planar, concave, degenerate, huge, clipped, deep overdraw) through the tile path: eight shooters are
    given distinct energies so that the clean top-k list is exactly that list; item buffers equal the oracle's, and the
    radiosity after the batch equals the oracle's batch."""
    v = random_soup(seed, n, size)
    P = v.shape[0]
    rng = np.random.default_rng(100 + seed)
    shooters = [int(x) for x in rng.choice(P, 8, replace=False)]
    c = np.full((P, 3), 0.5, np.float32); r = np.zeros((P, 3), np.float32); il = np.zeros((P, 3), np.float32)
    for j, s in enumerate(shooters):
        r[s] = 10.0 * (8 - j)
    monkeypatch.setenv("RAD_RASTER", "tiles")
    ctx = api.Context(N, 8, P, select_mode=api.SELECT_TOPK, flags=api.FLAG_KEEP_ITEMBUFFER)
    ctx.set_formfactors(api.formfactors(N))
    ctx.upload_scene(v, c, r, il)
    st = ctx.shoot(1)
    assert st.shots_done == 8 and st.queue_overflow == 0
    for hi, sh in enumerate(shooters):
        got = ctx.read_itembuffer(hi)
        exp = orc.render_hemicube(v, sh, N)
        assert (got == exp).all(), (seed, sh, int((got != exp).sum()))
    rad, illum = ctx.download_state()
    orad, oillum, *_ = orc.shoot(v, c, r, il, N, 8, 1, select_mode=1)
    assert rel_l2(rad, orad) < 1e-5 and rel_l2(illum, oillum) < 1e-6
    ctx.close()


def test_tiles_config2(api, orc, monkeypatch):
    """P = 16 469, hemicube 512 (16 x 24 tiles per hemicube): two batches of 8 against the default path."""
    scene = orc.scene_cornell(0.014)
    N = 512; k = 8
    out = []
    for mode in ("keys", "tiles"):
        monkeypatch.setenv("RAD_RASTER", mode)
        ctx = make_ctx(api, orc, scene, N, k=k, select_mode=api.SELECT_REFERENCE, flags=api.FLAG_KEEP_ITEMBUFFER)
        st = ctx.shoot(2)
        assert st.batches_done == 2 and st.queue_overflow == 0
        out.append((ctx.download_state(), [ctx.read_itembuffer(h) for h in range(k)], st.gpu_ms))
        ctx.close()
    (r0, i0), items0, ms0 = out[0]
    (r1, i1), items1, ms1 = out[1]
    print(f"config 2, 2 batches of 8: keys {ms0:.3f} ms, tiles {ms1:.3f} ms")
    assert rel_l2(r1, r0) < 1e-5 and rel_l2(i1, i0) < 1e-5
    for h in range(k):
        assert (items0[h] == items1[h]).all(), h
