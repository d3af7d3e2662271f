"""The oracle's rasteriser against an INDEPENDENT restatement of the raster rules of DESIGN.md §2 in plain Python:
float32 numpy scalars for the reference-order float arithmetic, unbounded Python ints for the edge functions.  The
OpenGL driver the reference ran on cannot be executed here (parity with it stays unpinned); what this pins is that
oracle/oracle.cpp — the definition the CUDA path is held to bit for bit — implements exactly the stated rules:
clip-space transform order, near-plane clipping with inside-vertex interpolation, reciprocal-w projection, 8-bit
sub-pixel round-to-nearest-even snapping, back-face culling by signed area, pixel-centre sampling with the top-left
rule, barycentric 24-bit depth, GL_LESS with draw order = patch id, viewports and scissors of the five faces."""
import numpy as np
import pytest

F = np.float32
FACE_TO_LOOK = (1, 2, 3, 4, 0)          # atlas faces UP, DOWN, LEFT, RIGHT, FRONT -> Camera::PatchLook (Main.h:210-211)


def face_window(f, N):
    """viewport origin and scissor (x, y, w, h) of atlas face f (Main.cpp:314-389)"""
    return {0: ((0, N), (0, N, N, N // 2)), 1: ((N, N // 2), (N, N, N, N // 2)), 2: ((-(N // 2), 0), (0, 0, N // 2, N)),
            3: ((N + N // 2, 0), (N + N // 2, 0, N // 2, N)), 4: ((N // 2, 0), (N // 2, 0, N, N))}[f]


def xform(m, p):
    x, y, z = F(p[0]), F(p[1]), F(p[2])
    return [F(F(F(m[r] * x) + F(m[4 + r] * y)) + F(m[8 + r] * z)) + m[12 + r] for r in range(4)]   # column-major m[c*4+r]


def lerp(a, b, da, db):
    t = F(da / F(da - db))
    return [F(a[i] + F(t * F(b[i] - a[i]))) for i in range(4)]


def rne(x):
    return int(np.rint(np.float64(x)))      # numpy rint rounds half to even


def project(c, hw, ox, oy):
    iw = F(F(1.0) / c[3])
    sx = F(F(F(c[0] * iw) * hw) + ox) * F(256.0); sy = F(F(F(c[1] * iw) * hw) + oy) * F(256.0)
    lim = F(536870912.0)
    X = rne(min(max(F(sx), -lim), lim)); Y = rne(min(max(F(sy), -lim), lim))
    Z = F(F(F(c[2] * iw) * F(0.5)) + F(0.5))
    return X, Y, Z


def edge(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def bias(ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    return 0 if (dy < 0 or (dy == 0 and dx < 0)) else -1


def draw_triangle(keys, a, b, c, sc, id1):
    (X0, Y0, Z0), (X1, Y1, Z1), (X2, Y2, Z2) = a, b, c
    area2 = edge(X0, Y0, X1, Y1, X2, Y2)
    if area2 <= 0:
        return
    scx, scy, scw, sch = sc
    px0 = max((min(X0, X1, X2) - 128 + 255) >> 8, scx); px1 = min((max(X0, X1, X2) - 128) >> 8, scx + scw - 1)
    py0 = max((min(Y0, Y1, Y2) - 128 + 255) >> 8, scy); py1 = min((max(Y0, Y1, Y2) - 128) >> 8, scy + sch - 1)
    inv = F(F(1.0) / F(area2)); dz1 = F(Z1 - Z0); dz2 = F(Z2 - Z0)
    b0, b1, b2 = bias(X1, Y1, X2, Y2), bias(X2, Y2, X0, Y0), bias(X0, Y0, X1, Y1)
    for py in range(py0, py1 + 1):
        for px in range(px0, px1 + 1):
            cx, cy = px * 256 + 128, py * 256 + 128
            e0, e1, e2 = edge(X1, Y1, X2, Y2, cx, cy), edge(X2, Y2, X0, Y0, cx, cy), edge(X0, Y0, X1, Y1, cx, cy)
            if e0 + b0 < 0 or e1 + b1 < 0 or e2 + b2 < 0:
                continue
            l1 = F(F(e1) * inv); l2 = F(F(e2) * inv)
            z = F(F(Z0 + F(l1 * dz1)) + F(l2 * dz2))
            z = min(max(z, F(0.0)), F(1.0))
            dq = rne(F(z * F(16777215.0)))
            if dq < 0xFFFFFF:
                key = (dq << 32) | id1
                if key < keys[py][px]:
                    keys[py][px] = key


def python_hemicube(orc, verts, shooter, N):
    P = verts.shape[0]
    W, H = 2 * N, N + N // 2
    EMPTY = (1 << 62)
    keys = [[EMPTY] * W for _ in range(H)]
    hw = F(N) * F(0.5)
    for f in range(5):
        m = orc.mvp(verts[shooter], FACE_TO_LOOK[f])
        (vpx, vpy), sc = face_window(f, N)
        ox, oy = F(F(vpx) + hw), F(F(vpy) + hw)
        for p in range(P):
            q = verts[p].reshape(4, 3)
            c = [xform(m, q[k]) for k in range(4)]
            for t in range(2):                                   # (0,1,2), (0,2,3)  (ModelContainer.cpp:112-117)
                tri = [c[0], c[t + 1], c[t + 2]]
                d = [F(v[2] + v[3]) for v in tri]
                ins = [x >= 0 for x in d]
                code = ins[0] * 1 + ins[1] * 2 + ins[2] * 4
                i0, i1, i2 = tri; d0, d1, d2 = d
                poly = {7: [i0, i1, i2], 0: [],
                        1: [i0, lerp(i0, i1, d0, d1), lerp(i0, i2, d0, d2)] if code == 1 else None,
                        2: [lerp(i1, i0, d1, d0), i1, lerp(i1, i2, d1, d2)] if code == 2 else None,
                        4: [lerp(i2, i1, d2, d1), i2, lerp(i2, i0, d2, d0)] if code == 4 else None,
                        3: [i0, i1, lerp(i1, i2, d1, d2), lerp(i0, i2, d0, d2)] if code == 3 else None,
                        6: [lerp(i1, i0, d1, d0), i1, i2, lerp(i2, i0, d2, d0)] if code == 6 else None,
                        5: [i0, lerp(i0, i1, d0, d1), lerp(i2, i1, d2, d1), i2] if code == 5 else None}[code]
                if not poly:
                    continue
                pv = [project(v, hw, ox, oy) for v in poly]
                draw_triangle(keys, pv[0], pv[1], pv[2], sc, p + 1)
                if len(pv) == 4:
                    draw_triangle(keys, pv[0], pv[2], pv[3], sc, p + 1)
    ids = np.array([[0 if k == EMPTY else (k & 0xFFFFFFFF) for k in row] for row in keys], np.uint32)
    dep = np.array([[0xFFFFFF if k == EMPTY else (k >> 32) for k in row] for row in keys], np.uint32)
    return ids, dep


def soup(seed, n, size):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.5, 3.5, (n, 1, 3))
    a = rng.normal(size=(n, 3)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.normal(size=(n, 3)); b -= (b * a).sum(1, keepdims=True) * a; b /= np.linalg.norm(b, axis=1, keepdims=True)
    s = size * rng.uniform(0.3, 1.5, (n, 1))
    v = c + np.stack([-a - b, a - b, a + b, -a + b], 1) * 0.5 * s[:, None] + rng.normal(scale=0.1, size=(n, 4, 3)) * s[:, None]
    big = rng.random(n) < 0.05
    v[big] = c[big] + (v[big] - c[big]) * 10.0                     # near-plane clipping, patches across several faces
    return np.ascontiguousarray(v.reshape(n, 12), np.float32)


@pytest.mark.parametrize("seed,n,size,N", [(11, 120, 0.8, 16), (12, 60, 1.5, 32)])
def test_oracle_raster_equals_the_stated_rules_on_random_quads(orc, seed, n, size, N):
    v = soup(seed, n, size)
    rng = np.random.default_rng(seed + 1000)
    for shooter in (int(x) for x in rng.integers(0, n, 3)):
        ids, dep = python_hemicube(orc, v, shooter, N)
        oids, odep = orc.render_hemicube(v, shooter, N, want_depth=True)
        assert (ids == oids).all(), (shooter, int((ids != oids).sum()))
        assert (dep == odep).all(), shooter
        assert (ids != 0).mean() > 0.2                             # the case really draws something


def test_oracle_raster_equals_the_stated_rules_on_the_box(orc):
    v, c, r, il = orc.scene_cornell(0.5)                           # P = 502: axis-aligned, shared edges, clipped neighbours
    for shooter in (323, 0, 77):
        ids, dep = python_hemicube(orc, v, shooter, 16)
        oids, odep = orc.render_hemicube(v, shooter, 16, want_depth=True)
        assert (ids == oids).all() and (dep == odep).all(), shooter
        assert (ids == 0).sum() == 0                               # closed box
