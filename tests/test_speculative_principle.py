"""CPU: the principle the speculative k = 1 path rests on (radiosity_b200/csrc/select_update.cu spec_apply_kernel), pinned on the
oracle.  F_h is a function of the shooter's geometry alone, so the reference's strict one-shot-at-a-time loop (Main.cpp:1137-1309,
HEMICUBES_CNT = 1) can be replayed over hemicubes rendered AHEAD of the selection that asks for them: render the 64 strongest
patches' hemicubes from the state at hand, then shoot one at a time — argmax of the CURRENT |B|^2 (last index among equals), S of
that moment, B_i += ((S F[i]) rho) c, emitter update — and stop at the first shooter that was not rendered.  The replay must
reproduce the oracle's own strict loop shot for shot."""
import numpy as np

from util import rel_l2

f32 = np.float32


def _F(orc, v, shooter, N, ff, P):
    ids = orc.render_hemicube(v, int(shooter), N)
    return orc.process_ids(ids, ff, N, P).astype(f32)


def test_render_ahead_replay_equals_the_strict_loop(orc):
    v, c, r, il = orc.scene_cornell(0.5)
    P, N, M, rho = v.shape[0], 32, 16, f32(0.3)
    ff = orc.formfactors(N)
    rad, illum = r.copy(), il.copy()
    shot_order, batches = [], 0
    total = 40
    while len(shot_order) < total:
        e = (rad.astype(f32) ** 2).sum(1, dtype=f32)
        order = sorted(range(P), key=lambda i: (e[i], i), reverse=True)[:M]          # the argmax's tie rule: higher id first
        ahead = {i: _F(orc, v, i, N, ff, P) for i in order if e[i] > 0}
        batches += 1
        while len(shot_order) < total:
            e = (rad.astype(f32) ** 2).sum(1, dtype=f32)
            best = max(range(P), key=lambda i: (e[i], i))
            if best not in ahead:
                break                                                                  # not rendered ahead: the batch ends, select again
            F = ahead.pop(best)                                                        # a rendered hemicube serves one shot
            S = rad[best].copy()
            rad = (rad + ((S[None, :] * F[:, None]) * rho) * c[best][None, :]).astype(f32)
            illum[best] += S; rad[best] -= S
            shot_order.append(best)
    orad, oillum, sched, *_ = orc.shoot(v, c, r, il, N, 1, total)
    assert [int(x) for x in sched.ravel()[:total]] == shot_order
    assert rel_l2(rad, orad) < 1e-5 and rel_l2(illum, oillum) < 1e-5
    assert batches < total                                                             # (and it does batch: several shots per render-ahead)


def _key(e, i):
    return (int(np.float32(e).view(np.uint32)) << 32) | i if e > 0 else 0


def _simulate_verify_commit(orc, area, N, M, total, pick_pool):
    """spec_sim_kernel / spec_walk_kernel<verify> / spec_walk_kernel<commit> in numpy: the loop replayed over the rendered-ahead
    set's own radiosities, every patch walked through the recorded shots to find the first one an outside patch would have taken
    (key > the recorded key), exactly those shots applied."""
    v, c, r, il = orc.scene_cornell(area)
    P, rho, ff = v.shape[0], f32(0.3), orc.formfactors(N)
    rad, illum = r.copy(), il.copy()
    order, truncated = [], 0
    while len(order) < total:
        R = total - len(order)
        e = (rad.astype(f32) ** 2).sum(1, dtype=f32)
        ids = [i for i in pick_pool(e, P, M) if e[i] > 0]
        assert ids
        F = {i: _F(orc, v, i, N, ff, P) for i in ids}
        Bm = {i: rad[i].copy() for i in ids}
        used, steps = set(), []
        while len(steps) < R:                                          # simulate: the set's members only
            ks = {i: _key((Bm[i].astype(f32) ** 2).sum(dtype=f32), i) for i in ids}
            w = max(ids, key=lambda i: ks[i])
            if ks[w] == 0 or w in used:
                break
            S, K = Bm[w].copy(), ks[w]
            for m in ids:
                Bm[m] = (Bm[m] + ((S * F[w][m]) * rho) * c[w]).astype(f32)
            Bm[w] = (Bm[w] - S).astype(f32)
            used.add(w); steps.append((w, S, K))
        b, jstar = rad.copy(), len(steps)                              # verify: every patch, the loop's own arithmetic
        for j, (w, S, K) in enumerate(steps):
            e = (b.astype(f32) ** 2).sum(1, dtype=f32)
            if max(_key(e[i], i) for i in range(P)) > K:
                jstar = j
                break
            b = (b + ((S[None, :] * F[w][:, None]) * rho) * c[w][None, :]).astype(f32); b[w] -= S
        truncated += jstar < len(steps)
        assert jstar > 0                                               # the first shooter is the set's first entry
        for w, S, K in steps[:jstar]:                                  # commit
            rad = (rad + ((S[None, :] * F[w][:, None]) * rho) * c[w][None, :]).astype(f32); illum[w] += S; rad[w] -= S
            order.append(w)
    orad, oillum, sched, *_ = orc.shoot(v, c, r, il, N, 1, total)
    assert [int(x) for x in sched.ravel()[:total]] == order
    assert rel_l2(rad, orad) < 1e-6 and rel_l2(illum, oillum) < 1e-6
    return truncated


def test_simulate_verify_commit_equals_the_strict_loop(orc):
    top = lambda e, P, M: sorted(range(P), key=lambda i: (e[i], i), reverse=True)[:M]
    _simulate_verify_commit(orc, 0.5, 32, 16, 50, top)
    # a set with a hole behind its strongest entries: a patch outside MUST overtake, the verification has to cut the batch there
    holed = lambda e, P, M: [x for k, x in enumerate(sorted(range(P), key=lambda i: (e[i], i), reverse=True)[:M + 1]) if k != 3]
    assert _simulate_verify_commit(orc, 0.5, 32, 8, 24, holed) > 0
