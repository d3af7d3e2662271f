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
