"""CPU: the claim behind the tie-free fast path of RAD_SELECT_REFERENCE (select_update.cu, topk_level_kernel ref_mode = 1), pinned
on the oracle's list (ModelContainer.cpp:259-299 restated in oracle.cpp): the list only rejects a patch whose energy is below its
last entry, whose energy never decreases and starts as patch 0's — so its final SET is the top-k of
S = {0} + {i : |B_i|^2 > 0 and >= |B_0|^2}, and whenever the top-(k + 1) energies of S are pairwise different its ORDER is by energy."""
import numpy as np

f32 = np.float32


def _rule(rad, k):
    e = (rad.astype(f32) ** 2).sum(1, dtype=f32)
    S = [i for i in range(1, len(e)) if e[i] > 0 and e[i] >= e[0]]
    if e[0] > 0:
        S.append(0)
    S.sort(key=lambda i: e[i], reverse=True)
    top = S[:k + 1]
    if len({float(e[i]) for i in top}) != len(top):
        return None                                   # a tie inside the list or across its end: the emulation's business
    ids = S[:k]
    if e[0] == 0 and len(ids) < k:
        ids = ids + [0]                               # the seeded dark patch 0 stays behind the positive ones while there is room
    return ids


def test_tie_free_list_is_the_top_k_of_the_candidates(orc):
    rng = np.random.default_rng(5)
    P, hits = 700, 0
    for trial in range(60):
        k = int(rng.choice([3, 10, 64]))
        rad = (rng.random((P, 3), dtype=f32) + f32(0.05)).astype(f32)
        mode = trial % 5
        if mode == 1:
            rad[0] = 0                                # dark patch 0
        elif mode == 2:
            rad[0] = f32(0.9)                         # bright patch 0: few candidates
        elif mode == 3:
            keep = rng.choice(np.arange(1, P), int(rng.integers(1, k + 3)), replace=False)
            m = np.zeros(P, bool); m[keep] = True; rad[~m] = 0       # fewer candidates than slots (patch 0 dark)
        elif mode == 4:
            j = int(rng.integers(1, P)); rad[j] = rad[int(rng.integers(1, P))]   # a planted tie somewhere
        exp, nul = orc.select(rad, k, 0)
        ids = _rule(rad, k)
        if ids is None:
            continue
        hits += 1
        want = [int(x) for x, n in zip(exp, nul) if not n]
        assert ids == want, (trial, k, mode)
    assert hits > 40
