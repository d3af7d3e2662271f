"""CPU: the rule of the rasteriser's conservative culling stage (restated in numpy, tests/cull_rule.py) against the
oracle's item buffers — a (patch, face) pair may only be dropped if no pixel of it is visible on that face.

Regression for a bug this check found: the reference builds a face's view matrix from LookAt(eye, target + eye, up) in
float32 (Camera.cpp:19-52, Transform.cpp:26-46).  For a small shooter |target| is tiny against |eye| (a side face's
target is n x u, ~edge^3), so "target + eye - eye" turns the face by up to a degree or more; the MVPs reproduce that bit
for bit, but the culls work in the ideal shooter frame and, with a fixed 2e-3 margin, dropped visible patches along the
face borders (0.05 % of the pixels at 1 M patches).  The margin now follows the measured deviation (RadEmitter::ctol):
largest |real - ideal| over the faces' basis vectors — for patches a thousand times smaller than the scene the real bases
can even come out with their axes swapped (a GPU parity sweep found two such shooters), which switches the frustum culls
off altogether.  The CUDA source's own margin function (camera.cuh cull_margin2, compiled for the host) must equal the
restatement."""
import os
import subprocess

import numpy as np
import pytest

import cull_rule
from test_gpu_parity import random_soup


def _count(orc, v, shooters, N):
    old = new = 0
    for sh in shooters:
        old += cull_rule.wrongly_culled(orc, v, int(sh), N, np.float32(4e-6))
        new += cull_rule.wrongly_culled(orc, v, int(sh), N, cull_rule.ctol(v, int(sh)))
    return old, new


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# shooters a GPU parity sweep (scripts/fuzz_parity.py) caught with the first, axis-agnostic deviation measure
PATHOLOGICAL = {5: [23181], 6: [6044]}


@pytest.mark.parametrize("seed,n,size,N", [(4, 20000, 0.05, 256), (6, 8000, 0.02, 128), (2, 4000, 0.12, 128), (5, 30000, 0.03, 256)])
def test_cull_rule_on_quad_soups(orc, seed, n, size, N):
    v = random_soup(seed, n, size)
    shooters = list(np.random.default_rng(seed).choice(v.shape[0], 12, replace=False)) + PATHOLOGICAL.get(seed, [])
    old, new = _count(orc, v, shooters, N)
    assert new == 0
    if size <= 0.05:
        assert old > 0          # the fixed margin did lose pixels on these small shooters: the check has teeth


def test_cull_rule_on_the_250k_patch_box(orc):
    v, c, r, il = orc.scene_cornell(0.0009)
    shooters = [158100] + list(np.random.default_rng(7).choice(v.shape[0], 5, replace=False))
    old, new = _count(orc, v, shooters, 256)
    assert new == 0 and old > 0
    assert cull_rule.frame_deviation(v, 158100) > 1e-3          # 0.13 degrees
    # at 16 k patches (the bench scene) the deviation is far inside the old margin: nothing changes there
    v2 = orc.scene_cornell(0.014)[0]
    assert max(cull_rule.frame_deviation(v2, int(s)) for s in np.random.default_rng(1).choice(v2.shape[0], 20, replace=False)) < 5e-4


def test_cuda_source_margin_equals_the_restatement(orc, tmp_path):
    """camera.cuh compiled for the host (tests/cpu/cull_margin_check.cpp shims the few CUDA built-ins): cull_margin2 of the
    product source == cull_rule.ctol for shooters of every kind, pathological ones included."""
    exe = str(tmp_path / "cull_margin_check")
    src = os.path.join(ROOT, "tests", "cpu", "cull_margin_check.cpp")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-o", exe, src], check=True)
    for v, shooters in ((random_soup(6, 8000, 0.02), list(range(0, 8000, 37)) + [6044]),
                        (random_soup(5, 30000, 0.03), [23181, 5, 77]),
                        (orc.scene_cornell(0.0009)[0], list(range(0, 250063, 1997)) + [158100]),
                        (orc.scene_cornell(0.014)[0], list(range(0, 16469, 197)))):
        inp = np.ascontiguousarray(v[shooters], np.float32).tobytes()
        got = np.frombuffer(subprocess.run([exe], input=inp, stdout=subprocess.PIPE, check=True).stdout, np.float32)
        exp = np.array([cull_rule.ctol(v, int(s)) for s in shooters], np.float32)
        assert got.shape == exp.shape and np.allclose(got, exp, rtol=1e-4, atol=0)
    # an axis-swapped camera switches the frustum culls off: margin^2 >= 2
    assert cull_rule.ctol(random_soup(5, 30000, 0.03), 23181) > 2.0
