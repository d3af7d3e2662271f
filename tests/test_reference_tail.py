"""CPU: the host-side tail of a batch — snapshot, record gather, energy transfer, emitter update, stop test
(Main.cpp:1161,1251-1303) — pinned on the reference's own TEXT.  oracle/ref_build.sh prints those lines of Main.cpp into
its temp dir and compiles them inside oracle/ref_probe.cpp's refp_main_tail (Main.cpp as a whole needs Win32/GL/CL and
cannot be built); tests/golden/make_golden.py ran whole batches that way (emitters from the reference's ModelContainer,
records from the reference's kernel text) and committed the digests.  The oracle's restatement of the loop must land on
the same bits."""
import importlib.util
import os

import numpy as np
import pytest

from util import sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return list(m.tail_cases())


@pytest.mark.parametrize("via_codec", [True, False])
def test_oracle_loop_equals_the_reference_text(orc, golden, via_codec):
    g = golden["reference"]["tail"]
    n = 0
    for name, area, N, k, nb, rad0, il0 in _cases():
        v, c, _, _ = orc.scene_cornell(area)
        rad, illum, sched, done, last = orc.shoot(v, c, rad0, il0, N, k, nb, select_mode=0, via_codec=via_codec, stop_test=True)
        e = g[name]
        assert int(done) == e["batches_done"] and (done < nb) == e["stopped"], name
        assert int(np.float32(last).view(np.uint32)) == e["last_bits"], name
        assert sha(rad) == e["rad_sha256"] and sha(illum) == e["illum_sha256"], name
        n += 1
    assert n == len(g)


def test_reference_text_live(orc, ref, golden):
    """the same batches on the reference build of this container (skipped where oracle/_ref did not travel)"""
    if not hasattr(ref, "refp_main_tail"):
        pytest.skip("libref_host.so built before the tail was added")
    g = golden["reference"]["tail"]
    for name, area, N, k, nb, rad0, il0 in _cases():
        rad, illum, done, last, stopped = orc.reference_tail_batches(area, rad0, il0, N, k, nb)
        e = g[name]
        assert done == e["batches_done"] and stopped == e["stopped"] and sha(rad) == e["rad_sha256"] and sha(illum) == e["illum_sha256"], name
