"""GPU (`-m gpu`), opt-in: tuning variants that were written when no GPU time was left to confirm them.  They sit behind
environment knobs that default to off; these tests run only with RAD_TEST_EXPERIMENTAL=1 (first thing to do with the next
GPU minutes: `RAD_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -q`, then
`scripts/gpu_quick.sh RAD_QUEUE_PREFETCH=1` for the bench line)."""
import os

import numpy as np
import pytest

from util import rel_l2
from test_gpu_parity import make_ctx

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RAD_TEST_EXPERIMENTAL") != "1", reason="opt-in: set RAD_TEST_EXPERIMENTAL=1")]


def test_queue_prefetch_variant_is_bit_identical(api, orc, monkeypatch):
    """raster_queue_kernel<true> (RAD_QUEUE_PREFETCH=1: the next step's records fetched into shared memory with cp.async):
    same item buffers as the default kernel and as the oracle, staged and fused.  The knob is read once per process (a static
    in the launcher), so the variant runs in a child process."""
    import subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from radiosity_b200 import api
        from oracle import orc
        v, c, r, il = orc.scene_cornell(0.014)
        N, k = 256, 8
        ctx = api.Context(N, k, v.shape[0], select_mode=api.SELECT_TOPK, flags=api.FLAG_KEEP_ITEMBUFFER)
        ctx.set_formfactors(api.formfactors(N)); ctx.upload_scene(v, c, r, il)
        shooters = [0, 5000, 9000, 12000, 16000, 323, 8977, 16468]
        ctx.set_emitters(shooters); ctx.render()
        for h, sh in enumerate(shooters):
            assert (ctx.read_itembuffer(h) == orc.render_hemicube(v, sh, N)).all(), sh
        ctx.upload_state(r, il)
        st = ctx.shoot(20)
        assert st.batches_done == 20 and st.queue_overflow == 0
        rad, illum = ctx.download_state()
        np.save(sys.argv[1], np.concatenate([rad.ravel(), illum.ravel()]))
        print("ok")
    """ % (root, os.path.join(root, "tests")))
    out = []
    for pf in ("0", "1"):
        env = dict(os.environ, RAD_QUEUE_PREFETCH=pf)
        path = os.path.join(root, "gpurun_out", f"exp_pf{pf}.npy")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        p = subprocess.run([sys.executable, "-c", code, path], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-2000:]
        out.append(np.load(path))
    assert rel_l2(out[1], out[0]) < 1e-5
