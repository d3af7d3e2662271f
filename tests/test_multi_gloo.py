"""CPU, world_size 2 over gloo: the host-side logic of the batched multi-GPU mode (shooter partition, one all-reduce of
dB per batch, identical update on every rank).  The GPU kernels are replaced by an oracle-backed engine here — the
orchestration code under test (radiosity_b200/multi.py) is the same one the GPU path uses."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np
sys.path.insert(0, os.environ["RAD_ROOT"]); sys.path.insert(0, os.path.join(os.environ["RAD_ROOT"], "tests"))
import torch, torch.distributed as dist
from oracle import orc
from radiosity_b200 import multi

class OracleEngine:
    """batch_partial/read_delta/write_delta/batch_finish over the CPU oracle (test stand-in for the CUDA context)."""
    def __init__(self, side, k, rank, world):
        self.v, self.c, self.rad, self.il = orc.scene_cornell(0.5)
        self.P = self.v.shape[0]; self.side, self.k = side, k
        self.h0, self.h1 = multi.shooter_range(k, rank, world)
        self.ff = orc.formfactors(side)
    def batch_partial(self):
        self.ids, self.nul = orc.select(self.rad, self.k, 1)            # clean top-k: same list on every rank
        self.S = self.rad[self.ids].copy()
        dB = np.zeros((self.P, 3), np.float32)
        for h in range(self.h0, self.h1):
            if self.nul[h]: continue
            F = orc.process_ids(orc.render_hemicube(self.v, self.ids[h], self.side), self.ff, self.side, self.P)
            dB += ((self.S[h][None, :] * F[:, None]) * np.float32(0.3)) * self.c[self.ids[h]][None, :]
        self.dB = dB
    def read_delta(self): return self.dB
    def write_delta(self, dB): self.dB = np.array(dB, np.float32)
    def batch_finish(self):
        self.rad = self.rad + self.dB
        last = 0.0
        for h in range(self.k):
            if self.nul[h]: continue
            e = self.ids[h]
            last = float(np.sqrt((self.rad[e].astype(np.float32) ** 2).sum()))
            self.il[e] += self.S[h]; self.rad[e] -= self.S[h]
        return last

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["RAD_PORT"], rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
eng = OracleEngine(32, 6, rank, world)
multi.shoot_batches_hosted(eng, dist, 3)
# replicas must stay in lock-step: identical state on every rank after the all-reduce
t = torch.from_numpy(eng.rad.copy()); lst = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(lst, t)
assert all(torch.equal(lst[0], x) for x in lst), "ranks diverged"
if rank == 0:
    np.save(os.environ["RAD_OUT"], np.stack([eng.rad, eng.il]))
dist.destroy_process_group()
'''


def test_shooter_range_covers_batch():
    sys.path.insert(0, ROOT)
    from radiosity_b200 import multi
    for k in (1, 6, 10, 64):
        for world in (1, 2, 3, 4, 8):
            r = [multi.shooter_range(k, i, world) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == k
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_equals_single_rank(tmp_path, orc):
    import subprocess
    torch = pytest.importorskip("torch")
    out = str(tmp_path / "state.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RAD_ROOT=ROOT, RAD_PORT=port, RANK=str(rank), WORLD_SIZE="2", RAD_OUT=out, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o
    two = np.load(out)
    # the same batched schedule on one rank: the oracle's own loop in top-k mode
    v, c, r, il = orc.scene_cornell(0.5)
    rad, illum, sched, done, last = orc.shoot(v, c, r, il, 32, 6, 3, select_mode=1)
    from util import rel_l2
    assert rel_l2(two[0], rad) < 1e-5 and rel_l2(two[1], illum) < 1e-6      # summation order differs only
