#!/usr/bin/env python
"""bench.py — hemicubes/s (shots/s) of the radiosity shooting loop on N B200s, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2|config1|config3|config4|config2_k1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], the one the metric is quoted on): built-in Cornell box subdivided to P = 16 469
patches (`area 0.014`), hemicube 512 (atlas 1024 x 768), shot from the fresh scene (light B = 100) in batches of k = 64
(the reference batches too, `hemicubes`, default 10) with the clean top-k schedule: 1024 shots = 16 batches per run.  One
"step" = ten such runs (10 240 shots; the fresh state is restored on the device before each).  `value` = whole-job shots/s with the scene resident in HBM (device
time, CUDA events on the launching stream inside rad_shoot, max over ranks); `e2e` = the same through the C ABI with HOST
buffers: the step's input state (B, I) uploaded from page-locked host memory, the shoot, and the resulting state read
back, inside the timed region (the geometry is uploaded once, like the reference's VBO).  Shots are counted from the
library's own counter (a batch whose list is not full counts what it shot).

With N > 1 the emitters of every batch are sharded over the ranks and dB is combined once per batch (fused peer-memory
update kernel over NVLink by default, `--exchange nccl` for ncclAllReduce).  Default `--scaling weak`: every rank keeps
k = 64 emitters per batch, i.e. the batch is the top-(64 N) list (per-GPU work fixed); `--scaling strong` keeps the batch
at k = 64.  Every N > 1 run also (a) checks the replicas bit-identical and the result against the same schedule on one GPU
(`multichip_check`) and (b) times BASELINE configs 3 (and 4 at N = 8) at fixed k = 64 on one GPU and sharded
(`strong_config3` / `strong_config4`).

`--impl reference` times the reference's own algorithm on the host cores (the CPU oracle port — GL/CL cannot run in
this image) on a bounded sample of the same workload, all host threads.
"""
import argparse
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (area, hemicube side, k, batches per run, description); a run starts from the fresh scene
    "config1": (0.5, 128, 10, 10, "built-in Cornell box, area 0.5 (P=502), hemicube 128, 100 shots (k=10 x 10 batches, the reference's default `hemicubes 10`)"),
    "config2": (0.014, 512, 64, 16, "built-in Cornell box, area 0.014 (P=16469), hemicube 512, 1024 shots (k=64 x 16 batches from the fresh scene, top-k schedule)"),
    "config3": (0.0009, 1024, 64, 8, "built-in Cornell box, area 0.0009 (P=250063), hemicube 1024, 512 shots (k=64 x 8 batches from the fresh scene, top-k schedule)"),
    "config4": (0.00022, 1024, 64, 8, "built-in Cornell box, area 0.00022 (P=1021554), hemicube 1024, 512 shots (k=64 x 8 batches from the fresh scene, top-k schedule)"),
    "config2_k1": (0.014, 512, 1, 1000, "built-in Cornell box, area 0.014 (P=16469), hemicube 512, 1000 shots (k=1, strict progressive, reference schedule)"),
}
# runs per step: a step repeats the workload's run (restore the fresh state, shoot) so that the timed region of the default
# K = 20 steps is well over a second (the state restore is a 400 KB device copy outside the device-timed rad_shoot)
RUNS_PER_STEP = {"config1": 50, "config2": 10, "config3": 2, "config4": 2, "config2_k1": 10}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """threads the CPU arm may use: the cores this process is allowed on (torchrun exports OMP_NUM_THREADS=1, which must
    not shrink the CPU baseline of an N > 1 run)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def weak_desc(desc, k):
    return re.sub(r"(\d+) shots \(k=64 x (\d+) batches", lambda m: f"up to {k * int(m.group(2))} shots (k={k} = 64 per rank x {m.group(2)} batches", desc)


def config_dict(desc, P, N, k, schedule):
    """identical in both arms (the driver compares them)"""
    return {"workload": desc, "patches": int(P), "hemicube": int(N), "atlas": [2 * N, N + N // 2], "k": int(k), "schedule": schedule,
            "l2": "GPU arm: 256 MB written between timed iterations (L2 flush), not timed",
            "timing": "GPU arm: CUDA events on the launching stream inside rad_shoot, max over ranks; CPU arm: wall clock"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region.  The sampler is started before the warm-up
    (nvidia-smi's own start-up perturbs the GPUs for tens of ms, worst with 8 of them) and only samples whose timestamp
    falls inside [mark_start, mark_end] are summarised."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device, enabled=True):
        self.device = device; self.proc = None; self.path = None; self.enabled = enabled
        self.t0 = self.t1 = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def mark_start(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                if self.t0 and self.t1 and not (self.t0 - datetime.timedelta(milliseconds=100) <= ts <= self.t1 + datetime.timedelta(milliseconds=100)):
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower() == "active":
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            hi = [x for x in sm if x >= 0.5 * max(sm)] or sm          # samples under load
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the only code of this file that executes oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def reference_arm(args, wl):
    """The reference's own algorithm on the host cores (CPU oracle port; see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    area, N, k, batches, desc = wl
    scaling = "strong"
    if args.gpus > 1 and k > 1:
        scaling = args.scaling
        if scaling == "weak":
            k = k * args.gpus                    # the same batch as the GPU arm at N ranks: the top-(64 N) list
            desc = weak_desc(desc, k)
    v, c, r, il = orc.scene_cornell(area)
    P = v.shape[0]
    threads = host_threads()
    sample_batches = 1 if k > 1 else 16      # a bounded sample per step keeps K+W steps within minutes
    rad, illum = r.copy(), il.copy()

    def step():
        nonlocal rad, illum
        rad, illum, sched, *_ = orc.shoot(v, c, rad, illum, N, k, sample_batches, select_mode=1 if k > 1 else 0, threads=threads)
        return int((sched != 0xFFFFFFFF).sum())          # NULL slots of a list that is not full are not shots

    for _ in range(args.warmup):
        step()
    shots = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        shots += step()
    dt = time.perf_counter() - t0
    value = shots / dt
    line = {"impl": "reference", "metric": "hemicubes_per_sec", "value": value, "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(desc, P, N, k, "topk" if k > 1 else "reference"),
            "note": "reference's algorithm restated on the CPU (oracle port: reference host code semantics + GL raster / CL kernel restatement); the reference's GL+CL stack cannot run in this image",
            "cpu_baseline": {"value": value, "unit": "shots/s", "cores": threads, "kind": "port",
                             "sample": f"{sample_batches} batch(es) of k={k} per step, consecutive batches of the same run from the fresh scene"},
            "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(wl, budget_s=20.0):
    """Bounded CPU sample of the same workload on the box's host cores (rank 0, N = 1)."""
    from oracle import orc
    area, N, k, batches, desc = wl
    v, c, r, il = orc.scene_cornell(area)
    threads = host_threads()
    mode = 1 if k > 1 else 0
    res = {}
    for th in sorted({1, threads}):
        t0 = time.perf_counter()
        done = 0
        rad, illum = r.copy(), il.copy()
        nb = 1 if k > 1 else 16
        while True:
            rad, illum, sched, *_ = orc.shoot(v, c, rad, illum, N, k, nb, select_mode=mode, threads=th)
            done += int((sched != 0xFFFFFFFF).sum())
            if time.perf_counter() - t0 > budget_s / 2 or done >= 256:
                break
        res[th] = (done / (time.perf_counter() - t0), done)
    best = max(res, key=lambda t: res[t][0])
    return {"value": res[best][0], "unit": "shots/s", "cores": best, "kind": "port",
            "sample": f"{res[best][1]} shots of the same workload (first batches from the fresh scene)",
            "single_thread_value": res[1][0], "host_threads_available": threads}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        from radiosity_b200 import api, multi
        self.torch, self.api, self.multi, self.args = torch, api, multi, args
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1")); self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
        self.exchange = args.exchange
        self._keep = []

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def flush(self):
        self.flush_buf.fill_(1); self.torch.cuda.synchronize()                        # L2 flush between timed iterations (not timed)

    def context(self, arrays, N, k, sharded):
        """context with the scene resident; sharded: this rank's partition of every batch + the dB exchange"""
        api = self.api
        v, c, r, il = arrays
        P = v.shape[0]
        mode = api.SELECT_TOPK if k > 1 else api.SELECT_REFERENCE

        def make():
            cx = api.Context(N, k, P, device=self.local, select_mode=mode)
            cx.set_formfactors(api.formfactors(N))
            cx.upload_scene(v, c, r, il)
            return cx
        ctx = make()
        if sharded and self.world > 1:
            if self.exchange == "peer":
                # CUDA IPC mapping of the peers' exchange buffers; if any rank cannot map them (container without IPC between
                # the ranks), every rank falls back to the in-library NCCL all-reduce together
                ok = 1
                try:
                    self.multi.init_peer(ctx, self.dist)
                except Exception as e:               # noqa: BLE001
                    print(f"bench.py: rank {self.rank}: peer-memory exchange unavailable ({e}); falling back to NCCL", file=sys.stderr)
                    ok = 0
                t_ok = self.torch.tensor([ok], device="cuda")
                self.dist.all_reduce(t_ok, op=self.dist.ReduceOp.MIN)
                if int(t_ok[0]) == 0:
                    self.dist.barrier()
                    ctx.close()
                    ctx = make()
                    self.exchange = "nccl"
            if self.exchange == "nccl":
                self.multi.init_nccl(ctx, self.dist)
        ctx.save_state()
        return ctx

    def close(self, ctx, sharded):
        if sharded and self.dist is not None:
            self.dist.barrier()                  # nobody unmaps its exchange buffer while a peer may still read it
        ctx.close()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def timed(self, ctx, batches, steps, warmup, collective=True, runs=1):
        """W untimed + K timed steps of `runs` runs each (restore the saved state, shoot `batches` batches); returns
        (shots/s, ms/step, shots, launches).  collective=False: this rank alone (no barrier, no max over ranks)."""
        def step():
            self.flush()
            ms, shots, launches = 0.0, 0, 0
            for _ in range(runs):
                ctx.restore_state()
                st = ctx.shoot(batches)
                assert st.batches_done == batches and st.queue_overflow == 0
                ms += st.gpu_ms; shots += st.shots_done; launches += st.kernel_launches
            return ms, shots, launches
        for _ in range(warmup):
            step()
        if collective:
            self.barrier()
        ms, shots, launches = 0.0, 0, 0
        for _ in range(steps):
            a, b, c = step()
            ms += a; shots += b; launches += c
        if collective:
            self.barrier()
            (ms,) = self.max_over_ranks(ms)
        return shots / (ms * 1e-3), ms / steps, shots, launches

    def state_digest(self, ctx):
        rad, illum = ctx.download_state()
        return rad, illum, hashlib.sha256(rad.tobytes() + illum.tobytes()).hexdigest()

    def replicas_identical(self, digest):
        if self.dist is None:
            return True
        lst = [None] * self.world
        self.dist.all_gather_object(lst, digest)
        return all(x == lst[0] for x in lst)

    def pinned_state(self, r, il):
        """page-locked host arrays for the e2e legs: the step's input state and its output"""
        torch = self.torch
        keep = [torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory() for a in (r, il)]
        outs = [torch.empty((r.shape[0], 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._keep += keep + outs
        return keep[0].numpy(), keep[1].numpy(), (outs[0].numpy(), outs[1].numpy())

    def e2e(self, ctx, r, il, batches, reps, skip, collective=True, runs=1):
        """per run: state up + shoot + state down; per step `runs` of them; wall clock (max over ranks) -> shots/s"""
        pr, pil, outs = self.pinned_state(r, il)
        t, shots = 0.0, 0
        for i in range(skip + reps):
            self.flush()
            if collective:
                self.barrier()
            t1 = time.perf_counter()
            n = 0
            for _ in range(runs):
                ctx.upload_state(pr, pil)
                st = ctx.shoot(batches)
                ctx.download_state(out=outs)
                n += st.shots_done
            if i >= skip:
                t += time.perf_counter() - t1; shots += n
        if collective:
            (t,) = self.max_over_ranks(t)
        return shots / t


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def multichip_check(B, arrays, N, k, ctx, nb):
    """Replicas in lock-step (bit-identical state on every rank) and the sharded result against the SAME batched schedule on
    one GPU (north_star: 1e-3 relative L2): nb batches from the fresh scene on both."""
    ctx.restore_state()
    st = ctx.shoot(nb)
    rad, illum, dig = B.state_digest(ctx)
    same = B.replicas_identical(dig)
    res = {"batches": nb, "k": k, "replicas_bit_identical": bool(same), "sha256_rank0": dig[:16]}
    if B.rank == 0:
        one = B.context(arrays, N, k, sharded=False)
        s1 = one.shoot(nb)
        r1, i1, _ = B.state_digest(one)
        one.close()
        res.update(rel_l2_radiosity_vs_one_gpu=rel_l2(rad, r1), rel_l2_illumination_vs_one_gpu=rel_l2(illum, i1),
                   shots=int(st.shots_done), shots_one_gpu=int(s1.shots_done))
        res["ok"] = bool(same and res["rel_l2_radiosity_vs_one_gpu"] < 1e-3 and res["rel_l2_illumination_vs_one_gpu"] < 1e-3 and st.shots_done == s1.shots_done)
    B.barrier()
    return res


def strong_scaling(B, name, steps=4, warmup=2):
    """BASELINE configs 3 / 4: batched top-64 shooting at FIXED k = 64 on one GPU (rank 0 alone) and sharded over all ranks."""
    area, N, k, batches, desc = WORKLOADS[name]
    runs = RUNS_PER_STEP[name]
    scene = B.api.Scene(area)
    v, _, c, r, il = scene.arrays()
    arrays = (v, c, r, il)
    res = {"workload": desc, "patches": int(scene.P), "hemicube": N, "k": k, "batches_per_run": batches, "runs_per_step": runs, "steps": steps, "warmup": warmup}
    v1 = r1 = i1 = None
    if B.rank == 0:
        one = B.context(arrays, N, k, sharded=False)
        v1, ms1, _, _ = B.timed(one, batches, steps, warmup, collective=False, runs=runs)
        res.update(v1=v1, ms_per_step_1gpu=ms1, e2e_1gpu=B.e2e(one, r, il, batches, 2, 1, collective=False, runs=runs))
        one.restore_state()
        stage1 = np.zeros(6, np.float64)
        for _ in range(4):
            stage1 += one.profile_batch()
        stage1 /= 4
        res["stages_1gpu_ms_per_batch"] = {"select+camera": float(stage1[0]), "raster_setup": float(stage1[1]), "raster_queue": float(stage1[2]),
                                           "process_hemicube": float(stage1[4]), "update": float(stage1[5])}
        one.restore_state(); one.shoot(batches - 1)
        r_less, _, _ = B.state_digest(one)              # (sanity of the comparison below: one batch less must NOT agree)
        one.restore_state(); one.shoot(batches)
        r1, i1, _ = B.state_digest(one)
        one.close()
    B.barrier()
    ctx = B.context(arrays, N, k, sharded=True)
    vN, msN, _, _ = B.timed(ctx, batches, steps, warmup, runs=runs)
    e2eN = B.e2e(ctx, r, il, batches, 2, 1, runs=runs)
    # where the sharded batch goes: the stages of one batch un-graphed, back to back (collective: every rank profiles the same batches)
    ctx.restore_state()
    stage = np.zeros(6, np.float64)
    for _ in range(4):
        stage += ctx.profile_batch()
    stage /= 4
    res["stages_sharded_ms_per_batch"] = {"select+camera": float(stage[0]), "raster_setup": float(stage[1]), "raster_queue": float(stage[2]), "local_dB+exchange": float(stage[3]),
                                          "process_hemicube": float(stage[4]), "update(incl. wait for peers)": float(stage[5])}
    ctx.restore_state(); ctx.shoot(batches)
    rad, illum, dig = B.state_digest(ctx)
    same = B.replicas_identical(dig)
    B.close(ctx, True)
    res.update(vN=vN, ms_per_step_sharded=msN, n_gpus=B.world, replicas_bit_identical=bool(same), e2e_sharded=e2eN,
               rel_l2_radiosity_vs_fresh_scene=rel_l2(rad, r))      # (sanity: the compared state is not the start state)
    if B.rank == 0:
        res.update(efficiency=vN / (B.world * v1), speedup=vN / v1, rel_l2_radiosity_vs_one_gpu=rel_l2(rad, r1), rel_l2_illumination_vs_one_gpu=rel_l2(illum, i1),
                   values_differing_from_one_gpu=int(np.count_nonzero(rad != r1)), rel_l2_radiosity_vs_one_batch_less=rel_l2(rad, r_less))
        res["ok"] = bool(same and res["rel_l2_radiosity_vs_one_gpu"] < 1e-3 and res["rel_l2_illumination_vs_one_gpu"] < 1e-3)
    return res


def k2_extremes(ctx, P, reps=10):
    """ProcessHemicube alone (item-buffer form) on the k item buffers of a real batch and on the two synthetic extremes of
    SURVEY.md 8d: constant-ID atlas (maximum contention) and id = hash(px) mod P (no coherence).  ms per launch."""
    RES, k = ctx.RES, ctx.k
    out = {"real": ctx.bench_process(reps)}
    const = np.full(RES, (P // 2) + 1, np.uint32)
    for h in range(k):
        ctx.write_itembuffer(h, const)
    out["constant_id"] = ctx.bench_process(reps)
    px = np.arange(RES, dtype=np.uint64)
    for h in range(k):
        hsh = ((px + np.uint64(h * 7919)) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(33)
        ctx.write_itembuffer(h, (hsh % np.uint64(P)).astype(np.uint32) + np.uint32(1))
    out["hashed_id"] = ctx.bench_process(reps)
    return out


def sub_bench(B, name, select_mode=None, steps=5, warmup=3):
    """another schedule / workload on this GPU, device-timed: value + ms per step (+ e2e)"""
    area, N, k, batches, desc = WORKLOADS[name]
    runs = RUNS_PER_STEP[name]
    scene = B.api.Scene(area)
    v, _, c, r, il = scene.arrays()
    api = B.api
    mode = select_mode if select_mode is not None else (api.SELECT_TOPK if k > 1 else api.SELECT_REFERENCE)
    cx = api.Context(N, k, scene.P, device=B.local, select_mode=mode)
    cx.set_formfactors(api.formfactors(N)); cx.upload_scene(v, c, r, il); cx.save_state()
    val, ms, shots, _ = B.timed(cx, batches, steps, warmup, collective=False, runs=runs)
    e2e = B.e2e(cx, r, il, batches, 3, 1, collective=False, runs=runs)
    cx.close()
    return {"workload": desc, "value": val, "unit": "shots/s", "ms_per_step": ms, "shots_per_step": shots // steps, "runs_per_step": runs, "steps": steps, "warmup": warmup,
            "e2e": e2e, "schedule": "reference" if mode == api.SELECT_REFERENCE else "topk"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-results (k=1, reference schedule, K2 extremes, strong-scaling configs)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: dB combined by the fused peer-memory kernel (default) or by ncclAllReduce")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: emitters per batch grow with N (weak) or stay at k (strong)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, wl)

    B = Bench(args)
    api, rank, world, dist = B.api, B.rank, B.world, B.dist
    area, N, k, batches, desc = wl
    runs = RUNS_PER_STEP[args.workload]
    k_rank = k                                   # emitters per batch and rank
    scaling = "strong" if (world == 1 or k == 1) else args.scaling
    if world > 1 and scaling == "weak":
        k = k * world                            # the batch is the top-(k N) list, every rank renders k of it
        desc = weak_desc(desc, k)
    elif world > 1:
        k_rank = k // world
    scene = api.Scene(area)
    v, _, c, r, il = scene.arrays()
    arrays = (v, c, r, il)
    P = scene.P
    RES = 3 * N * N
    ctx = B.context(arrays, N, k, sharded=True)
    extras = not args.no_extras

    with ClockSampler(B.local, enabled=(rank == 0)) as clk:
        time.sleep(0.5)                                                   # let nvidia-smi finish starting before anything is timed
        for _ in range(args.warmup):                                      # (warm-up outside the clock window)
            B.flush(); ctx.restore_state(); ctx.shoot(batches)
        B.barrier()
        clk.mark_start()
        t0 = time.perf_counter()
        value, ms_per_step, shots_timed, launches = B.timed(ctx, batches, args.steps, 0, runs=runs)
        wall = time.perf_counter() - t0
        # e2e: host buffers in, host buffers out, through the C ABI: the step's input state from page-locked host memory,
        # the shoot, the resulting state back into page-locked host memory (geometry stays resident: it is not a step input)
        e2e_value = B.e2e(ctx, r, il, batches, min(args.steps, 10), 2, runs=runs)
        clk.mark_end()

        # per-kernel shares: the same batches un-graphed with CUDA events between the launches
        ctx.restore_state()
        stage = np.zeros(6, np.float64)
        nprof = 32
        for _ in range(nprof):
            stage += ctx.profile_batch()
        stage /= nprof
        # ProcessHemicube alone on the item buffers of a real batch
        ctx.restore_state()
        _, k2_valid = ctx.select(); ctx.render()
        lo = rank * k_rank if world > 1 else 0
        k2_slots = int(np.count_nonzero(k2_valid[lo:lo + k_rank]))      # NULL emitters render nothing and are skipped by the kernel
        k2_ms = ctx.bench_process(20)
        k2x = k2_extremes(ctx, P) if (extras and world == 1 and k > 1) else None
        # display stage: Colors::smoothShadePatch for every patch on the device
        ctx.upload_neighbours(scene.neighbours())
        ctx.shade_vertices()
        k5_ms = min(ctx.shade_vertices()[1] for _ in range(5))
        # RED.MIN.64 rate of the rasteriser's access pattern over this context's key buffers (k x RES x 8 B)
        red_rate_ref = max(ctx.bench_atomics(0, 1 << 27) for _ in range(2))
    clocks = clk.summary()

    mc = multichip_check(B, arrays, N, k, ctx, min(batches, 16)) if world > 1 else None
    B.close(ctx, True)
    strong = {}
    if world > 1 and extras and args.workload == "config2":
        strong["strong_config3"] = strong_scaling(B, "config3")
        if world >= 8 or os.environ.get("RAD_BENCH_CONFIG4"):
            strong["strong_config4"] = strong_scaling(B, "config4")
    subs = {}
    if world == 1 and extras and args.workload == "config2" and rank == 0:
        subs["k1"] = sub_bench(B, "config2_k1")
        subs["reference_schedule_k64"] = sub_bench(B, "config2", select_mode=api.SELECT_REFERENCE, steps=3)
        subs["config3"] = sub_bench(B, "config3", steps=4, warmup=2)

    if rank == 0:
        peak, peak_src = measured_peak()
        nslots = k_rank
        # stage times of one batch: [0] select+camera, [1] raster set-up, [2] raster queues, [4] fused ProcessHemicube, [5] apply
        # algorithmic bytes per launch: K3 12 B/patch; K1 48 B/patch/hemicube + 4 B/pixel (set-up + queue kernels together);
        # K2 8 B/pixel + 4 B/patch/hemicube (F); K4 36 B/patch + 4 B/patch/hemicube
        ring = stage[4] == 0 and stage[2] > 0           # ring path: walk + ProcessHemicube run in ONE kernel (raster_ring_kernel), timed together in [2]
        k1n = "raster + process_hemicube (K1 + K2: raster_setup + raster_ring)" if ring else "raster (K1: raster_setup + raster_queue)"
        stages = {"select+camera": (stage[0], 12.0 * P)}
        if ring:
            stages[k1n] = (stage[1] + stage[2], nslots * (48.0 * P + 4.0 * RES) + nslots * (8.0 * RES + 4.0 * P))
        else:
            stages[k1n] = (stage[1] + stage[2], nslots * (48.0 * P + 4.0 * RES))
            stages["process_hemicube (K2, fused key form)"] = (stage[4], nslots * (8.0 * RES + 4.0 * P))
        if world > 1:
            stages["local dB + exchange (multi-GPU)"] = (stage[3], 12.0 * P + 4.0 * P * nslots)
        stages["apply_update (K4)"] = (stage[5], 36.0 * P + (4.0 * P * nslots if world == 1 else 12.0 * P * world))
        total_stage = float(stage.sum())
        kern = {}
        for n_, (ms, b) in stages.items():
            kern[n_] = {"ms_per_batch": float(ms), "share": float(ms / total_stage), "algorithmic_bytes": b,
                        "achieved_gbs": float(b / (ms * 1e-3) / 1e9) if ms > 0 and b > 0 else None}
        kern[k1n]["setup_ms"] = float(stage[1])
        kern[k1n]["ring_ms" if ring else "queue_ms"] = float(stage[2])
        # every non-empty pixel needed at least one RED.MIN.64 (oracle statistics: 1.1 covered fragments per pixel on this scene)
        red_rate = nslots * RES / (float(stage[2]) * 1e-3) / 1e9 if stage[2] > 0 else 0.0
        kern[k1n]["red_rate"] = {
            "unit": "1e9 RED.MIN.64/s", "achieved_lower_bound": red_rate, "micro_benchmark": red_rate_ref, "ratio": red_rate / red_rate_ref if red_rate_ref else None,
            "key_footprint_mb": k * RES * 8 / 1e6,
            "note": "NOT a roofline: micro_benchmark = rad_bench_atomics (quarter warps walking random 8x8 boxes over all k key buffers, no other work); "
                    "achieved = atlas pixels of a batch / walk time" + (" (ring path: the walk shares the kernel with ProcessHemicube and its keys stay in a few L2-resident buffers)" if ring else "")}
        dom = max(kern, key=lambda n_: kern[n_]["ms_per_batch"])
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get(args.workload, {}).get("raster_ring" if ring else dom.split(" ")[0]) if world == 1 else None
                traffic_src = tj.get("source")
            except Exception:
                traffic = None
        ach = kern[dom]["achieved_gbs"] or 0.0
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": (traffic_src or "none") + " — static: from the committed ncu capture of the same command, not measured in this run",
                    "peak_source": peak_src,
                    "note": "the dominant kernel (K1 raster walk, on the ring path fused with K2 in one launch) is bound by instruction issue and the RED.MIN.64 rate of L2, not by HBM: "
                            "its algorithmic bytes are tiny; K2 (ProcessHemicube) alone is reported in process_hemicube against the same peak"}
        raster_mode = "tiles" if os.environ.get("RAD_RASTER") == "tiles" else "keys"
        if raster_mode == "tiles":      # opt-in tile-binned rasteriser: the stage slots hold other kernels
            roofline["note"] = ("RAD_RASTER=tiles: in `kernels`, queue_ms = bin_kernel x2 + bin_scan_kernel, 'process_hemicube (K2, fused key form)' = tile_kernel "
                                "(visibility in shared memory + fused ProcessHemicube)")
        kk = k2_slots                                    # item buffers one launch of this rank covers (non-NULL emitters only)
        k2_bytes = kk * 8.0 * RES + kk * 4.0 * P
        k2 = {"gpix_per_s": kk * RES / (k2_ms * 1e-3) / 1e9, "ms_per_launch": k2_ms, "pixels_per_launch": kk * RES, "slots": kk,
              "achieved_gbs": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak_gbs": peak, "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peak,
              "dram_frac": (kk * 4.0 * RES) / (k2_ms * 1e-3) / 1e9 / peak,
              "algorithmic_bytes_per_pixel": 8, "itembuffer_bytes": kk * RES * 4,
              "fused_form": ({"ring": True, "note": "ring path: ProcessHemicube runs inside raster_ring_kernel on key buffers that stay in L2 (no HBM key traffic); see kernels"} if ring else
                             {"ms_per_batch": float(stage[4]), "key_bytes": nslots * 8.0 * RES,
                              "dram_frac": nslots * 8.0 * RES / (float(stage[4]) * 1e-3) / 1e9 / peak if stage[4] > 0 else None}),
              "note": "rad_bench_process: item buffers of a real batch (%.0f MB, %s L2), uint32 ids + the dFF table shared by all hemicubes.  frac counts SURVEY 8d's 8 B/pixel "
                      "(4 B id + 4 B dFF); the 3 MB dFF table stays in L2 (ncu: DRAM reads = the id bytes), so dram_frac counts 4 B/pixel: the share of the HBM peak this kernel really uses"
                      % (kk * RES * 4 / 1e6, "larger than" if kk * RES * 4 > 126e6 else "fits in")}
        if k2x:
            k2["extremes"] = {n_: {"ms_per_launch": t, "gpix_per_s": k * RES / (t * 1e-3) / 1e9, "dram_frac": k * 4.0 * RES / (t * 1e-3) / 1e9 / peak} for n_, t in k2x.items()}
            k2["extremes"]["note"] = "SURVEY 8d: all k item buffers overwritten by one id (maximum contention) / by hash(px) mod P (no coherence: one red per pixel)"
        line = {"metric": "hemicubes_per_sec", "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": config_dict(desc, P, N, k, "topk" if k > 1 else "reference"),
                "run": {"batches_per_run": batches, "runs_per_step": runs, "shots_per_step": shots_timed // args.steps, "timed_region_s": ms_per_step * args.steps * 1e-3, "wall_s_incl_flush": wall,
                        "raster": raster_mode + ("+ring" if ring else ""),
                        "parallelism": (f"{k_rank} of the batch's {k} shooters per rank, dB combined once per batch by " +
                                        ("the fused peer-memory update kernel (NVLink, CUDA IPC)" if B.exchange == "peer" else "ncclAllReduce") if world > 1 else "1gpu")},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": int(P * 24 * runs), "d2h_bytes_per_step": int(P * 24 * runs),
                        "note": "per run (%d per step): rad_upload_state (B, I from page-locked host arrays) + rad_shoot + rad_download_state (B, I into page-locked host arrays), wall clock, max over ranks; "
                                "the geometry is uploaded once (rad_upload_scene), like the reference's vertex buffer" % runs},
                "gpu_launches": int(launches),
                "roofline": roofline, "kernels": kern, "process_hemicube": k2,
                "display_stage": {"ms": k5_ms, "algorithmic_bytes": 116 * P, "achieved_gbs": 116.0 * P / (k5_ms * 1e-3) / 1e9,
                                  "note": "K5 smoothShadePatch gather: 36 B state + 32 B neighbour ids + 48 B vertex colours per patch"}}
        if mc is not None:
            line["multichip_check"] = mc
        line.update(strong)
        line.update(subs)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line), flush=True)
        if mc is not None and not mc.get("ok", False):
            print("bench.py: multichip_check FAILED: " + json.dumps(mc), file=sys.stderr)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and mc is not None and not mc.get("ok", False):
        raise SystemExit(3)


if __name__ == "__main__":
    main()
