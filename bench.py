#!/usr/bin/env python
"""bench.py — hemicubes/s (shots/s) of the radiosity shooting loop on N B200s, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2|config1|config3|config4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], the one the metric is quoted on): built-in Cornell box subdivided to P = 16 469
patches (`area 0.014`), hemicube 512 (atlas 1024 x 768), 1024 shots per step from the fresh scene (light B = 100),
shot in batches of k = 64 (the reference batches too, `hemicubes`, default 10) with the clean top-k schedule.  One
"step" = one rad_shoot() call of 16 batches.  `value` = whole-job shots/s with the scene resident in HBM (device time,
CUDA events on the launching stream inside rad_shoot, max over ranks); `e2e` = the same through the C ABI with HOST
buffers: scene upload + shoot + state download inside the timed region.  With N > 1 the emitters of every batch
are sharded over the ranks and dB is combined by one NCCL all-reduce per batch.  Default `--scaling weak`: every rank
keeps k = 64 emitters per batch, i.e. the batch is the top-(64 N) list (per-GPU work fixed); `--scaling strong` keeps
the batch at k = 64 and gives every rank 64 / N of it.  Shots are counted from the library's own counter (a batch
whose list is not full — the first one of a fresh scene with 99 light patches — counts what it shot).

`--impl reference` times the reference's own algorithm on the host cores (the CPU oracle port — GL/CL cannot run in
this image) on a bounded sample of the same workload, all host threads.
"""
import argparse
import json
import os
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
    # name: (area, hemicube side, k, batches per step, description)
    "config1": (0.5, 128, 10, 10, "built-in Cornell box, area 0.5 (P=502), hemicube 128, 100 shots (k=10 x 10 batches)"),
    "config2": (0.014, 512, 64, 16, "built-in Cornell box, area 0.014 (P=16469), hemicube 512, 1024 shots (k=64 x 16 batches, top-k schedule)"),
    "config3": (0.0009, 1024, 64, 2, "built-in Cornell box, area 0.0009 (P=250063), hemicube 1024, 128 shots (k=64 x 2 batches, top-k schedule)"),
    "config4": (0.00022, 1024, 64, 1, "built-in Cornell box, area 0.00022 (P=1021554), hemicube 1024, 64 shots (k=64 x 1 batch, top-k schedule)"),
    "config2_k1": (0.014, 512, 1, 1000, "built-in Cornell box, area 0.014 (P=16469), hemicube 512, 1000 shots (k=1, strict progressive, reference schedule)"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def reference_arm(args, wl):
    """The reference's own algorithm on the host cores (CPU oracle port; see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    area, N, k, batches, desc = wl
    if args.gpus > 1 and args.scaling == "weak" and k > 1:
        import re
        k = k * args.gpus                        # the same batch as the GPU arm at N ranks: the top-(64 N) list
        desc = re.sub(r"(\d+) shots \(k=64 x (\d+) batch", lambda m: f"up to {k * int(m.group(2))} shots (k={k} = 64 per rank x {m.group(2)} batch", desc)
    v, c, r, il = orc.scene_cornell(area)
    P = v.shape[0]
    threads = orc.max_threads()
    sample_batches = 1                       # one batch of k shots per step keeps K+W steps within minutes
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
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling if (args.gpus > 1 and k > 1) else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "patches": P, "hemicube": N, "k": k, "schedule": "topk" if k > 1 else "reference",
                       "note": "reference's algorithm restated on the CPU (oracle port: reference host code semantics + GL raster / CL kernel restatement); the reference's GL+CL stack cannot run in this image"},
            "cpu_baseline": {"value": value, "unit": "shots/s", "cores": threads, "kind": "port",
                             "sample": f"{sample_batches} batch of k={k} shots per step, consecutive batches of the same run"},
            "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(wl, budget_s=20.0):
    """Bounded CPU sample of the same workload on the box's host cores (rank 0, N = 1)."""
    from oracle import orc
    area, N, k, batches, desc = wl
    v, c, r, il = orc.scene_cornell(area)
    threads = orc.max_threads()
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: dB combined by the fused peer-memory kernel (default) or by ncclAllReduce")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: emitters per batch grow with N (weak) or stay at k (strong)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, wl)

    import torch
    from radiosity_b200 import api, multi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    area, N, k, batches, desc = wl
    k_rank = k                                   # emitters per batch and rank
    scaling = "strong" if (world == 1 or k == 1) else args.scaling
    if world > 1 and scaling == "weak":
        k = k * world                            # the batch is the top-(k N) list, every rank renders k of it
        import re
        desc = re.sub(r"(\d+) shots \(k=64 x (\d+) batch", lambda m: f"up to {k * int(m.group(2))} shots (k={k} = 64 per rank x {m.group(2)} batch", desc)
    elif world > 1:
        k_rank = k // world
    scene = api.Scene(area)
    v, _, c, r, il = scene.arrays()
    P = scene.P
    mode = api.SELECT_TOPK if k > 1 else api.SELECT_REFERENCE
    def make_context():
        cx = api.Context(N, k, P, device=local, select_mode=mode)
        cx.set_formfactors(api.formfactors(N))
        cx.upload_scene(v, c, r, il)
        return cx

    ctx = make_context()
    exchange = args.exchange
    if world > 1:
        if exchange == "peer":
            # CUDA IPC mapping of the peers' exchange buffers; if any rank cannot map them (container without IPC between
            # the ranks), every rank falls back to the in-library NCCL all-reduce together
            ok = 1
            try:
                multi.init_peer(ctx, dist)
            except Exception as e:               # noqa: BLE001
                print(f"bench.py: rank {rank}: peer-memory exchange unavailable ({e}); falling back to NCCL", file=sys.stderr)
                ok = 0
            t_ok = torch.tensor([ok], device="cuda")
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            if int(t_ok[0]) == 0:
                dist.barrier()
                ctx.close()
                ctx = make_context()
                exchange = "nccl"
        if exchange == "nccl":
            multi.init_nccl(ctx, dist)
    ctx.save_state()
    shots_per_step = batches * k
    RES = 3 * N * N

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.fill_(1); torch.cuda.synchronize()                          # L2 flush between timed iterations (not timed)
        ctx.restore_state()
        return ctx.shoot(batches)

    launches = 0
    shots_timed = 0
    gpu_ms = []
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        time.sleep(0.5)                                                   # let nvidia-smi finish starting before anything is timed
        for _ in range(args.warmup):
            st = step_device()
        barrier()
        clk.mark_start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = step_device()
            assert st.batches_done == batches and st.queue_overflow == 0
            gpu_ms.append(st.gpu_ms); launches += st.kernel_launches; shots_timed += st.shots_done
        barrier()
        wall = time.perf_counter() - t0

        # e2e: host buffers in, host buffers out, through the C ABI
        # the step's inputs and outputs live in page-locked host memory (the contract's e2e definition)
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory()
            return t, t.numpy()
        keep_alive = [pinned(a) for a in (v, c, r, il)]
        pv, pc, pr, pil = (x[1] for x in keep_alive)
        out_t = [torch.empty((P, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        out_np = (out_t[0].numpy(), out_t[1].numpy())
        e2e_t = []; e2e_shots = 0
        for i in range(2 + min(args.steps, 10)):
            flush.fill_(1); torch.cuda.synchronize()
            barrier()
            t1 = time.perf_counter()
            ctx.upload_scene(pv, pc, pr, pil)
            st = ctx.shoot(batches)
            rad, illum = ctx.download_state(out=out_np)
            e2e_t.append(time.perf_counter() - t1)
            if i >= 2:
                e2e_shots += st.shots_done
        e2e_t = e2e_t[2:]

        # per-kernel shares: the same batches un-graphed with CUDA events between the launches
        ctx.restore_state()
        stage = np.zeros(6, np.float64)
        nprof = batches * 2
        for _ in range(nprof):
            stage += ctx.profile_batch()
        stage /= nprof
        # ProcessHemicube alone on the k item buffers of a real batch
        ctx.restore_state()
        _, k2_valid = ctx.select(); ctx.render()
        lo = rank * k_rank if world > 1 else 0
        k2_slots = int(np.count_nonzero(k2_valid[lo:lo + k_rank]))      # NULL emitters render nothing and are skipped by the kernel
        k2_ms = ctx.bench_process(20)
        # display stage: Colors::smoothShadePatch for every patch on the device
        ctx.upload_neighbours(scene.neighbours())
        ctx.shade_vertices()
        k5_ms = min(ctx.shade_vertices()[1] for _ in range(5))
        # measured RED.MIN.64 roofline for the rasteriser's access pattern over this context's key buffers (k x RES x 8 B)
        red_peak = max(ctx.bench_atomics(0, 1 << 27) for _ in range(2))
        clk.mark_end()
    clocks = clk.summary()

    total_ms = float(sum(gpu_ms)); e2e_s = float(sum(e2e_t))
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    value = shots_timed / (total_ms * 1e-3)
    e2e_value = e2e_shots / e2e_s
    shots_per_step = shots_timed // args.steps

    if rank == 0:
        peak, peak_src = measured_peak()
        h = ctx.lib  # noqa: F841
        nslots = k_rank
        # stage times of one batch: [0] select+camera, [1] raster set-up, [2] raster queues, [4] fused ProcessHemicube, [5] apply
        # algorithmic bytes per launch: K3 12 B/patch; K1 48 B/patch/hemicube + 4 B/pixel (set-up + queue kernels together);
        # K2 8 B/pixel + 4 B/patch/hemicube (F); K4 36 B/patch + 4 B/patch/hemicube
        stages = {"select+camera": (stage[0], 12.0 * P),
                  "raster (K1: raster_setup + raster_queue)": (stage[1] + stage[2], nslots * (48.0 * P + 4.0 * RES)),
                  "process_hemicube (K2, fused key form)": (stage[4], nslots * (8.0 * RES + 4.0 * P)),
                  "apply_update (K4)": (stage[5], 36.0 * P + 4.0 * P * nslots)}
        total_stage = float(stage.sum())
        kern = {}
        for n_, (ms, b) in stages.items():
            kern[n_] = {"ms_per_batch": float(ms), "share": float(ms / total_stage), "algorithmic_bytes": b,
                        "achieved_gbs": float(b / (ms * 1e-3) / 1e9) if ms > 0 and b > 0 else None}
        kern["raster (K1: raster_setup + raster_queue)"]["setup_ms"] = float(stage[1])
        kern["raster (K1: raster_setup + raster_queue)"]["queue_ms"] = float(stage[2])
        # every non-empty pixel needed at least one RED.MIN.64 (oracle statistics: 1.1 covered fragments per pixel on this scene)
        red_rate = nslots * RES / (float(stage[2]) * 1e-3) / 1e9 if stage[2] > 0 else 0.0
        kern["raster (K1: raster_setup + raster_queue)"]["atomic_roofline"] = {
            "bound": "l2_atomic", "unit": "1e9 RED.MIN.64/s", "achieved_lower_bound": red_rate, "peak": red_peak, "frac": red_rate / red_peak if red_peak else None,
            "key_footprint_mb": k * RES * 8 / 1e6,
            "note": "peak = rad_bench_atomics(pattern 0: quarter warps walking random 8x8 boxes) over the same key buffers; achieved = atlas pixels of a batch / raster_queue time"}
        dom = max(kern, key=lambda n_: kern[n_]["ms_per_batch"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload, {}).get(dom.split(" ")[0])
            except Exception:
                traffic = None
        ach = kern[dom]["achieved_gbs"] or 0.0
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src,
                    "note": "the dominant kernel (K1 raster) is bound by set-up arithmetic, instruction issue and L2 atomics, not by HBM: its algorithmic bytes are tiny; "
                            "the HBM-bound kernel of the path is K2 (ProcessHemicube), reported in process_hemicube against the same peak"}
        raster_mode = "tiles" if os.environ.get("RAD_RASTER") == "tiles" else "keys"
        if raster_mode == "tiles":      # opt-in tile-binned rasteriser: the stage slots hold other kernels
            roofline["note"] = ("RAD_RASTER=tiles: in `kernels`, queue_ms = bin_kernel x2 + bin_scan_kernel, 'process_hemicube (K2, fused key form)' = tile_kernel "
                                "(visibility in shared memory + fused ProcessHemicube); atomic_roofline does not apply")
        kk = k2_slots                                    # item buffers one launch of this rank covers (non-NULL emitters only)
        k2_bytes = kk * 8.0 * RES + kk * 4.0 * P
        k2 = {"gpix_per_s": kk * RES / (k2_ms * 1e-3) / 1e9, "ms_per_launch": k2_ms, "pixels_per_launch": kk * RES,
              "achieved_gbs": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak_gbs": peak, "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peak,
              "algorithmic_bytes_per_pixel": 8, "itembuffer_bytes": kk * RES * 4,
              "note": "rad_bench_process: k item buffers of a real batch (%.0f MB, %s L2), uint32 ids + shared dFF table" % (kk * RES * 4 / 1e6, "larger than" if kk * RES * 4 > 126e6 else "fits in")}
        line = {"metric": "hemicubes_per_sec", "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": desc, "patches": P, "hemicube": N, "atlas": [2 * N, N + N // 2], "k": k, "batches_per_step": batches,
                           "shots_per_step": shots_per_step, "schedule": "topk" if k > 1 else "reference",
                           "parallelism": (f"{k_rank} of the batch's {k} shooters per rank, dB combined once per batch by " +
                                           ("the fused peer-memory update kernel (NVLink, CUDA IPC)" if exchange == "peer" else "ncclAllReduce") if world > 1 else "1gpu"),
                           "l2": "256 MB write between timed iterations (flush)", "timing": "CUDA events on the launching stream inside rad_shoot, max over ranks",
                           "wall_s_incl_flush": wall, "raster": raster_mode},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": int(P * 84 + 0), "d2h_bytes_per_step": int(P * 24),
                        "note": "rad_upload_scene (page-locked host arrays -> HBM, layout conversion on the GPU) + rad_shoot + rad_download_state (-> page-locked host arrays) per step, wall clock"},
                "gpu_launches": int(launches),
                "roofline": roofline, "kernels": kern, "process_hemicube": k2,
                "display_stage": {"ms": k5_ms, "algorithmic_bytes": 116 * P, "achieved_gbs": 116.0 * P / (k5_ms * 1e-3) / 1e9,
                                  "note": "K5 smoothShadePatch gather: 36 B state + 32 B neighbour ids + 48 B vertex colours per patch"}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()                           # nobody unmaps its exchange buffer while a peer may still read it
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
