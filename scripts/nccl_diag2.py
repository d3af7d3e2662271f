"""8-rank diagnosis (2): which part of bench.py's step makes the sharded step slow — flush / barrier / clock sampler."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosity_b200 import api, multi
from bench import WORKLOADS, ClockSampler
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
area, N, k, batches, desc = WORKLOADS["config3"]
scene = api.Scene(area)
ctx = api.context_for_scene(scene, N, k, device=local, select_mode=api.SELECT_TOPK)
multi.init_nccl(ctx, dist)
ctx.save_state()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(tag, do_flush, do_barrier, sampler):
    res = []
    cm = ClockSampler(local) if sampler else None
    if cm: cm.__enter__()
    for i in range(6):
        if do_flush: flush.fill_(1); torch.cuda.synchronize()
        ctx.restore_state()
        if do_barrier: dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); st = ctx.shoot(batches); res.append((round(st.gpu_ms, 2), round((time.perf_counter() - t0) * 1e3, 2)))
    if cm: cm.__exit__()
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print(tag, "rank0", out[0], "rank7", out[-1], flush=True)
run("warm", False, True, False)
run("noflush_barrier", False, True, False)
run("noflush_nobarrier", False, False, False)
run("flush_nobarrier", True, False, False)
run("flush_nobarrier_sampler", True, False, True)
run("noflush_nobarrier_sampler", False, False, True)
dist.barrier(); dist.destroy_process_group()
