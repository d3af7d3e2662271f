"""8-rank diagnosis of the per-batch all-reduce: torch's communicator vs the library's own (same libnccl instance)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosity_b200 import api, multi
from bench import WORKLOADS
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for mb in (0.2, 3, 12):
    t = torch.ones(int(mb * 1e6 / 4), device="cuda")
    for _ in range(5): dist.all_reduce(t)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): dist.all_reduce(t)
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f"torch all_reduce {mb} MB: {a.elapsed_time(b)/20*1e3:.1f} us", flush=True)
area, N, k, batches, desc = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config3"]
scene = api.Scene(area)
ctx = api.context_for_scene(scene, N, k, device=local, select_mode=api.SELECT_TOPK)
multi.init_nccl(ctx, dist)
ctx.save_state()
for i in range(4):
    ctx.restore_state(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); st = ctx.shoot(batches); dt = time.perf_counter() - t0
    if rank == 0: print(f"lib shoot {batches} batches: gpu_ms {st.gpu_ms:.3f} wall_ms {dt*1e3:.3f}", flush=True)
# host-mediated through torch's communicator, dB kept on the device? (read/write through host here)
ctx.restore_state(); ctx.set_partition(rank, world)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
multi.shoot_batches_hosted(ctx, dist, batches)
if rank == 0: print(f"hosted {batches} batches wall_ms {(time.perf_counter()-t0)*1e3:.3f}", flush=True)
dist.barrier(); dist.destroy_process_group()
