"""Batched multi-GPU shooting: host-side orchestration (one process per GPU, torch.distributed for the plumbing).

The path shards at the shooter level: scene and state are replicated, every rank runs the same
deterministic selection, rank r renders and processes emitters [r*k/G, (r+1)*k/G) of the batch, the received energy
dB[P][3] is summed over ranks, and every rank applies the identical update.

Three ways to combine dB:
  * fused over peer memory (Context.peer_init + Context.shoot): the ranks' exchange buffers are mapped into each other
    over NVLink (CUDA IPC); the update kernel waits for the peers' release flags and sums their dB planes itself —
    no collective call at all (the default of bench.py);
  * in-library NCCL (Context.comm_init + Context.shoot): one ncclAllReduce per batch on the context's stream —
    the production path, no host round trip;
  * host-mediated (shoot_batches_hosted below): read dB, all_reduce through any torch.distributed backend, write it
    back — used by the CPU `gloo` tests of the orchestration and as a fallback when NCCL cannot be loaded.
"""
import numpy as np


def shooter_range(k, rank, world):
    """Emitter slots [h0, h1) of a k-emitter batch owned by `rank` — must match rad_set_partition() in rad_cuda.cu."""
    return (k * rank) // world, (k * (rank + 1)) // world


def init_nccl(ctx, dist):
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    return rank, world


def init_peer(ctx, dist):
    """Fused exchange over peer memory: all-gather the ranks' CUDA IPC handles, map the peers' exchange buffers.
    Raises on EVERY rank if any rank could not export its buffer (the all-gather itself always completes); a rank whose
    mapping of a peer fails raises on its own — callers that want a common fallback agree on it afterwards (bench.py)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    try:
        mine = ctx.peer_handle()
    except Exception:                            # noqa: BLE001 — reported below, on all ranks
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    if any(h is None for h in handles):
        raise RuntimeError("rad_peer_handle failed on rank(s) " + ", ".join(str(i) for i, h in enumerate(handles) if h is None))
    ctx.peer_init(rank, world, handles)
    return rank, world


def shoot_batches_hosted(engine, dist, n_batches):
    """Host-mediated batches.  `engine` needs batch_partial(), read_delta(), write_delta(dB), batch_finish()."""
    import torch
    last = 0.0
    for _ in range(n_batches):
        engine.batch_partial()
        dB = torch.from_numpy(np.ascontiguousarray(engine.read_delta(), dtype=np.float32))
        if dist is not None and dist.get_world_size() > 1:
            if dist.get_backend() == "nccl":          # NCCL reduces device tensors only
                dB = dB.cuda()
            dist.all_reduce(dB, op=dist.ReduceOp.SUM)
            dB = dB.cpu()
        engine.write_delta(dB.numpy())
        last = engine.batch_finish()
    return last
