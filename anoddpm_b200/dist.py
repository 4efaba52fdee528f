"""Data-parallel plumbing for the detection path: one process per GPU, images sharded by global index,
no collective inside the sampling loop (each image's chain depends only on its own x, t, seed and the
replicated weights -- reference detection.py:209-227 iterates images independently).  Only scalar
results (timings, counts) are reduced at the end.  Training adds the one gradient exchange per step of SURVEY
section 8(e): optim.GradReducer averages the flat gradient in buckets, tail first, overlapped with the native backward.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """(rank, world, local_rank).  Initialises torch.distributed when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous block of global indices owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def seeds_for_step(master_seed, step, global_indices, channels=1):
    """Simplex seeds of one sampling step for the given GLOBAL image indices, (index-major, channel-minor).
    A function of (master_seed, step, global index, channel) only, so results do not depend on the world size
    or on which rank owns an image."""
    out = []
    for gi in global_indices:
        for c in range(channels):
            ss = np.random.SeedSequence([int(master_seed) & 0xFFFFFFFF, int(step), int(gi), int(c)])
            v = int(ss.generate_state(1, dtype=np.uint64)[0] % 20000000000) - 10000000000
            out.append(v if v != 0 else 1)
    return out


def max_over_ranks(value, device):
    """max of a python float over ranks (device-timed numbers are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def shutdown():
    """tear the process group down (no-op when torch.distributed was never initialised)"""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
