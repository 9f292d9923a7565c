"""Multi-GPU plumbing of the hot path: frames are independent, so N GPUs run N disjoint batches (one process per
GPU, torchrun-style env) and there is NO collective on the data path.  torch.distributed is used only for the
barrier and for the MAX-over-ranks reduction of measured times."""
from __future__ import annotations

import os

import torch


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str, device=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (no-op for one process)."""
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def rank_seed(base_seed: int, rank: int) -> int:
    """Each rank synthesises its own frames: partition = batch dimension, disjoint by construction."""
    return int(base_seed) + int(rank)


def shard_frames(n_frames: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a global list of frames for this rank (strong-scaling callers, e.g. a
    DistributedSampler-free evaluation loop)."""
    per, rem = divmod(int(n_frames), int(world))
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (times are always reported as the slowest rank's)."""
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def aggregate_rate(units_per_rank_per_step: int, world: int, steps: int, ms_total: float) -> float:
    """Whole-job throughput: the units all ranks processed divided by the slowest rank's time."""
    return units_per_rank_per_step * world * steps / (ms_total * 1e-3)
