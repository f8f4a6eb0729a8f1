"""Batch-parallel sharding helpers: clips are independent units, so ranks take a strided subset
exactly like the reference's validation loop ``range(rank, len(dataset), world_size)``
(/root/reference/basicsr/models/video_base_model.py:44).  No data-path collective exists for inference; the only
exchanges are the timing reduction (max over ranks) and, optionally, metric sums."""
import torch
import torch.distributed as dist


def clip_shard(num_clips, rank, world_size):
    """Indices of the clips this rank processes."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_clips, world_size))


def reduce_max(value, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, device=None):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def max_over_ranks(value, world_size=None):
    """Device-timed milliseconds -> the slowest rank's (bench.py: every multi-GPU number is the max over ranks).
    On an NCCL group the reduction runs on the current CUDA device."""
    if not (dist.is_available() and dist.is_initialized()) or (world_size is not None and world_size <= 1):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
    return reduce_max(value, device=dev)
