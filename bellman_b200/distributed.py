"""Multi-GPU plumbing for the base-range sharded MSM: one process per GPU,
torch.distributed for the single exchange step.  The data path has exactly one collective per
proof: an all-gather of the per-shard partial sums (960 bytes per rank)."""
import numpy as np

from . import PARTIALS_BYTES


def shard_range(length, index, count):
    """Contiguous range `index` of `count` over `length` items:
    [length*index/count, length*(index+1)/count) -- the split bb_crs_create applies to the base
    vectors (after it has set aside up to 4 window groups, see prover.cu: shard_policy)."""
    return length * index // count, length * (index + 1) // count


def all_gather_partials(partials, group=None):
    """All-gather one 960-byte partial-sum blob per rank; returns the list ordered by rank.
    Works on NCCL (blob staged through a CUDA tensor) and on gloo (CPU tensor)."""
    import torch
    import torch.distributed as dist

    assert len(partials) == PARTIALS_BYTES
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(partials), dtype=torch.uint8).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().numpy()) for t in out]


def max_over_ranks(seconds, group=None):
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
