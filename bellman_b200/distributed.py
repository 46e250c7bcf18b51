"""Multi-GPU plumbing for the base-range sharded MSM: one process per GPU,
torch.distributed for the single exchange step.  The data path has exactly one collective per
proof: an all-gather of the per-shard partial sums (960 bytes per rank)."""
import numpy as np

from . import PARTIALS_BYTES

MSG_BYTES = 159          # error text a failing rank shares with the others (960 + 1 + 159 = 1120 bytes per rank)


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


def create_proof_sharded(assignment, params, full_vk_params, r, s, device_ptrs=None, group=None):
    """One proof over all ranks of `group` (create_proof, groth16/src/prover.rs:217-360, with the
    MSMs sharded): every rank runs prove_partials on its shard, the 960-byte blobs AND each rank's
    status are all-gathered in the same collective, and rank 0 finalises.  A SynthesisError on any
    rank is re-raised on every rank (nobody is left waiting in the all-gather).  Returns the 192
    proof bytes on rank 0, None elsewhere.  `full_vk_params` is any Parameters object of this rank
    (only its verifying-key elements are read by finalize)."""
    import torch
    import torch.distributed as dist

    import threading

    from . import BackendError, _ERRORS, finalize, finalize_static, prove_partials

    # rank 0: the scalar multiplications of the finalisation that need no MSM result run on a host
    # thread while the devices compute the partial sums
    ahead = {}
    worker_thread = None
    if dist.get_rank(group) == 0:
        def _static():
            try:
                ahead["static"] = finalize_static(full_vk_params, r, s)
            except Exception as e:                       # reported by finalize() below if it matters
                ahead["error"] = e
        worker_thread = threading.Thread(target=_static)
        worker_thread.start()
    status, blob, msg = 0, bytes(PARTIALS_BYTES), ""
    try:
        blob = prove_partials(assignment, params, device_ptrs)
    except Exception as e:                               # whatever it is, the other ranks must not be left in the all-gather
        msg = str(e)
        status = next((code for code, cls in _ERRORS.items() if isinstance(e, cls)), 17 if isinstance(e, BackendError) else 255)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    # one collective: partial sums, status and (on failure) the failing rank's own message
    text = msg.encode("utf-8", "replace")[:MSG_BYTES].ljust(MSG_BYTES, b"\0")
    mine = torch.frombuffer(bytearray(blob + bytes([status]) + text), dtype=torch.uint8).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    gathered = [bytes(t.cpu().numpy()) for t in out]
    if worker_thread is not None:
        worker_thread.join()
    for rank, g in enumerate(gathered):
        if g[PARTIALS_BYTES]:
            code = g[PARTIALS_BYTES]
            what = g[PARTIALS_BYTES + 1:].rstrip(b"\0").decode("utf-8", "replace")
            raise _ERRORS.get(code, BackendError)(f"rank {rank} failed with status {code}" + (f": {what}" if what else ""))
    if dist.get_rank(group) == 0:
        return finalize(full_vk_params, [g[:PARTIALS_BYTES] for g in gathered], r, s, static=ahead.get("static"))
    return None
