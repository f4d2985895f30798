"""Batch-of-sequences data parallelism across the GPUs of one node.

The per-frame path shards over independent sequences only (the reference itself runs one
process per GPU with no communication, lib/test/evaluation/running.py:97-100,170).  One process per GPU, rank r owns
a contiguous block of sequences, there is no per-frame collective, and the run ends with ONE all-gather of the
per-sequence trajectories [S_local, T, 4] -> [S_total, T, 4] on every rank (NCCL over NVLink on the B200 box, gloo in
the CPU tests).
"""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_sequences: int, rank: int, world: int):
    """Contiguous, balanced block of sequence indices owned by `rank` (the first n % world ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_sequences, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_process_group(backend: str | None = None):
    """Join the torchrun rendezvous (MASTER_ADDR/PORT, RANK, WORLD_SIZE from the environment)."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if world == 1:
        return rank, world, local
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}  # bind the communicator to the GPU
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_trajectories(local, n_sequences: int | None = None):
    """All-gather per-rank trajectories [S_local, T, C] into [S_total, T, C] on every rank, in rank order: the ONE
    collective of a run.

    Ranks may own different numbers of sequences; the shard sizes follow from ``shard_range`` (pass ``n_sequences`` = the
    global sequence count), so no second collective is needed to exchange them.  Blocks are padded to the largest shard
    for the collective and trimmed afterwards.  With one process this is the identity."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    if n_sequences is None:
        n_sequences = local.shape[0] * world  # equal shards
    counts = [shard_range(n_sequences, r, world) for r in range(world)]
    counts = [b - a for a, b in counts]
    if counts[rank] != local.shape[0]:
        raise RuntimeError(f"rank {rank} holds {local.shape[0]} sequences, shard_range gives {counts[rank]} of {n_sequences}")
    smax = max(counts)
    if local.shape[0] == smax:
        send = local.contiguous()
    else:
        send = local.new_zeros((smax,) + tuple(local.shape[1:]))
        send[: local.shape[0]] = local
    if dist.get_backend() == "nccl":
        recv = local.new_empty((world * smax,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(recv, send)
    else:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send)
        recv = torch.cat(parts, dim=0)
    if all(c == smax for c in counts):
        return recv
    return torch.cat([recv[r * smax: r * smax + counts[r]] for r in range(world)], dim=0)


def warmup_gather(like):
    """One all-gather of the same shape / dtype as the run's final ``gather_trajectories`` call, issued BEFORE a timed
    region: NCCL builds its communicator, channels and kernels lazily at the first collective (several milliseconds),
    which is set-up cost of the process, not of the run."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    gather_trajectories(torch.zeros_like(like))
    if like.is_cuda:
        torch.cuda.synchronize()
