"""Multi-GPU plumbing: environments are independent, so they shard statically over ranks
(one process per GPU, no data-path collective) and the only exchange is one all-gather of a
small per-rank metrics vector at the end of a run.

Mirrors the reference's episode sharding `episodes[i::n_processes]` -> GPU i
(experiments/eval_policy_parallel.py:266-279), with torch.distributed (NCCL on GPUs, gloo in
the CPU tests) in place of multiprocessing.Pool + the filesystem."""
from __future__ import annotations

from typing import List, Sequence


def shard_envs(total_envs: int, world: int, rank: int) -> range:
    """Global env ids owned by `rank`: contiguous blocks, sizes differing by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total_envs, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def interleave_scene_types(counts: dict, world: int) -> List[List[str]]:
    """Mixed workloads (BASELINE configs[4]): give every rank the same scene-type ratio.
    counts = {"rope": 1024, "sloth": 512, "tblock": 512} -> per-rank list of scene names."""
    out = [[] for _ in range(world)]
    for name in sorted(counts):
        for r in range(world):
            out[r] += [name] * len(shard_envs(counts[name], world, r))
    return out


def gather_metrics(local: Sequence[float], device=None) -> List[List[float]]:
    """All-gather a fixed-size float64 vector from every rank (rank order).  Works on an
    initialised torch.distributed group of any backend; without one it returns [local]."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [list(map(float, local))]
    t = torch.tensor(list(local), dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
