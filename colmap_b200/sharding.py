"""Multi-GPU sharding of the two paths.

PatchMatch: one reference image = one independent problem; COLMAP runs one host thread per GPU pulling problems
from a pool (src/colmap/mvs/patch_match.cc:176-205).  Here: one process per GPU (torch.distributed), problems
assigned largest-first to the least loaded rank, no data-path collective in the photometric phase; timings are
reduced with MAX over ranks.
"""
from typing import List, Sequence


def assign_problems(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment: returns, per rank, the list of problem indices (ascending).
    Deterministic: ties go to the lower rank / lower problem index."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (device time in ms) over the default process group; identity without one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count: int, device=None) -> List[int]:
    """All-gather of an integer per rank (e.g. reference pixels processed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(count)]
    t = torch.tensor([count], dtype=torch.int64, device=device if device is not None else "cpu")
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [int(x.item()) for x in outs]
