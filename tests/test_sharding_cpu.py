"""N > 1 host logic on CPU: problem assignment and the max-over-ranks / gather reductions with world_size 2 (gloo)."""
import os
import socket
import sys

import torch.multiprocessing as mp

from colmap_b200.sharding import assign_problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_problems_balances_and_covers():
    costs = [8.3, 2.1, 2.1, 8.3, 4.0, 1.0, 8.3]
    for world in (1, 2, 4, 8):
        parts = assign_problems(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)          # LPT bound
    assert assign_problems([], 3) == [[], [], []]
    assert assign_problems([1.0] * 200, 8) == [list(range(r, 200, 8)) for r in range(8)]   # 200-image workspace, 8 GPUs


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from colmap_b200.sharding import gather_counts, max_over_ranks
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = max_over_ranks(10.0 + rank)
    counts = gather_counts(100 * (rank + 1))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, m, counts))


def test_reductions_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 11.0
    assert res[0][2] == res[1][2] == [100, 200]
