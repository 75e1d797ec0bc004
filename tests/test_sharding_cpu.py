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


def _ws_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    from colmap_b200.patch_match import Image, PatchMatchOptions
    from colmap_b200.workspace import run_two_phase
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    K = np.eye(3, dtype=np.float32)
    imgs = [Image(bitmap=np.full((6, 8), i, np.uint8), K=K, R=K, T=np.zeros(3, np.float32)) for i in range(5)]
    srcs = [[(i + 1) % 5, (i + 2) % 5] for i in range(5)]
    calls = []

    def fake_runner(options, problem):   # stands in for the GPU: depth = image index (+100 in the geometric phase)
        i = problem.ref_image_idx
        if options.geom_consistency:
            # every source depth map must have arrived through the exchange, the reference's own maps are its init
            for j in problem.src_image_idxs:
                assert np.all(problem.depth_maps[j] == float(j)), (i, j)
            assert np.all(problem.depth_maps[i] == float(i)) and np.all(problem.normal_maps[i] == float(i))
            calls.append(("geom", i))
            return np.full((6, 8), 100.0 + i, np.float32), np.zeros((3, 6, 8), np.float32)
        calls.append(("photo", i))
        return np.full((6, 8), float(i), np.float32), np.full((3, 6, 8), float(i), np.float32)
    out = run_two_phase(imgs, srcs, PatchMatchOptions(depth_min=1, depth_max=2), rank, world, runner=fake_runner)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, sorted(out.keys()), calls))


def test_two_phase_workspace_exchange_world_size_2_gloo():
    """The photometric -> (all-gather of depth maps) -> geometric schedule with two ranks and a fake GPU runner."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ws_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = res[0][1] + res[1][1]
    assert sorted(owned) == [0, 1, 2, 3, 4]                     # every image processed exactly once
    for _, keys, calls in res:
        photo = [i for k, i in calls if k == "photo"]; geom = [i for k, i in calls if k == "geom"]
        assert sorted(photo) == sorted(geom) == keys
        assert calls.index(("geom", geom[0])) > calls.index(("photo", photo[-1]))   # phase barrier
