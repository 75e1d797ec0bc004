"""Two-GPU test of the two-phase PatchMatch schedule with the NCCL depth-map exchange: the result must equal the
single-process run of the same schedule bit for bit (each problem is deterministic).  Needs >= 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene():
    from colmap_b200.synthetic import make_patch_match_scene
    sc = make_patch_match_scene(96, 72, 3, seed=6)
    images = sc["images"]                      # 4 views; every view becomes a reference image with the others as sources
    srcs = [[j for j in range(4) if j != i] for i in range(4)]
    return sc, images, srcs


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from colmap_b200.patch_match import PatchMatchOptions
    from colmap_b200.workspace import run_two_phase
    sc, images, srcs = _scene()
    o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], num_iterations=1, gpu_index=str(rank))
    out = run_two_phase(images, srcs, o, rank, world, device=torch.device("cuda", rank))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, {i: (d, n) for i, (d, n) in out.items()}))


def test_two_phase_nccl_exchange_matches_single_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from colmap_b200.patch_match import PatchMatchOptions
    from colmap_b200.workspace import run_two_phase
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    merged = {}
    for _, d in res:
        merged.update(d)
    sc, images, srcs = _scene()
    o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], num_iterations=1, gpu_index="0")
    ref = run_two_phase(images, srcs, o, 0, 1)
    assert sorted(merged.keys()) == [0, 1, 2, 3]
    for i in range(4):
        assert np.array_equal(merged[i][0].view(np.uint32), ref[i][0].view(np.uint32))
        assert np.array_equal(merged[i][1].view(np.uint32), ref[i][1].view(np.uint32))
        assert (merged[i][0] > 0).mean() > 0.5
