"""Two-phase (photometric -> geometric) PatchMatch over a set of images, sharded over GPUs.

Mirrors the schedule of mvs::PatchMatchController::Run (src/colmap/mvs/patch_match.cc:170-207): photometric problems for
every reference image, a barrier, then the geometric problems that read the photometric depth / normal maps of their
source images.  The reference exchanges those maps through the file system (`*.photometric.bin` + CachedWorkspace,
patch_match.cc:182-197,226); here every rank keeps its maps on its GPU and the exchange is ONE all-gather over NCCL
(torch.distributed) between the phases — the only collective of the path.

Workspace file I/O (`patch-match.cfg`, `.bin` maps, image decoding) is the controller's job and stays out of scope.
"""
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .patch_match import Image, PatchMatch, PatchMatchOptions, Problem
from .sharding import assign_problems


def _default_runner(options: PatchMatchOptions, problem: Problem):
    pm = PatchMatch(options, problem)
    pm.Run()
    out = (pm.GetDepthMap(), pm.GetNormalMap())
    pm.close()
    return out


def all_gather_maps(local: Dict[int, np.ndarray], num_images: int, shape, device=None) -> List[np.ndarray]:
    """All-gather of per-image float32 maps keyed by image index (every image owned by exactly one rank).
    With an initialised torch.distributed group the payload travels as device tensors over NCCL (or gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local[i] for i in range(num_images)]
    world = dist.get_world_size()
    owned = sorted(local.keys())
    counts = torch.tensor([len(owned)], dtype=torch.int64, device=device or "cpu")
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    cap = int(max(int(c.item()) for c in all_counts))
    idx = torch.full((cap,), -1, dtype=torch.int64, device=device or "cpu")
    buf = torch.zeros((cap,) + tuple(shape), dtype=torch.float32, device=device or "cpu")
    for k, i in enumerate(owned):
        idx[k] = i
        buf[k] = torch.from_numpy(np.ascontiguousarray(local[i], np.float32)).to(buf.device)
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    buf_all = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(idx_all, idx)
    dist.all_gather(buf_all, buf)       # the depth-map exchange
    out: List[Optional[np.ndarray]] = [None] * num_images
    for ids, b in zip(idx_all, buf_all):
        ids = ids.cpu().numpy()
        b = b.cpu().numpy()
        for k, i in enumerate(ids):
            if i >= 0:
                out[int(i)] = b[k]
    return out  # type: ignore


def run_two_phase(images: Sequence[Image], source_lists: Sequence[Sequence[int]], options: PatchMatchOptions,
                  rank: int = 0, world: int = 1, device=None,
                  runner: Callable = _default_runner):
    """Photometric + geometric PatchMatch for every image (image i uses source_lists[i]).  Returns
    {image index: (depth, normal)} of the geometric phase for the images this rank owns."""
    n = len(images)
    costs = [images[i].GetWidth() * images[i].GetHeight() * max(len(source_lists[i]), 1) for i in range(n)]
    mine = assign_problems(costs, world)[rank]
    photo = PatchMatchOptions(**{**options.__dict__, "geom_consistency": False, "filter": False})
    depth_local, normal_local = {}, {}
    for i in mine:  # phase 1 (patch_match.cc:182-195)
        d, nrm = runner(photo, Problem(ref_image_idx=i, src_image_idxs=list(source_lists[i]), images=list(images)))
        depth_local[i], normal_local[i] = d, nrm
    shape = (images[0].GetHeight(), images[0].GetWidth())
    depth_all = all_gather_maps(depth_local, n, shape, device)    # barrier + exchange (replaces the disk round trip)
    geom = PatchMatchOptions(**{**options.__dict__, "geom_consistency": True})
    out = {}
    for i in mine:  # phase 2 (patch_match.cc:199-204)
        normals = [normal_local.get(j, np.zeros((3,) + shape, np.float32)) for j in range(n)]  # only the reference's normal map is read
        prob = Problem(ref_image_idx=i, src_image_idxs=list(source_lists[i]), images=list(images),
                       depth_maps=depth_all, normal_maps=normals)
        out[i] = runner(geom, prob)
    return out
