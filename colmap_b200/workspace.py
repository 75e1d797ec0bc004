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


def _device_runner(device):
    """Runner that leaves the maps in HBM: CUDA tensors filled by b200pm_get_*_device."""
    import torch

    def run(options: PatchMatchOptions, problem: Problem):
        pm = PatchMatch(options, problem)
        pm.Run()
        h, w, _ = pm._dims
        d = pm.GetDepthMapDevice(torch.empty((h, w), dtype=torch.float32, device=device))
        n = pm.GetNormalMapDevice(torch.empty((3, h, w), dtype=torch.float32, device=device))
        pm.close()
        return d, n
    return run


def all_gather_maps(local: Dict[int, "np.ndarray"], num_images: int, shape, device=None) -> List:
    """All-gather of per-image float32 maps keyed by image index (every image owned by exactly one rank).
    `local` holds numpy arrays or CUDA tensors; with CUDA tensors nothing touches the host: the owned maps are packed
    into one device buffer, ONE all-gather over NCCL moves them, and the result is a list of device views.  (numpy in,
    numpy out: the gloo / single-process form used by the CPU tests.)"""
    import torch
    import torch.distributed as dist
    on_device = any(hasattr(v, "data_ptr") for v in local.values())
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local[i] for i in range(num_images)]
    world = dist.get_world_size()
    dev = device if (device is not None and on_device) else (device or "cpu")
    owned = sorted(local.keys())
    counts = torch.tensor([len(owned)], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    cap = int(max(int(c.item()) for c in all_counts))
    idx = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    buf = torch.zeros((cap,) + tuple(shape), dtype=torch.float32, device=dev)
    for k, i in enumerate(owned):
        idx[k] = i
        v = local[i]
        buf[k].copy_(v if hasattr(v, "data_ptr") else torch.from_numpy(np.ascontiguousarray(v, np.float32)))   # D2D when resident
    idx_all = torch.empty((world * cap,), dtype=idx.dtype, device=dev)                    # rank-major concatenation
    buf_all = torch.empty((world * cap,) + tuple(shape), dtype=buf.dtype, device=dev)
    dist.all_gather_into_tensor(idx_all, idx)
    dist.all_gather_into_tensor(buf_all, buf)       # the depth-map exchange
    out: List = [None] * num_images
    ids = idx_all.cpu().numpy()
    for k in range(world * cap):
        i = int(ids[k])
        if i >= 0:
            out[i] = buf_all[k] if on_device else buf_all[k].cpu().numpy()
    return out


def run_two_phase(images: Sequence[Image], source_lists: Sequence[Sequence[int]], options: PatchMatchOptions,
                  rank: int = 0, world: int = 1, device=None,
                  runner: Optional[Callable] = None, resident: Optional[bool] = None):
    """Photometric + geometric PatchMatch for every image (image i uses source_lists[i]).  Returns
    {image index: (depth, normal)} of the geometric phase for the images this rank owns (numpy arrays).

    With a CUDA `device` and the default runner the photometric depth / normal maps never leave HBM: they are exported
    device-to-device (b200pm_get_*_device), all-gathered over NCCL and handed to the geometric problems as device
    pointers (b200pm_problem::maps_on_device)."""
    n = len(images)
    if resident is None:
        resident = runner is None and device is not None and str(device).startswith("cuda")
    if runner is None:
        runner = _device_runner(device) if resident else _default_runner
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
    zero_normal = None
    for i in mine:  # phase 2 (patch_match.cc:199-204)
        if resident:
            import torch
            if zero_normal is None:
                zero_normal = torch.zeros((3,) + shape, dtype=torch.float32, device=device)
            normals = [normal_local.get(j, zero_normal) for j in range(n)]      # only the reference's normal map is read
        else:
            normals = [normal_local.get(j, np.zeros((3,) + shape, np.float32)) for j in range(n)]
        prob = Problem(ref_image_idx=i, src_image_idxs=list(source_lists[i]), images=list(images),
                       depth_maps=depth_all, normal_maps=normals)
        d, nrm = runner(geom, prob)
        out[i] = (d.cpu().numpy(), nrm.cpu().numpy()) if hasattr(d, "data_ptr") else (d, nrm)
    return out
