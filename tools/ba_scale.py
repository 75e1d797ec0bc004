"""B3 point-sharded over the ranks of a torchrun launch: LM iterations/s with the peer-memory all-reduce and with NCCL, then
the device-timeline phase breakdown of one solve (B200BA_PROFILE).  torchrun --nproc-per-node N tools/ba_scale.py"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from colmap_b200.bundle_adjustment import (ITERATIVE_SCHUR, SIMPLE_RADIAL, BAComm, BundleAdjustmentOptions, shard_flat_problem,
                                           solve_flat, solve_flat_sharded)
from colmap_b200.synthetic import synthesize_ba_problem

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
gt, noisy = synthesize_ba_problem(500, 300000, 7, models=(SIMPLE_RADIAL,), seed=42, num_obs=2000000)
noisy.pose_constant = noisy.pose_constant.copy(); noisy.pose_fixed_dim = noisy.pose_fixed_dim.copy()
noisy.pose_constant[0] = 1
noisy.pose_fixed_dim[1] = int(np.argmax(np.abs(noisy.poses[1, 4:] - noisy.poses[0, 4:])))
o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR, gpu_index=lr)


def new_comm():
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.tensor(list(BAComm.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, src=0)
    return BAComm(bytes(idt.cpu().tolist()), rank, world)


def run(comm, n):
    out = []
    for _ in range(n):
        if world > 1:
            local = shard_flat_problem(noisy, rank, world)
            torch.cuda.synchronize(); dist.barrier()
            t = time.time(); s = solve_flat_sharded(o, local, comm); torch.cuda.synchronize(); dist.barrier()
        else:
            local = noisy.copy(); local.pose_constant, local.pose_fixed_dim = noisy.pose_constant, noisy.pose_fixed_dim
            t = time.time(); s = solve_flat(o, local)
        out.append((s.num_successful_steps + s.num_unsuccessful_steps, s.solve_ms, (time.time() - t) * 1e3, s.final_cost))
    return out


modes = ("peer", "nccl") if world > 1 else ("single",)
for mode in modes:
    if mode == "nccl":
        os.environ["B200BA_NO_P2P"] = "1"
    else:
        os.environ.pop("B200BA_NO_P2P", None)
    comm = new_comm() if world > 1 else None
    r = run(comm, 4)[1:]
    if rank == 0:
        lm = sum(x[0] for x in r); dev = sum(x[1] for x in r); wall = sum(x[2] for x in r)
        print(f"world {world} {mode:6s} peer_memory={comm.peer_memory() if comm else None}: {lm / dev * 1e3:8.1f} LM it/s device, "
              f"{lm / wall * 1e3:8.1f} wall, cost {r[-1][3]:.6f}", flush=True)
    if mode != "nccl":
        os.environ["B200BA_PROFILE"] = "1"
        run(comm, 1)
        os.environ.pop("B200BA_PROFILE")
    if comm:
        comm.close()
if world > 1:
    dist.destroy_process_group()
