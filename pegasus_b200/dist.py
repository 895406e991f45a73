"""View-parallel multi-GPU plumbing (one process per GPU, torch.distributed).

The reference is single-process; frames are independent given the poses, so the scene
is replicated per GPU, frame f goes to rank f % world, and the only data-path communication is a
broadcast of the per-frame pose packets (K x 412 bytes) from the rank that owns the physics
trajectory.  The pose kernel reads the broadcast's receive buffer directly.
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> tuple:
    """Returns (rank, world, local_rank); initialises the process group when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_frames(num_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame ownership: frame f -> rank f % world."""
    return list(range(rank, num_frames, world))


def broadcast_pose_packets(packets: torch.Tensor, src: int = 0) -> torch.Tensor:
    """In-place broadcast of a (frames, K, 103) or (K, 103) float32 packet tensor from `src`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(packets, src=src)
    return packets


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()
