"""World-size-2 CPU (gloo) coverage of the view-parallel plumbing in pegasus_b200/dist.py: round-robin
frame sharding, the pose-packet broadcast from the rank that owns the trajectory, and the
max-over-ranks timing reduction bench.py uses.  (The reference is single-process)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, K, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pegasus_b200 import dist as pgd
    from pegasus_b200.sh_rotation import generate_pose_packets, quat_xyzw_to_rotation
    r, w, _ = pgd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    packets = torch.zeros((n_frames, K, 103), dtype=torch.float32)
    if rank == 0:  # only the physics rank knows the trajectory
        rng = np.random.default_rng(5)
        for f in range(n_frames):
            poses = [(quat_xyzw_to_rotation(rng.normal(size=4)), rng.normal(size=3)) for _ in range(K)]
            packets[f] = torch.from_numpy(generate_pose_packets(poses, np.zeros((K, 3), np.float32)))
    pgd.broadcast_pose_packets(packets, src=0)
    mine = pgd.shard_frames(n_frames, rank, world)
    # every rank reports a checksum of the packets of ITS frames + its frame list
    np.save(os.path.join(out_dir, f"frames_{rank}.npy"), np.asarray(mine))
    np.save(os.path.join(out_dir, f"packets_{rank}.npy"), packets.numpy())
    t = pgd.max_over_ranks(10.0 + rank)
    s = pgd.sum_over_ranks(len(mine))
    assert t == 10.0 + (world - 1)
    assert s == n_frames
    pgd.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_pose_broadcast(tmp_path):
    world, n_frames, K = 2, 7, 3
    mp.spawn(_worker, args=(world, _free_port(), n_frames, K, str(tmp_path)), nprocs=world, join=True)
    f0, f1 = (np.load(tmp_path / f"frames_{r}.npy") for r in range(2))
    assert sorted(list(f0) + list(f1)) == list(range(n_frames))      # a partition of the frames
    assert list(f0) == [0, 2, 4, 6] and list(f1) == [1, 3, 5]           # frame f -> rank f % world
    p0, p1 = (np.load(tmp_path / f"packets_{r}.npy") for r in range(2))
    assert np.array_equal(p0, p1) and np.abs(p0).sum() > 0              # rank 1 received rank 0's poses
    # packets are well formed: R orthonormal, q unit, rotate_sh flag set
    R = p1[..., 0:9].reshape(-1, 3, 3)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-5
    assert np.abs(np.linalg.norm(p1[..., 15:19], axis=-1) - 1).max() < 1e-6
    assert (p1[..., 102].view(np.int32) == 1).all()


def test_single_process_is_a_noop():
    from pegasus_b200 import dist as pgd
    assert pgd.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]
    t = torch.ones(2, 103)
    assert pgd.broadcast_pose_packets(t) is t
    assert pgd.max_over_ranks(3.5) == 3.5
