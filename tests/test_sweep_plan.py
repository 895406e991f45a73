"""Host logic of the dataset sweep (BASELINE.json configs[4]: 30 object clouds x 4 environments, 10 000 views
sharded over 1/2/4/8 ranks): every (scene, view) pair is rendered exactly once, ranks are balanced to one view,
each rank walks a contiguous run of scenes, and the plan itself does not depend on the number of ranks.
Includes the world-size-2 gloo run of the same plan."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from pegasus_b200 import sweep


def covered(scenes, world):
    seen = {}
    for r in range(world):
        for it in sweep.shard_views(scenes, r, world):
            for v in range(it.first_view, it.first_view + it.n_views):
                key = (it.scene.scene_id, v)
                assert key not in seen, f"{key} rendered by ranks {seen[key]} and {r}"
                seen[key] = r
    return seen


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_configs4_sweep_is_a_balanced_partition(world):
    scenes = sweep.plan_scenes(n_envs=4, n_objects=30, total_views=10_000, views_per_scene=300, seed=7)
    assert sum(s.n_views for s in scenes) == 10_000 and len(scenes) == 34 and scenes[-1].n_views == 100
    for s in scenes:
        assert s.env == s.scene_id % 4 and 3 <= len(s.objects) <= 6
        assert len(set(s.objects)) == len(s.objects) and all(0 <= o < 30 for o in s.objects)
    seen = covered(scenes, world)
    assert len(seen) == 10_000                                            # nothing skipped, nothing twice
    per_rank = np.bincount(list(seen.values()), minlength=world)
    assert per_rank.max() - per_rank.min() <= 1                            # balanced to one view
    for r in range(world):
        items = sweep.shard_views(scenes, r, world)
        ids = [it.scene.scene_id for it in items]
        assert ids == list(range(ids[0], ids[0] + len(ids)))              # a contiguous run of scenes
        for it in items[1:-1]:                                            # only the end scenes can be partial
            assert it.first_view == 0 and it.n_views == it.scene.n_views
    assert len(scenes) <= sweep.scene_loads(scenes, world) <= len(scenes) + world - 1


def test_plan_is_independent_of_the_world_size_and_seeded():
    a = sweep.plan_scenes(4, 30, 2_000, 250, seed=3)
    b = sweep.plan_scenes(4, 30, 2_000, 250, seed=3)
    c = sweep.plan_scenes(4, 30, 2_000, 250, seed=4)
    assert a == b and a != c
    assert len({s.objects for s in a}) > 1                                # scenes differ in their object subsets
    # more ranks than views, empty sweeps, argument errors
    tiny = sweep.plan_scenes(1, 2, 3, 10, k_min=1, k_max=2)
    assert sum(len(sweep.shard_views(tiny, r, 8)) for r in range(8)) == 3
    assert sweep.plan_scenes(4, 30, 0, 300) == [] and sweep.shard_views([], 0, 2) == []
    with pytest.raises(ValueError):
        sweep.plan_scenes(4, 3, 100, 10, k_min=2, k_max=5)
    with pytest.raises(ValueError):
        sweep.shard_views(tiny, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pegasus_b200 import dist as pgd
    r, w, _ = pgd.init_from_env(backend="gloo")
    scenes = sweep.plan_scenes(4, 30, 1_001, 120, seed=11)    # every rank derives the same plan, nothing is sent
    items = sweep.shard_views(scenes, r, w)
    mine = sum(it.n_views for it in items)
    assert pgd.sum_over_ranks(mine) == 1_001
    assert pgd.max_over_ranks(mine) - mine <= 1
    np.save(os.path.join(out_dir, f"views_{rank}.npy"),
            np.asarray([(it.scene.scene_id, it.first_view, it.n_views) for it in items]))
    pgd.barrier()
    dist.destroy_process_group()


def test_two_ranks_derive_disjoint_work_lists_from_the_same_plan(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    v0, v1 = (np.load(tmp_path / f"views_{r}.npy") for r in range(2))
    assert v0[:, 2].sum() + v1[:, 2].sum() == 1_001
    assert v0[-1, 0] <= v1[0, 0]                               # rank 0's scenes come before rank 1's
    if v0[-1, 0] == v1[0, 0]:                                  # the scene the cut falls into is split, not duplicated
        assert v0[-1, 1] + v0[-1, 2] == v1[0, 1]
