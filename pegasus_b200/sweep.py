"""Dataset sweeps over many scenes (BASELINE.json configs[4]; SURVEY §8 e): which rank renders which views.

The reference builds ONE scene per run (`pegasus.py` `__main__`: one environment, a hand-picked object list,
one camera path) and is started again for the next dataset.  A sweep — e.g. 30 object clouds x 4
environments, 10 000 camera views — is a list of such scenes, each with its own random subset of the object
clouds, and the unit of distribution is the (scene, view) pair: frames are independent given the poses, so
nothing is exchanged between ranks.  What is NOT free is switching scenes: a rank has to load the clouds,
build the ComposedScene (activations, canonical object arrays) and calibrate the pair capacity.  The plan
therefore cuts the flattened (scene, view) sequence into `world` CONTIGUOUS parts of equal size (±1 view): a
rank walks a contiguous run of scenes, and at most world - 1 scenes are built twice (those a cut falls into).

Everything here is host-side planning (deterministic, seeded, no torch): `plan_scenes` decides the scenes,
`shard_views` gives a rank its work list, `scene_loads` counts what a plan costs in scene switches.  The
rendering of one work item is the DatasetGenerator loop of pegasus_b200/generate.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class SceneSpec:
    scene_id: int               # BOP scene directory train/<scene_id:06d>
    env: int                    # index of the environment cloud
    objects: Tuple[int, ...]    # indices of the object clouds merged into it, in merge order (bullet ids 1..K)
    n_views: int                # camera views of this scene
    seed: int                   # seed of this scene's camera path / pose draw


@dataclass(frozen=True)
class WorkItem:
    scene: SceneSpec
    first_view: int             # views [first_view, first_view + n_views) of the scene, in path order
    n_views: int


def plan_scenes(n_envs: int, n_objects: int, total_views: int, views_per_scene: int, k_min: int = 3, k_max: int = 6,
                seed: int = 0) -> List[SceneSpec]:
    """Scene s uses environment s % n_envs and a random subset of k in [k_min, k_max] object clouds (drawn
    without replacement from a generator seeded with (seed, s), so the plan does not depend on the number of
    ranks); every scene has `views_per_scene` views, the last one the remainder."""
    if n_envs <= 0 or n_objects <= 0 or total_views < 0 or views_per_scene <= 0:
        raise ValueError("n_envs, n_objects, views_per_scene must be positive and total_views non-negative")
    if not (0 <= k_min <= k_max <= n_objects):
        raise ValueError("need 0 <= k_min <= k_max <= n_objects")
    scenes, left, s = [], int(total_views), 0
    while left > 0:
        rng = np.random.default_rng([int(seed), s])
        k = int(rng.integers(k_min, k_max + 1))
        objs = tuple(int(o) for o in rng.choice(n_objects, size=k, replace=False))
        n = min(left, int(views_per_scene))
        scenes.append(SceneSpec(scene_id=s, env=s % n_envs, objects=objs, n_views=n, seed=int(rng.integers(1 << 31))))
        left -= n
        s += 1
    return scenes


def shard_views(scenes: Sequence[SceneSpec], rank: int, world: int) -> List[WorkItem]:
    """Rank `rank`'s part of the flattened (scene, view) sequence: global views [lo, hi) with
    lo = floor(rank * V / world), hi = floor((rank + 1) * V / world), cut into one WorkItem per scene touched."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    total = sum(s.n_views for s in scenes)
    lo, hi = rank * total // world, (rank + 1) * total // world
    items, base = [], 0
    for s in scenes:
        a, b = max(lo, base), min(hi, base + s.n_views)
        if b > a:
            items.append(WorkItem(scene=s, first_view=a - base, n_views=b - a))
        base += s.n_views
    return items


def scene_loads(scenes: Sequence[SceneSpec], world: int) -> int:
    """Scene builds summed over all ranks (>= len(scenes), <= len(scenes) + world - 1)."""
    return sum(len(shard_views(scenes, r, world)) for r in range(world))
