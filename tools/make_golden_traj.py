#!/usr/bin/env python
"""Body 1 of the reference's recorded PyBullet run (/root/reference/src/engine/simulation_steps.json: the only pose
fixture the reference ships) -> tests/golden/simulation_body1.npz: t (steps, 3), q xyzw (steps, 4),
float64, in the file's own step order.  The dynamic workload (BASELINE.json configs[2]) replays it with per-object
time / space offsets; the GPU box has no /root/reference, so the excerpt is committed.

Run in the build container:  python tools/make_golden_traj.py"""
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "simulation_body1.npz")

d = json.load(open(os.path.join(REF, "src/engine/simulation_steps.json")))
body = d["trajectory"]["1"]
keys = list(body.keys())  # insertion order = step order (static_object_pose takes keys[-1])
t = np.array([body[k]["t"] for k in keys], dtype=np.float64)
q = np.array([body[k]["q"] for k in keys], dtype=np.float64)
np.savez_compressed(OUT, t=t, q=q, steps=np.array([int(k) for k in keys]))
print(OUT, t.shape, q.shape, "z from", t[0, 2], "to", t[-1, 2])
