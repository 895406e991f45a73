#!/usr/bin/env python
"""Per-stage kernel times of the bench workload (configs[1]) with one frame in flight — a quicker look than
bench.py for kernel A/B work.  Usage: python tools/stage_times.py [--masks 0|1] [--frames N] [--tag name]
Environment selects kernel variants (PG_COMP_VARIANT, PG_NUMERICS, PG_LIB_PATH ...)."""
import argparse
import colorsys
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--masks", type=int, default=1)
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--env-n", type=int, default=2_000_000)
    ap.add_argument("--objects", type=int, default=5)
    ap.add_argument("--obj-n", type=int, default=200_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    from pegasus_b200 import Camera, ComposedScene, _lib, synth
    dev = torch.device("cuda", 0)
    L = _lib.load()
    env = synth.make_env(args.env_n, seed=1000)
    objs = {i + 1: synth.make_object(args.obj_n, seed=2000 + i) for i in range(args.objects)}
    cams_h = synth.orbit_cameras(100, args.width, args.height, seed=3000)
    ncol = max(args.objects, 1)
    colors = np.asarray([colorsys.hls_to_rgb(i / ncol, 0.6, 0.7)[::-1] for i in range(ncol)], dtype=np.float32)
    scene = ComposedScene(env, objs, colors, device=dev, sh_mode="rotate")
    scene.set_poses(synth.static_poses(args.objects, seed=4000))
    cams = [Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], args.width, args.height, device=dev) for c in cams_h]
    bg = torch.zeros(3, device=dev)
    masks = bool(args.masks)
    out = scene.alloc_outputs(args.width, args.height, masks=masks)
    cap = 0
    for i in range(0, 100, 7):
        o = scene.render(cams[i], bg, masks=masks, out=out, sync_check=True)
        cap = max(cap, o["num_stored"])
    cap = int(cap * 1.1) + 4096
    for i in range(5):
        scene.render(cams[i], bg, masks=masks, out=out, sync_check=False, pair_capacity=cap)
    torch.cuda.synchronize()
    _lib.check(L.pg_profile_enable(args.frames), "pg_profile_enable")
    for i in range(args.frames):
        scene.render(cams[(5 + i) % 100], bg, masks=masks, out=out, sync_check=False, pair_capacity=cap)
    torch.cuda.synchronize()
    st = np.zeros((args.frames, _lib.NUM_STAGES), dtype=np.float32)
    buf = (C.c_float * _lib.NUM_STAGES)()
    for f in range(int(L.pg_profile_frames())):
        _lib.check(L.pg_profile_read(f, buf), "pg_profile_read")
        st[f] = np.frombuffer(buf, dtype=np.float32)
    L.pg_profile_enable(0)
    m = st.mean(axis=0)
    res = {n: round(float(v), 4) for n, v in zip(_lib.STAGE_NAMES, m)}
    res["sum"] = round(float(m.sum()), 4)
    if masks:
        acc = {}
        for i in range(0, 100, 13):
            scene.render(cams[i], bg, masks=True, out=out, sync_check=True, pair_capacity=cap, debug=2)
            for k, v in scene.read_stats().items():
                acc[k] = acc.get(k, 0) + v / 8.0
        res["stats"] = {k: round(v) for k, v in acc.items()}
    env_desc = {k: v for k, v in os.environ.items() if k.startswith("PG_")}
    print(json.dumps({"tag": args.tag, "masks": masks, "env": env_desc, "ms": res}))


if __name__ == "__main__":
    main()
