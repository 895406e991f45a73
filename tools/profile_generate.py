#!/usr/bin/env python
"""cProfile of the main thread of DatasetGenerator.generate with the writer (bench scene, png_on_gpu): where the
host time per frame goes once PNG encoding is off the CPU.  Usage: python tools/profile_generate.py [--frames 100]"""
import argparse
import colorsys
import cProfile
import os
import pstats
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--host-png", action="store_true")
    ap.add_argument("--no-writer", action="store_true")
    args = ap.parse_args()
    import torch
    from pegasus_b200 import BOPDatasetWriter, Camera, ComposedScene, DatasetGenerator, ObjectMeta, synth
    dev = torch.device("cuda", 0)
    W, H = 1920, 1080
    env = synth.make_env(2_000_000, seed=1000)
    objs = {i + 1: synth.make_object(200_000, seed=2000 + i) for i in range(5)}
    colors = np.asarray([colorsys.hls_to_rgb(i / 5, 0.6, 0.7)[::-1] for i in range(5)], dtype=np.float32)
    scene = ComposedScene(env, objs, colors, device=dev, sh_mode="canonical")
    poses = synth.static_poses(5, seed=4000)
    scene.set_poses(poses)
    cams_h = synth.orbit_cameras(args.frames, W, H, seed=3000)
    cams = [Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], W, H, device=dev) for c in cams_h]
    metas = [ObjectMeta.from_points(k, objs[k]["xyz"]) for k in sorted(objs)]
    gen = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=args.threads, png_on_gpu=not args.host_png)
    gen.calibrate(cams[::16], margin=1.25)
    out = "/dev/shm/pg_prof"
    shutil.rmtree(out, ignore_errors=True)
    fx = 0.5 * W / np.tan(0.5 * np.deg2rad(72.28))
    wr = None if args.no_writer else BOPDatasetWriter("p", out, fx, fx, W, H, W, H, scene_id=0)
    gen.generate(cams[:8], poses=poses, writer=wr, metas=metas)
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    gen.generate(cams, poses=poses, writer=wr, metas=metas)
    pr.disable()
    dt = time.perf_counter() - t
    if wr is not None:
        wr.close()
    print(f"{args.frames / dt:.1f} frames/s, {1e3 * dt / args.frames:.3f} ms per frame (main thread wall)")
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
    shutil.rmtree(out, ignore_errors=True)


if __name__ == "__main__":
    main()
