#!/usr/bin/env python
"""BASELINE.json configs[4] on hardware WITH the consumer (SURVEY §8 e; /root/reference/pegasus.py:333-365 writes
every frame's PNGs + BOP JSON): a bounded slice of the dataset sweep — `pegasus_b200.sweep.plan_scenes` ->
per scene a `ComposedScene` + `DatasetGenerator` -> `BOPDatasetWriter` into a tmpfs directory — reporting, separately,

  e2e_no_writer     frames/s of the generator loop alone (products land in pinned host memory: bench.py's `e2e`),
  e2e_with_writer   frames/s with PNG files and the JSON files, per number of writer threads, for both encoders: the host
                    one (OpenCV / libpng, one frame per writer thread) and pg_png_encode (streams made on the GPU, the
                    host only frames them) -> the core count at which the host stops being the limiter,
  sweep             the whole slice with the best thread count, scene switches (cloud synthesis excluded, scene build,
                    calibration) inside the clock.

One JSON line on stdout.  Under torchrun every rank renders its `sweep.shard_views` part and rank 0 reports the
aggregate (max wall time over ranks).  Synthetic clouds of the configured sizes (no assets offline).

Usage: python tools/c5_sweep.py [--total-views 400] [--views-per-scene 100] [--threads 2,4,8,16] [--out /dev/shm/pg_c5]
"""
import argparse
import colorsys
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dir_bytes(path):
    tot = 0
    for r, _, fs in os.walk(path):
        for f in fs:
            tot += os.path.getsize(os.path.join(r, f))
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--objects", type=int, default=30)
    ap.add_argument("--env-n", type=int, default=2_000_000)
    ap.add_argument("--obj-n", type=int, default=200_000)
    ap.add_argument("--total-views", type=int, default=400)
    ap.add_argument("--views-per-scene", type=int, default=100)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--threads", default="2,4,8,16,32")
    ap.add_argument("--probe-views", type=int, default=100, help="views of scene 0 used for the writer-thread probe")
    ap.add_argument("--out", default="/dev/shm/pg_c5")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    import torch
    from pegasus_b200 import (BOPDatasetWriter, Camera, ComposedScene, DatasetGenerator, ObjectMeta, dist as pgd,
                              sweep, synth)
    rank, world, local = pgd.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    W, H = args.width, args.height
    cores = os.cpu_count()
    scenes = sweep.plan_scenes(args.envs, args.objects, args.total_views, args.views_per_scene, seed=args.seed)
    items = sweep.shard_views(scenes, rank, world)
    out_root = os.path.join(args.out, f"rank{rank}")
    shutil.rmtree(out_root, ignore_errors=True)
    os.makedirs(out_root, exist_ok=True)

    env_cache, obj_cache = {}, {}

    def env_cloud(i):
        if i not in env_cache:
            env_cache.clear()  # one environment cloud on the host at a time (0.5 GB each)
            env_cache[i] = synth.make_env(args.env_n, seed=1000 + i)
        return env_cache[i]

    def obj_cloud(i):
        if i not in obj_cache:
            obj_cache[i] = synth.make_object(args.obj_n, seed=2000 + i)
        return obj_cache[i]

    def build(spec):
        """Scene switch: ComposedScene (activations, canonical arrays) + cameras + static poses + metas."""
        K = len(spec.objects)
        colors = np.asarray([colorsys.hls_to_rgb(i / max(K, 1), 0.6, 0.7)[::-1] for i in range(max(K, 1))], dtype=np.float32)
        objs = {k + 1: obj_cloud(o) for k, o in enumerate(spec.objects)}
        scene = ComposedScene(env_cloud(spec.env), objs, colors, device=dev, sh_mode="canonical")
        cams_h = synth.orbit_cameras(spec.n_views, W, H, seed=spec.seed)
        cams = [Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], W, H, device=dev) for c in cams_h]
        poses = synth.static_poses(K, seed=spec.seed + 1)
        metas = [ObjectMeta.from_points(o, objs[k + 1]["xyz"]) for k, o in enumerate(spec.objects)]
        return scene, cams, poses, metas

    def writer_for(spec, tag):
        fx = 0.5 * W / np.tan(0.5 * np.deg2rad(72.28))
        return BOPDatasetWriter(tag, out_root, fx, fx, W, H, W, H, scene_id=spec.scene_id)

    res = {"config": {"workload": f"configs[4] slice: {len(scenes)} scenes ({args.envs} envs x {args.env_n}, "
                                  f"{args.objects} object clouds x {args.obj_n}, 3-6 objects per scene), "
                                  f"{args.total_views} views {W}x{H}", "world": world, "host_cores": cores,
                      "out": args.out, "png": "OpenCV imwrite defaults, 13 PNGs per 5-object frame"}}

    # ---- probe on this rank's first scene: generator alone, then with the writer per thread count
    spec0 = items[0].scene
    for o in spec0.objects:  # cloud synthesis is not part of any clock
        obj_cloud(o)
    env_cloud(spec0.env)
    scene, cams, poses, metas = build(spec0)
    n_probe = min(args.probe_views, len(cams))
    gen = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=4)
    scene.set_poses(poses)
    gen.calibrate(cams[::max(1, len(cams) // 16)], margin=1.15)  # a frame that still overflows is re-rendered (regrow)
    gen.generate(cams[:8], poses=poses)  # warm-up
    torch.cuda.synchronize()
    pgd.barrier()
    t = time.perf_counter()
    gen.generate(cams[:n_probe], poses=poses)
    torch.cuda.synchronize()
    no_writer = n_probe / (pgd.max_over_ranks((time.perf_counter() - t) * 1e3, device=dev) / 1e3)
    probe = []
    for on_gpu, nt in [(g, int(x)) for g in (False, True) for x in args.threads.split(",")]:
        g2 = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=nt, png_on_gpu=on_gpu)
        g2.pair_capacity = gen.pair_capacity
        if on_gpu:
            g2._calibrate_png(cams[::max(1, len(cams) // 16)], None)  # untimed set-up, like the pair capacity
            torch.cuda.synchronize()
        tag = f"probe_t{nt}"
        wr = writer_for(spec0, tag)
        pgd.barrier()  # all ranks run the same configuration at the same time: they share the host cores
        t = time.perf_counter()
        g2.generate(cams[:n_probe], poses=poses, writer=wr, metas=metas)
        wr.close()  # joins the writer threads, flushes scene_camera.json / scene_gt.json
        dt = pgd.max_over_ranks((time.perf_counter() - t) * 1e3, device=dev) / 1e3
        nbytes = dir_bytes(os.path.join(out_root, tag))
        probe.append({"png": "gpu" if on_gpu else "host", "writer_threads_per_rank": nt, "frames_per_s": world * n_probe / dt,
                      "file_mb_per_frame": nbytes / n_probe / 1e6, "d2h_mb_per_frame": g2.d2h_bytes_per_frame / 1e6,
                      "png_fallbacks": g2.png_fallbacks})
        shutil.rmtree(os.path.join(out_root, tag), ignore_errors=True)
        del g2
    best = max(probe, key=lambda p: p["frames_per_s"])
    del gen, scene
    torch.cuda.empty_cache()

    # ---- the slice itself with the best thread count: scene switches inside the clock
    pgd.barrier()
    t_all = time.perf_counter()
    build_s, synth_s, frames, regrown = 0.0, 0.0, 0, 0
    for it in items:
        ts = time.perf_counter()
        for o in it.scene.objects:
            obj_cloud(o)
        env_cloud(it.scene.env)
        synth_s += time.perf_counter() - ts  # stands for reading PLY files: not part of the clock
        tb = time.perf_counter()
        scene, cams, poses, metas = build(it.scene)
        g = DatasetGenerator(scene, W, H, frames_in_flight=3, writer_threads=best["writer_threads_per_rank"],
                             png_on_gpu=best["png"] == "gpu")
        fr = list(range(it.first_view, it.first_view + it.n_views))
        scene.set_poses(poses)
        g.calibrate([cams[f] for f in fr[::max(1, len(fr) // 16)]], margin=1.15)
        torch.cuda.synchronize()
        build_s += time.perf_counter() - tb
        wr = writer_for(it.scene, "sweep")
        st = g.generate(cams, poses=poses, writer=wr, metas=metas, frames=fr)
        wr.close()  # joins the writer threads, flushes scene_camera.json / scene_gt.json
        frames += st["frames"]
        regrown += st["regrown"]
        shutil.rmtree(os.path.join(out_root, "sweep"), ignore_errors=True)  # bounded tmpfs use
        del g, scene
        torch.cuda.empty_cache()
    wall = time.perf_counter() - t_all - synth_s
    wall_max = pgd.max_over_ranks(wall * 1e3, device=dev) / 1e3
    total = pgd.sum_over_ranks(frames, device=dev)
    if rank == 0:
        res.update({
            "e2e_no_writer": {"value": no_writer * world, "unit": "frames/s", "views": n_probe,
                              "what": "DatasetGenerator alone, every rank on its first scene at the same time"},
            "e2e_with_writer": probe, "best": {"png": best["png"], "writer_threads_per_rank": best["writer_threads_per_rank"]},
            "sweep": {"value": total / wall_max, "unit": "frames/s", "frames": int(total), "wall_s": wall_max,
                      "scene_builds_rank0": len(items), "scene_build_s_rank0": build_s, "capacity_regrowths_rank0": regrown,
                      "what": "whole slice incl. scene builds + calibration, cloud synthesis excluded"},
        })
        print(json.dumps(res))
    shutil.rmtree(out_root, ignore_errors=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
