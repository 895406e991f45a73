#!/usr/bin/env python
"""Time pg_png_encode on one frame of the bench workload (configs[1]): CUDA events around N back-to-back calls on
resident products, stream sizes, and — for scale — the host encoder on the same frame.
Usage: python tools/png_time.py [--n 50]"""
import argparse
import colorsys
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    import torch
    from pegasus_b200 import Camera, ComposedScene, DatasetGenerator, synth
    dev = torch.device("cuda", 0)
    W, H = args.width, args.height
    env = synth.make_env(2_000_000, seed=1000)
    objs = {i + 1: synth.make_object(200_000, seed=2000 + i) for i in range(5)}
    colors = np.asarray([colorsys.hls_to_rgb(i / 5, 0.6, 0.7)[::-1] for i in range(5)], dtype=np.float32)
    scene = ComposedScene(env, objs, colors, device=dev, sh_mode="rotate")
    scene.set_poses(synth.static_poses(5, seed=4000))
    cams_h = synth.orbit_cameras(100, W, H, seed=3000)
    cams = [Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], W, H, device=dev) for c in cams_h]
    gen = DatasetGenerator(scene, W, H, frames_in_flight=1, png_on_gpu=True)
    gen.calibrate(cams[::16])
    enc, st = gen.png_enc[0], torch.cuda.current_stream(dev)
    scene.render(cams[3], gen.bg, masks=True, out=gen.outs[0], sync_check=True, pair_capacity=gen.pair_capacity)
    gen._pack_slot(0, st)
    for _ in range(3):
        enc.encode(st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(args.n):
        enc.encode(st)
    b.record(st)
    b.synchronize()
    ms = a.elapsed_time(b) / args.n
    streams = enc.streams(enc.arena.cpu(), enc.result.cpu())
    sizes = {k: len(v) for k, v in streams.items()}
    raw = W * H * (3 + 2 + 3 + 2 * 5)
    # the host encoder on the same products
    import cv2
    rgb = gen.packs[0]["rgb"].cpu().numpy()
    d16 = gen.packs[0]["depth"].cpu().numpy().view(np.uint16)
    sem = gen.outs[0]["sem_seg"].cpu().numpy()
    masks = [gen.outs[0][n][k].cpu().numpy() * 255 for n in ("silhouette", "visible") for k in range(5)]
    t = time.perf_counter()
    host_bytes = 0
    for img in [rgb[:, :, ::-1].copy(), d16, sem[:, :, ::-1].copy()] + masks:
        ok, buf = cv2.imencode(".png", img)
        host_bytes += len(buf)
    host_ms = (time.perf_counter() - t) * 1e3
    print(json.dumps({"pg_png_encode_ms_per_frame": ms, "images": len(sizes), "stream_bytes": sizes,
                      "total_stream_bytes": int(sum(sizes.values())), "raw_bytes": raw,
                      "source_gb_per_s": raw / ms / 1e6, "d2h_bytes_per_frame": gen.d2h_bytes_per_frame,
                      "host_opencv_ms_per_frame_one_core": host_ms, "host_opencv_bytes": host_bytes}))


if __name__ == "__main__":
    main()
