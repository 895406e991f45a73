import sys, os, colorsys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pegasus_b200 import Camera, ComposedScene, synth
from pegasus_b200.scene import export_binning
dev = torch.device('cuda', 0)
W, H = 1920, 1080
env = synth.make_env(2_000_000, seed=1000)
objs = {i + 1: synth.make_object(200_000, seed=2000 + i) for i in range(5)}
cams_h = synth.orbit_cameras(100, W, H, seed=3000)
colors = np.asarray([colorsys.hls_to_rgb(i / 5, 0.6, 0.7)[::-1] for i in range(5)], dtype=np.float32)
scene = ComposedScene(env, objs, colors, device=dev, sh_mode="rotate")
scene.set_poses(synth.static_poses(5, seed=4000))
bg = torch.zeros(3, device=dev)
for vi in (5, 40):
    c = cams_h[vi]
    cam = Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], W, H, device=dev)
    out = scene.render(cam, bg, masks=True, sync_check=True)
    keys, plist, ranges = export_binning(dev, scene.P, W, H, out["pair_capacity"], out["num_stored"])
    ln = (ranges[:, 1].astype(np.int64) - ranges[:, 0].astype(np.int64))
    isobj = plist >= scene.n_env
    cs = np.concatenate([[0], np.cumsum(isobj)])
    nobj = cs[ranges[:, 1]] - cs[ranges[:, 0]]
    order = np.argsort(-ln)
    print("view", vi, "tiles", len(ln), "stored", out["num_stored"], "mean len", ln.mean(), "max len", ln.max(), "p99", np.percentile(ln, 99), "p999", np.percentile(ln, 99.9))
    print(" object entries total", int(nobj.sum()), "tiles with objects", int((nobj > 0).sum()), "max obj entries in a tile", int(nobj.max()))
    print(" top tiles (len, nobj):", [(int(ln[t]), int(nobj[t])) for t in order[:12]])
    oo = np.argsort(-nobj)
    print(" top by nobj (len, nobj):", [(int(ln[t]), int(nobj[t])) for t in oo[:12]])
    print(" hist nobj:", np.histogram(nobj[nobj > 0], bins=[1, 100, 500, 1000, 2000, 4000, 8000, 16000, 64000])[0])
