import sys, colorsys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pegasus_b200 import Camera, ComposedScene, synth
dev = torch.device('cuda', 0)
W, H = 1920, 1080
env = synth.make_env(2_000_000, seed=1000)
objs = {i + 1: synth.make_object(200_000, seed=2000 + i) for i in range(5)}
cams_h = synth.orbit_cameras(100, W, H, seed=3000)
colors = np.asarray([colorsys.hls_to_rgb(i / 5, 0.6, 0.7)[::-1] for i in range(5)], dtype=np.float32)
scene = ComposedScene(env, objs, colors, device=dev, sh_mode="rotate")
scene.set_poses(synth.static_poses(5, seed=4000))
bg = torch.zeros(3, device=dev)
for vi in (5, 40, 77):
    c = cams_h[vi]
    cam = Camera(c["R"], c["T"], c["FoVx"], c["FoVy"], W, H, device=dev)
    a = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.render(cam, bg, masks=True, numerics="exact").items()}
    b = scene.render(cam, bg, masks=True, numerics="fast")
    dc = (a["color"] - b["color"]).abs()
    dd = ((a["depth"] - b["depth"]).abs() / a["depth"].abs().clamp_min(1e-6))
    print("view", vi, "rgb max", float(dc.max()), "n>1e-3", int((dc > 1e-3).sum()), "n>1e-4", int((dc > 1e-4).sum()), "n>1e-5", int((dc>1e-5).sum()),
          "| depth rel max", float(dd.max()), "n>1e-4", int((dd > 1e-4).sum()),
          "| T max", float((a["final_T"] - b["final_T"]).abs().max()),
          "| vis diff", int((a["visible"] != b["visible"]).sum()), "sil diff", int((a["silhouette"] != b["silhouette"]).sum()),
          "sem diff", int((a["sem_seg"] != b["sem_seg"]).sum()), "seg max", float((a["seg_color"] - b["seg_color"]).abs().max()))
