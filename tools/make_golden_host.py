"""Generate tests/golden/reference_host_rows.json (+ ref_cloud.ply) by RUNNING the reference's own
Python for the rows either side of the renderer.

Runs only in the build container (needs /root/reference); nothing here runs on the GPU box.
Same technique as tools/make_golden.py: the reference's functions are pulled out of their files'
ASTs and exec'd against small stubs for the packages this image lacks (plyfile, open3d), so the code
that runs is the reference's own text.

What gets pinned (file:line in /root/reference):
  camera path   src/gs/pegasus_setup.py:85-143     create_camera_trajectory, modes random / random+zoom
  PLY layout    src/gs/gaussian_model.py:193-288   construct_list_of_attributes, save_ply, load_ply
                (stub plyfile: PlyElement.describe / PlyData.write / PlyData.read over numpy
                structured arrays, binary little endian — the PLY standard, which is all plyfile does here)
  camera.json   src/tools/pegasus_working.py:348-370  write_camera_json
  scene_camera  src/tools/pegasus_working.py:440-455  add_scene_camera_json
  scene_gt      src/tools/pegasus_working.py:457-566  add_scene_gt_json (stub open3d mesh with fixed
                box points / centres; cv2.convertPointsFromHomogeneous is the real one)
  image packing pegasus.py:345-355                  (rgb * 255).astype(uint8), (depth * 1000).astype(uint16)
"""
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (patches torch's device= arguments for CPU, provides extract())

REF, GSP, REPO = mg.REF, mg.GSP, mg.REPO
_orig_tensor = torch.tensor
torch.tensor = lambda *a, **k: _orig_tensor(*a, **mg._strip(k))  # load_ply says device="cuda"


# ---- plyfile stand-in: binary-little-endian PLY over numpy structured arrays -------------------
class _Prop:
    def __init__(self, name):
        self.name = name


class PlyElement:
    def __init__(self, data, name):
        self.data, self.name = data, name
        self.properties = [_Prop(n) for n in data.dtype.names]

    @staticmethod
    def describe(data, name):
        return PlyElement(data, name)

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def write(self, path):
        el = self.elements[0]
        names = {"f4": "float", "f8": "double", "u1": "uchar", "i4": "int"}
        head = ["ply", "format binary_little_endian 1.0", f"element {el.name} {len(el.data)}"]
        for n in el.data.dtype.names:
            head.append(f"property {names[el.data.dtype[n].str[1:]]} {n}")
        head.append("end_header")
        with open(path, "wb") as f:
            f.write(("\n".join(head) + "\n").encode("ascii"))
            f.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        rev = {"float": "<f4", "double": "<f8", "uchar": "u1", "int": "<i4"}
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply"
            props, n, name = [], 0, "vertex"
            while True:
                tok = f.readline().decode().split()
                if tok[0] == "element":
                    name, n = tok[1], int(tok[2])
                elif tok[0] == "property":
                    props.append((tok[2], rev[tok[1]]))
                elif tok[0] == "end_header":
                    break
            data = np.fromfile(f, dtype=np.dtype(props), count=n)
        return PlyData([PlyElement(data, name)])


def main():
    from scipy.spatial.transform import Rotation
    from typing import Literal
    G = {}
    rng = np.random.default_rng(77)

    # ---------------- camera path ----------------
    sys.path.insert(0, GSP)
    sys.path.insert(0, REF)
    from src.utility.pose_interpolation import interpolate_pose
    from utils.graphics_utils import focal2fov, fov2focal
    cl = mg.extract(os.path.join(GSP, "scene/colmap_loader.py"), ["qvec2rotmat"], dict(np=np))

    class CamRec:
        def __init__(self, colmap_id, R, T, FoVx, FoVy, image, gt_alpha_mask, image_name, uid, data_device):
            self.R, self.T, self.FoVx, self.FoVy = np.array(R), np.array(T), float(FoVx), float(FoVy)
            self.shape = tuple(image.shape)

    ps = mg.extract(os.path.join(REF, "src/gs/pegasus_setup.py"), ["create_camera_trajectory"],
                    dict(np=np, torch=torch, Literal=Literal, qvec2rotmat=cl["qvec2rotmat"],
                         interpolate_pose=interpolate_pose, focal2fov=focal2fov, Camera=CamRec),
                    cls="PegasusSetup")

    class Ext:
        def __init__(self, q, t):
            self.qvec, self.tvec = q, t

    class Intr:
        width, height = 1600, 1200
        params = [1234.5, 1250.25, 800.0, 600.0]

    ext = {}
    for i, key in enumerate([11, 3, 7, 19, 5, 2, 13, 17]):
        q = Rotation.from_rotvec(rng.normal(size=3) * 0.4).as_quat()  # xyzw
        ext[key] = Ext(np.array([q[3], q[0], q[1], q[2]]), rng.normal(size=3) * 0.3 + np.array([0, 0, 1.5]))
    G["cam_ext_keys"] = list(ext.keys())
    G["cam_ext_qvec"] = [ext[k].qvec.tolist() for k in ext]
    G["cam_ext_tvec"] = [ext[k].tvec.tolist() for k in ext]
    G["cam_intr"] = dict(width=Intr.width, height=Intr.height, fx=Intr.params[0], fy=Intr.params[1])

    class SetupStub:
        cam_extr = ext
        cam_intr = {1: Intr}
        camera_data = [{"fx": Intr.params[0]}]
        render_width, render_height = 640, 480

    SetupStub.create_camera_trajectory = ps["create_camera_trajectory"]
    for mode in ("random", "random+zoom", "sequence"):
        np.random.seed(1234)
        cams = SetupStub().create_camera_trajectory(num_cameras=3, num_interpolation_steps=4, mode=mode)
        G[f"campath_{mode}"] = dict(R=[c.R.tolist() for c in cams], T=[c.T.tolist() for c in cams],
                                    FoVx=cams[0].FoVx, FoVy=cams[0].FoVy, image_shape=list(cams[0].shape))

    # ---------------- PLY ----------------
    gm = mg.extract(os.path.join(REF, "src/gs/gaussian_model.py"),
                    ["construct_list_of_attributes", "save_ply", "load_ply"],
                    dict(np=np, torch=torch, nn=torch.nn, os=os, mkdir_p=lambda p: os.makedirs(p, exist_ok=True),
                         PlyData=PlyData, PlyElement=PlyElement), cls="GaussianModelBase")

    class GM:
        max_sh_degree = 3

    for k, f in gm.items():
        setattr(GM, k, f)
    g = GM()
    n = 7
    g._xyz = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32))
    g._features_dc = torch.from_numpy(rng.normal(size=(n, 1, 3)).astype(np.float32))
    g._features_rest = torch.from_numpy(rng.normal(size=(n, 15, 3)).astype(np.float32))
    g._opacity = torch.from_numpy(rng.normal(size=(n, 1)).astype(np.float32))
    g._scaling = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32))
    g._rotation = torch.from_numpy(rng.normal(size=(n, 4)).astype(np.float32))
    ply_path = os.path.join(REPO, "tests/golden/ref_cloud.ply")
    g.save_ply(ply_path)
    h = GM()
    h.load_ply(ply_path)
    G["ply_attribute_names"] = g.construct_list_of_attributes()
    for k in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"):
        assert torch.equal(getattr(g, k), getattr(h, k).detach()), k  # the reference's own round trip
        G["ply" + k] = getattr(h, k).detach().numpy().tolist()

    # ---------------- BOP writer ----------------
    import cv2

    class FakeBox:
        def __init__(self, pts, ctr):
            self._p, self._c = pts, ctr

        def get_box_points(self):
            return self._p

        def get_center(self):
            return self._c

    class FakeMesh:
        def __init__(self, seed):
            r = np.random.default_rng(seed)
            self._p = r.normal(size=(8, 3)) * 0.05
            self._bc = r.normal(size=3) * 0.01
            self._mc = r.normal(size=3) * 0.01

        def get_minimal_oriented_bounding_box(self, robust=True):
            return FakeBox(self._p, self._bc)

        def get_center(self):
            return self._mc

    o3d = types.SimpleNamespace(io=types.SimpleNamespace(read_triangle_mesh=lambda path: FakeMesh(int(path))))
    bw = mg.extract(os.path.join(REF, "src/tools/pegasus_working.py"),
                    ["write_camera_json", "add_scene_camera_json", "add_scene_gt_json"],
                    dict(np=np, json=json, Path=Path, o3d=o3d, cv2=cv2, focal2fov=focal2fov, fov2focal=fov2focal,
                         plt=None), cls="PegasusBOPDatasetWriter")

    class W:
        pass

    for k, f in bw.items():
        setattr(W, k, f)
    w = W()
    w.camera_intr = {1: Intr}
    w.render_width, w.render_height = 640, 480
    w.scene_camera_json, w.scene_gt_json = {}, {}
    with tempfile.TemporaryDirectory() as td:
        w.dataset_path = Path(td)
        w.write_camera_json(file_name="camera.json")
        G["bop_camera_json_text"] = open(os.path.join(td, "camera.json")).read()
    w.add_scene_camera_json(frame_id=4)
    G["bop_scene_camera"] = {str(k): v for k, v in w.scene_camera_json.items()}

    class Obj:
        def __init__(self, seed, cls_id):
            r = np.random.default_rng(seed)
            self.meta_info = types.SimpleNamespace(urdf_obj_path=str(seed), ID=cls_id)
            self.R_init = torch.from_numpy(Rotation.from_rotvec(r.normal(size=3)).as_matrix().astype(np.float32))
            self.t_init = torch.from_numpy((r.normal(size=3) * 0.2).astype(np.float32))

    objs = {3: Obj(101, 12), 1: Obj(102, 5), 2: Obj(103, 12)}
    cam = types.SimpleNamespace(R=Rotation.from_rotvec([0.2, -0.4, 0.1]).as_matrix(), T=np.array([0.05, -0.1, 1.3]))
    w.add_scene_gt_json(time_step=4, gs_object_list=objs, cam=cam, rgb_image=None, debug=False)
    G["bop_objects"] = [dict(bullet_id=k, seed=int(o.meta_info.urdf_obj_path), obj_id=o.meta_info.ID,
                             R_init=o.R_init.numpy().tolist(), t_init=o.t_init.numpy().tolist(),
                             o3d_box_points=FakeMesh(int(o.meta_info.urdf_obj_path))._p.tolist(),
                             box_center=FakeMesh(int(o.meta_info.urdf_obj_path))._bc.tolist(),
                             mesh_center=FakeMesh(int(o.meta_info.urdf_obj_path))._mc.tolist())
                        for k, o in objs.items()]
    G["bop_cam"] = dict(R=cam.R.tolist(), T=cam.T.tolist())
    G["bop_scene_gt"] = json.loads(json.dumps(w.scene_gt_json))
    G["bop_scene_gt_text"] = json.dumps(w.scene_gt_json, indent=1)

    # ---------------- image packing (pegasus.py:345-355) ----------------
    rgb = rng.uniform(0, 1, size=(5, 7, 3)).astype(np.float32)
    rgb[0, 0] = [0.0, 1.0, 0.99999994]
    depth = rng.uniform(0, 3.5, size=(5, 7, 1)).astype(np.float32)
    depth[0, 0, 0] = 0.0
    depth[0, 1, 0] = 65.534
    G["pack_rgb_in"] = rgb.tolist()
    G["pack_depth_in"] = depth.tolist()
    G["pack_rgb_u8"] = (np.ascontiguousarray(rgb) * 255).astype("uint8").tolist()
    G["pack_depth_u16"] = (torch.from_numpy(depth).numpy() * 1000).astype(np.uint16)[..., 0].tolist()

    out = os.path.join(REPO, "tests/golden/reference_host_rows.json")
    with open(out, "w") as f:
        json.dump(G, f)
    print("wrote", out, os.path.getsize(out), "bytes;", len(G), "entries")


if __name__ == "__main__":
    main()
