"""Gaussian-splatting PLY files -> the raw cloud dicts ComposedScene takes.

Mirrors GaussianModel.load_ply / save_ply of /root/reference/src/gs/gaussian_model.py:207-288
without plyfile: the header is parsed here and the vertex block is read with one numpy structured
view.  File layout written by the reference (``construct_list_of_attributes``, :193-205): per vertex
62 little-endian float32 ``x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3``.
``f_rest`` is CHANNEL-major in the file — (N, 3, 15) flattened — and becomes (N, 15, 3) in memory
(:256-263 reshape, :281-283 transpose); ``f_dc`` likewise (N, 3, 1) -> (N, 1, 3).  Values are the
raw pre-activation parameters; ComposedScene applies sigmoid / exp / normalize once at load.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

CLOUD_KEYS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def attribute_names(max_sh_degree: int = 3) -> List[str]:
    """Property order the reference writes (gaussian_model.py:193-205)."""
    n_rest = 3 * ((max_sh_degree + 1) ** 2 - 1)
    return (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] +
            [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"] + [f"scale_{i}" for i in range(3)] +
            [f"rot_{i}" for i in range(4)])


def _read_header(f) -> Tuple[str, int, List[Tuple[str, str]], int]:
    """-> (format, vertex count, [(name, numpy type)] of the vertex element, header size in bytes)."""
    if f.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt, count, props, in_vertex = None, None, [], False
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header is not terminated")
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                count = int(tok[2])
            elif count is None:
                raise ValueError("elements before 'vertex' are not supported")
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise ValueError("list properties on the vertex element are not supported")
            props.append((tok[2], _PLY_TYPES[tok[1]]))
        elif tok[0] == "end_header":
            break
    if fmt is None or count is None:
        raise ValueError("PLY header lacks format / vertex element")
    return fmt, count, props, f.tell()


def read_vertices(path: str) -> Dict[str, np.ndarray]:
    """All vertex properties of a PLY file as 1-D arrays keyed by property name."""
    with open(path, "rb") as f:
        fmt, n, props, off = _read_header(f)
        if fmt == "ascii":
            table = np.loadtxt(f, dtype=np.float64, max_rows=n, ndmin=2)
            if table.shape != (n, len(props)):
                raise ValueError(f"ascii PLY: expected {n}x{len(props)} values, got {table.shape}")
            return {name: table[:, i] for i, (name, _) in enumerate(props)}
        order = {"binary_little_endian": "<", "binary_big_endian": ">"}.get(fmt)
        if order is None:
            raise ValueError(f"unsupported PLY format {fmt!r}")
        dt = np.dtype([(name, order + t) for name, t in props])
        if os.path.getsize(path) - off < n * dt.itemsize:
            raise ValueError("PLY vertex block is truncated")
        rec = np.fromfile(f, dtype=dt, count=n)
    return {name: rec[name] for name, _ in props}


def load_ply(path: str, max_sh_degree: int = 3) -> Dict[str, np.ndarray]:
    """Raw cloud dict (float32): xyz (N,3), features_dc (N,1,3), features_rest (N,(d+1)^2-1,3),
    opacity (N,1), scaling (N,3), rotation (N,4) — the tensors GaussianModel.load_ply builds."""
    v = read_vertices(path)

    def numbered(prefix):
        names = sorted((k for k in v if k.startswith(prefix)), key=lambda s: int(s.split("_")[-1]))
        return names

    n_coef = (max_sh_degree + 1) ** 2 - 1
    rest_names = numbered("f_rest_")
    if len(rest_names) != 3 * n_coef:   # the reference asserts the same (gaussian_model.py:259)
        raise ValueError(f"{path}: {len(rest_names)} f_rest properties, SH degree {max_sh_degree} needs {3 * n_coef}")
    f32 = lambda names: np.stack([np.asarray(v[k], dtype=np.float32) for k in names], axis=1)
    xyz = f32(["x", "y", "z"])
    n = xyz.shape[0]
    dc = f32(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(n, 3, 1).transpose(0, 2, 1)
    rest = f32(rest_names).reshape(n, 3, n_coef).transpose(0, 2, 1)
    return dict(xyz=xyz, features_dc=np.ascontiguousarray(dc), features_rest=np.ascontiguousarray(rest),
                opacity=f32(["opacity"]), scaling=f32(numbered("scale_")), rotation=f32(numbered("rot_")))


def save_ply(path: str, cloud: Dict[str, np.ndarray]) -> None:
    """Write a cloud dict in the reference's layout (binary little endian, normals zero)."""
    xyz = np.asarray(cloud["xyz"], dtype=np.float32)
    n = xyz.shape[0]
    dc = np.asarray(cloud["features_dc"], dtype=np.float32).reshape(n, 1, 3).transpose(0, 2, 1).reshape(n, 3)
    rest = np.asarray(cloud["features_rest"], dtype=np.float32)
    n_coef = rest.shape[1] if rest.ndim == 3 else rest.size // max(3 * n, 1)  # (n, coef, 3)
    rest = rest.reshape(n, n_coef, 3).transpose(0, 2, 1).reshape(n, 3 * n_coef)
    cols = np.concatenate([xyz, np.zeros_like(xyz), dc, rest,
                           np.asarray(cloud["opacity"], dtype=np.float32).reshape(n, 1),
                           np.asarray(cloud["scaling"], dtype=np.float32).reshape(n, 3),
                           np.asarray(cloud["rotation"], dtype=np.float32).reshape(n, 4)], axis=1)
    deg = int(round((n_coef + 1) ** 0.5)) - 1
    names = attribute_names(deg)
    assert cols.shape[1] == len(names)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + \
             [f"property float {k}" for k in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(cols, dtype="<f4").tobytes())
