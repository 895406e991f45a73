"""Conservativeness of the tile-row culling used by the binning stage (pegasus_b200/csrc/tile_cull.h).

The header is compiled for the host (identical IEEE operations) and, for random and adversarial
Gaussians, every pixel whose float32 `power` (the compositing kernel's exact operation order) reaches
the Gaussian's cut must lie in a tile the row test keeps.  Dropping a kept tile would change the image."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "cull_host.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "native", "_build")
OUT = os.path.join(OUT_DIR, "libcull_host.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", OUT, SRC], check=True)
    L = C.CDLL(OUT)
    L.cull_runs.restype = C.c_int
    L.cull_runs.argtypes = [C.c_float] * 6 + [C.c_int] * 6 + [np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    return L


def power_f32(gx, gy, qa, qb, qc, px, py):
    """composite.cu's operation order in float32 (px, py integer grids)."""
    f = np.float32
    dx = (f(gx) - px.astype(f)).astype(f)
    dy = (f(gy) - py.astype(f)).astype(f)
    w = (dy * (f(qc) * dy).astype(f)).astype(f)
    sq = (dx.astype(np.float64) * (f(qa) * dx).astype(f).astype(np.float64) + w.astype(np.float64)).astype(f)  # fma
    bxy = ((f(qb) * dx).astype(f) * dy).astype(f)
    return (sq.astype(np.float64) * -0.5 - bxy.astype(np.float64)).astype(f)                                   # fma


def check_one(L, gx, gy, cov, opacity, W, H, rng):
    a, b, c = cov
    det = a * c - b * b
    if not det > 0:
        return 0, 0
    f = np.float32
    qa, qb, qc = f(c / det), f(-b / det), f(a / det)
    mid = 0.5 * (a + c)
    lam = mid + np.sqrt(max(0.1, mid * mid - det))
    radius = int(np.ceil(3.0 * np.sqrt(lam)))
    g = (W + 15) // 16, (H + 15) // 16
    rx0 = min(g[0], max(0, int((gx - radius) / 16)))
    ry0 = min(g[1], max(0, int((gy - radius) / 16)))
    rx1 = min(g[0], max(0, int((gx + radius + 15) / 16)))
    ry1 = min(g[1], max(0, int((gy + radius + 15) / 16)))
    if (rx1 - rx0) * (ry1 - ry0) == 0:
        return 0, 0
    cut = f(-80.0)
    if opacity > 0:
        cut = f(min(-(np.log(f(255.0) * f(opacity)) + 0.01), -1e-6))
    if not cut > -80:
        cut = f(-80.0)
    cut = np.frombuffer(np.uint32((np.frombuffer(f(cut).tobytes(), np.uint32)[0] & ~np.uint32(63)) | np.uint32(rng.integers(0, 6))).tobytes(), f)[0]
    runs = np.zeros(2 * (ry1 - ry0), np.int32)
    kept = L.cull_runs(f(gx), f(gy), qa, qb, qc, cut, W, H, rx0, ry0, rx1, ry1, runs)
    # brute force over the pixels of the rectangle
    px = np.arange(rx0 * 16, min(rx1 * 16, W))
    py = np.arange(ry0 * 16, min(ry1 * 16, H))
    PX, PY = np.meshgrid(px, py)
    pw = power_f32(gx, gy, qa, qb, qc, PX, PY)
    contrib = ~(pw > 0) & ~(pw < cut)
    ys, xs = np.nonzero(contrib)
    for y, x in zip(PY[ys, xs], PX[ys, xs]):
        ta, tb = runs[2 * (y // 16 - ry0)], runs[2 * (y // 16 - ry0) + 1]
        assert ta <= x // 16 < tb, (gx, gy, cov, opacity, int(x), int(y), int(ta), int(tb))
    assert (runs[0::2] >= rx0).all() and (runs[1::2] <= rx1).all()
    return kept, (rx1 - rx0) * (ry1 - ry0)


def random_cov(rng, smin, smax, max_ratio):
    s1 = np.exp(rng.uniform(np.log(smin), np.log(smax)))
    s2 = max(s1 / np.exp(rng.uniform(0, np.log(max_ratio))), 0.0)
    th = rng.uniform(0, np.pi)
    c, s = np.cos(th), np.sin(th)
    a = c * c * s1 * s1 + s * s * s2 * s2 + 0.3
    b = c * s * (s1 * s1 - s2 * s2)
    d = s * s * s1 * s1 + c * c * s2 * s2 + 0.3
    return a, b, d


@pytest.mark.parametrize("W,H", [(640, 480), (333, 205), (1920, 1080)])
def test_row_runs_never_drop_a_contributing_pixel(lib, W, H):
    rng = np.random.default_rng(W)
    kept = total = 0
    n = 1500 if W < 1000 else 600
    for i in range(n):
        gx, gy = rng.uniform(-40, W + 40), rng.uniform(-40, H + 40)
        cov = random_cov(rng, 0.2, 60.0, 40.0)
        op = [0.999, 0.9, 0.5, 0.1, 0.02, 0.005, 1 / 255.0 + 1e-4][i % 7]
        k, t = check_one(lib, gx, gy, cov, op, W, H, rng)
        kept += k
        total += t
    assert 0 < kept < 0.75 * total   # the test is not vacuous: a large share of the rectangles is culled


def test_row_runs_adversarial_shapes(lib):
    rng = np.random.default_rng(7)
    W, H = 1920, 1080
    for i in range(400):
        gx, gy = rng.uniform(0, W), rng.uniform(0, H)
        kind = i % 4
        if kind == 0:      # needle: extreme anisotropy at 45 degrees
            cov = random_cov(rng, 100.0, 400.0, 1e3)
        elif kind == 1:    # axis aligned, qb == 0 exactly
            s1, s2 = rng.uniform(0.1, 80), rng.uniform(0.1, 80)
            cov = (s1 * s1 + 0.3, 0.0, s2 * s2 + 0.3)
        elif kind == 2:    # minimum-size splats (dilation only)
            cov = random_cov(rng, 0.01, 0.3, 3.0)
        else:              # centred exactly on pixel / tile boundaries
            gx, gy = float(16 * rng.integers(0, W // 16)), float(16 * rng.integers(0, H // 16)) - 0.5
            cov = random_cov(rng, 1.0, 30.0, 10.0)
        check_one(lib, gx, gy, cov, [0.99, 0.3, 0.01][i % 3], W, H, rng)
