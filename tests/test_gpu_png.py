"""GPU tests of pg_png_encode (csrc/png.cu) through the C ABI: the streams it writes are, bit for bit, those of the
tokenizer restated in tests/png_model.py with the same table; stock zlib inflates them to the Sub-filtered
scanlines; OpenCV decodes the framed files to the source pixels; the histogram it accumulates equals the model's;
a stream that does not fit its capacity is reported, never truncated silently."""
import zlib

import numpy as np
import pytest
import torch

from pegasus_b200 import png_codec as pc
from pegasus_b200.png_gpu import FramePngEncoder, PngOverflow, PngTables
from tests import png_model as pm

pytestmark = pytest.mark.gpu


def make_images(rng, H, W):
    yy, xx = np.mgrid[0:H, 0:W]
    rgb_smooth = np.stack([(xx * 3 + yy) % 256, (xx + yy * 2) % 256, (xx * yy) % 256], -1).astype(np.uint8)
    rgb_noise = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    sem = np.zeros((H, W, 3), np.uint8)
    sem[H // 4:H // 2, W // 5:W // 2] = (10, 200, 30)
    sem[H // 2:, W // 2:] = (250, 0, 7)
    depth = ((np.sin(xx / 17.0) + np.cos(yy / 11.0) + 2.5) * 9000).astype(np.uint16)
    depth[rng.integers(0, H, 50), rng.integers(0, W, 50)] = 65535
    mask = np.zeros((H, W), np.uint8)
    mask[H // 3:2 * H // 3, W // 4:3 * W // 4] = 1
    mask[rng.integers(0, H, 30), rng.integers(0, W, 30)] = 1
    return [("rgb_smooth", pc.KIND_RGB8, "rgb", rgb_smooth), ("rgb_noise", pc.KIND_RGB8, "rgb", rgb_noise),
            ("sem", pc.KIND_RGB8, "sem", sem), ("depth", pc.KIND_GRAY16, "depth", depth),
            ("mask", pc.KIND_MASK8, "mask", mask), ("mask_empty", pc.KIND_MASK8, "mask", np.zeros((H, W), np.uint8)),
            ("mask_full", pc.KIND_MASK8, "mask", np.full((H, W), 7, np.uint8))]


def to_dev(kind, a, dev):
    if kind == pc.KIND_GRAY16:
        return torch.from_numpy(a.view(np.int16).copy()).to(dev)
    return torch.from_numpy(a.copy()).to(dev)


@pytest.mark.parametrize("H,W", [(37, 53), (24, 64), (5, 1030)])
def test_streams_equal_model_bit_for_bit(H, W):
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(H * 1000 + W)
    imgs = make_images(rng, H, W)
    tables = PngTables(dev, ["rgb", "sem", "depth", "mask"])
    enc = FramePngEncoder(tables, W, H, [(n, k, g, to_dev(k, a, dev)) for n, k, g, a in imgs])
    st = torch.cuda.current_stream(dev)
    # 1. the flat (uncalibrated) table already round-trips
    out = enc.encode_unbounded(st)
    for n, k, g, a in imgs:
        assert zlib.decompress(out[n]) == pm.scanlines(k, a).tobytes(), n
    # 2. the histogram the kernel accumulates is the model's
    enc.accumulate_hist(st)
    h = tables.hist.cpu().numpy().astype(np.int64)
    for g in tables.groups:
        want = sum(pm.token_hist(k, [a]) for n, k, gg, a in imgs if gg == g)
        assert np.array_equal(h[tables.index[g]], want), g
    tables.rebuild_from_hist()
    host_tables = tables.dev.cpu().numpy().view(np.uint32)
    # 3. calibrated: bit-identical to the model, decodes to the pixels
    out = enc.encode_unbounded(st)
    for n, k, g, a in imgs:
        want = pm.encode(k, a, host_tables[tables.index[g]])
        assert out[n] == want, n
        got = pm.decode_png(k, pc.png_file(k, W, H, out[n]))
        ref = a if k != pc.KIND_MASK8 else (a != 0).astype(np.uint8) * 255
        assert got.dtype == ref.dtype and np.array_equal(got, ref), n


def test_bounded_capacity_and_overflow_flag():
    dev = torch.device("cuda", 0)
    H, W = 64, 256
    rng = np.random.default_rng(5)
    imgs = make_images(rng, H, W)[:2]
    tables = PngTables(dev, ["rgb"])
    enc = FramePngEncoder(tables, W, H, [(n, k, g, to_dev(k, a, dev)) for n, k, g, a in imgs])
    st = torch.cuda.current_stream(dev)
    enc.accumulate_hist(st)
    tables.rebuild_from_hist()
    sizes = enc.measured_sizes(st)
    enc.set_capacities([s + 16 for s in sizes])
    enc.encode(st)
    streams = enc.streams(enc.arena.cpu(), enc.result.cpu())
    for (n, k, g, a), s in zip(imgs, sizes):
        assert len(streams[n]) == s
        assert zlib.decompress(bytes(streams[n])) == pm.scanlines(k, a).tobytes()
    # too small for the noisy image: flagged, and nothing is written past the capacity
    guard = torch.full((enc.arena_bytes + 4096,), 0x5A, dtype=torch.uint8, device=dev)
    enc.set_capacities([sizes[0] + 16, sizes[1] // 2])
    enc.encode(st)
    with pytest.raises(PngOverflow):
        enc.streams(enc.arena.cpu(), enc.result.cpu())
    del guard


def test_full_hd_frame_round_trip_and_ratio():
    """1080p: noisy RGB / depth and ten mask planes in one call; sizes in the range the host encoder reaches."""
    dev = torch.device("cuda", 0)
    H, W = 1080, 1920
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:H, 0:W]
    base = (np.sin(xx / 40.0) * np.cos(yy / 33.0) * 0.5 + 0.5)
    rgb = np.clip(base[..., None] * np.array([200, 180, 220]) + rng.normal(0, 5, (H, W, 3)), 0, 255).astype(np.uint8)
    depth = np.clip(base * 4000 + 500 + rng.normal(0, 3, (H, W)), 0, 65535).astype(np.uint16)
    masks = []
    for k in range(10):
        m = np.zeros((H, W), np.uint8)
        m[100 + 60 * k:400 + 60 * k, 200 + 100 * k:700 + 100 * k] = 1
        masks.append(m)
    sem = np.zeros((H, W, 3), np.uint8)
    for k, m in enumerate(masks[:5]):
        sem[m != 0] = (20 * k + 5, 255 - 30 * k, 40 + k)
    imgs = [("rgb", pc.KIND_RGB8, "rgb", rgb), ("depth", pc.KIND_GRAY16, "depth", depth), ("sem", pc.KIND_RGB8, "sem", sem)]
    imgs += [(f"mask{k}", pc.KIND_MASK8, "mask", m) for k, m in enumerate(masks)]
    tables = PngTables(dev, ["rgb", "depth", "sem", "mask"])
    enc = FramePngEncoder(tables, W, H, [(n, k, g, to_dev(k, a, dev)) for n, k, g, a in imgs])
    st = torch.cuda.current_stream(dev)
    enc.accumulate_hist(st)
    tables.rebuild_from_hist()
    sizes = enc.measured_sizes(st)
    enc.set_capacities([int(s * 1.25) + 4096 for s in sizes])
    enc.encode(st)
    streams = enc.streams(enc.arena.cpu(), enc.result.cpu())
    for n, k, g, a in imgs:
        assert zlib.decompress(bytes(streams[n])) == pm.scanlines(k, a).tobytes(), n
    raw = {n: a.nbytes for n, k, g, a in imgs}
    assert len(streams["rgb"]) < 0.8 * raw["rgb"]          # Huffman-coded Sub residuals
    assert len(streams["mask0"]) < 0.01 * raw["mask0"]     # runs collapse into matches
    assert len(streams["sem"]) < 0.01 * raw["sem"]
    got = pm.decode_png(pc.KIND_GRAY16, pc.png_file(pc.KIND_GRAY16, W, H, streams["depth"]))
    assert np.array_equal(got, depth)


def test_more_images_than_one_launch_batch():
    """A frame of 32 objects has 3 + 64 images; the library encodes them in batches of 24 per launch pair."""
    dev = torch.device("cuda", 0)
    H, W = 33, 200
    rng = np.random.default_rng(3)
    imgs = []
    for k in range(53):
        m = np.zeros((H, W), np.uint8)
        m[k % H:, (3 * k) % W:] = 1
        m[rng.integers(0, H, 5), rng.integers(0, W, 5)] ^= 1
        imgs.append((f"m{k}", pc.KIND_MASK8, "mask", m))
    tables = PngTables(dev, ["mask"])
    enc = FramePngEncoder(tables, W, H, [(n, k, g, to_dev(k, a, dev)) for n, k, g, a in imgs])
    st = torch.cuda.current_stream(dev)
    enc.accumulate_hist(st)
    tables.rebuild_from_hist()
    out = enc.encode_unbounded(st)
    for n, k, g, a in imgs:
        assert zlib.decompress(out[n]) == pm.scanlines(k, a).tobytes(), n


def test_strided_source_rows_and_4k_width():
    """Rows may be strided (a view into a larger buffer); a 3840-wide RGB row still fits the shared-memory staging."""
    dev = torch.device("cuda", 0)
    H, W = 6, 3840
    rng = np.random.default_rng(4)
    big = rng.integers(0, 256, (H, W + 64, 3), dtype=np.uint8)
    big[:, 1000:3000] = (9, 9, 200)
    t = torch.from_numpy(big).to(dev)[:, 32:32 + W]          # pitch (W + 64) * 3, offset 96 bytes
    d16 = rng.integers(0, 65536, (H, W + 2)).astype(np.uint16)
    td = torch.from_numpy(d16.view(np.int16).copy()).to(dev)[:, 1:1 + W]   # 2-byte aligned only: the byte path
    tables = PngTables(dev, ["rgb", "depth"])
    enc = FramePngEncoder(tables, W, H, [("rgb", pc.KIND_RGB8, "rgb", t), ("depth", pc.KIND_GRAY16, "depth", td)])
    st = torch.cuda.current_stream(dev)
    enc.accumulate_hist(st)
    tables.rebuild_from_hist()
    out = enc.encode_unbounded(st)
    assert zlib.decompress(out["rgb"]) == pm.scanlines(pc.KIND_RGB8, big[:, 32:32 + W]).tobytes()
    assert zlib.decompress(out["depth"]) == pm.scanlines(pc.KIND_GRAY16, d16[:, 1:1 + W]).tobytes()


def test_run_lengths_around_chunk_and_thread_boundaries():
    """Rows made of random runs with lengths around 3 / 258 / 259 / 516 (match chunking) that start and end anywhere
    relative to the threads' parts of a row: the kernel's streams must equal the restated tokenizer's bit for bit."""
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(321)
    lens = np.array([1, 2, 3, 4, 5, 17, 63, 64, 65, 257, 258, 259, 260, 516, 517, 600, 1500])
    imgs = []
    W, H = 2047, 12
    for kind, name in ((pc.KIND_MASK8, "mask"), (pc.KIND_RGB8, "rgb"), (pc.KIND_GRAY16, "depth")):
        rows = []
        for _ in range(H):
            out = []
            while len(out) < W:
                out += [int(rng.integers(0, 4))] * int(rng.choice(lens))
            rows.append(out[:W])
        a = np.asarray(rows)
        if kind == pc.KIND_RGB8:
            img = np.repeat(a[..., None], 3, axis=2).astype(np.uint8) * 60
        elif kind == pc.KIND_GRAY16:
            img = (a * 21845).astype(np.uint16)
        else:
            img = (a > 1).astype(np.uint8)
        imgs.append((name, kind, name, img))
    tables = PngTables(dev, ["mask", "rgb", "depth"])
    enc = FramePngEncoder(tables, W, H, [(n, k, g, to_dev(k, a, dev)) for n, k, g, a in imgs])
    st = torch.cuda.current_stream(dev)
    enc.accumulate_hist(st)
    h = tables.hist.cpu().numpy().astype(np.int64)
    for n, k, g, a in imgs:
        assert np.array_equal(h[tables.index[g]], pm.token_hist(k, [a])), n
    tables.rebuild_from_hist()
    host_tables = tables.dev.cpu().numpy().view(np.uint32)
    out = enc.encode_unbounded(st)
    for n, k, g, a in imgs:
        assert out[n] == pm.encode(k, a, host_tables[tables.index[g]]), n
        assert zlib.decompress(out[n]) == pm.scanlines(k, a).tobytes(), n
