"""CPU tests of the host side of the on-GPU PNG encoder (pegasus_b200/png_codec.py): the Huffman tables and block
headers it builds must make the token stream of tests/png_model.py (the kernel's tokenizer, restated) a zlib stream
that stock zlib inflates to the filtered scanlines, and the framed file a PNG that OpenCV decodes to the pixels."""
import zlib

import numpy as np
import pytest

from tests import png_model as pm
from pegasus_b200 import png_codec as pc


def images(kind, rng, H=37, W=53):
    if kind == pc.KIND_RGB8:
        smooth = (np.add.outer(np.arange(H), np.arange(W))[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)
        noise = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        flat = np.zeros((H, W, 3), np.uint8)
        flat[5:20, 7:40] = (10, 200, 30)
        return [smooth, noise, flat]
    if kind == pc.KIND_GRAY16:
        return [rng.integers(0, 65536, (H, W)).astype(np.uint16), np.full((H, W), 1234, np.uint16),
                (np.add.outer(np.arange(H), np.arange(W)) * 37 % 65536).astype(np.uint16)]
    m = np.zeros((H, W), np.uint8)
    m[10:30, 3:50] = 1
    return [m, np.zeros((H, W), np.uint8), np.ones((H, W), np.uint8), rng.integers(0, 2, (H, W)).astype(np.uint8)]


@pytest.mark.parametrize("kind", [pc.KIND_RGB8, pc.KIND_GRAY16, pc.KIND_MASK8])
def test_model_stream_inflates_and_png_decodes(kind):
    rng = np.random.default_rng(kind)
    imgs = images(kind, rng)
    table = pc.build_table(pm.token_hist(kind, imgs))
    for img in imgs:
        z = pm.encode(kind, img, table)
        assert zlib.decompress(z) == pm.scanlines(kind, img).tobytes()
        H, W = img.shape[:2]
        assert len(z) <= pc.worst_case_bytes(kind, W, H)
        got = pm.decode_png(kind, pc.png_file(kind, W, H, z))
        want = img if kind != pc.KIND_MASK8 else (img != 0).astype(np.uint8) * 255
        assert got.dtype == want.dtype and np.array_equal(got, want)


def test_table_from_unrelated_statistics_still_decodes():
    """The table is static per scene: an image whose statistics differ from the sample must still round-trip."""
    rng = np.random.default_rng(7)
    table = pc.build_table(pm.token_hist(pc.KIND_MASK8, [np.zeros((8, 300), np.uint8)]))
    img = rng.integers(0, 256, (9, 31, 3), dtype=np.uint8)
    assert zlib.decompress(pm.encode(pc.KIND_RGB8, img, table)) == pm.scanlines(pc.KIND_RGB8, img).tobytes()


def test_length_limit_and_completeness():
    """Skewed histograms need the depth limit; the code must stay complete (zlib rejects incomplete sets)."""
    h = np.zeros(pc.N_LITLEN, np.int64)
    h[:40] = [2 ** min(i, 50) for i in range(40)]
    lens = pc.huffman_lengths(np.maximum(h, 1).tolist(), pc.MAX_BITS)
    assert max(lens) <= pc.MAX_BITS and min(lens) >= 1
    assert sum(2.0 ** -n for n in lens) == 1.0
    table = pc.build_table(h)
    img = np.arange(6 * 200 * 3, dtype=np.uint32).reshape(6, 200, 3).astype(np.uint8)
    assert zlib.decompress(pm.encode(pc.KIND_RGB8, img, table)) == pm.scanlines(pc.KIND_RGB8, img).tobytes()


def test_long_runs_are_cut_at_258():
    f = np.zeros(1 + 258 * 3 + 2, np.uint8)
    toks = pm.tokenize_row(f)
    assert toks == [("lit", 0), ("len", 258), ("len", 258), ("len", 258), ("lit", 0), ("lit", 0)]
    assert [pc.length_symbol(v)[0] for v in (3, 10, 11, 12, 257, 258)] == [257, 264, 265, 265, 284, 285]


def test_random_run_structures_round_trip():
    """Rows made of random runs (lengths around the 3 / 258 / 259 boundaries) over every kind and odd widths."""
    rng = np.random.default_rng(123)
    lens = np.array([1, 2, 3, 4, 5, 17, 257, 258, 259, 260, 516, 517, 600])
    for kind, W in ((pc.KIND_MASK8, 1031), (pc.KIND_RGB8, 347), (pc.KIND_GRAY16, 521)):
        rows = []
        for _ in range(6):
            vals, out = rng.integers(0, 4, 64), []
            for v in vals:
                out += [int(v)] * int(rng.choice(lens))
            rows.append(out[:W] + [0] * max(0, W - len(out)))
        a = np.asarray(rows)
        if kind == pc.KIND_RGB8:
            img = np.repeat(a[..., None], 3, axis=2).astype(np.uint8) * 60
        elif kind == pc.KIND_GRAY16:
            img = (a * 21845).astype(np.uint16)
        else:
            img = (a > 1).astype(np.uint8)
        table = pc.build_table(pm.token_hist(kind, [img]))
        z = pm.encode(kind, img, table)
        assert zlib.decompress(z) == pm.scanlines(kind, img).tobytes()
        got = pm.decode_png(kind, pc.png_file(kind, W, img.shape[0], z))
        want = img if kind != pc.KIND_MASK8 else (img != 0).astype(np.uint8) * 255
        assert np.array_equal(got, want)
