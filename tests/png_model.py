"""TEST INFRASTRUCTURE — a slow restatement of the tokenizer of csrc/png.cu (pg_png_encode) in plain Python, used to
check pegasus_b200.png_codec's tables / headers on the CPU and the kernel's streams on the GPU.  Never imported by the
product."""
import zlib

import numpy as np

from pegasus_b200 import png_codec as pc


def scanlines(kind, img):
    """[H][1 + bpp W] bytes: filter type 1 + the Sub-filtered big-endian samples of every row."""
    img = np.asarray(img)
    if kind == pc.KIND_RGB8:
        raw = img.reshape(img.shape[0], -1).astype(np.uint8)
    elif kind == pc.KIND_GRAY16:
        raw = img.astype(">u2").view(np.uint8).reshape(img.shape[0], -1)
    else:
        raw = np.where(img != 0, 255, 0).astype(np.uint8)
    bpp = pc.BPP[kind]
    sub = raw.copy()
    sub[:, bpp:] = raw[:, bpp:] - raw[:, :-bpp]
    return np.concatenate([np.ones((raw.shape[0], 1), np.uint8), sub], axis=1)


def tokenize_row(f):
    """The kernel's rule: a run of equal bytes is its first byte as a literal, then the repeats cut into chunks of
    258; a chunk of >= 3 bytes is one match (distance 1), a shorter one is literals.  Runs do not cross rows."""
    L, i, out = len(f), 0, []
    while i < L:
        j = i
        while j + 1 < L and f[j + 1] == f[i]:
            j += 1
        out.append(("lit", int(f[i])))
        reps = j - i
        while reps > 0:
            c = min(258, reps)
            if c >= 3:
                out.append(("len", c))
            else:
                out.extend([("lit", int(f[i]))] * c)
            reps -= c
        i = j + 1
    return out


def token_hist(kind, imgs):
    h = np.zeros(pc.N_LITLEN, np.int64)
    for img in imgs:
        for row in scanlines(kind, img):
            for t, v in tokenize_row(row):
                h[v if t == "lit" else pc.length_symbol(v)[0]] += 1
    h[256] += len(imgs)
    return h


def encode(kind, img, table):
    """zlib stream of the image with the given table, bit for bit what pg_png_encode must produce."""
    lines = scanlines(kind, img)
    acc, n = 0x0178, 16  # zlib header bytes 0x78 0x01
    hb = int(table[pc.T_HDR_BITS])
    for k in range((hb + 31) // 32):
        acc |= int(table[pc.T_HDR + k]) << (n + 32 * k)
    n += hb
    for row in lines:
        for t, v in tokenize_row(row):
            w = int(table[pc.T_LIT + v] if t == "lit" else table[pc.T_LEN + v - 3])
            acc |= (w & 0xFFFFFF) << n
            n += w >> 24
    w = int(table[pc.T_EOB])
    acc |= (w & 0xFFFFFF) << n
    n += w >> 24
    nbytes = (n + 7) // 8
    body = acc.to_bytes(nbytes, "little")
    return body + zlib.adler32(lines.tobytes()).to_bytes(4, "big")


def decode_png(kind, data):
    import cv2
    img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    if kind == pc.KIND_RGB8:
        img = img[:, :, ::-1]
    return img
