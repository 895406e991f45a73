"""Host side of the on-GPU PNG encoder (the files PEGASUS writes per frame,
/root/reference/pegasus.py:340-358 -> imageio PNGs; /root/reference/src/tools/pegasus_working.py:407-438).

Encoding a 1080p frame's 13 PNGs costs ~170 ms of host CPU (libpng through OpenCV), so 16 host cores sustain 75
frames/s next to a GPU that renders 500+.  `pg_png_encode` (csrc/png.cu) therefore produces the complete zlib
stream of every image on the device — PNG Sub filter, run-length matches (distance = 1 in the filtered stream) and
Huffman coding with a code table that is STATIC per scene — and the host only frames it: signature, IHDR, one IDAT
chunk with its CRC-32, IEND.  The files decode to exactly the pixels the reference's writer stores.

This module owns what is cheap and sequential:
  * `build_table(hist)`    token histogram (286 literal/length symbols) -> length-limited canonical Huffman code,
                           the per-token bit patterns the kernel looks up, and the dynamic-block header bits
                           (RFC 1951 §3.2.7), packed as the u32 table `pg_png_encode` reads;
  * `png_file(...)`        zlib stream -> PNG file bytes.
No compression happens here and there is no CPU encoder in the product: `tests/png_model.py` holds a slow
restatement of the kernel's tokenizer, used only to check tables and streams.
"""
from __future__ import annotations

import heapq
import struct
import zlib
from typing import List, Sequence, Tuple

import numpy as np

# image kinds of pg_png_image.kind (include/pegasus_b200.h)
KIND_RGB8, KIND_GRAY16, KIND_MASK8 = 0, 1, 2
BPP = {KIND_RGB8: 3, KIND_GRAY16: 2, KIND_MASK8: 1}
COLOR_TYPE = {KIND_RGB8: (8, 2), KIND_GRAY16: (16, 0), KIND_MASK8: (8, 0)}  # (bit depth, PNG colour type)

N_LITLEN = 286                 # literals 0..255, end of block 256, length symbols 257..285
MAX_BITS = 15
# layout of the u32 table the kernel reads (PNG_TABLE_* in csrc/png.cu)
T_LIT, T_LEN, T_EOB, T_HDR_BITS, T_HDR = 0, 256, 512, 513, 514
HDR_WORDS = 96
TABLE_WORDS = T_HDR + HDR_WORDS

# RFC 1951 §3.2.5: length symbol, extra bits and base of every match length 3..258
_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195,
             227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


def length_symbol(length: int) -> Tuple[int, int, int]:
    """(symbol 257..285, number of extra bits, extra value) of a match length 3..258."""
    if not 3 <= length <= 258:
        raise ValueError("match length out of range")
    if length == 258:
        return 285, 0, 0
    i = max(k for k in range(28) if _LEN_BASE[k] <= length)
    return 257 + i, _LEN_EXTRA[i], length - _LEN_BASE[i]


def huffman_lengths(freq: Sequence[int], max_bits: int) -> List[int]:
    """Code lengths of a complete prefix code over the symbols with freq > 0 (at least two), none longer than
    max_bits.  Plain Huffman; when the tree is too deep the frequencies are halved (floor 1) and it is rebuilt — a
    few percent from optimal in the rare case it triggers, always complete, always decodable."""
    f = [int(x) for x in freq]
    used = [i for i, x in enumerate(f) if x > 0]
    if len(used) < 2:
        raise ValueError("need at least two used symbols")
    while True:
        heap = [(f[i], i, None, None) for i in used]
        heapq.heapify(heap)
        tie = len(f)
        while len(heap) > 1:
            a = heapq.heappop(heap)
            b = heapq.heappop(heap)
            heapq.heappush(heap, (a[0] + b[0], tie, a, b))
            tie += 1
        lengths = [0] * len(f)
        stack = [(heap[0], 0)]
        while stack:
            node, d = stack.pop()
            if node[2] is None:
                lengths[node[1]] = d
            else:
                stack.append((node[2], d + 1))
                stack.append((node[3], d + 1))
        if max(lengths) <= max_bits:
            return lengths
        f = [max(1, x // 2) if x > 0 else 0 for x in f]


def canonical_codes(lengths: Sequence[int]) -> List[int]:
    """RFC 1951 §3.2.2 canonical codes (MSB-first values) of the given code lengths."""
    max_len = max(lengths)
    bl_count = [0] * (max_len + 2)
    for n in lengths:
        if n:
            bl_count[n] += 1
    code, next_code = 0, [0] * (max_len + 2)
    for bits in range(1, max_len + 1):
        code = (code + bl_count[bits - 1]) << 1
        next_code[bits] = code
    out = [0] * len(lengths)
    for i, n in enumerate(lengths):
        if n:
            out[i] = next_code[n]
            next_code[n] += 1
    return out


def _rev(code: int, nbits: int) -> int:
    """Huffman codes enter deflate's LSB-first bit stream most-significant bit first."""
    r = 0
    for _ in range(nbits):
        r = (r << 1) | (code & 1)
        code >>= 1
    return r


class _Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, value: int, nbits: int):
        self.acc |= (value & ((1 << nbits) - 1)) << self.n
        self.n += nbits


def build_table(hist: Sequence[int]) -> np.ndarray:
    """hist[286]: how often the tokenizer produced each literal / end-of-block / length symbol on sample images
    (`pg_png_image.hist`).  Every symbol gets a code (absent ones count as 1), so the table stays valid for any
    later image; only the compression ratio depends on how typical the sample was.  Returns the u32[TABLE_WORDS]
    table: [T_LIT + b] token of literal b, [T_LEN + (len - 3)] token of a match of that length at distance 1,
    [T_EOB] end of block — each `bits | nbits << 24`, ready to be OR-ed into an LSB-first stream —, [T_HDR_BITS]
    the number of header bits and [T_HDR ...] the header itself (BFINAL = 1, BTYPE = dynamic, the code lengths)."""
    h = np.asarray(hist, dtype=np.int64).reshape(-1)
    if h.size != N_LITLEN:
        raise ValueError(f"histogram must have {N_LITLEN} entries")
    freq = np.maximum(h, 1)
    freq[256] = 1
    lens = huffman_lengths(freq.tolist(), MAX_BITS)
    codes = canonical_codes(lens)
    tok = lambda sym: (_rev(codes[sym], lens[sym]), lens[sym])
    table = np.zeros(TABLE_WORDS, dtype=np.uint32)
    for b in range(256):
        bits, n = tok(b)
        table[T_LIT + b] = bits | (n << 24)
    for length in range(3, 259):
        sym, ne, ev = length_symbol(length)
        bits, n = tok(sym)
        bits |= ev << n          # extra bits follow the length code, least-significant bit first
        n += ne
        n += 1                   # the one distance code (distance 1) is a single 0 bit
        table[T_LEN + length - 3] = bits | (n << 24)
    bits, n = tok(256)
    table[T_EOB] = bits | (n << 24)

    # ---- block header: lengths of the 286 literal/length codes + 1 distance code, each sent as itself with the
    # code-length code (no repeat symbols 16..18: 287 x <= 7 bits, once per image)
    seq = list(lens) + [1]
    cl_freq = [0] * 19
    for v in seq:
        cl_freq[v] += 1
    if sum(1 for x in cl_freq if x) < 2:
        cl_freq[0 if cl_freq[0] == 0 else 1] += 1  # a complete code needs two symbols
    cl_lens = huffman_lengths(cl_freq, 7)
    cl_codes = canonical_codes(cl_lens)
    hclen = max(i for i in range(19) if cl_lens[_CL_ORDER[i]]) + 1
    hclen = max(hclen, 4)
    w = _Bits()
    w.put(1, 1)                  # BFINAL
    w.put(2, 2)                  # BTYPE = 10 dynamic Huffman
    w.put(N_LITLEN - 257, 5)     # HLIT
    w.put(0, 5)                  # HDIST: one distance code
    w.put(hclen - 4, 4)
    for i in range(hclen):
        w.put(cl_lens[_CL_ORDER[i]], 3)
    for v in seq:
        w.put(_rev(cl_codes[v], cl_lens[v]), cl_lens[v])
    if w.n > 32 * HDR_WORDS:
        raise RuntimeError("deflate block header does not fit the table")
    table[T_HDR_BITS] = w.n
    for k in range((w.n + 31) // 32):
        table[T_HDR + k] = (w.acc >> (32 * k)) & 0xFFFFFFFF
    return table


def worst_case_bytes(kind: int, width: int, height: int) -> int:
    """Upper bound of a stream: 2 bytes zlib header, the block header, <= 15 bits per filtered byte (every match
    token of <= 21 bits stands for >= 3 bytes), end of block, padding, Adler-32."""
    row = BPP[kind] * width + 1
    return 2 + (32 * HDR_WORDS + 15 * row * height + MAX_BITS + 7) // 8 + 4 + 8


def _chunk(tag: bytes, data) -> List[bytes]:
    crc = zlib.crc32(data, zlib.crc32(tag))
    return [struct.pack(">I", len(data)), tag, data, struct.pack(">I", crc)]


_SIG = b"\x89PNG\r\n\x1a\n"


def png_parts(kind: int, width: int, height: int, zstream) -> List[bytes]:
    """The pieces of the PNG file around an already compressed zlib stream (bytes-like; not copied)."""
    depth, ctype = COLOR_TYPE[kind]
    ihdr = struct.pack(">IIBBBBB", width, height, depth, ctype, 0, 0, 0)
    return [_SIG] + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zstream) + _chunk(b"IEND", b"")


def png_file(kind: int, width: int, height: int, zstream) -> bytes:
    return b"".join(bytes(p) if not isinstance(p, bytes) else p for p in png_parts(kind, width, height, zstream))


def write_png(path: str, kind: int, width: int, height: int, zstream) -> None:
    with open(path, "wb") as f:
        for p in png_parts(kind, width, height, zstream):
            f.write(p)
