// png.cu — the zlib streams of a frame's PNG files, produced on the device (the reference converts and
// PNG-encodes every product on the host, /root/reference/pegasus.py:340-358 + imageio; ~150 ms of CPU per 1080p frame
// against a 1.6 ms frame).  Three launches encode all images of a frame:
//
//   png_size_kernel   per (image, row): stage the row, PNG Sub filter (one SIMD subtraction per 4 bytes), find the run
//                     starts (one bit per byte in a 64-bit mask per thread), tokenize, count the row's bits and each
//                     thread's share of them, the row's Adler-32 partial sums (and, when calibrating, the token
//                     histogram); zero-fills the output buffers.
//   png_scan_kernel   per image: exclusive scan of the rows' bit counts = every row's place in the stream; writes what
//                     is not a row: zlib + deflate block header, end-of-block code, Adler-32, the stream's length.
//   png_write_kernel  per (image, row): the same tokens again, now placed: tokens are packed into a shared-memory
//                     bit stream through a 64-bit register accumulator and written out as whole words (first / last
//                     word of a row: atomic OR, the neighbouring rows share them).
//
// Tokens (one dynamic-Huffman deflate block per image, code table STATIC per scene, built on the host from a
// calibration histogram: pegasus_b200/png_codec.py): a run of equal bytes of the filtered row stream is its first
// byte as a literal, then the repeats in chunks of 258 — a chunk of >= 3 bytes is ONE match (length, distance 1),
// a shorter one literals.  Piecewise-constant images (masks, semantic colours) become zeros under the Sub filter
// and collapse into matches; noisy ones (RGB, depth) are Huffman-coded residuals.  Every byte's token follows from
// its position in its run alone, so a row is tokenized by 256 threads in parallel: a block-wide max-scan gives each
// thread the last run start before its bytes, a min-scan the first one after them, and inside its own bytes it
// visits run starts and chunk boundaries only (dense words — four run starts — take four literals at once).
// Byte / integer work, instruction-bound; bit-exact against tests/png_model.py and stock zlib.
#include "pg_common.cuh"

namespace pg {

constexpr int PNG_THREADS = 256;
constexpr int PNG_WARPS = PNG_THREADS / 32;
constexpr int PNG_BATCH = 24;  // images per launch (kernel parameter space)
// table layout = pegasus_b200/png_codec.py
constexpr int T_LIT = 0, T_LEN = 256, T_EOB = 512, T_HDR_BITS = 513, T_HDR = 514;
constexpr int PNG_TOKENS = 513;
static_assert(T_HDR + 96 == PG_PNG_TABLE_WORDS, "table layout");

struct PngImage {
    const void* src;
    uint8_t* out;
    const uint32_t* table;
    uint32_t* row_bits;             // [H]
    unsigned long long* row_adler;  // [H][2]: sum of the row's stream bytes, sum of (L - i) * byte i
    uint16_t* thread_bits;          // [H][PNG_THREADS] bits of each thread's tokens (size kernel -> write kernel)
    uint32_t* row_off;              // [H + 1] exclusive scan of row_bits (scan kernel -> write kernel)
    uint32_t* hist;
    uint32_t* result;
    uint32_t out_capacity;
    int kind, src_pitch;
};
struct PngBatch {
    PngImage img[PNG_BATCH];
    int n, width, height;
};

__device__ __forceinline__ int png_bpp(int kind) { return kind == PG_PNG_RGB8 ? 3 : (kind == PG_PNG_GRAY16 ? 2 : 1); }

// length 3..258 -> deflate length symbol 257..285 (RFC 1951 §3.2.5); histogram only
__device__ __forceinline__ int png_len_symbol(int len) {
    if (len == 258) return 285;
    if (len <= 10) return 254 + len;
    const int l = len - 3, e = 29 - __clz(l);  // e = floor(log2 l) - 2 extra bits
    return 257 + 4 * e + (l >> e);
}

// block-wide exclusive scans over one value per thread; `tmp` holds PNG_WARPS + 1 words
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* tmp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // tmp may still be read from an earlier call
    if (lane == 31) tmp[warp] = x;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PNG_WARPS; ++w) {
        const uint32_t c = tmp[w];
        if (w < warp) base += c;
        tot += c;
    }
    *total = tot;
    return base + x - v;
}
__device__ __forceinline__ int block_excl_max(int v, int* tmp) {  // max over the threads before this one, -1 if none
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = max(x, y);
    }
    __syncthreads();
    if (lane == 31) tmp[warp] = x;
    __syncthreads();
    int base = -1;
#pragma unroll
    for (int w = 0; w < PNG_WARPS; ++w)
        if (w < warp) base = max(base, tmp[w]);
    int prev = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) prev = -1;
    return max(base, prev);
}
__device__ __forceinline__ int block_excl_min_after(int v, int none, int* tmp) {  // min over the threads after this one
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_down_sync(0xffffffffu, x, o);
        if (lane + o < 32) x = min(x, y);
    }
    __syncthreads();
    if (lane == 0) tmp[warp] = x;
    __syncthreads();
    int base = none;
#pragma unroll
    for (int w = 0; w < PNG_WARPS; ++w)
        if (w > warp) base = min(base, tmp[w]);
    int next = __shfl_down_sync(0xffffffffu, x, 1);
    if (lane == 31) next = none;
    return min(base, next);
}
__device__ __forceinline__ unsigned long long block_sum64(unsigned long long v, unsigned long long* tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) tmp[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < PNG_WARPS; ++w) t += tmp[w];
    return t;
}

// Shared-memory layout of a row.  The PNG stream of a row is L = 1 + bpp * W bytes: position 0 = the filter type
// (1 = Sub), position 1 + p = raw[p] - raw[p - bpp] (raw = big-endian samples).  It is kept at byte OFFSET o =
// position + 3 of the word array `fw`, so that the filtered samples start on a word boundary and the filter is one
// SIMD subtraction per 4 bytes: fw[w + 1] = raw_w[w] - (raw bytes bpp earlier).  Offsets 0..2 are padding.
// Returns false (fw unwritten) for a row whose samples are all zero: most rows of a mask or a semantic map; their
// tokens have a closed form (ZeroRow).
__device__ __forceinline__ bool png_stage_row(const PngImage& im, int row, int width, uint32_t* raw_w, uint32_t* fw) {
    const int kind = im.kind, bpp = png_bpp(kind), nraw = bpp * width;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(im.src) + (size_t)row * im.src_pitch;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)im.src_pitch) & 3) == 0;
    const int nw = aligned ? nraw / 4 : 0, nwr = (nraw + 3) / 4;
    uint8_t* raw_s = reinterpret_cast<uint8_t*>(raw_w);
    const uint32_t* src_w = reinterpret_cast<const uint32_t*>(src);
    uint32_t nz = 0;
    for (int w = threadIdx.x; w < nw; w += PNG_THREADS) {
        uint32_t v = __ldg(src_w + w);
        if (kind == PG_PNG_GRAY16) v = __byte_perm(v, 0, 0x2301);   // two u16 -> big-endian bytes
        else if (kind == PG_PNG_MASK8) v = __vcmpne4(v, 0u);         // non-zero -> 255
        raw_w[w] = v;
        nz |= v;
    }
    for (int p = 4 * nw + threadIdx.x; p < 4 * nwr; p += PNG_THREADS) {  // unaligned source / the last partial word
        uint8_t v = 0;
        if (p < nraw) {
            if (kind == PG_PNG_GRAY16) {
                const uint16_t s = reinterpret_cast<const uint16_t*>(src)[p >> 1];
                v = (p & 1) ? (uint8_t)(s & 255u) : (uint8_t)(s >> 8);
            } else {
                v = src[p];
                if (kind == PG_PNG_MASK8) v = v ? 255 : 0;
            }
        }
        raw_s[p] = v;
        nz |= v;
    }
    if (!__syncthreads_or(nz != 0)) return false;
    const int sh = 32 - 8 * bpp;
    for (int w = threadIdx.x; w < nwr; w += PNG_THREADS) {
        const uint32_t cur = raw_w[w], before = w ? raw_w[w - 1] : 0u;
        fw[w + 1] = __vsub4(cur, __funnelshift_r(before, cur, sh));  // bytes 4w-bpp .. 4w-bpp+3
    }
    if (threadIdx.x == 0) fw[0] = 0x01000000u;  // offset 3 = position 0 = filter type Sub
    __syncthreads();
    return true;
}

// The tokens of an all-zero row, stream bytes [1, 0, 0, ... 0] (L bytes): literal 1, literal 0, then the L - 2 repeats
// of the zero as q matches of 258 and a tail of r < 258 (one match when r >= 3, else r literals).
struct ZeroRow {
    int q, r, n;            // n = number of tokens
    uint32_t lit1, lit0, m258, tail;  // token words (bits | nbits << 24)
    __device__ __forceinline__ ZeroRow(const uint32_t* tok, int L) {
        const int reps = L - 2;
        q = reps / 258;
        r = reps - 258 * q;
        lit1 = tok[T_LIT + 1];
        lit0 = tok[T_LIT + 0];
        m258 = tok[T_LEN + 255];
        tail = r >= 3 ? tok[T_LEN + r - 3] : lit0;
        n = 2 + q + (r >= 3 ? 1 : r);
    }
    __device__ __forceinline__ uint32_t bits() const {
        return (lit1 >> 24) + (lit0 >> 24) + (uint32_t)q * (m258 >> 24) + (r >= 3 ? 1u : (uint32_t)r) * (tail >> 24);
    }
    // token t in [0, n): its word and its bit offset inside the row
    __device__ __forceinline__ uint32_t token(int t, uint32_t* off) const {
        if (t == 0) { *off = 0; return lit1; }
        if (t == 1) { *off = lit1 >> 24; return lit0; }
        const uint32_t head = (lit1 >> 24) + (lit0 >> 24);
        if (t < 2 + q) { *off = head + (uint32_t)(t - 2) * (m258 >> 24); return m258; }
        *off = head + (uint32_t)q * (m258 >> 24) + (uint32_t)(t - 2 - q) * (tail >> 24);
        return tail;
    }
};

// ---- tokens of one thread's part of a row --------------------------------------------------------------------
// A thread owns the words [w0, w1) of `fw`, i.e. offsets [4 w0, 4 w1).  `mask` has bit j set when offset 4 w0 + j
// starts a run (its byte differs from the one before it; position 0 always does).  With the last start before the
// part (`s`, from a block-wide max-scan) and the first one after it (`nxt`, min-scan) every token follows from the
// mask alone — no byte is compared twice and no position inside a long run is visited.
struct PngPart {
    int o0, o1;               // offsets [o0, o1) that are stream positions of this thread
    unsigned long long mask;  // run starts, bit j <-> offset 4 w0 + j
    int base;                 // 4 w0
};

__device__ __forceinline__ PngPart png_part(const uint32_t* fw, int Lo) {
    const int nwo = (Lo + 3) / 4;
    const int sw = (nwo + PNG_THREADS - 1) / PNG_THREADS;   // words per thread (<= 16, checked by the launcher)
    const int w0 = min(nwo, (int)threadIdx.x * sw), w1 = min(nwo, w0 + sw);
    PngPart p;
    p.base = 4 * w0;
    p.o0 = max(p.base, 3);
    p.o1 = min(4 * w1, Lo);
    unsigned long long m = 0;
    uint32_t prevw = w0 > 0 ? fw[w0 - 1] : 0u;
    for (int w = w0; w < w1; ++w) {
        const uint32_t cur = fw[w];
        const uint32_t ne = __vcmpne4(cur, (cur << 8) | (prevw >> 24));      // 0xFF where byte != the byte before it
        const uint32_t nib = (((ne & 0x01010101u) * 0x01020408u) >> 24) & 15u;  // one bit per byte
        m |= (unsigned long long)nib << (4 * (w - w0));
        prevw = cur;
    }
    if (w0 == 0 && w1 > 0) m = (m & ~7ull) | 8ull;            // offsets 0..2 are padding, offset 3 always starts a run
    const int n = p.o1 - p.base;                              // drop the offsets beyond the row
    if (n < 64) m &= n > 0 ? ((1ull << n) - 1ull) : 0ull;
    p.mask = m;
    return p;
}

// repeats of a run: offsets [a, b) all continue the run that started at offset s and ends before offset e
template <class Lit, class Len>
__device__ __forceinline__ void png_repeats(uint32_t v, int s, int a, int b, int e, Lit&& lit, Len&& len) {
    int cs = s + 1 + ((a - s - 1) / 258) * 258;     // start of the 258-chunk that holds offset a
    for (; cs < b; cs += 258) {
        const int cl = min(258, e - cs);
        if (cl >= 3) {
            if (cs >= a) len(cl);                   // one match, emitted by the chunk's first offset
        } else {
            for (int o = max(cs, a); o < min(cs + cl, b); ++o) lit(v);
        }
    }
}

template <class Lit, class Len>
__device__ __forceinline__ void png_walk(const uint32_t* fw, const PngPart& p, int s, int nxt, Lit&& lit, Len&& len) {
    const uint8_t* f8 = reinterpret_cast<const uint8_t*>(fw);
    int pos = p.o0;
    while (pos < p.o1) {
        const unsigned long long m = p.mask >> (pos - p.base);
        // dense data (photographic rows: every byte differs from its neighbour): four literals straight from the word
        if ((pos & 3) == 0 && pos + 4 <= p.o1 && ((uint32_t)m & 15u) == 15u) {
            uint32_t word = fw[pos >> 2];
#pragma unroll 1
            for (int j = 0; j < 4; ++j, word >>= 8) lit(word & 255u);
            s = pos + 3;
            pos += 4;
            continue;
        }
        const int next = m ? pos + __ffsll((long long)m) - 1 : p.o1;   // next run start at or after pos
        if (next > pos) png_repeats(f8[s], s, pos, next, next < p.o1 ? next : nxt, lit, len);
        if (next < p.o1) {
            s = next;
            lit(f8[next]);
        }
        pos = next + 1;
    }
}

struct PngRowSmem {
    uint32_t tok[PNG_TOKENS];
    uint32_t tmp[PNG_WARPS + 1];
    unsigned long long tmp64[PNG_WARPS];
};

// dynamic shared memory: PngRowSmem | raw words | fw words | (write kernel) stream words
__device__ __forceinline__ int png_row_words(int L) { return ((L + 3 + 3) / 4 + 4 + 3) & ~3; }

// sum of the stream bytes of the thread's part and of (L - position) * byte (Adler-32 partial sums)
__device__ __forceinline__ void png_adler_part(const uint32_t* fw, const PngPart& p, int L, uint32_t* sumA,
                                               unsigned long long* sumB) {
    uint32_t a = 0;
    unsigned long long b = 0;
    for (int o = p.base; o < p.o1; o += 4) {
        uint32_t word = fw[o >> 2];
        if (o < 3) word &= 0xFF000000u;                                  // padding offsets
        if (o + 4 > p.o1) word &= 0xFFFFFFFFu >> (8 * (o + 4 - p.o1));   // offsets beyond the part
        const uint32_t sa = __dp4a(word, 0x01010101u, 0u);
        a += sa;
        // weight of offset o + j: L - (o + j - 3)
        b += (unsigned long long)(L + 3 - o) * sa - __dp4a(word, 0x03020100u, 0u);
    }
    *sumA = a;
    *sumB = b;
}

template <bool HIST>
__global__ void __launch_bounds__(PNG_THREADS) png_size_kernel(const PngBatch B) {
    extern __shared__ __align__(16) unsigned char png_smem[];
    const PngImage& im = B.img[blockIdx.y];
    const int row = blockIdx.x, H = B.height;
    const int L = png_bpp(im.kind) * B.width + 1, Lo = L + 3, RW = png_row_words(L);
    PngRowSmem& sm = *reinterpret_cast<PngRowSmem*>(png_smem);
    uint32_t* raw_w = reinterpret_cast<uint32_t*>(png_smem + sizeof(PngRowSmem));
    uint32_t* fw = raw_w + RW;
    // zero-fill this row's share of the output buffer (the write kernel ORs into it)
    {
        const uint32_t n16 = im.out_capacity / 16, per = (n16 + H - 1) / H;
        uint4* o = reinterpret_cast<uint4*>(im.out);
        const uint32_t a = row * per, b = min(n16, a + per);
        for (uint32_t i = a + threadIdx.x; i < b; i += PNG_THREADS) o[i] = make_uint4(0, 0, 0, 0);
    }
    uint32_t* hist = HIST ? im.hist : nullptr;  // calibration launches only
    if (!png_stage_row(im, row, B.width, raw_w, fw)) {
        if (threadIdx.x == 0) {
            const ZeroRow z(im.table, L);
            if (row == 0) im.result[0] = im.result[1] = 0u;
            im.row_bits[row] = z.bits();
            im.row_adler[2 * row] = 1ull;                       // the filter-type byte
            im.row_adler[2 * row + 1] = (unsigned long long)L;  // (L - 0) * 1
            if (hist) {
                atomicAdd(&hist[1], 1u);
                atomicAdd(&hist[0], 1u + (z.r < 3 ? (uint32_t)z.r : 0u));
                if (z.q) atomicAdd(&hist[285], (uint32_t)z.q);
                if (z.r >= 3) atomicAdd(&hist[png_len_symbol(z.r)], 1u);
                if (row == 0) atomicAdd(&hist[256], 1u);
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < PNG_TOKENS; i += PNG_THREADS) sm.tok[i] = __ldg(im.table + i);
    const PngPart part = png_part(fw, Lo);
    const int my_last = part.mask ? part.base + 63 - __clzll((long long)part.mask) : -1;
    const int my_first = part.mask ? part.base + __ffsll((long long)part.mask) - 1 : Lo;
    const int s0 = block_excl_max(my_last, reinterpret_cast<int*>(sm.tmp));   // also orders sm.tok
    const int nx = block_excl_min_after(my_first, Lo, reinterpret_cast<int*>(sm.tmp));
    uint32_t bits = 0;
    png_walk(fw, part, s0, nx,
             [&](uint32_t b) {
                 bits += sm.tok[T_LIT + b] >> 24;
                 if (HIST && hist) atomicAdd(&hist[b], 1u);
             },
             [&](int cl) {
                 bits += sm.tok[T_LEN + cl - 3] >> 24;
                 if (HIST && hist) atomicAdd(&hist[png_len_symbol(cl)], 1u);
             });
    uint32_t sumA;
    unsigned long long sumB;
    png_adler_part(fw, part, L, &sumA, &sumB);
    uint32_t total;
    block_excl_sum(bits, sm.tmp, &total);
    im.thread_bits[(size_t)row * PNG_THREADS + threadIdx.x] = (uint16_t)bits;
    const unsigned long long A = block_sum64(sumA, sm.tmp64);
    const unsigned long long Bs = block_sum64(sumB, sm.tmp64);
    if (threadIdx.x == 0) {
        if (row == 0) im.result[0] = im.result[1] = 0u;
        im.row_bits[row] = total;
        im.row_adler[2 * row] = A;
        im.row_adler[2 * row + 1] = Bs;
        if (hist && row == 0) atomicAdd(&hist[256], 1u);
    }
}

// OR `nbits` (<= 32) bits of `v` into the byte stream `out` at bit position `pos`, staying below `cap` bytes
__device__ __forceinline__ bool png_or_bits(uint8_t* out, uint32_t cap, unsigned long long pos, uint32_t v, int nbits) {
    if (nbits == 0) return true;
    if ((pos + nbits + 7) / 8 > cap) return false;
    uint32_t* w = reinterpret_cast<uint32_t*>(out);
    const unsigned long long x = (unsigned long long)v << (pos & 31);
    atomicOr(&w[pos >> 5], (uint32_t)x);
    if ((uint32_t)(x >> 32)) atomicOr(&w[(pos >> 5) + 1], (uint32_t)(x >> 32));
    return true;
}

// One CTA per image between the two row kernels: exclusive scan of the rows' bit counts (every row's place in the
// stream), and everything of the stream that is not a row: zlib header, deflate block header, end-of-block code,
// padding, Adler-32 combined from the rows' partial sums, the stream's length.
__global__ void __launch_bounds__(PNG_THREADS) png_scan_kernel(const PngBatch B) {
    __shared__ uint32_t tmp[PNG_WARPS + 1];
    __shared__ unsigned long long tmp64[PNG_WARPS];
    const PngImage& im = B.img[blockIdx.x];
    const int H = B.height, L = png_bpp(im.kind) * B.width + 1;
    const uint32_t cap = im.out_capacity;
    // rows [r0, r1) of this thread, in order
    const int per = (H + PNG_THREADS - 1) / PNG_THREADS;
    const int r0 = min(H, (int)threadIdx.x * per), r1 = min(H, r0 + per);
    uint32_t mine = 0;
    unsigned long long a = 0, b = 0;
    for (int r = r0; r < r1; ++r) {
        mine += im.row_bits[r];
        const unsigned long long ar = im.row_adler[2 * r];
        a += ar;
        b += (im.row_adler[2 * r + 1] + (unsigned long long)(H - 1 - r) * (unsigned long long)L % 65521ull * (ar % 65521ull)) % 65521ull;
    }
    uint32_t total;
    uint32_t run = block_excl_sum(mine, tmp, &total);
    for (int r = r0; r < r1; ++r) {
        im.row_off[r] = run;
        run += im.row_bits[r];
    }
    a = block_sum64(a, tmp64);
    b = block_sum64(b, tmp64);
    const uint32_t hdr_bits = __ldg(im.table + T_HDR_BITS);
    bool ok = true;
    for (int k = threadIdx.x; 32 * k < (int)hdr_bits; k += PNG_THREADS)
        ok &= png_or_bits(im.out, cap, 16ull + 32ull * k, __ldg(im.table + T_HDR + k), min(32, (int)hdr_bits - 32 * k));
    if (threadIdx.x == 0) {
        im.row_off[H] = total;
        ok &= png_or_bits(im.out, cap, 0, 0x0178u, 16);  // zlib header: CM 8, 32 K window, level 0
        const unsigned long long end = 16ull + hdr_bits + total;
        const uint32_t eob = __ldg(im.table + T_EOB);
        ok &= png_or_bits(im.out, cap, end, eob & 0xFFFFFFu, (int)(eob >> 24));
        const unsigned long long nbytes = (end + (eob >> 24) + 7) / 8;
        const unsigned long long N = (unsigned long long)H * L;
        const uint32_t A = (uint32_t)((1ull + a) % 65521ull), Bv = (uint32_t)((N + b) % 65521ull);
        const uint32_t adler = (Bv << 16) | A;
        for (int q = 0; q < 4; ++q) ok &= png_or_bits(im.out, cap, 8ull * (nbytes + q), (adler >> (24 - 8 * q)) & 255u, 8);
        im.result[0] = (uint32_t)min(nbytes + 4ull, 0xFFFFFFFFull);
    }
    if (!ok) im.result[1] = 1u;
}

__global__ void __launch_bounds__(PNG_THREADS) png_write_kernel(const PngBatch B) {
    extern __shared__ __align__(16) unsigned char png_smem[];
    const PngImage& im = B.img[blockIdx.y];
    const int row = blockIdx.x;
    const int L = png_bpp(im.kind) * B.width + 1, Lo = L + 3, RW = png_row_words(L);
    PngRowSmem& sm = *reinterpret_cast<PngRowSmem*>(png_smem);
    uint32_t* raw_w = reinterpret_cast<uint32_t*>(png_smem + sizeof(PngRowSmem));
    uint32_t* fw = raw_w + RW;
    uint32_t* out_s = fw + RW;
    const uint32_t cap = im.out_capacity;
    // where this row starts: 16 bits of zlib header + the block header + the rows before it (png_scan_kernel)
    const unsigned long long base = 16ull + __ldg(im.table + T_HDR_BITS) + im.row_off[row];
    const uint32_t my_bits = im.row_bits[row];
    const int shift = (int)(base & 31);
    const int nw = (int)((shift + (unsigned long long)my_bits + 31) / 32);
    for (int w = threadIdx.x; w < nw + 1; w += PNG_THREADS) out_s[w] = 0;
    bool ok = true;
    if (!png_stage_row(im, row, B.width, raw_w, fw)) {
        // all-zero row: a handful of tokens with closed-form offsets, OR-ed straight into the output
        const ZeroRow z(im.table, L);
        for (int t = threadIdx.x; t < z.n; t += PNG_THREADS) {
            uint32_t off;
            const uint32_t tok = z.token(t, &off);
            ok &= png_or_bits(im.out, cap, base + off, tok & 0xFFFFFFu, (int)(tok >> 24));
        }
    } else {
        for (int i = threadIdx.x; i < PNG_TOKENS; i += PNG_THREADS) sm.tok[i] = __ldg(im.table + i);
        const PngPart part = png_part(fw, Lo);
        const int my_last = part.mask ? part.base + 63 - __clzll((long long)part.mask) : -1;
        const int my_first = part.mask ? part.base + __ffsll((long long)part.mask) - 1 : Lo;
        const int s0 = block_excl_max(my_last, reinterpret_cast<int*>(sm.tmp));  // its barriers also order sm.tok / out_s
        const int nx = block_excl_min_after(my_first, Lo, reinterpret_cast<int*>(sm.tmp));
        const uint32_t bits = im.thread_bits[(size_t)row * PNG_THREADS + threadIdx.x];  // counted by the size kernel
        uint32_t total;
        const uint32_t excl = block_excl_sum(bits, sm.tmp, &total);
        // second walk: place the tokens.  64-bit accumulator; a word is stored plainly once this thread has produced
        // it up to its last bit and did not start inside it, else OR-ed (neighbouring threads share those words).
        {
            const uint32_t pos = (uint32_t)shift + excl;
            unsigned long long acc = 0;
            int fill = (int)(pos & 31), w = (int)(pos >> 5);
            bool first = true;
            auto put = [&](uint32_t tok) {
                acc |= (unsigned long long)(tok & 0xFFFFFFu) << fill;
                fill += (int)(tok >> 24);
                if (fill >= 32) {
                    if (first) atomicOr(&out_s[w], (uint32_t)acc);
                    else out_s[w] = (uint32_t)acc;
                    first = false;
                    acc >>= 32;
                    fill -= 32;
                    ++w;
                }
            };
            png_walk(fw, part, s0, nx, [&](uint32_t b) { put(sm.tok[T_LIT + b]); },
                     [&](int cl) { put(sm.tok[T_LEN + cl - 3]); });
            if (fill > 0 && acc) atomicOr(&out_s[w], (uint32_t)acc);
        }
        __syncthreads();
        // shared-memory stream -> global: whole words, the row's first and last word shared with its neighbours
        uint32_t* g = reinterpret_cast<uint32_t*>(im.out) + (base >> 5);
        const unsigned long long w0 = base >> 5;
        for (int w = threadIdx.x; w < nw; w += PNG_THREADS) {
            if ((w0 + w + 1) * 4 > cap) { ok = false; continue; }
            const uint32_t v = out_s[w];
            if (w == 0 || w == nw - 1) { if (v) atomicOr(&g[w], v); }
            else g[w] = v;
        }
    }
    if (!ok) im.result[1] = 1u;
}

size_t png_dynamic_smem(int width, bool write) {
    const size_t L = 3 * (size_t)width + 1, rw = ((L + 3 + 3) / 4 + 4 + 3) & ~(size_t)3;   // = png_row_words(L)
    size_t b = sizeof(PngRowSmem) + 2 * rw * 4;
    if (write) b += ((15 * L + 31) / 32 + 4) * 4;
    return (b + 15) / 16 * 16;
}

int launch_png_encode(int n_images, const pg_png_image* images, int width, int height, cudaStream_t stream) {
    if (n_images == 0 || width == 0 || height == 0) return PG_OK;
    const size_t smem_size = png_dynamic_smem(width, false), smem_write = png_dynamic_smem(width, true);
    // a thread's part of a row is at most 16 words (its run starts live in one 64-bit mask)
    if (smem_write > 200 * 1024 || (3 * (size_t)width + 1 + 3 + 3) / 4 > 16 * (size_t)PNG_THREADS) {
        set_error("pg_png_encode: image width %d is not supported (rows of at most %d stream bytes)", width,
                  64 * PNG_THREADS - 8);
        return PG_ERR_INVALID;
    }
    int rc = ensure_dynamic_smem(png_size_kernel<false>, smem_size);
    if (rc != PG_OK) return rc;
    rc = ensure_dynamic_smem(png_size_kernel<true>, smem_size);
    if (rc != PG_OK) return rc;
    rc = ensure_dynamic_smem(png_write_kernel, smem_write);
    if (rc != PG_OK) return rc;
    for (int first = 0; first < n_images; first += PNG_BATCH) {
        PngBatch B;
        B.n = min(PNG_BATCH, n_images - first);
        B.width = width;
        B.height = height;
        for (int i = 0; i < B.n; ++i) {
            const pg_png_image& s = images[first + i];
            PngImage& d = B.img[i];
            d.src = s.src;
            d.out = s.out;
            d.table = s.table;
            d.row_bits = reinterpret_cast<uint32_t*>(s.scratch);
            unsigned char* sc = reinterpret_cast<unsigned char*>(s.scratch);
            d.row_adler = reinterpret_cast<unsigned long long*>(sc + (((size_t)height * 4 + 15) / 16 * 16));
            d.thread_bits = reinterpret_cast<uint16_t*>(sc + (((size_t)height * 4 + 15) / 16 * 16) + (size_t)height * 16);
            d.row_off = reinterpret_cast<uint32_t*>(sc + (((size_t)height * 4 + 15) / 16 * 16) + (size_t)height * (16 + 512));
            d.hist = s.hist;
            d.result = s.result;
            d.out_capacity = s.out_capacity;
            d.kind = s.kind;
            d.src_pitch = s.src_pitch;
        }
        const dim3 grid((unsigned)height, (unsigned)B.n);
        bool any_hist = false;
        for (int i = 0; i < B.n; ++i) any_hist |= B.img[i].hist != nullptr;
        if (any_hist) png_size_kernel<true><<<grid, PNG_THREADS, smem_size, stream>>>(B);
        else png_size_kernel<false><<<grid, PNG_THREADS, smem_size, stream>>>(B);
        PG_CUDA_CHECK(cudaGetLastError());
        PG_CUDA_CHECK(cudaGetLastError());
        png_scan_kernel<<<(unsigned)B.n, PNG_THREADS, 0, stream>>>(B);
        PG_CUDA_CHECK(cudaGetLastError());
        png_write_kernel<<<grid, PNG_THREADS, smem_write, stream>>>(B);
        PG_CUDA_CHECK(cudaGetLastError());
        count_launch(3);
    }
    return PG_OK;
}

}  // namespace pg
