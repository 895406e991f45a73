// png.cu — the zlib streams of a frame's PNG files, produced on the device (the reference converts and
// PNG-encodes every product on the host, /root/reference/pegasus.py:340-358 + imageio; ~170 ms of CPU per 1080p frame
// against a 1.6 ms frame).  One launch pair encodes all images of a frame:
//
//   png_size_kernel   per (image, row): stage the row, PNG Sub filter, tokenize, count the row's bits, its Adler-32
//                     partial sums (and, when calibrating, the token histogram); zero-fills the output buffers.
//   png_write_kernel  per (image, row): the same tokens again, now placed: the row's bit offset is the sum of the
//                     earlier rows' counts, tokens are packed into a shared-memory bit stream through a 64-bit
//                     register accumulator and written out as whole words (first / last word of a row: atomic OR).
//                     Row 0 adds the zlib + deflate block header, the last row the end-of-block code and Adler-32.
//
// Tokens (one dynamic-Huffman deflate block per image, code table STATIC per scene, built on the host from a
// calibration histogram: pegasus_b200/png_codec.py): a run of equal bytes of the filtered row stream is its first
// byte as a literal, then the repeats in chunks of 258 — a chunk of >= 3 bytes is ONE match (length, distance 1),
// a shorter one literals.  Piecewise-constant images (masks, semantic colours) become zeros under the Sub filter
// and collapse into matches; noisy ones (RGB, depth) are Huffman-coded residuals.  Every byte decides its token
// from its position in its run alone, so the rows are tokenized by 256 threads in parallel, no sequential pass.
// Byte / integer work, bound by shared-memory traffic; bit-exact against tests/png_model.py and stock zlib.
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

// Row `row` of the image as PNG stream bytes in shared memory: f[0] = 1 (filter type Sub), f[1 + p] = raw[p] -
// raw[p - bpp].  raw = big-endian samples; `raw_s` is staging of the unfiltered bytes.  Returns false (and leaves f
// unwritten) for a row whose samples are all zero: most rows of a mask or a semantic map; their tokens have a closed
// form (ZeroRow).
__device__ __forceinline__ bool png_stage_row(const PngImage& im, int row, int width, uint8_t* raw_s, uint8_t* f) {
    const int kind = im.kind, bpp = png_bpp(kind), nraw = bpp * width;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(im.src) + (size_t)row * im.src_pitch;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)im.src_pitch) & 3) == 0;
    const int nw = aligned ? nraw / 4 : 0;
    uint32_t* raw_w = reinterpret_cast<uint32_t*>(raw_s);
    const uint32_t* src_w = reinterpret_cast<const uint32_t*>(src);
    uint32_t nz = 0;
    for (int w = threadIdx.x; w < nw; w += PNG_THREADS) {
        uint32_t v = __ldg(src_w + w);
        if (kind == PG_PNG_GRAY16) v = __byte_perm(v, 0, 0x2301);   // two u16 -> big-endian bytes
        else if (kind == PG_PNG_MASK8) v = __vcmpne4(v, 0u);         // non-zero -> 255
        raw_w[w] = v;
        nz |= v;
    }
    for (int p = 4 * nw + threadIdx.x; p < nraw; p += PNG_THREADS) {
        uint8_t v;
        if (kind == PG_PNG_GRAY16) {
            const uint16_t s = reinterpret_cast<const uint16_t*>(src)[p >> 1];
            v = (p & 1) ? (uint8_t)(s & 255u) : (uint8_t)(s >> 8);
        } else {
            v = src[p];
            if (kind == PG_PNG_MASK8) v = v ? 255 : 0;
        }
        raw_s[p] = v;
        nz |= v;
    }
    if (!__syncthreads_or(nz != 0)) return false;
    for (int p = threadIdx.x; p < nraw; p += PNG_THREADS)
        f[1 + p] = (uint8_t)(raw_s[p] - (p >= bpp ? raw_s[p - bpp] : 0));
    if (threadIdx.x == 0) f[0] = 1;
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

// The tokens of stream positions [i0, i1) of a row of L bytes; s = last run start before i0 (-1: none), nxt = first
// run start at or after i1 (L: none) — with it the length of a chunk is found inside the thread's own positions.
template <class Lit, class Len>
__device__ __forceinline__ void png_walk(const uint8_t* f, int L, int i0, int i1, int s, int nxt, Lit&& lit, Len&& len) {
    if (i0 >= i1) return;
    const uint32_t* fw = reinterpret_cast<const uint32_t*>(f);  // i0 is a multiple of 4: four positions per load
    uint32_t prev = i0 > 0 ? f[i0 - 1] : 0x100u;                // 0x100: no byte equals it, position 0 starts a run
    for (int wb = i0; wb < i1; wb += 4) {
        const uint32_t word = fw[wb >> 2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = wb + j;
            if (i >= i1) break;
            const uint32_t b = (word >> (8 * j)) & 255u;
            if (b != prev) s = i;
            prev = b;
            const int k = i - s;
            if (k == 0) { lit(i, b); continue; }
            const int off = (k - 1) % 258;        // position inside its chunk of the run's repeats
            if (off >= 3) continue;               // inside a match of >= 4 bytes
            const int cs = i - off;               // chunk start
            const bool ge3 = cs + 2 < L && f[cs + 1] == b && f[cs + 2] == b;
            if (!ge3) { lit(i, b); continue; }    // chunk of 1 or 2 bytes: literals
            if (off == 0) {
                int e = i + 1;                    // first position after the run: inside this thread's positions, else nxt
                while (e < i1 && f[e] == b) ++e;
                if (e == i1) e = nxt;
                len(i, min(258, e - cs));
            }
        }
    }
}

struct PngRowSmem {
    uint32_t tok[PNG_TOKENS];
    uint32_t tmp[PNG_WARPS + 1];
    unsigned long long tmp64[PNG_WARPS];
};

// dynamic shared memory: PngRowSmem | raw[Lpad] | f[Lpad] | (write kernel) stream words
__device__ __forceinline__ int png_lpad(int L) { return (L + 15) / 16 * 16; }

// last (-1: none) and first (L: none) run start among positions [i0, i1)
__device__ __forceinline__ void png_starts(const uint8_t* f, int L, int i0, int i1, int* last, int* first) {
    int ls = -1, fs = L;
    if (i0 < i1) {
        const uint32_t* fw = reinterpret_cast<const uint32_t*>(f);
        uint32_t prev = i0 > 0 ? f[i0 - 1] : 0x100u;
        for (int wb = i0; wb < i1; wb += 4) {
            const uint32_t word = fw[wb >> 2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = wb + j;
                const uint32_t b = (word >> (8 * j)) & 255u;
                if (i < i1 && b != prev) {
                    ls = i;
                    fs = min(fs, i);
                }
                prev = b;
            }
        }
    }
    *last = ls;
    *first = fs;
}

// positions per thread: the row split evenly over the CTA, rounded up to whole words
__device__ __forceinline__ int png_seg(int L) { return ((L + PNG_THREADS - 1) / PNG_THREADS + 3) & ~3; }

__global__ void __launch_bounds__(PNG_THREADS) png_size_kernel(const PngBatch B) {
    extern __shared__ __align__(16) unsigned char png_smem[];
    const PngImage& im = B.img[blockIdx.y];
    const int row = blockIdx.x, H = B.height;
    const int L = png_bpp(im.kind) * B.width + 1, Lp = png_lpad(L);
    PngRowSmem& sm = *reinterpret_cast<PngRowSmem*>(png_smem);
    uint8_t* raw_s = png_smem + sizeof(PngRowSmem);
    uint8_t* f = raw_s + Lp;
    // zero-fill this row's share of the output buffer (the write kernel ORs into it)
    {
        const uint32_t n16 = im.out_capacity / 16, per = (n16 + H - 1) / H;
        uint4* o = reinterpret_cast<uint4*>(im.out);
        const uint32_t a = row * per, b = min(n16, a + per);
        for (uint32_t i = a + threadIdx.x; i < b; i += PNG_THREADS) o[i] = make_uint4(0, 0, 0, 0);
    }
    uint32_t* hist = im.hist;
    if (!png_stage_row(im, row, B.width, raw_s, f)) {
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
    __syncthreads();
    const int S = png_seg(L);
    const int i0 = min(L, (int)threadIdx.x * S), i1 = min(L, i0 + S);
    int my_last, my_first;
    png_starts(f, L, i0, i1, &my_last, &my_first);
    const int s0 = block_excl_max(my_last, reinterpret_cast<int*>(sm.tmp));
    const int nx = block_excl_min_after(my_first, L, reinterpret_cast<int*>(sm.tmp));
    uint32_t bits = 0, sumA = 0;
    unsigned long long sumB = 0;
    png_walk(f, L, i0, i1, s0, nx,
             [&](int, uint8_t b) {
                 bits += sm.tok[T_LIT + b] >> 24;
                 if (hist) atomicAdd(&hist[b], 1u);
             },
             [&](int, int cl) {
                 bits += sm.tok[T_LEN + cl - 3] >> 24;
                 if (hist) atomicAdd(&hist[png_len_symbol(cl)], 1u);
             });
    for (int i = i0; i < i1; ++i) {
        sumA += f[i];
        sumB += (unsigned long long)(L - i) * f[i];
    }
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

__global__ void __launch_bounds__(PNG_THREADS) png_write_kernel(const PngBatch B) {
    extern __shared__ __align__(16) unsigned char png_smem[];
    const PngImage& im = B.img[blockIdx.y];
    const int row = blockIdx.x, H = B.height;
    const int L = png_bpp(im.kind) * B.width + 1, Lp = png_lpad(L);
    PngRowSmem& sm = *reinterpret_cast<PngRowSmem*>(png_smem);
    uint8_t* raw_s = png_smem + sizeof(PngRowSmem);
    uint8_t* f = raw_s + Lp;
    uint32_t* out_s = reinterpret_cast<uint32_t*>(f + Lp);
    const uint32_t cap = im.out_capacity;
    // where this row starts: 16 bits of zlib header + the block header + the rows before it
    unsigned long long before = 0;
    for (int r = threadIdx.x; r < row; r += PNG_THREADS) before += im.row_bits[r];
    const uint32_t hdr_bits = __ldg(im.table + T_HDR_BITS);
    const unsigned long long base = 16ull + hdr_bits + block_sum64(before, sm.tmp64);
    const uint32_t my_bits = im.row_bits[row];
    const int shift = (int)(base & 31);
    const int nw = (int)((shift + (unsigned long long)my_bits + 31) / 32);
    for (int w = threadIdx.x; w < nw + 1; w += PNG_THREADS) out_s[w] = 0;
    bool ok = true;
    if (!png_stage_row(im, row, B.width, raw_s, f)) {
        // all-zero row: a handful of tokens with closed-form offsets, OR-ed straight into the output
        const ZeroRow z(im.table, L);
        for (int t = threadIdx.x; t < z.n; t += PNG_THREADS) {
            uint32_t off;
            const uint32_t tok = z.token(t, &off);
            ok &= png_or_bits(im.out, cap, base + off, tok & 0xFFFFFFu, (int)(tok >> 24));
        }
    } else {
        for (int i = threadIdx.x; i < PNG_TOKENS; i += PNG_THREADS) sm.tok[i] = __ldg(im.table + i);
        __syncthreads();  // tokens staged; out_s is clear as well
        const int S = png_seg(L);
        const int i0 = min(L, (int)threadIdx.x * S), i1 = min(L, i0 + S);
        int my_last, my_first;
        png_starts(f, L, i0, i1, &my_last, &my_first);
        const int s0 = block_excl_max(my_last, reinterpret_cast<int*>(sm.tmp));
        const int nx = block_excl_min_after(my_first, L, reinterpret_cast<int*>(sm.tmp));
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
            png_walk(f, L, i0, i1, s0, nx, [&](int, uint8_t b) { put(sm.tok[T_LIT + b]); },
                     [&](int, int cl) { put(sm.tok[T_LEN + cl - 3]); });
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
    if (row == 0) {
        if (threadIdx.x == 0) ok &= png_or_bits(im.out, cap, 0, 0x0178u, 16);  // zlib header: CM 8, 32 K window, level 0
        for (int k = threadIdx.x; 32 * k < (int)hdr_bits; k += PNG_THREADS)
            ok &= png_or_bits(im.out, cap, 16ull + 32ull * k, __ldg(im.table + T_HDR + k), min(32, (int)hdr_bits - 32 * k));
    }
    if (row == H - 1) {
        // end of block, padding to a byte, Adler-32 of the uncompressed stream (all rows' partial sums)
        unsigned long long a = 0, b = 0;
        for (int r = threadIdx.x; r < H; r += PNG_THREADS) {
            const unsigned long long ar = im.row_adler[2 * r];
            a += ar;
            b += (im.row_adler[2 * r + 1] + (unsigned long long)(H - 1 - r) * (unsigned long long)L % 65521ull * (ar % 65521ull)) % 65521ull;
        }
        a = block_sum64(a, sm.tmp64);
        b = block_sum64(b, sm.tmp64);
        if (threadIdx.x == 0) {
            const unsigned long long end = base + my_bits;
            const uint32_t eob = __ldg(im.table + T_EOB);
            ok &= png_or_bits(im.out, cap, end, eob & 0xFFFFFFu, (int)(eob >> 24));
            const unsigned long long nbytes = (end + (eob >> 24) + 7) / 8;
            const unsigned long long N = (unsigned long long)H * L;
            const uint32_t A = (uint32_t)((1ull + a) % 65521ull), Bv = (uint32_t)((N + b) % 65521ull);
            const uint32_t adler = (Bv << 16) | A;
            for (int q = 0; q < 4; ++q)
                ok &= png_or_bits(im.out, cap, 8ull * (nbytes + q), (adler >> (24 - 8 * q)) & 255u, 8);
            im.result[0] = (uint32_t)min(nbytes + 4ull, 0xFFFFFFFFull);
        }
    }
    if (!ok) im.result[1] = 1u;
}

size_t png_dynamic_smem(int width, bool write) {
    const size_t L = 3 * (size_t)width + 1, Lp = (L + 15) / 16 * 16;
    size_t b = sizeof(PngRowSmem) + 2 * Lp;
    if (write) b += ((15 * L + 31) / 32 + 4) * 4;
    return (b + 15) / 16 * 16;
}

int launch_png_encode(int n_images, const pg_png_image* images, int width, int height, cudaStream_t stream) {
    if (n_images == 0 || width == 0 || height == 0) return PG_OK;
    const size_t smem_size = png_dynamic_smem(width, false), smem_write = png_dynamic_smem(width, true);
    if (smem_write > 200 * 1024) {
        set_error("pg_png_encode: image width %d needs %zu bytes of shared memory per row", width, smem_write);
        return PG_ERR_INVALID;
    }
    int rc = ensure_dynamic_smem(png_size_kernel, smem_size);
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
            d.hist = s.hist;
            d.result = s.result;
            d.out_capacity = s.out_capacity;
            d.kind = s.kind;
            d.src_pitch = s.src_pitch;
        }
        const dim3 grid((unsigned)height, (unsigned)B.n);
        png_size_kernel<<<grid, PNG_THREADS, smem_size, stream>>>(B);
        PG_CUDA_CHECK(cudaGetLastError());
        png_write_kernel<<<grid, PNG_THREADS, smem_write, stream>>>(B);
        PG_CUDA_CHECK(cudaGetLastError());
        count_launch(2);
    }
    return PG_OK;
}

}  // namespace pg
