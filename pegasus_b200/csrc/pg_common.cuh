// pg_common.cuh — shared device helpers and the workspace layout of libpegasus_b200.so (sm_100a).
//
// Numerical contract: every float op on the result path is an explicit
// round-to-nearest intrinsic (__fmul_rn / __fadd_rn / __fsub_rn / __fmaf_rn / __fdiv_rn / __fsqrt_rn)
// in the order nvcc -fmad=true contracts the upstream rasterizer's expressions into, so results are
// reproducible bit-for-bit by the CPU oracle and independent of compiler contraction decisions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pegasus_b200.h"

#define PG_TILE 16
#define PG_SM_COUNT 148

namespace pg {

// ---- explicit IEEE helpers -------------------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }

// exp(x), x <= 0: Cody-Waite + degree-6 polynomial out of FMAs only (no MUFU), <= 1.3 ulp.
// Same sequence as orc_expf in oracle/pegasus_oracle.c.
__device__ __forceinline__ float expf_exact(float x) {
    const float L2E = 1.44269502162933349609375f;
    const float MAGIC = 12582912.0f;
    const float LN2_HI = 0.693145751953125f;
    const float LN2_LO = 1.428606765330187045e-06f;
    if (x < -80.0f) return 0.0f;
    float z = fma(x, L2E, MAGIC);
    float n = sub(z, MAGIC);
    float r = fma(n, -LN2_HI, x);
    r = fma(n, -LN2_LO, r);
    float p = 0x1.6b5016p-10f;
    p = fma(p, r, 0x1.126caep-7f);
    p = fma(p, r, 0x1.55578ep-5f);
    p = fma(p, r, 0x1.55540cp-3f);
    p = fma(p, r, 0x1.fffffcp-2f);
    p = fma(p, r, 1.0f);
    p = fma(p, r, 1.0f);
    int ni = __float_as_int(z) - 0x4B400000;
    return __int_as_float(__float_as_int(p) + (ni << 23));
}

// Bit 31 of a point-list value: the binning stage proved that this Gaussian cannot reach alpha >= 1/255
// at any pixel centre of this tile; compositing drops the entry without fetching its record.
#define PG_CULL_FLAG 0x80000000u

// Conservative test: can a Gaussian with centre (gx,gy), conic (qa,qb,qc) and cut `cut`
// (alpha < 1/255 wherever power < cut) be skipped for every pixel centre of [x0,x1]x[y0,y1]?
// q(d) = 0.5*qa*dx^2 + qb*dx*dy + 0.5*qc*dy^2 is convex for a positive-definite conic, so its minimum
// over the rectangle is 0 if the centre is inside, else it lies on the edges facing the centre,
// where it has a closed form.  The 1e-4 relative + 1e-3 absolute margin dwarfs the rounding of both
// this test and the compositing arithmetic; indefinite conics and NaNs are never culled.
__device__ __forceinline__ bool block_culled(float gx, float gy, float qa, float qb, float qc, float cut,
                                             float x0, float x1, float y0, float y1) {
    const float xl = x0 - gx, xh = x1 - gx, yl = y0 - gy, yh = y1 - gy;
    const float cx = fminf(fmaxf(0.0f, xl), xh), cy = fminf(fmaxf(0.0f, yl), yh);
    const float dy1 = fminf(fmaxf(-qb * cx * __frcp_rn(qc), yl), yh);
    const float dx2 = fminf(fmaxf(-qb * cy * __frcp_rn(qa), xl), xh);
    const float q1 = 0.5f * (qa * cx * cx + qc * dy1 * dy1) + qb * cx * dy1;
    const float q2 = 0.5f * (qa * dx2 * dx2 + qc * cy * cy) + qb * dx2 * cy;
    const float qmin = fminf(q1, q2);
    const bool pd = qa > 0.0f && qc > 0.0f && qa * qc - qb * qb > 0.0f;
    return pd && (qmin * 0.9999f - 1e-3f > -cut);
}

// Bit 6 of GeomRec::b.w: the record needs the compositing kernel's general path (opacity > 0.99, where the
// reference's min(0.99, .) can bind, or a non-finite field); all other records take the lean path.
#define PG_REC_GENERAL 0x40
#define PG_REC_FLAGS 0x7F

// Tile rectangle of a Gaussian: [min, max) tile columns / rows touched by the square of side 2 radius + 1
// around the pixel centre, clamped to the grid.  Used by preprocess (radii, pair counts) and by the binning stage,
// which re-derives the rectangle from the record instead of gathering it from a second array.
__device__ __forceinline__ ushort4 tile_rect(float pixx, float pixy, int radius, int gx, int gy) {
    const float fr = (float)radius;
    int rminx = (int)div(sub(pixx, fr), 16.0f), rminy = (int)div(sub(pixy, fr), 16.0f);
    int rmaxx = (int)div(add(add(pixx, fr), 15.0f), 16.0f), rmaxy = (int)div(add(add(pixy, fr), 15.0f), 16.0f);
    rminx = min(gx, max(0, rminx)); rminy = min(gy, max(0, rminy));
    rmaxx = min(gx, max(0, rmaxx)); rmaxy = min(gy, max(0, rmaxy));
    return make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx, (unsigned short)rmaxy);
}

// ---- per-Gaussian record staged into shared memory by the compositing kernel (48 B) ----------
struct __align__(16) GeomRec {
    float4 a;  // x, y, conic.x, conic.y
    float4 b;  // conic.z, opacity, depth, cut (power below which alpha < 1/255; mantissa bits 0-5 = object id,
               // bit 6 = PG_REC_GENERAL)
    float4 c;  // r, g, b, radius in pixels (as int bits: with the pixel centre it gives the tile rectangle, tile_rect())
};

// ---- status block at the head of the workspace ------------------------------------------------
struct Counters {
    uint32_t num_rendered;   // the reference's R (every tile of every rectangle), saturating
    uint32_t overflow;
    uint32_t num_visible;
    uint32_t sort_n;         // pairs stored = min(pairs kept by the binning stage, pair capacity): what the tile sort processes
    uint32_t tile_counter[8];  // dynamic CTA-tile tickets: [0..3] depth-sort passes, [5..6] tile-sort passes,
                               // [7] depth-key compaction
    uint32_t run_ovf;          // rows beyond RUN_FIX of tall rectangles, allocated in runs_ovf by count_kernel
    unsigned long long stats[8];  // debug&2: pairs evaluated, pairs reaching exp, pairs blended (all chains), pixel slots walked,
                                  // warp-hits: environment, object while a main chain lives, object afterwards; cull passes afterwards
    unsigned long long rendered_full;  // sum of all tile-rectangle areas = the reference's num_rendered
};

// Sticky status behind the Counters (never cleared by a forward; pg_workspace_init clears it): pg_status is
// {Counters' first four words, Sticky}.
struct Sticky {
    uint32_t overflow_frames;   // forwards whose stored pairs exceeded the pair capacity
    uint32_t max_pairs_needed;  // largest stored-pair demand seen
};

// Onesweep tile geometry
constexpr int SORT_THREADS = 256;
#ifndef PG_SORT_IPT
#define PG_SORT_IPT 16
#endif
constexpr int SORT_IPT = PG_SORT_IPT;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;  // 4096 items per CTA-tile
constexpr int RADIX = 256;
constexpr int SCAN_TILE = 2048;  // items per CTA of the depth-key compaction
constexpr int OBJ_SPREAD = 8;    // partial per-tile object counters (hot words serialise in L2)
constexpr int RUN_FIX = 8;       // tile rows per rectangle with a fixed slot in the run table (taller ones spill)

struct Layout {
    // all offsets in bytes from the workspace base; every region 256-B aligned
    size_t sticky;       // Sticky (first, so that it lies outside the cleared range)
    size_t counters;     // Counters
    size_t zero_begin;   // [zero_begin, zero_end) is cleared at the start of every forward
    size_t hist_depth;   // u32[4][256]  depth-sort digit histograms -> exclusive bases
    size_t hist_tile;    // u32[2][256] digit histograms of the two tile-sort passes
    size_t ranges;       // uint2[tiles] (cleared: tiles without pairs keep (0,0))
    size_t tile_obj_count; // u32[tiles][OBJ_SPREAD] un-culled pairs whose Gaussian belongs to an object (partial counts:
                           // the last sort pass spreads its atomics over OBJ_SPREAD words per tile)
    size_t status_depth; // u32[4][tilesP][256]
    size_t status_compact; // u32[tilesC] look-back of the depth-key compaction
    size_t status_tile;  // u32[2][tilesR][256]
    size_t zero_end;
    size_t bins_tile;    // u32[2][256] exclusive bases of the two tile-sort passes
    size_t tile_order;   // u32[tiles] tile ids by descending list length (compositing launch order)
    size_t recs;         // GeomRec[P]
    size_t srect;        // ushort4[P] tile rectangles of the visible Gaussians in depth order (count_kernel -> emit_kernel)
    size_t ovf_base;     // u32[P] where the rows beyond RUN_FIX of a tall rectangle live in runs_ovf
    size_t grp_loc;      // u32[ceil(P / 32)] pairs of the count_kernel CTA's groups (32 sorted positions) before each group
    size_t cta_pairs, cta_base;  // u32[ceil(P / 256)] pairs per count_kernel CTA and their exclusive scan
    size_t rnd_off;      // u32[ceil(P / 32)][gy] pairs of a group's rounds (32 tile rows) before each round
    size_t runs_fix;     // u32[P][RUN_FIX] ta | tb << 11 of the first RUN_FIX tile rows of every visible rectangle (depth order)
    size_t runs_ovf;     // u32[R_cap] the rows beyond RUN_FIX (more rows than the pair capacity count as an overflow)
    size_t dkey_a, dkey_b, dval_a, dval_b;  // u32[P] depth-sort ping-pong
    size_t tkey_a, tkey_b, tval_a, tval_b;  // u32[R_cap] tile-sort ping-pong
    size_t total;
    uint32_t tiles, tilesP, tilesR, tilesC;
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline Layout make_layout(int P, int W, int H, uint64_t R_cap) {
    Layout L;
    uint32_t gx = (W + PG_TILE - 1) / PG_TILE, gy = (H + PG_TILE - 1) / PG_TILE;
    L.tiles = gx * gy;
    L.tilesP = (uint32_t)((P + SORT_TILE - 1) / SORT_TILE);
    if (L.tilesP == 0) L.tilesP = 1;
    L.tilesR = (uint32_t)((R_cap + SORT_TILE - 1) / SORT_TILE);
    if (L.tilesR == 0) L.tilesR = 1;
    L.tilesC = (uint32_t)((P + SCAN_TILE - 1) / SCAN_TILE);
    if (L.tilesC == 0) L.tilesC = 1;
    size_t o = 0;
    L.sticky = o; o = align_up(o + sizeof(Sticky));
    L.counters = o; o = align_up(o + sizeof(Counters));
    L.zero_begin = L.counters;
    L.hist_depth = o; o = align_up(o + 4 * RADIX * 4);
    L.hist_tile = o; o = align_up(o + 2 * RADIX * 4);
    L.ranges = o; o = align_up(o + (size_t)L.tiles * 8);
    L.tile_obj_count = o; o = align_up(o + (size_t)L.tiles * OBJ_SPREAD * 4);
    L.status_depth = o; o = align_up(o + (size_t)4 * L.tilesP * RADIX * 4);
    L.status_compact = o; o = align_up(o + (size_t)L.tilesC * 4);
    L.status_tile = o; o = align_up(o + (size_t)2 * L.tilesR * RADIX * 4);
    L.zero_end = o;
    L.bins_tile = o; o = align_up(o + 2 * RADIX * 4);
    L.tile_order = o; o = align_up(o + (size_t)L.tiles * 4);
    L.recs = o; o = align_up(o + (size_t)P * sizeof(GeomRec));
    L.srect = o; o = align_up(o + (size_t)P * 8);
    L.ovf_base = o; o = align_up(o + (size_t)P * 4);
    L.grp_loc = o; o = align_up(o + ((size_t)P / 32 + 1) * 4);
    L.cta_pairs = o; o = align_up(o + ((size_t)P / 256 + 1) * 4);
    L.cta_base = o; o = align_up(o + ((size_t)P / 256 + 1) * 4);
    L.rnd_off = o; o = align_up(o + ((size_t)P / 32 + 1) * (size_t)gy * 4);
    L.runs_fix = o; o = align_up(o + (size_t)P * RUN_FIX * 4);
    L.runs_ovf = o; o = align_up(o + (size_t)R_cap * 4);
    L.dkey_a = o; o = align_up(o + (size_t)P * 4);
    L.dkey_b = o; o = align_up(o + (size_t)P * 4);
    L.dval_a = o; o = align_up(o + (size_t)P * 4);
    L.dval_b = o; o = align_up(o + (size_t)P * 4);
    L.tkey_a = o; o = align_up(o + (size_t)R_cap * 4);
    L.tkey_b = o; o = align_up(o + (size_t)R_cap * 4);
    L.tval_a = o; o = align_up(o + (size_t)R_cap * 4);
    L.tval_b = o; o = align_up(o + (size_t)R_cap * 4);
    L.total = o;
    return L;
}

// number of bits needed to index `tiles` tile ids (>= 1)
inline int tile_bits(uint32_t tiles) {
    int b = 1;
    while ((1u << b) < tiles) ++b;
    return b;
}

void set_error(const char* fmt, ...);
void count_launch(int n);

#if defined(__CUDACC__)
// Decoupled look-back by ONE WARP, 32 predecessors per step: lane l inspects tile (t - l); the usable prefix of the
// window ends at the first tile that has published nothing yet or at the first inclusive prefix.  Returns the
// exclusive prefix of `tile` (all lanes).  Status word: value | flag << SHIFT (0 = nothing, 1 = aggregate, 2 = inclusive).
template <typename T, int SHIFT>
__device__ __forceinline__ T warp_lookback(const volatile T* status, uint32_t tile, int lane) {
    const T kVal = (T(1) << SHIFT) - 1;
    T prev = 0;
    int t = (int)tile - 1;
    while (true) {
        const int mine = t - lane;
        const T sv = mine >= 0 ? status[mine] : (T(2) << SHIFT);
        const uint32_t f = (uint32_t)(sv >> SHIFT);
        const uint32_t not_ready = __ballot_sync(0xffffffffu, f == 0);
        const uint32_t incl = __ballot_sync(0xffffffffu, f == 2);
        const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
        const int first_in = incl ? __ffs(incl) - 1 : 32;
        const int take = min(first_nr, first_in + 1);  // lanes [0, take)
        T v = lane < take ? (sv & kVal) : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        prev += v;
        if (first_in < first_nr) break;
        t -= take;
    }
    return prev;
}

#endif

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: what has been granted is
// remembered per (kernel, device) — keyed by the function's address, because kernels with the same signature
// share one instantiation of this template — behind a mutex, so several host threads / devices are fine.
int smem_granted(const void* kernel, int device, int bytes, bool record);
#if defined(__CUDACC__)
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, bool max_carveout = false) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const void* key = reinterpret_cast<const void*>(kernel);
    if (smem_granted(key, dev, bytes, false) >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    if (max_carveout) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
    }
    smem_granted(key, dev, bytes, true);
    return cudaSuccess;
}
#endif

}  // namespace pg

#define PG_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            pg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PG_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)
