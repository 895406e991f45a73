// binning.cu — tile binning after the depth sort (re-designed as count -> scan -> write):
//   count_kernel     : the only gather of the stage (one 48-byte record per visible Gaussian, in depth order); one lane
//                      per rectangle ROW computes the run of tiles that can contribute and leaves it, with the
//                      round / group / CTA pair counts, in depth-ordered tables.  No dependency between CTAs.
//   pair_scan_kernel : one CTA scans the per-CTA totals; finalises the pair count, the overflow flag, sticky status.
//   emit_kernel      : key duplication in DEPTH order as a pure writer: every warp knows its output offset on entry,
//                      reads the runs back coalesced and writes (tile id, Gaussian index) pairs pair-parallel: 8 B
//                      written per pair.  Also accumulates the digit histograms of both tile-sort passes
//                      (shared-memory reductions: per pair for the low digit, per run for the high digit).
//   tile_scan_kernel : exclusive digit bases of both tile-sort passes.
//   tile_order_kernel: normalises the tile ranges the last sort pass reduced (identifyTileRanges happens
//                      inside that pass) and orders the tiles by descending list length for compositing.
//   export_keys_kernel (tests only): rebuilds the reference's 64-bit tile|depth keys.
#include "pg_common.cuh"
#include "tile_cull.h"

namespace pg {

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;

struct EmitWarpSmem {  // count_scan_kernel: one warp's 32 depth-ordered Gaussians
    float4 cga[32];            // CullGauss: gx, gy, qb, thr2qa
    float4 cgb[32];            //            det_lo, inv_qa, umax, k
    ushort4 rect[32];
    uint32_t g[32];            // Gaussian index | EMIT_OK
    uint32_t rowbase[33];      // exclusive scan of the rectangles' row counts
};
struct EmitWriteSmem {  // emit_kernel: one group of 32 depth-ordered Gaussians
    ushort4 rect[32];
    uint32_t g[32];            // Gaussian index
    uint32_t ovf[32];          // where the rows beyond RUN_FIX of a tall rectangle live in runs_ovf
    uint32_t rowbase[33];      // exclusive scan of the rectangles' row counts
};
struct EmitSmem {
    EmitWriteSmem w[EMIT_WARPS];   // the CTA's groups
    uint32_t rounds[EMIT_WARPS + 1];  // exclusive scan of the groups' round counts
    uint32_t gbase[EMIT_WARPS];    // output offset of each group
    uint32_t next;                 // round ticket
    uint8_t widx[EMIT_WARPS][32];  // per warp, write phase: rank among the non-empty rows of a round -> lane
    uint32_t hist[2][RADIX];       // digit histograms of this CTA's stored pairs (both tile-sort passes)
};
constexpr uint32_t EMIT_OK = 0x80000000u;  // g bit 31: the row test may cull (CullGauss::ok)

// One stored pair.  flag: the tile cannot receive a contribution (KEEP_ALL lists only).
__device__ __forceinline__ void emit_pair(bool valid, uint32_t dst, uint32_t tile, uint32_t g, bool flag,
                                          uint32_t* __restrict__ tkeys, uint32_t* __restrict__ tvals,
                                          uint32_t* __restrict__ hist_lo, uint32_t mask_lo) {
    if (valid) {
        // low digit of the tile sort: the tiles of a run are consecutive, so the lanes of a warp hit distinct
        // counters (shared-memory reduction without return value)
        atomicAdd(&hist_lo[tile & mask_lo], 1u);
        tkeys[dst] = tile;
        tvals[dst] = flag ? (g | PG_CULL_FLAG) : g;
    }
}

// (0) of count_kernel: lane = one of the warp's 32 consecutive depth-ordered Gaussians.  Stages the row
// test's constants, the rectangle and the index in the warp's shared memory, scans the rectangles' row counts
// (rowbase[0..32]) and returns the warp's total number of tile rows.
__device__ __forceinline__ uint32_t emit_gather(EmitWarpSmem& ws, const uint32_t* __restrict__ perm,
                                                const GeomRec* __restrict__ recs, ushort4* __restrict__ srect,
                                                uint32_t p, uint32_t n_vis, int gx, int gy, int lane) {
    uint32_t g = 0, rows = 0;
    ushort4 r = make_ushort4(0, 0, 0, 0);
    if (p < n_vis) {
        g = perm[p];
        // ONE gather per Gaussian: the 48-byte record carries everything the binning needs (the tile rectangle
        // follows from the pixel centre and the integer radius in c.w exactly as preprocess derived it)
        const float4 ra = recs[g].a, rb = recs[g].b;
        const int radius = __float_as_int(recs[g].c.w);
        r = tile_rect(ra.x, ra.y, radius, gx, gy);
        srect[p] = r;  // in depth order for the emit kernel (coalesced)
        const CullGauss cg = cull_setup(ra.x, ra.y, ra.z, ra.w, rb.x, rb.w);
        ws.cga[lane] = make_float4(cg.gx, cg.gy, cg.qb, cg.thr2qa);
        ws.cgb[lane] = make_float4(cg.det_lo, cg.inv_qa, cg.umax, cg.k);
        if (cg.ok) g |= EMIT_OK;
        rows = (r.z > r.x) ? (uint32_t)(r.w - r.y) : 0u;
    }
    ws.g[lane] = g;
    ws.rect[lane] = r;
    uint32_t x = rows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    ws.rowbase[lane] = x - rows;
    if (lane == 31) ws.rowbase[32] = x;
    __syncwarp();
    return __shfl_sync(0xffffffffu, x, 31);
}

// (1) of count_kernel: tile row number t of the warp's 32 rectangles (t < total rows): which Gaussian q
// (largest q with rowbase[q] <= t), which tile row ty, and its run [ta, tb) of tiles that can contribute.
// Returns the number of pairs stored for the row.
template <bool KEEP_ALL>
__device__ __forceinline__ uint32_t emit_row(const EmitWarpSmem& ws, uint32_t t, int W, int H, int& q, int& ty,
                                             int& ta, int& tb, int& x0, uint32_t& gi) {
    q = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
        if (ws.rowbase[q + step] <= t) q += step;
    const ushort4 rr = ws.rect[q];
    ty = (int)rr.y + (int)(t - ws.rowbase[q]);
    const float4 ca = ws.cga[q], cb = ws.cgb[q];
    CullGauss cg;
    cg.gx = ca.x; cg.gy = ca.y; cg.qb = ca.z; cg.thr2qa = ca.w;
    cg.det_lo = cb.x; cg.inv_qa = cb.y; cg.umax = cb.z; cg.k = cb.w;
    cg.qa = 0.0f; cg.qc = 0.0f;  // not used by the row test
    const uint32_t gq = ws.g[q];
    cg.ok = (gq & EMIT_OK) != 0;
    gi = gq & ~EMIT_OK;
    const int c = cull_row_run(cg, ty, rr.x, rr.z, W, H, &ta, &tb);
    x0 = KEEP_ALL ? (int)rr.x : ta;
    return KEEP_ALL ? (uint32_t)(rr.z - rr.x) : (uint32_t)c;
}

// count_kernel: the tile-row runs of every visible Gaussian, computed ONCE, row-parallel, in DEPTH order, and kept
// for the emit kernel.  A CTA owns 8 GROUPS of 32 consecutive sorted positions; their rounds (32 tile rows each) are
// taken by whichever warp is free; no dependency between CTAs (no look-back).  This is the only kernel of the
// binning stage that gathers: one 48-byte record per Gaussian; everything it leaves behind is in depth order, so
// the emit kernel reads coalesced:
//   srect[p]           = tile rectangle of sorted position p;
//   runs_fix[p][i]     = ta | tb << 11 of row i < RUN_FIX of sorted position p (run [ta, tb) of tiles that can
//                        contribute; gx <= 2047);
//   runs_ovf[b + i - RUN_FIX] for the rows beyond RUN_FIX of tall rectangles, b = ovf_base[p] allocated with one
//                        global atomic per tall rectangle (no order needed; capacity = pair capacity);
//   rnd_off[group][k]  = stored pairs of the group's rounds before round k (rounds per group <= gy);
//   grp_loc[group]     = stored pairs of the CTA's groups before this one;
//   cta_pairs[cta]     = stored pairs of the CTA  -> pair_scan_kernel -> cta_base[cta].
template <bool KEEP_ALL>
// bounded for 6 CTAs per SM: count + emit 0.208 vs 0.219 ms unbounded (46 / 29 registers), 0.214 at 8
__global__ void __launch_bounds__(EMIT_THREADS, 6)
count_kernel(const uint32_t* __restrict__ perm, ushort4* __restrict__ srect,
             const GeomRec* __restrict__ recs, uint32_t P, int W, int H, int gx, int gy,
             uint32_t* __restrict__ runs_fix, uint32_t* __restrict__ runs_ovf, uint32_t* __restrict__ ovf_base,
             uint32_t R_cap, uint32_t* __restrict__ rnd_off, uint32_t* __restrict__ grp_loc,
             uint32_t* __restrict__ cta_pairs, Counters* __restrict__ counters) {
    __shared__ EmitWarpSmem sw[EMIT_WARPS];
    __shared__ uint32_t s_ovf[EMIT_WARPS][32];
    __shared__ uint32_t s_rounds[EMIT_WARPS + 1];  // exclusive scan of the groups' round counts
    __shared__ uint32_t s_tall[EMIT_WARPS], s_tall_base, s_gtot[EMIT_WARPS];
    __shared__ uint32_t s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_vis = min(counters->num_visible, P);
    const uint32_t num_groups = (n_vis + 31u) / 32u;
    const uint32_t grp0 = blockIdx.x * EMIT_WARPS;
    if (grp0 >= num_groups) return;  // CTA-uniform
    // ---- (0) warp w stages group grp0 + w
    uint32_t my_rows = 0, tall = 0, tall_before = 0;  // tall: rows beyond RUN_FIX of this lane's rectangle
    const uint32_t p = (grp0 + warp) * 32u + lane;
    {
        EmitWarpSmem& ws = sw[warp];
        my_rows = emit_gather(ws, perm, recs, srect, p, n_vis, gx, gy, lane);
        const uint32_t rows = ws.rowbase[lane + 1] - ws.rowbase[lane];
        tall = rows > RUN_FIX ? rows - RUN_FIX : 0u;
        uint32_t x = tall;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        tall_before = x - tall;
        if (lane == 31) s_tall[warp] = x;
    }
    if (lane == 0) s_rounds[warp + 1] = (my_rows + 31u) / 32u;
    if (tid == 0) { s_rounds[0] = 0; s_next = 0; }
    __syncthreads();
    if (tid == 0) {
        uint32_t t_all = 0;
        for (int w = 0; w < EMIT_WARPS; ++w) {
            s_rounds[w + 1] += s_rounds[w];
            const uint32_t c = s_tall[w];
            s_tall[w] = t_all;
            t_all += c;
        }
        // tall rectangles take space for their rows beyond RUN_FIX: ONE allocation per CTA (an atomic per rectangle
        // on one word serialises in L2)
        s_tall_base = t_all ? atomicAdd(&counters->run_ovf, t_all) : 0u;
    }
    __syncthreads();
    {
        const uint32_t b = s_tall_base + s_tall[warp] + tall_before;
        s_ovf[warp][lane] = b;
        if (p < n_vis) ovf_base[p] = b;
    }
    __syncthreads();
    // ---- (1) the CTA's rounds (32 tile rows of one group each) are taken by whichever warp is free: a group of the
    // nearest Gaussians has ~60 rounds, most groups have ~5
    const uint32_t total_rounds = s_rounds[EMIT_WARPS];
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&s_next, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= total_rounds) break;
        int w = 0;
#pragma unroll
        for (int i = 1; i < EMIT_WARPS; ++i) w += (s_rounds[i] <= u) ? 1 : 0;
        const uint32_t k = u - s_rounds[w];
        const EmitWarpSmem& ws = sw[w];
        const uint32_t total_rows = ws.rowbase[32];
        const uint32_t t = k * 32u + lane;
        uint32_t len = 0, gi;
        int q = 0, ty, ta = 0, tb = 0, x0;
        if (t < total_rows) {
            len = emit_row<KEEP_ALL>(ws, t, W, H, q, ty, ta, tb, x0, gi);
            const uint32_t i = t - ws.rowbase[q];
            const uint32_t run = (uint32_t)ta | ((uint32_t)tb << 11);
            if (i < RUN_FIX) runs_fix[((size_t)(grp0 + w) * 32u + q) * RUN_FIX + i] = run;
            else {
                const uint32_t slot = s_ovf[w][q] + (i - RUN_FIX);
                if (slot < R_cap) runs_ovf[slot] = run;  // beyond the table: an overflow frame (pair_scan_kernel flags it)
            }
        }
        const uint32_t L = __reduce_add_sync(0xffffffffu, len);
        if (lane == 0) rnd_off[(size_t)(grp0 + w) * gy + k] = L;  // pairs of this round; scanned in place below
    }
    __syncthreads();
    // ---- (2) warp w: exclusive scan of its group's per-round pair counts -> where each round's pairs start
    uint32_t carry = 0;
    if (grp0 + warp < num_groups) {
        const uint32_t nr = s_rounds[warp + 1] - s_rounds[warp];
        uint32_t* ro = rnd_off + (size_t)(grp0 + warp) * gy;
        for (uint32_t k0 = 0; k0 < nr; k0 += 32) {
            const uint32_t k = k0 + lane;
            const uint32_t v = k < nr ? ro[k] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (k < nr) ro[k] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
    if (lane == 0) s_gtot[warp] = carry;  // < 32 * 65536 tiles
    __syncthreads();
    // ---- (3) the groups' offsets inside the CTA, the CTA's pair total (-> pair_scan_kernel -> cta_base)
    if (tid < EMIT_WARPS) {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < EMIT_WARPS; ++w) {
            const uint32_t c = s_gtot[w];
            if (w < tid) before += c;
            total += c;
        }
        if (grp0 + tid < num_groups) grp_loc[grp0 + tid] = before;
        if (tid == 0) cta_pairs[blockIdx.x] = total;  // < 256 * 65536 tiles
    }
}

// pair_scan_kernel: exclusive scan of the per-CTA pair counts of count_kernel (one CTA; at most P / 256 values, 1024 per
// sweep), and the frame's pair count: sort_n, overflow, sticky status.  cta_base saturates at 2^32 - 1 (nothing is
// ever written at or beyond the pair capacity <= 2^30).
__global__ void __launch_bounds__(1024)
pair_scan_kernel(const uint32_t* __restrict__ cta_pairs, uint32_t* __restrict__ cta_base, uint32_t P, uint32_t R_cap,
                 Counters* __restrict__ counters, Sticky* __restrict__ sticky) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_vis = min(counters->num_visible, P);
    const uint32_t n = (n_vis + EMIT_THREADS - 1) / EMIT_THREADS;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 1024) {
        const uint32_t i = i0 + tid;
        const unsigned long long v = i < n ? cta_pairs[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = s_warp[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += y;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long excl = s_carry + s_warp[warp] + x - v;
        if (i < n) cta_base[i] = (uint32_t)min(excl, 0xFFFFFFFFull);
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned long long R = s_carry;
        const unsigned long long ovf = counters->run_ovf;  // rows of tall rectangles beyond RUN_FIX: same capacity
        counters->sort_n = (uint32_t)min(R, (unsigned long long)R_cap);
        if (R > R_cap || ovf > R_cap) {
            counters->overflow = 1;
            atomicAdd(&sticky->overflow_frames, 1u);
        }
        atomicMax(&sticky->max_pairs_needed, (uint32_t)min(max(R, ovf), 1ull << 30));
    }
}

// KEEP_ALL = false (default): only (tile, Gaussian) pairs that can contribute are stored — per tile a
//            subsequence, in the same order, of the reference's list.
// KEEP_ALL = true : every tile of every rectangle is stored exactly as the reference's duplicateWithKeys
//            does (pairs that cannot contribute carry PG_CULL_FLAG): pg_export_binning / n_contrib parity.
//
// A pure writer: count_kernel computed every tile row's run, pair_scan_kernel the frame-wide offsets, so every WARP knows
// where its output starts and works on its own — no look-back, no row test, no CTA-wide barrier between set-up and
// the final histogram flush.  Per warp, 32 consecutive depth-ordered Gaussians:
//   (0) gather index, rectangle and offsets; scan the rectangles' row counts;
//   (1) the unit of work is the TILE ROW of a rectangle (rectangles span 1..gy rows): rows are numbered across
//       the 32 rectangles and taken 32 at a time, lane = row, its run [ta, tb) read back from the run tables;
//   (2) the runs of a group of 32 rows are written pair-parallel (slot j of the concatenated runs -> lane
//       j % 32): coalesced stores, no lane walks a long run alone.
// Output order = depth order, rows top to bottom, tiles left to right = the reference's order.
template <bool KEEP_ALL>
// bounded for 6 CTAs per SM: count + emit 0.208 vs 0.219 ms unbounded (46 / 29 registers), 0.214 at 8
__global__ void __launch_bounds__(EMIT_THREADS, 6)
emit_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ cta_base, const uint32_t* __restrict__ grp_loc,
            const uint32_t* __restrict__ rnd_off,
            const uint32_t* __restrict__ runs_fix, const uint32_t* __restrict__ runs_ovf,
            const uint32_t* __restrict__ ovf_base, const ushort4* __restrict__ srect, uint32_t P, uint32_t gx, int gy,
            uint32_t* __restrict__ tkeys, uint32_t* __restrict__ tvals, uint32_t R_cap,
            const Counters* __restrict__ counters, int bits_lo, uint32_t* __restrict__ hist /*[2][RADIX]*/) {
    __shared__ EmitSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask_lo = (1u << bits_lo) - 1u;
    for (int i = tid; i < 2 * RADIX; i += EMIT_THREADS) (&sm.hist[0][0])[i] = 0;
    // only the first num_visible sorted positions exist (the depth sort moved the visible Gaussians only)
    const uint32_t n_vis = min(counters->num_visible, P);
    const uint32_t num_groups = (n_vis + 31u) / 32u;
    const uint32_t grp0 = blockIdx.x * EMIT_WARPS;
    if (grp0 >= num_groups) return;  // CTA-uniform; nothing to flush
    // ---- (0) warp w stages group grp0 + w: index, rectangle (in depth order: count_kernel left it there), tall-rectangle
    // slot — all coalesced
    {
        EmitWriteSmem& ws = sm.w[warp];
        const uint32_t p = (grp0 + warp) * 32u + lane;
        uint32_t g = 0, rows = 0;
        ushort4 r = make_ushort4(0, 0, 0, 0);
        if (p < n_vis) {
            g = perm[p];
            r = srect[p];
            rows = (r.z > r.x) ? (uint32_t)(r.w - r.y) : 0u;
        }
        ws.g[lane] = g;
        ws.rect[lane] = r;
        ws.ovf[lane] = rows > RUN_FIX ? ovf_base[p] : 0u;
        uint32_t xs = rows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, xs, o);
            if (lane >= o) xs += y;
        }
        ws.rowbase[lane] = xs - rows;
        if (lane == 31) {
            ws.rowbase[32] = xs;
            sm.rounds[warp + 1] = (xs + 31u) / 32u;
            const uint64_t b64 = grp0 + warp < num_groups ? (uint64_t)cta_base[blockIdx.x] + grp_loc[grp0 + warp] : 0ull;
            sm.gbase[warp] = (uint32_t)min(b64, (uint64_t)0xFFFFFFFFu);
        }
    }
    if (tid == 0) { sm.rounds[0] = 0; sm.next = 0; }
    __syncthreads();
    if (tid == 0)
        for (int w = 0; w < EMIT_WARPS; ++w) sm.rounds[w + 1] += sm.rounds[w];
    __syncthreads();
    // ---- the CTA's rounds (32 tile rows of one group each; count_kernel left every round's output offset) are taken
    // by whichever warp is free: a group of the nearest Gaussians has ~60 rounds of up to 32 x gx pairs, most groups
    // have ~5 rounds of a few dozen pairs, and a warp that owned a whole group would keep its CTA waiting
    const uint32_t total_rounds = sm.rounds[EMIT_WARPS];
    uint8_t* const widx = sm.widx[warp];
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&sm.next, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= total_rounds) break;
        int w = 0;
#pragma unroll
        for (int i = 1; i < EMIT_WARPS; ++i) w += (sm.rounds[i] <= u) ? 1 : 0;
        const uint32_t k = u - sm.rounds[w];
        const EmitWriteSmem& ws = sm.w[w];
        const uint32_t total_rows = ws.rowbase[32];
        uint32_t off;
        {
            const uint64_t o64 = (uint64_t)sm.gbase[w] + rnd_off[(size_t)(grp0 + w) * gy + k];
            off = (uint32_t)min(o64, (uint64_t)0xFFFFFFFFu);  // saturated offsets only occur beyond R_cap
        }
        // ---- (1) lane = row t: which Gaussian (largest q with rowbase[q] <= t), which tile row, its run
        const uint32_t t = k * 32u + lane;
        uint32_t len = 0, gi = 0;
        int ta = 0, tb = 0, ty = 0, x0 = 0;
        if (t < total_rows) {
            int q = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (ws.rowbase[q + step] <= t) q += step;
            const ushort4 rr = ws.rect[q];
            const uint32_t i = t - ws.rowbase[q];
            uint32_t run;
            if (i < RUN_FIX) run = runs_fix[((size_t)(grp0 + w) * 32u + q) * RUN_FIX + i];
            else {
                const uint32_t slot = ws.ovf[q] + (i - RUN_FIX);
                run = slot < R_cap ? runs_ovf[slot] : 0u;  // beyond the table: an overflow frame
            }
            ty = (int)rr.y + (int)i;
            gi = ws.g[q];
            ta = (int)(run & 2047u);
            tb = (int)(run >> 11);
            x0 = KEEP_ALL ? (int)rr.x : ta;
            len = KEEP_ALL ? (uint32_t)(rr.z - rr.x) : (uint32_t)(tb - ta);
        }
        // exclusive scan of the run lengths of this round
        uint32_t o0 = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, o0, o);
            if (lane >= o) o0 += y;
        }
        const uint32_t L = __shfl_sync(0xffffffffu, o0, 31);
        o0 -= len;
        if (L == 0) continue;  // warp-uniform
        const uint32_t tile0 = (uint32_t)ty * gx + (uint32_t)x0;
        const uint32_t dst0 = off + o0;
        const uint32_t room = dst0 < R_cap ? R_cap - dst0 : 0u;
        // high digit of the tile sort: once per run (a run crosses a digit boundary at most every
        // 2^bits_lo tiles); only pairs that are really stored count
        for (uint32_t tcur = tile0, left = min(len, room); left > 0;) {
            const uint32_t n1 = min(left, (((tcur >> bits_lo) + 1u) << bits_lo) - tcur);
            atomicAdd(&sm.hist[1][(tcur >> bits_lo) & 255u], n1);
            tcur += n1;
            left -= n1;
        }
        // ---- (2) pair-parallel write.  Row of a slot: non-empty rows have distinct start offsets; per batch of
        // 32 slots the heads falling into it form a bit mask (one warp reduction), a popcount gives every slot
        // the rank of its row among the non-empty rows, and widx maps the rank back to the lane holding the row.
        const uint32_t NE = __ballot_sync(0xffffffffu, len > 0);
        const uint32_t tms = tile0 - o0;  // tile of slot j in this row = tms + j
        const int xms = x0 - (int)o0;
        if (len > 0) widx[__popc(NE & ((1u << lane) - 1u))] = (uint8_t)lane;
        __syncwarp();
        const uint32_t groom = off < R_cap ? R_cap - off : 0u;
        uint32_t base_rank = 0;  // non-empty rows that start before the current batch
        for (uint32_t jb = 0; jb < L; jb += 32) {
            const uint32_t rel = o0 - jb;  // wraps for rows that started earlier
            const uint32_t hm = __reduce_or_sync(0xffffffffu, (len > 0 && rel < 32u) ? (1u << rel) : 0u);
            const uint32_t j = jb + lane;
            const uint32_t rank = base_rank + __popc(hm & (0xFFFFFFFFu >> (31 - lane))) - 1u;
            const int src = widx[rank & 31u];
            const uint32_t p_tms = __shfl_sync(0xffffffffu, tms, src);
            const uint32_t p_g = __shfl_sync(0xffffffffu, gi, src);
            bool flag = false;
            if (KEEP_ALL) {
                const int xx = __shfl_sync(0xffffffffu, xms, src) + (int)j;
                const int p_ta = __shfl_sync(0xffffffffu, ta, src), p_tb = __shfl_sync(0xffffffffu, tb, src);
                flag = !(xx >= p_ta && xx < p_tb);
            }
            emit_pair(j < L && j < groom, off + j, p_tms + j, p_g, flag, tkeys, tvals, sm.hist[0], mask_lo);
            base_rank += __popc(hm);
        }
        __syncwarp();  // widx is rewritten by the warp's next round
    }
    __syncthreads();
    for (int i = tid; i < 2 * RADIX; i += EMIT_THREADS) {
        const uint32_t c = (&sm.hist[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// exclusive scans of the two tile-sort digit histograms -> bases; finalises pg_status.num_rendered.
// One CTA of 512 threads: warps 0..7 scan the low-digit row, warps 8..15 the high-digit row.
__global__ void __launch_bounds__(512)
tile_scan_kernel(const uint32_t* __restrict__ hist /*[2][256]*/, uint32_t* __restrict__ bins /*[2][256]*/,
                 Counters* __restrict__ counters) {
    __shared__ uint32_t s_scan[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const unsigned long long full = counters->rendered_full;  // emit has finished (stream order)
        counters->num_rendered = full > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)full;
    }
    uint32_t v = hist[tid], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    uint32_t wb = 0;
    const int w0 = tid < RADIX ? 0 : 8;
    for (int w = w0; w < warp; ++w) wb += s_scan[w];
    bins[tid] = wb + x - v;
}

// Compositing launch order: tile ids by descending list length (longest-processing-time-first keeps
// the SMs evenly loaded to the end of the kernel; natural order leaves a ~14 % idle tail).  One CTA:
// counting sort on length / 16 (4096 buckets, longer lists clipped into the first bucket); the order
// inside a bucket is irrelevant.
constexpr int ORDER_BUCKETS = 4096;
__global__ void __launch_bounds__(1024)
tile_order_kernel(uint2* __restrict__ ranges, uint32_t tiles, uint32_t* __restrict__ order) {
    __shared__ uint32_t s_cnt[ORDER_BUCKETS];
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < ORDER_BUCKETS; i += 1024) s_cnt[i] = 0;
    __syncthreads();
    for (uint32_t t = tid; t < tiles; t += 1024) {
        // the last sort pass left (~first, last + 1), or (0, 0) for a tile without pairs: normalise to [first, last + 1)
        uint2 r = ranges[t];
        r = r.y == 0 ? make_uint2(0u, 0u) : make_uint2(~r.x, r.y);
        ranges[t] = r;
        const uint32_t b = (ORDER_BUCKETS - 1) - min((r.y - r.x) >> 4, (uint32_t)(ORDER_BUCKETS - 1));
        atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    // exclusive scan of the 4096 bucket counts: 4 consecutive buckets per thread
    uint32_t c[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j] = s_cnt[tid * 4 + j]; sum += c[j]; }
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t base = s_warp[warp] + x - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_cnt[tid * 4 + j] = base; base += c[j]; }
    __syncthreads();
    for (uint32_t t = tid; t < tiles; t += 1024) {
        const uint2 r = ranges[t];
        const uint32_t b = (ORDER_BUCKETS - 1) - min((r.y - r.x) >> 4, (uint32_t)(ORDER_BUCKETS - 1));
        order[atomicAdd(&s_cnt[b], 1u)] = t;
    }
}

// tests only: keys[i] = (tile << 32) | depth_bits[point_list[i]], tile from ranges
__global__ void export_keys_kernel(const uint2* __restrict__ ranges, uint32_t tiles,
                                   const uint32_t* __restrict__ point_list,
                                   const GeomRec* __restrict__ recs, uint64_t* __restrict__ keys,
                                   uint32_t* __restrict__ point_list_out, uint32_t* __restrict__ ranges_out) {
    uint32_t tile = blockIdx.x;
    if (tile >= tiles) return;
    uint2 r = ranges[tile];
    if (threadIdx.x == 0) { ranges_out[2 * tile] = r.x; ranges_out[2 * tile + 1] = r.y; }
    for (uint32_t i = r.x + threadIdx.x; i < r.y; i += blockDim.x) {
        uint32_t g = point_list[i] & ~PG_CULL_FLAG;
        keys[i] = ((uint64_t)tile << 32) | __float_as_uint(recs[g].b.z);
        point_list_out[i] = g;
    }
}

int launch_count(bool keep_all, const uint32_t* perm, ushort4* srect, const GeomRec* recs, uint32_t P, int W, int H,
                 uint32_t* runs_fix, uint32_t* runs_ovf, uint32_t* ovf_base, uint32_t R_cap, uint32_t* rnd_off,
                 uint32_t* grp_loc, uint32_t* cta_pairs, uint32_t* cta_base, Counters* counters, Sticky* sticky,
                 cudaStream_t stream) {
    if (P == 0) return PG_OK;
    const uint32_t blocks = (P + EMIT_THREADS - 1) / EMIT_THREADS;  // sized for P; CTAs beyond num_visible exit at once
    const int gx = (W + PG_TILE - 1) / PG_TILE, gy = (H + PG_TILE - 1) / PG_TILE;
    if (keep_all)
        count_kernel<true><<<blocks, EMIT_THREADS, 0, stream>>>(perm, srect, recs, P, W, H, gx, gy, runs_fix, runs_ovf, ovf_base,
                                                                 R_cap, rnd_off, grp_loc, cta_pairs, counters);
    else
        count_kernel<false><<<blocks, EMIT_THREADS, 0, stream>>>(perm, srect, recs, P, W, H, gx, gy, runs_fix, runs_ovf, ovf_base,
                                                                  R_cap, rnd_off, grp_loc, cta_pairs, counters);
    PG_CUDA_CHECK(cudaGetLastError());
    pair_scan_kernel<<<1, 1024, 0, stream>>>(cta_pairs, cta_base, P, R_cap, counters, sticky);
    count_launch(2);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_emit(bool keep_all, const uint32_t* perm, const uint32_t* cta_base, const uint32_t* grp_loc, const uint32_t* rnd_off,
                const uint32_t* runs_fix, const uint32_t* runs_ovf, const uint32_t* ovf_base, const ushort4* srect, uint32_t P,
                uint32_t gx, int gy, uint32_t* tkeys, uint32_t* tvals, uint32_t R_cap, Counters* counters, int bits_lo,
                uint32_t* hist_tile, cudaStream_t stream) {
    if (P == 0) return PG_OK;
    const uint32_t blocks = (P + EMIT_THREADS - 1) / EMIT_THREADS;  // sized for P; CTAs beyond num_visible exit at once
    if (keep_all)
        emit_kernel<true><<<blocks, EMIT_THREADS, 0, stream>>>(perm, cta_base, grp_loc, rnd_off, runs_fix, runs_ovf, ovf_base, srect, P,
                                                                gx, gy, tkeys, tvals, R_cap, counters, bits_lo, hist_tile);
    else
        emit_kernel<false><<<blocks, EMIT_THREADS, 0, stream>>>(perm, cta_base, grp_loc, rnd_off, runs_fix, runs_ovf, ovf_base, srect, P,
                                                                 gx, gy, tkeys, tvals, R_cap, counters, bits_lo, hist_tile);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_tile_scan(const uint32_t* hist_tile, uint32_t* bins, Counters* counters, cudaStream_t stream) {
    tile_scan_kernel<<<1, 512, 0, stream>>>(hist_tile, bins, counters);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_tile_order(uint2* ranges, uint32_t tiles, uint32_t* order, cudaStream_t stream) {
    tile_order_kernel<<<1, 1024, 0, stream>>>(ranges, tiles, order);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_export_keys(const uint2* ranges, uint32_t tiles, const uint32_t* point_list, const GeomRec* recs,
                       uint64_t* keys, uint32_t* point_list_out, uint32_t* ranges_out, cudaStream_t stream) {
    export_keys_kernel<<<tiles, 128, 0, stream>>>(ranges, tiles, point_list, recs, keys, point_list_out, ranges_out);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

}  // namespace pg
