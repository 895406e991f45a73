// binning.cu — tile binning after the depth sort (re-designed):
//   emit_kernel      : fused scan of the stored pairs per Gaussian (decoupled look-back over chunks of
//                      512 Gaussians, in DEPTH order) + key duplication, one thread per rectangle ROW:
//                      writes (tile id, Gaussian index) pairs.  12 B read per Gaussian, 8 B written per pair.
//                      also accumulates the digit histograms of both tile-sort passes (shared-memory
//                      reductions: per pair for the low digit, per run for the high digit).
//   tile_scan_kernel : exclusive digit bases of both tile-sort passes.
//   tile_order_kernel: normalises the tile ranges the last sort pass reduced (identifyTileRanges happens
//                      inside that pass) and orders the tiles by descending list length for compositing.
//   export_keys_kernel (tests only): rebuilds the reference's 64-bit tile|depth keys.
#include "pg_common.cuh"
#include "tile_cull.h"

namespace pg {

constexpr uint32_t E_FLAG_AGG = 1u << 30;
constexpr uint32_t E_FLAG_INCL = 2u << 30;
constexpr uint32_t E_VAL_MASK = (1u << 30) - 1;

#ifndef PG_EMIT_THREADS
#define PG_EMIT_THREADS 256
#endif
constexpr int EMIT_THREADS = PG_EMIT_THREADS;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;
constexpr int EMIT_EPT = EMIT_CHUNK / EMIT_THREADS;  // entries per thread (blocked = depth order)
constexpr int EMIT_RPT = 14;                         // rows per thread in the row scan
constexpr int EMIT_ROWCAP = EMIT_THREADS * EMIT_RPT; // tile rows of one window (3584)
constexpr uint32_t EMIT_OK = 0x80000000u;            // s_g bit 31: the row test may cull (CullGauss::ok)

struct EmitSmem {
    // per entry of the chunk (depth order)
    float4 cga[EMIT_CHUNK];            // CullGauss: gx, gy, qb, thr2qa
    float4 cgb[EMIT_CHUNK];            //            det_lo, inv_qa, umax, k
    ushort4 rect[EMIT_CHUNK];
    uint32_t g[EMIT_CHUNK];            // Gaussian index | EMIT_OK
    uint32_t rowbase[EMIT_CHUNK + 1];  // exclusive scan of the rectangles' row counts
    // per tile row of the current window
    uint32_t rowinfo[EMIT_ROWCAP];     // ta | tb << 11 | entry << 22   (run [ta, tb) of this row; gx <= 2047)
    uint32_t rowoff[EMIT_ROWCAP + 1];  // exclusive scan of the rows' stored pairs (window-relative)
    uint32_t scan[EMIT_WARPS];
    uint32_t chunk, base;
    uint32_t hist[2][RADIX];           // digit histograms of this chunk's stored pairs (both tile-sort passes)
    uint8_t widx[EMIT_WARPS][32];      // write phase: rank among a warp's non-empty rows -> lane
};
static_assert(EMIT_CHUNK <= 1024 && EMIT_CHUNK % EMIT_THREADS == 0, "entry id is packed into 10 bits");

// One stored pair.  flag: the tile cannot receive a contribution (KEEP_ALL lists only).
__device__ __forceinline__ void emit_pair(bool valid, uint32_t dst, uint32_t tile, uint32_t g, bool flag,
                                          uint32_t n_env, uint32_t* __restrict__ tkeys,
                                          uint32_t* __restrict__ tvals, uint32_t* __restrict__ tile_obj_count,
                                          uint32_t* __restrict__ hist_lo, uint32_t mask_lo) {
    if (valid) {
        // low digit of the tile sort: the tiles of a run are consecutive, so the lanes of a warp hit distinct
        // counters (shared-memory reduction without return value)
        atomicAdd(&hist_lo[tile & mask_lo], 1u);
        tkeys[dst] = tile;
        tvals[dst] = flag ? (g | PG_CULL_FLAG) : g;
        if (g >= n_env && !flag) atomicAdd(&tile_obj_count[tile], 1u);
    }
}

// CTA-wide exclusive scan of one value per thread (EMIT_THREADS threads); returns the exclusive prefix, *total = sum.
__device__ __forceinline__ uint32_t cta_excl_scan(uint32_t v, uint32_t* s_scan, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // previous users of s_scan are done
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    uint32_t wb = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < EMIT_WARPS; ++w) {
        const uint32_t c = s_scan[w];
        if (w < warp) wb += c;
        tot += c;
    }
    *total = tot;
    return wb + x - v;
}

// KEEP_ALL = false (default): only (tile, Gaussian) pairs that can contribute are stored — per tile a
//            subsequence, in the same order, of the reference's list.
// KEEP_ALL = true : every tile of every rectangle is stored exactly as the reference's duplicateWithKeys
//            does (pairs that cannot contribute carry PG_CULL_FLAG): pg_export_binning / n_contrib parity.
//
// The unit of work is the TILE ROW of a rectangle, not the Gaussian: rectangles span 1..gy rows, and a
// thread that walks a tall one alone stalls its warp.  Per chunk of EMIT_CHUNK depth-ordered entries:
//   (0) gather rect + record, derive the row test's constants (tile_cull.h) once per entry;
//   (1) scan the row counts -> every (entry, row) gets a slot in a flat row list (windows of
//       EMIT_ROWCAP rows; one window is the common case and its runs stay cached in shared memory);
//   (2) one thread per row: closed-form run [ta, tb) of tiles that can contribute; scan of run lengths;
//   (3) decoupled look-back over chunks -> output offset of the chunk;
//   (4) every warp writes the runs of its 32 rows pair-parallel (slot j of the concatenated runs -> lane
//       j % 32): coalesced stores, no lane walks a long run alone.
// Output order = entry order (depth), rows top to bottom, tiles left to right = the reference's order.
template <bool KEEP_ALL>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(const uint32_t* __restrict__ sorted_dkey, const uint32_t* __restrict__ perm,
            const ushort4* __restrict__ rects, const GeomRec* __restrict__ recs, uint32_t P, uint32_t gx,
            int W, int H, uint32_t* __restrict__ tkeys,
            uint32_t* __restrict__ tvals, uint32_t R_cap, uint32_t* __restrict__ status,
            uint32_t n_env, uint32_t* __restrict__ tile_obj_count, Counters* __restrict__ counters,
            Sticky* __restrict__ sticky, int bits_lo, uint32_t* __restrict__ hist /*[2][RADIX]*/) {
    extern __shared__ __align__(16) unsigned char emit_smem_raw[];
    EmitSmem& sm = *reinterpret_cast<EmitSmem*>(emit_smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask_lo = (1u << bits_lo) - 1u;
    if (tid == 0) sm.chunk = atomicAdd(&counters->tile_counter[4], 1u);
    for (int i = tid; i < 2 * RADIX; i += EMIT_THREADS) (&sm.hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t chunk = sm.chunk;
    // Culled Gaussians carry the largest depth key and sort behind every visible one: only the first
    // ceil(num_visible / EMIT_CHUNK) chunks hold rectangles (num_visible is final: preprocess has finished).
    const uint32_t n_vis = min(counters->num_visible, P);
    const uint32_t num_chunks = max((n_vis + EMIT_CHUNK - 1) / EMIT_CHUNK, 1u);
    if (chunk >= num_chunks) return;

    // ---- (0) gather: thread owns EMIT_EPT consecutive sorted positions (blocked, the scan order)
    uint32_t my_rows = 0;
    unsigned long long local_full = 0;
#pragma unroll
    for (int j = 0; j < EMIT_EPT; ++j) {
        const int q = tid * EMIT_EPT + j;
        const uint32_t spos = chunk * EMIT_CHUNK + q;
        uint32_t g = 0;
        ushort4 r = make_ushort4(0, 0, 0, 0);
        if (spos < P && sorted_dkey[spos] != 0xFFFFFFFFu) {
            g = perm[spos];
            r = rects[g];
            const float4 ra = recs[g].a, rb = recs[g].b;
            const CullGauss cg = cull_setup(ra.x, ra.y, ra.z, ra.w, rb.x, rb.w);
            sm.cga[q] = make_float4(cg.gx, cg.gy, cg.qb, cg.thr2qa);
            sm.cgb[q] = make_float4(cg.det_lo, cg.inv_qa, cg.umax, cg.k);
            if (cg.ok) g |= EMIT_OK;
        }
        sm.g[q] = g;
        sm.rect[q] = r;
        const uint32_t nr = (r.z > r.x) ? (uint32_t)(r.w - r.y) : 0u;
        my_rows += nr;
        local_full += (unsigned long long)nr * (uint32_t)(r.z - r.x);
    }
    // the reference's R = sum of all rectangle areas (what pg_status.num_rendered reports)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_full += __shfl_xor_sync(0xffffffffu, local_full, o);
    if (lane == 0 && local_full) atomicAdd(&counters->rendered_full, local_full);

    // ---- (1) row slots
    uint32_t total_rows;
    {
        uint32_t rb = cta_excl_scan(my_rows, sm.scan, &total_rows);
#pragma unroll
        for (int j = 0; j < EMIT_EPT; ++j) {
            const int q = tid * EMIT_EPT + j;
            const ushort4 r = sm.rect[q];
            sm.rowbase[q] = rb;
            rb += (r.z > r.x) ? (uint32_t)(r.w - r.y) : 0u;
        }
        if (tid == EMIT_THREADS - 1) sm.rowbase[EMIT_CHUNK] = rb;
    }
    __syncthreads();
    const uint32_t nwin = (total_rows + EMIT_ROWCAP - 1) / EMIT_ROWCAP;

    // rows [w0, w0 + tw) -> sm.rowinfo / sm.rowoff; returns the window's stored pairs
    auto compute_window = [&](uint32_t w0, uint32_t tw) -> uint32_t {
        // (a) expansion: every entry writes its id into the slots of its rows that fall into the window
#pragma unroll
        for (int j = 0; j < EMIT_EPT; ++j) {
            const int q = tid * EMIT_EPT + j;
            const uint32_t b = sm.rowbase[q], e = sm.rowbase[q + 1];
            const uint32_t lo = max(b, w0), hi = min(e, w0 + tw);
            for (uint32_t t = lo; t < hi; ++t) sm.rowinfo[t - w0] = (uint32_t)q;
        }
        __syncthreads();
        // (b) one thread per row (adjacent lanes = adjacent rows)
        for (uint32_t t = tid; t < tw; t += EMIT_THREADS) {
            const uint32_t q = sm.rowinfo[t];
            const ushort4 r = sm.rect[q];
            const int ty = (int)r.y + (int)(w0 + t - sm.rowbase[q]);
            const float4 ca = sm.cga[q], cb = sm.cgb[q];
            CullGauss cg;
            cg.gx = ca.x; cg.gy = ca.y; cg.qb = ca.z; cg.thr2qa = ca.w;
            cg.det_lo = cb.x; cg.inv_qa = cb.y; cg.umax = cb.z; cg.k = cb.w;
            cg.qa = 0.0f; cg.qc = 0.0f;  // not used by the row test
            cg.ok = (sm.g[q] & EMIT_OK) != 0;
            int ta, tb;
            const int c = cull_row_run(cg, ty, r.x, r.z, W, H, &ta, &tb);
            sm.rowinfo[t] = (uint32_t)ta | ((uint32_t)tb << 11) | (q << 22);
            sm.rowoff[t] = KEEP_ALL ? (uint32_t)(r.z - r.x) : (uint32_t)c;
        }
        __syncthreads();
        // (c) exclusive scan of the run lengths, EMIT_RPT consecutive rows per thread
        uint32_t len[EMIT_RPT], sum = 0;
#pragma unroll
        for (int i = 0; i < EMIT_RPT; ++i) {
            const uint32_t t = tid * EMIT_RPT + i;
            len[i] = t < tw ? sm.rowoff[t] : 0u;
            sum += len[i];
        }
        uint32_t win_total;
        uint32_t o = cta_excl_scan(sum, sm.scan, &win_total);
#pragma unroll
        for (int i = 0; i < EMIT_RPT; ++i) {
            const uint32_t t = tid * EMIT_RPT + i;
            if (t < tw) sm.rowoff[t] = o;
            o += len[i];
        }
        if (tid == 0) sm.rowoff[tw] = win_total;
        __syncthreads();
        return win_total;
    };

    // ---- (2) sweep 1: count
    uint32_t total = 0;
    {
        uint64_t t64 = 0;
        for (uint32_t w = 0; w < nwin; ++w) {
            const uint32_t w0 = w * EMIT_ROWCAP;
            t64 += compute_window(w0, min((uint32_t)EMIT_ROWCAP, total_rows - w0));
        }
        total = (uint32_t)min(t64, (uint64_t)E_VAL_MASK);
    }
    // ---- (3) chunk-level decoupled look-back (single value): warp 0 inspects 32 predecessors per step
    if (warp == 0) {
        volatile uint32_t* st = status + chunk;
        uint32_t prev = 0;
        const uint32_t tot_c = total;
        if (chunk == 0) {
            if (lane == 0) *st = tot_c | E_FLAG_INCL;
        } else {
            if (lane == 0) *st = tot_c | E_FLAG_AGG;
            int t = (int)chunk - 1;  // lane l looks at chunk t - l
            while (true) {
                const int mine = t - lane;
                const uint32_t sv = mine >= 0 ? *(volatile uint32_t*)(status + mine) : (2u << 30);
                const uint32_t f = sv >> 30;
                const uint32_t not_ready = __ballot_sync(0xffffffffu, f == 0);
                const uint32_t incl = __ballot_sync(0xffffffffu, f == 2);
                // usable prefix of the window: lanes below the first not-ready one, up to the first inclusive one
                const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
                const int first_in = incl ? __ffs(incl) - 1 : 32;
                const int take = min(first_nr, first_in + 1);  // lanes [0, take)
                uint32_t v = lane < take ? (sv & E_VAL_MASK) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                prev = min(prev + v, E_VAL_MASK);
                if (first_in < first_nr) break;
                t -= take;
            }
            if (lane == 0) *st = min(prev + tot_c, E_VAL_MASK) | E_FLAG_INCL;
        }
        if (lane == 0) {
            sm.base = prev;
            if (chunk == num_chunks - 1) {
                uint64_t R = (uint64_t)prev + total;
                counters->sort_n = (uint32_t)min(R, (uint64_t)R_cap);
                if (R > R_cap) {
                    counters->overflow = 1;
                    atomicAdd(&sticky->overflow_frames, 1u);
                }
                atomicMax(&sticky->max_pairs_needed, (uint32_t)min(R, (uint64_t)E_VAL_MASK + 1u));
            }
        }
    }
    __syncthreads();
    // ---- (4) sweep 2: write
    uint32_t off = sm.base;
    for (uint32_t w = 0; w < nwin; ++w) {
        const uint32_t w0 = w * EMIT_ROWCAP;
        const uint32_t tw = min((uint32_t)EMIT_ROWCAP, total_rows - w0);
        uint32_t win_total;
        if (nwin > 1) win_total = compute_window(w0, tw);  // a single window is still cached
        else win_total = sm.rowoff[tw];
        for (uint32_t t0 = 0; t0 < tw; t0 += EMIT_THREADS) {
            const uint32_t t = t0 + tid;
            uint32_t info = 0, o0 = 0, len = 0, g = 0;
            int ty = 0, x0 = 0;
            if (t < tw) {
                info = sm.rowinfo[t];
                o0 = sm.rowoff[t];
                len = sm.rowoff[t + 1] - o0;
                const uint32_t q = info >> 22;
                const ushort4 r = sm.rect[q];
                g = sm.g[q] & ~EMIT_OK;
                ty = (int)r.y + (int)(w0 + t - sm.rowbase[q]);
                x0 = KEEP_ALL ? (int)r.x : (int)(info & 2047u);
            }
            const int ta = (int)(info & 2047u), tb = (int)((info >> 11) & 2047u);
            const uint32_t dst0 = off + o0;                    // may wrap only beyond R_cap (checked below)
            const uint32_t room = dst0 < R_cap ? R_cap - dst0 : 0u;
            const uint32_t tile0 = (uint32_t)ty * gx + (uint32_t)x0;
            // high digit of the tile sort: once per run (a run crosses a digit boundary at most every
            // 2^bits_lo tiles); only pairs that are really stored count
            for (uint32_t tcur = tile0, left = min(len, room); left > 0;) {
                const uint32_t n1 = min(left, (((tcur >> bits_lo) + 1u) << bits_lo) - tcur);
                atomicAdd(&sm.hist[1][(tcur >> bits_lo) & 255u], n1);
                tcur += n1;
                left -= n1;
            }
            // The warp writes the pairs of its 32 rows PAIR-parallel: slot j of the group's concatenated output
            // goes to lane j % 32, so stores are fully coalesced and no lane walks a long run alone.  Row of a
            // slot: non-empty rows have distinct start offsets s; per batch of 32 slots the heads falling into
            // it form a bit mask (one warp reduction), a popcount gives every slot the rank of its row among
            // the non-empty rows, and widx maps the rank back to the lane that holds the row.
            const uint32_t NE = __ballot_sync(0xffffffffu, len > 0);
            if (NE == 0) continue;  // warp-uniform
            const uint32_t gbase = __shfl_sync(0xffffffffu, o0, 0);  // NE != 0: the warp's first row exists
            const uint32_t L = __reduce_add_sync(0xffffffffu, len);
            const uint32_t srow = o0 - gbase;                        // start slot of this lane's row
            const uint32_t tms = tile0 - srow;                       // tile of slot j in this row = tms + j
            const int xms = x0 - (int)srow;
            if (len > 0) sm.widx[warp][__popc(NE & ((1u << lane) - 1u))] = (uint8_t)lane;
            __syncwarp();
            const uint32_t gdst = off + gbase;
            const uint32_t groom = gdst < R_cap ? R_cap - gdst : 0u;
            uint32_t base_rank = 0;  // non-empty rows that start before the current batch
            for (uint32_t jb = 0; jb < L; jb += 32) {
                const uint32_t rel = srow - jb;  // wraps for rows that started earlier
                const uint32_t hm = __reduce_or_sync(0xffffffffu, (len > 0 && rel < 32u) ? (1u << rel) : 0u);
                const uint32_t j = jb + lane;
                const uint32_t rank = base_rank + __popc(hm & (0xFFFFFFFFu >> (31 - lane))) - 1u;
                const int src = sm.widx[warp][rank & 31u];
                const uint32_t p_tms = __shfl_sync(0xffffffffu, tms, src);
                const uint32_t p_g = __shfl_sync(0xffffffffu, g, src);
                bool flag = false;
                if (KEEP_ALL) {
                    const int x = __shfl_sync(0xffffffffu, xms, src) + (int)j;
                    const int p_ta = __shfl_sync(0xffffffffu, ta, src), p_tb = __shfl_sync(0xffffffffu, tb, src);
                    flag = !(x >= p_ta && x < p_tb);
                }
                emit_pair(j < L && j < groom, gdst + j, p_tms + j, p_g, flag, n_env, tkeys, tvals, tile_obj_count,
                          sm.hist[0], mask_lo);
                base_rank += __popc(hm);
            }
            __syncwarp();  // widx is rewritten by the next group
        }
        off = (uint32_t)min((uint64_t)off + win_total, (uint64_t)0xFFFFFFFFu);
        if (nwin > 1) __syncthreads();  // the next window overwrites rowinfo / rowoff
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

int launch_emit(bool keep_all, const uint32_t* sorted_dkey, const uint32_t* perm, const ushort4* rects, const GeomRec* recs,
                uint32_t P, uint32_t gx, int W, int H, uint32_t* tkeys, uint32_t* tvals, uint32_t R_cap, uint32_t* status,
                uint32_t n_env, uint32_t* tile_obj_count, Counters* counters, Sticky* sticky, int bits_lo,
                uint32_t* hist_tile, cudaStream_t stream) {
    uint32_t chunks = (P + EMIT_CHUNK - 1) / EMIT_CHUNK;
    if (chunks == 0) return PG_OK;
    if (keep_all) {
        PG_CUDA_CHECK(ensure_dynamic_smem(emit_kernel<true>, (int)sizeof(EmitSmem)));
        emit_kernel<true><<<chunks, EMIT_THREADS, sizeof(EmitSmem), stream>>>(sorted_dkey, perm, rects, recs, P, gx, W, H, tkeys, tvals, R_cap, status,
                                                      n_env, tile_obj_count, counters, sticky, bits_lo, hist_tile);
    } else {
        PG_CUDA_CHECK(ensure_dynamic_smem(emit_kernel<false>, (int)sizeof(EmitSmem)));
        emit_kernel<false><<<chunks, EMIT_THREADS, sizeof(EmitSmem), stream>>>(sorted_dkey, perm, rects, recs, P, gx, W, H, tkeys, tvals, R_cap, status,
                                                       n_env, tile_obj_count, counters, sticky, bits_lo, hist_tile);
    }
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
