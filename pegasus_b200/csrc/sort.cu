// sort.cu — hand-written onesweep LSD radix sort (u32 keys, u32 values, 8-bit digits, single pass per
// digit with decoupled look-back) used twice per frame:
//   1. P Gaussians by depth bits (4 passes, values = Gaussian index generated on the fly), and
//   2. R (tile, Gaussian) pairs by tile id (2 passes of ceil(b/2)/floor(b/2) bits, b = bits of the
//      tile count) — the pairs are emitted in depth order, so a stable sort on the tile id alone
//      yields exactly the order the reference gets from sorting 64-bit tile|depth keys.
// HBM-bound: each pass reads 8 B and writes 8 B per item (the last tile pass writes values only).
//
// CTA-tile = 256 threads x 16 items, warp-blocked so that rank order == memory order (stability).
// Ranking: per-warp digit counters in shared memory + ballot multisplit (one vote per digit bit).
// CTA-tiles take tickets from an atomic counter, so a tile only ever waits on tiles that already run.
#include "pg_common.cuh"

namespace pg {

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INCL = 2u << 30;
constexpr uint32_t VAL_MASK = (1u << 30) - 1;

// ---- depth keys of the VISIBLE Gaussians, compacted in index order, + their digit histograms -----------------
// preprocess leaves 0xFFFFFFFF in the depth key of every culled Gaussian (41 % of the bench scene).  One pass over
// the P keys drops them — order preserved, so the stable sort that follows still breaks depth ties by Gaussian
// index as the reference's does — writes (key, index) pairs and accumulates the 4 x 256 digit histograms of what
// is kept.  The four onesweep passes then move num_visible items instead of P.
// CTA-tile = 256 threads x 8 keys, warp-blocked (rank order == memory order); single-value decoupled look-back.
constexpr int SCAN_IPT = SCAN_TILE / 256;

__global__ void __launch_bounds__(256) compact_hist_kernel(const uint32_t* __restrict__ keys_in, uint32_t n,
                                                            uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                            uint32_t* __restrict__ hist /*[4][256]*/,
                                                            uint32_t* __restrict__ status, uint32_t* __restrict__ ticket) {
    __shared__ uint32_t s_h[4][RADIX];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_tile, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < 4 * RADIX; i += 256) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t wbase = tile * SCAN_TILE + warp * (32 * SCAN_IPT) + lane;
    uint32_t keys[SCAN_IPT], before[SCAN_IPT];  // before: kept items of this warp ahead of item i of this lane
    uint32_t wcount = 0;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        const uint32_t idx = wbase + i * 32;
        keys[i] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        const bool keep = keys[i] != 0xFFFFFFFFu;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        before[i] = wcount + __popc(m & lt);
        wcount += __popc(m);
        if (keep) {
#pragma unroll
            for (int p = 0; p < 4; ++p) atomicAdd(&s_h[p][(keys[i] >> (8 * p)) & 255u], 1u);
        }
    }
    if (lane == 0) s_warp[warp] = wcount;
    __syncthreads();
    if (warp == 0) {
        // warp 0: exclusive scan of the 8 warp counts, then the look-back over the tiles with earlier tickets
        const uint32_t c = lane < 8 ? s_warp[lane] : 0u;
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, x, 7);
        __syncwarp();
        if (lane < 8) s_warp[lane] = x - c;
        volatile uint32_t* st = status;
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = total | FLAG_INCL;
        } else {
            if (lane == 0) st[tile] = total | FLAG_AGG;
            excl = warp_lookback<uint32_t, 30>(st, tile, lane);
            if (lane == 0) st[tile] = ((excl + total) & VAL_MASK) | FLAG_INCL;
        }
        if (lane == 0) s_base = excl;
    }
    __syncthreads();
    const uint32_t base = s_base + s_warp[warp];
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        if (keys[i] != 0xFFFFFFFFu) {
            keys_out[base + before[i]] = keys[i];
            vals_out[base + before[i]] = wbase + i * 32;
        }
    }
    for (int i = tid; i < 4 * RADIX; i += 256) {
        const uint32_t c = (&s_h[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// exclusive scan of each 256-entry row, in place. One CTA of 256 threads per row.
__global__ void __launch_bounds__(256) scan_rows_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_w[8];
    uint32_t* row = hist + blockIdx.x * RADIX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v = row[tid], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_w[w];
    row[tid] = base + x - v;
}

// MINB: CTAs per SM the registers are bounded for.  The tile sort (20 M items, issue-bound) wants 4 (64 registers;
// 3: 0.309 vs 0.282 ms); the depth sort (1.8 M items, one wave, latency-bound) 3 (85 registers, no spills: 0.105 vs 0.116 ms).
template <bool IOTA, bool WRITE_KEYS, int MINB>
__global__ void __launch_bounds__(SORT_THREADS, MINB)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                     const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                     const uint32_t* __restrict__ n_ptr, uint32_t n_imm, int begin_bit, int num_bits,
                     const uint32_t* __restrict__ bin_base, uint32_t* __restrict__ status,
                     uint32_t* __restrict__ ticket, uint2* __restrict__ ranges_raw, uint32_t n_env,
                     uint32_t* __restrict__ tile_obj_count) {
    constexpr int WARPS = SORT_THREADS / 32;
    // s_warp_pos[w][d]: first the number of digit-d items of warp w (early counts), then the running
    // position inside the CTA-tile's digit-sorted staging buffer where warp w's next digit-d item goes
    __shared__ uint32_t s_warp_pos[WARPS][RADIX];
    __shared__ uint32_t s_gbase[RADIX];
    __shared__ uint32_t s_keys[SORT_TILE];
    __shared__ uint32_t s_vals[SORT_TILE];
    __shared__ uint32_t s_scan[WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t n = n_ptr ? *n_ptr : n_imm;
    const uint32_t num_tiles = (n + SORT_TILE - 1) / SORT_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * RADIX; i += SORT_THREADS) (&s_warp_pos[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= num_tiles) return;
    const uint32_t base = tile * SORT_TILE;
    const uint32_t tile_n = min((uint32_t)SORT_TILE, n - base);
    const uint32_t mask = (1u << num_bits) - 1u;

    // ---- load (warp-blocked: item i of lane l is element wbase + 32 i, so rank order == memory order)
    uint32_t keys[SORT_IPT], vals[SORT_IPT];
    const uint32_t wbase = base + warp * (32 * SORT_IPT) + lane;
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        uint32_t idx = wbase + i * 32;
        bool valid = idx < n;
        keys[i] = valid ? keys_in[idx] : 0xFFFFFFFFu;
        if (IOTA) vals[i] = idx;
        else vals[i] = valid ? vals_in[idx] : 0u;
    }
    // ---- early counts: per-warp digit histogram (padding items count as digit `mask`, sorted last)
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        uint32_t idx = wbase + i * 32;
        uint32_t d = idx < n ? ((keys[i] >> begin_bit) & mask) : mask;
        atomicAdd(&s_warp_pos[warp][d], 1u);
    }
    __syncthreads();
    // ---- digit `tid`: CTA count, publish the aggregate EARLY (the ranking below overlaps the
    //      predecessors' progress, so the look-back after it mostly finds inclusive prefixes)
    uint32_t cta_count = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) cta_count += s_warp_pos[w][tid];
    const uint32_t pad = SORT_TILE - tile_n;
    const uint32_t pub = cta_count - ((uint32_t)tid == mask ? pad : 0u);
    volatile uint32_t* st = status + (size_t)tile * RADIX + tid;
    if ((uint32_t)tid <= mask) *st = pub | (tile == 0 ? FLAG_INCL : FLAG_AGG);
    // exclusive scan of cta_count over the digits -> start of each digit in the staging buffer
    uint32_t bin_start;
    {
        uint32_t x = cta_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; ++w) wb += s_scan[w];
        bin_start = wb + x - cta_count;
    }
    {
        uint32_t run = bin_start;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_warp_pos[w][tid];
            s_warp_pos[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    // ---- rank + scatter into the staging buffer: warp multisplit by explicit votes (one per digit
    //      bit; inline PTX because nvcc turns the C++ idiom into MATCH.ANY, which is slower when the
    //      32 digits differ), then one counter read per lane and one update per digit group
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_IPT; ++i) {
        uint32_t idx = wbase + i * 32;
        uint32_t d = idx < n ? ((keys[i] >> begin_bit) & mask) : mask;
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < num_bits) {
                uint32_t bal, bit = (d >> b) & 1u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\nvote.sync.ballot.b32 %0, p, 0xffffffff;\n}\n"
                             : "=r"(bal) : "r"(bit));
                peers &= bal ^ (bit - 1u);  // bit ? bal : ~bal
            }
        }
        const uint32_t below = __popc(peers & lt);
        const uint32_t pos = s_warp_pos[warp][d] + below;
        __syncwarp();
        if (below == 0) s_warp_pos[warp][d] = pos + __popc(peers);
        __syncwarp();
        s_keys[pos] = keys[i];
        s_vals[pos] = vals[i];
    }
    // ---- decoupled look-back, one digit per thread, LB_WIN predecessors in flight per step
    if ((uint32_t)tid <= mask) {
        uint32_t excl = 0;
        if (tile != 0) {
            constexpr int LB_WIN = 8;
            int t = (int)tile - 1;
            bool found = false;
            while (!found) {
                uint32_t sv[LB_WIN];
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j)
                    sv[j] = (t - j >= 0) ? *(volatile uint32_t*)(status + (size_t)(t - j) * RADIX + tid) : (2u << 30);
                int used = 0;
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j) {
                    if (!found && used == j) {
                        const uint32_t f = sv[j] >> 30;
                        if (f != 0) {
                            excl += sv[j] & VAL_MASK;
                            ++used;
                            if (f == 2) found = true;
                        }
                    }
                }
                t -= used;
            }
            *st = ((excl + pub) & VAL_MASK) | FLAG_INCL;
        }
        s_gbase[tid] = bin_base[tid] + excl - bin_start;
    }
    __syncthreads();
    // ---- coalesced store: staging position j of digit d goes to (global start of d) + (j - CTA start of d)
    for (uint32_t j0 = 0; j0 < tile_n; j0 += SORT_THREADS) {
        const uint32_t j = j0 + tid;
        const bool valid = j < tile_n;
        uint32_t k = 0, v = 0, dst = 0;
        if (valid) {
            k = s_keys[j];
            v = s_vals[j];
            const uint32_t d = (k >> begin_bit) & mask;
            dst = s_gbase[d] + j;
            if (WRITE_KEYS) keys_out[dst] = k;
            vals_out[dst] = v;
        }
        // Last pass of the tile sort (keys are whole tile ids): identifyTileRanges happens here.  Where the id
        // changes inside this CTA-tile's staging buffer, a run of that id starts / ends at a known output
        // position; the tile's range is the union over CTA-tiles: max of the ends, min of the starts (kept
        // bit-inverted, so that the per-frame clear to zero is the identity of both reductions).
        if (ranges_raw) {
            const bool head = valid && (j == 0 || s_keys[j - 1] != k);
            if (head) atomicMax(&ranges_raw[k].x, ~dst);
            if (valid && (j + 1 == tile_n || s_keys[j + 1] != k)) atomicMax(&ranges_raw[k].y, dst + 1u);
            // un-culled OBJECT pairs per tile (the compositing producer stops scanning a list once it has seen them
            // all): counted per run of equal ids inside the warp's 32 staging positions — one atomic per run that
            // holds objects instead of one per object pair (3.8 M atomics on ~1500 hot words cost the emit kernel
            // 0.1 ms when it did this)
            if (tile_obj_count) {
                const bool is_obj = valid && !(v & PG_CULL_FLAG) && v >= n_env;
                const uint32_t heads = __ballot_sync(0xffffffffu, head || (valid && lane == 0));
                const uint32_t objm = __ballot_sync(0xffffffffu, is_obj);
                if (valid && (heads >> lane & 1u)) {
                    const uint32_t later = heads & ~((2u << lane) - 1u);      // heads after this lane
                    const uint32_t seg = (later ? ((1u << (__ffs(later) - 1)) - 1u) : 0xFFFFFFFFu) & ~((1u << lane) - 1u);
                    const uint32_t c = __popc(objm & seg);
                    if (c) atomicAdd(&tile_obj_count[k * OBJ_SPREAD + ((tile + warp) & (OBJ_SPREAD - 1))], c);
                }
            }
        }
    }
}

int launch_compact_hist(const uint32_t* keys_in, uint32_t n, uint32_t* keys_out, uint32_t* vals_out, uint32_t* hist,
                        uint32_t* status, uint32_t* ticket, cudaStream_t stream) {
    if (n == 0) return PG_OK;
    compact_hist_kernel<<<(n + SCAN_TILE - 1) / SCAN_TILE, 256, 0, stream>>>(keys_in, n, keys_out, vals_out, hist, status, ticket);
    PG_CUDA_CHECK(cudaGetLastError());
    scan_rows_kernel<<<4, 256, 0, stream>>>(hist);
    count_launch(2);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

// One onesweep pass.  max_tiles bounds the grid (tiles beyond the device-side count exit at once).
int launch_onesweep_pass(bool iota, bool write_keys, const uint32_t* keys_in, uint32_t* keys_out,
                         const uint32_t* vals_in, uint32_t* vals_out, const uint32_t* n_ptr,
                         uint32_t n_imm, uint32_t max_tiles, int begin_bit, int num_bits,
                         const uint32_t* bin_base, uint32_t* status, uint32_t* ticket, uint2* ranges_raw,
                         uint32_t n_env, uint32_t* tile_obj_count, cudaStream_t stream) {
    if (max_tiles == 0) return PG_OK;
    dim3 grid(max_tiles), block(SORT_THREADS);
    // small grids (at most two waves of 3 CTAs per SM on 148 SMs: the depth sort of a few million Gaussians) are
    // latency-bound and run the spill-free 3-CTA build
    const bool depth_pass = write_keys && !ranges_raw && max_tiles <= 148u * 3u * 2u;
#define PG_ONESWEEP(I, W, M) onesweep_pass_kernel<I, W, M><<<grid, block, 0, stream>>>(keys_in, keys_out, vals_in, vals_out, n_ptr, n_imm, begin_bit, num_bits, bin_base, status, ticket, ranges_raw, n_env, tile_obj_count)
    if (iota && write_keys) PG_ONESWEEP(true, true, 4);
    else if (!iota && write_keys && depth_pass) PG_ONESWEEP(false, true, 3);
    else if (!iota && write_keys) PG_ONESWEEP(false, true, 4);
    else if (!iota && !write_keys) PG_ONESWEEP(false, false, 4);
    else PG_ONESWEEP(true, false, 4);
#undef PG_ONESWEEP
    PG_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return PG_OK;
}

}  // namespace pg
