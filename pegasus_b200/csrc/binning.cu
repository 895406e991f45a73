// binning.cu — tile binning after the depth sort (re-designed):
//   emit_kernel      : fused inclusive scan of tiles_touched (decoupled look-back over 1024-Gaussian
//                      chunks, in DEPTH order) + key duplication: writes (tile id, Gaussian index)
//                      pairs and counts pairs per tile.  12 B read per Gaussian, 8 B written per pair.
//   tile_hist_kernel : digit histograms of both tile-sort passes from the emitted tile ids (warp votes).
//   tile_scan_kernel : exclusive digit bases of both tile-sort passes.
//   ranges_kernel    : identifyTileRanges on the sorted tile ids.
//   export_keys_kernel (tests only): rebuilds the reference's 64-bit tile|depth keys.
#include "pg_common.cuh"
#include "tile_cull.h"

namespace pg {

constexpr uint32_t E_FLAG_AGG = 1u << 30;
constexpr uint32_t E_FLAG_INCL = 2u << 30;
constexpr uint32_t E_VAL_MASK = (1u << 30) - 1;

struct EmitSmem {
    float4 ga[EMIT_CHUNK];  // x, y, conic.x, conic.y
    ushort4 rect[EMIT_CHUNK];
    float2 gb[EMIT_CHUNK];  // conic.z, cut
    uint32_t g[EMIT_CHUNK];
    uint32_t cnt[EMIT_CHUNK];
    uint32_t off[EMIT_CHUNK];
    uint32_t rowpre[8][32];
    short2 rowrun[8][32];
    uint32_t scan[8];
    uint32_t chunk, base;
};

constexpr int EMIT_SMALL = 16;  // entries with at most this many pairs are written by their own thread

// One stored pair.  flag: the tile cannot receive a contribution (KEEP_ALL lists only).
__device__ __forceinline__ void emit_pair(bool valid, uint32_t dst, uint32_t tile, uint32_t g, bool flag,
                                          uint32_t n_env, uint32_t* __restrict__ tkeys,
                                          uint32_t* __restrict__ tvals, uint32_t* __restrict__ tile_obj_count) {
    if (valid) {
        tkeys[dst] = tile;
        tvals[dst] = flag ? (g | PG_CULL_FLAG) : g;
        if (g >= n_env && !flag) atomicAdd(&tile_obj_count[tile], 1u);
    }
}

// KEEP_ALL = false (default): only (tile, Gaussian) pairs that can contribute are stored — per tile a
//            subsequence, in the same order, of the reference's list.
// KEEP_ALL = true : every tile of every rectangle is stored exactly as the reference's duplicateWithKeys
//            does (pairs that cannot contribute carry PG_CULL_FLAG): pg_export_binning / n_contrib parity.
// Phases: (0) gather the chunk's 1024 depth-ordered entries into shared memory; (A) one thread per
// entry counts its pairs with the closed-form tile-row test of tile_cull.h; (B) CTA scan + decoupled
// look-back -> output offsets; (C1) entries with <= EMIT_SMALL pairs are written by their own thread,
// (C2) larger ones by their whole warp, lanes over pairs (coalesced), rows found by binary search.
template <bool KEEP_ALL>
__global__ void __launch_bounds__(256)
emit_kernel(const uint32_t* __restrict__ sorted_dkey, const uint32_t* __restrict__ perm,
            const ushort4* __restrict__ rects, const GeomRec* __restrict__ recs, uint32_t P, uint32_t gx,
            int W, int H, uint32_t* __restrict__ tkeys,
            uint32_t* __restrict__ tvals, uint32_t R_cap, uint32_t* __restrict__ status,
            uint32_t n_env, uint32_t* __restrict__ tile_obj_count, Counters* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char emit_smem_raw[];
    EmitSmem& sm = *reinterpret_cast<EmitSmem*>(emit_smem_raw);
    uint32_t* s_g = sm.g;
    ushort4* s_rect = sm.rect;
    uint32_t* s_cnt = sm.cnt;
    uint32_t* s_off = sm.off;
    float4* s_ga = sm.ga;
    float2* s_gb = sm.gb;
    uint32_t (*s_rowpre)[32] = sm.rowpre;
    short2 (*s_rowrun)[32] = sm.rowrun;
    uint32_t* s_scan = sm.scan;
    uint32_t& s_chunk = sm.chunk;
    uint32_t& s_base = sm.base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_chunk = atomicAdd(&counters->tile_counter[4], 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    const uint32_t num_chunks = (P + EMIT_CHUNK - 1) / EMIT_CHUNK;
    if (chunk >= num_chunks) return;

    // ---- (0) gather: thread owns 4 consecutive sorted positions (blocked, the scan order)
    const uint32_t s0 = chunk * EMIT_CHUNK + tid * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t s = s0 + j;
        uint32_t g = 0;
        ushort4 r = make_ushort4(0, 0, 0, 0);
        if (s < P && sorted_dkey[s] != 0xFFFFFFFFu) {
            g = perm[s];
            r = rects[g];
            const float4 ra = recs[g].a, rb = recs[g].b;
            s_ga[tid * 4 + j] = ra;
            s_gb[tid * 4 + j] = make_float2(rb.x, rb.w);
        }
        s_g[tid * 4 + j] = g;
        s_rect[tid * 4 + j] = r;
    }
    __syncthreads();
    // ---- (A) count
    uint32_t local = 0;
    unsigned long long local_full = 0;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const int q = tid * 4 + j;
        const ushort4 r = s_rect[q];
        const uint32_t n = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
        uint32_t cnt = n;
        if (!KEEP_ALL && n > 0) {
            const float4 ga = s_ga[q];
            const float2 gb = s_gb[q];
            const CullGauss cg = cull_setup(ga.x, ga.y, ga.z, ga.w, gb.x, gb.y);
            cnt = 0;
            for (int ty = r.y; ty < r.w; ++ty) {
                int ta, tb;
                cnt += (uint32_t)cull_row_run(cg, ty, r.x, r.z, W, H, &ta, &tb);
            }
        }
        s_cnt[q] = cnt;
        local += cnt;
        local_full += n;
    }
    // the reference's R = sum of all rectangle areas (what pg_status.num_rendered reports)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_full += __shfl_xor_sync(0xffffffffu, local_full, o);
    if (lane == 0 && local_full) atomicAdd(&counters->rendered_full, local_full);
    // ---- (B) block exclusive scan of `local`
    uint32_t x = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    uint32_t wb = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) wb += s_scan[w];
        total += s_scan[w];
    }
    uint32_t excl = wb + x - local;
    // chunk-level decoupled look-back (single value): warp 0 inspects 32 predecessors per step
    if (warp == 0) {
        volatile uint32_t* st = status + chunk;
        uint32_t prev = 0;
        const uint32_t tot_c = min(total, E_VAL_MASK);
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
                prev += v;
                if (first_in < first_nr) break;
                t -= take;
            }
            if (lane == 0) *st = min(prev + tot_c, E_VAL_MASK) | E_FLAG_INCL;
        }
        if (lane == 0) {
            s_base = prev;
            if (chunk == num_chunks - 1) {
                uint64_t R = (uint64_t)prev + total;
                counters->sort_n = (uint32_t)min(R, (uint64_t)R_cap);
                if (R > R_cap) counters->overflow = 1;
            }
        }
    }
    __syncthreads();
    const uint32_t base = s_base;
    {
        uint32_t o = base + excl;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s_off[tid * 4 + j] = o;
            o += s_cnt[tid * 4 + j];
        }
    }
    // ---- (C1) small entries: own thread, tile rows in order (y outer, x inner = the reference's order)
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const int q = tid * 4 + j;
        const uint32_t cnt = s_cnt[q];
        if (cnt == 0 || cnt > (uint32_t)EMIT_SMALL) continue;
        const ushort4 r = s_rect[q];
        const uint32_t g = s_g[q];
        const float4 ga = s_ga[q];
        const float2 gb = s_gb[q];
        const CullGauss cg = cull_setup(ga.x, ga.y, ga.z, ga.w, gb.x, gb.y);
        uint32_t dst = s_off[q];
        for (int ty = r.y; ty < r.w; ++ty) {
            int ta, tb;
            cull_row_run(cg, ty, r.x, r.z, W, H, &ta, &tb);
            const int xa = KEEP_ALL ? (int)r.x : ta, xb = KEEP_ALL ? (int)r.z : tb;
            for (int tx = xa; tx < xb; ++tx, ++dst)
                emit_pair(dst < R_cap, dst, (uint32_t)ty * gx + (uint32_t)tx, g, KEEP_ALL && !(tx >= ta && tx < tb),
                          n_env, tkeys, tvals, tile_obj_count);
        }
    }
    __syncthreads();  // s_off of every entry is visible to its warp mates
    // ---- (C2) large entries: the whole warp, lanes over pairs
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        uint32_t big = __ballot_sync(0xffffffffu, s_cnt[tid * 4 + j] > (uint32_t)EMIT_SMALL);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int q = (warp * 32 + src) * 4 + j;
            const ushort4 r = s_rect[q];
            const uint32_t g = s_g[q];
            const float4 ga = s_ga[q];
            const float2 gb = s_gb[q];
            const CullGauss cg = cull_setup(ga.x, ga.y, ga.z, ga.w, gb.x, gb.y);
            uint32_t off = s_off[q];
            const int rows = r.w - r.y;
            for (int row0 = 0; row0 < rows; row0 += 32) {
                const int row = row0 + lane;
                int ta = r.x, tb = r.x;
                uint32_t len = 0;
                if (row < rows) {
                    const int c = cull_row_run(cg, r.y + row, r.x, r.z, W, H, &ta, &tb);
                    len = KEEP_ALL ? (uint32_t)(r.z - r.x) : (uint32_t)c;
                }
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                s_rowpre[warp][lane] = incl - len;
                s_rowrun[warp][lane] = make_short2((short)ta, (short)tb);
                __syncwarp();
                for (uint32_t p0 = 0; p0 < tot; p0 += 32) {
                    const uint32_t p = p0 + lane;
                    const bool valid = p < tot && off + p < R_cap;
                    // largest i with rowpre[i] <= p (its row is non-empty)
                    int i = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1)
                        if (s_rowpre[warp][i + step] <= p) i += step;
                    const short2 run = s_rowrun[warp][i];
                    const int col = (int)(p - s_rowpre[warp][i]);
                    const int tx = (KEEP_ALL ? (int)r.x : (int)run.x) + col;
                    const int ty = r.y + row0 + i;
                    emit_pair(valid, off + p, (uint32_t)ty * gx + (uint32_t)tx, g,
                              KEEP_ALL && !(tx >= run.x && tx < run.y), n_env, tkeys, tvals, tile_obj_count);
                }
                off += tot;
                __syncwarp();
            }
        }
    }
}

// Digit histograms of both tile-sort passes from the emitted tile ids.  Shared-memory atomics cost
// ~64 cycles per warp instruction on this part, so counting is done with warp votes instead: one
// ballot per digit bit gives every lane the set of lanes with the same digit, the lowest of them adds
// the group size to a warp-private counter (plain load/store).  ~30 vote-pipe cycles per 32 keys.
constexpr int THIST_IPT = 16;
__global__ void __launch_bounds__(256)
tile_hist_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, int bits_lo, int bits_hi,
                 uint32_t* __restrict__ hist /*[2][256]*/) {
    __shared__ uint32_t s_h[8][2][RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8 * 2 * RADIX; i += 256) (&s_h[0][0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = *n_ptr;
    const uint32_t mask_lo = (1u << bits_lo) - 1u;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t tile_items = 256 * THIST_IPT;
    for (uint32_t base = blockIdx.x * tile_items; base < n; base += gridDim.x * tile_items) {
        const uint32_t wbase = base + warp * (32 * THIST_IPT) + lane;
        uint32_t k[THIST_IPT];
#pragma unroll
        for (int i = 0; i < THIST_IPT; ++i) k[i] = (wbase + i * 32 < n) ? keys[wbase + i * 32] : 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < THIST_IPT; ++i) {
            const bool valid = wbase + i * 32 < n;
            const uint32_t vm = __ballot_sync(0xffffffffu, valid);
            const uint32_t dlo = k[i] & mask_lo, dhi = (k[i] >> bits_lo) & 255u;
            uint32_t plo = vm, phi = vm;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (b < bits_lo) {
                    uint32_t bal, bit = (dlo >> b) & 1u;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\nvote.sync.ballot.b32 %0, p, 0xffffffff;\n}\n"
                                 : "=r"(bal) : "r"(bit));
                    plo &= bal ^ (bit - 1u);
                }
                if (b < bits_hi) {
                    uint32_t bal, bit = (dhi >> b) & 1u;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\nvote.sync.ballot.b32 %0, p, 0xffffffff;\n}\n"
                                 : "=r"(bal) : "r"(bit));
                    phi &= bal ^ (bit - 1u);
                }
            }
            if (valid && (plo & lt) == 0) s_h[warp][0][dlo] += __popc(plo);
            if (valid && (phi & lt) == 0) s_h[warp][1][dhi] += __popc(phi);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * RADIX; i += 256) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) c += (&s_h[w][0][0])[i];
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

// identifyTileRanges on the sorted tile ids: ranges[tile] = [first, last + 1); tiles without pairs keep
// the (0, 0) of the per-frame clear.  n lives on the device (stored pairs).
__global__ void __launch_bounds__(256)
ranges_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, uint2* __restrict__ ranges) {
    const uint32_t n = *n_ptr;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = keys[i];
        if (i == 0 || keys[i - 1] != k) ranges[k].x = i;
        if (i == n - 1 || keys[i + 1] != k) ranges[k].y = i + 1;
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
                uint32_t n_env, uint32_t* tile_obj_count, Counters* counters, cudaStream_t stream) {
    uint32_t chunks = (P + EMIT_CHUNK - 1) / EMIT_CHUNK;
    if (chunks == 0) return PG_OK;
    static bool attr_set = false;
    if (!attr_set) {
        PG_CUDA_CHECK(cudaFuncSetAttribute(emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem)));
        PG_CUDA_CHECK(cudaFuncSetAttribute(emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem)));
        attr_set = true;
    }
    if (keep_all)
        emit_kernel<true><<<chunks, 256, sizeof(EmitSmem), stream>>>(sorted_dkey, perm, rects, recs, P, gx, W, H, tkeys, tvals, R_cap, status,
                                                      n_env, tile_obj_count, counters);
    else
        emit_kernel<false><<<chunks, 256, sizeof(EmitSmem), stream>>>(sorted_dkey, perm, rects, recs, P, gx, W, H, tkeys, tvals, R_cap, status,
                                                       n_env, tile_obj_count, counters);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_tile_hist(const uint32_t* keys, const uint32_t* n_ptr, uint32_t max_n, int bits_lo, int bits_hi, uint32_t* hist,
                     cudaStream_t stream) {
    if (max_n == 0) return PG_OK;
    const uint32_t tiles = (max_n + 256 * THIST_IPT - 1) / (256 * THIST_IPT);
    const uint32_t blocks = min(tiles, (uint32_t)(PG_SM_COUNT * 8));
    tile_hist_kernel<<<blocks, 256, 0, stream>>>(keys, n_ptr, bits_lo, bits_hi, hist);
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

int launch_ranges(const uint32_t* sorted_tile_keys, const uint32_t* n_ptr, uint32_t max_n, uint2* ranges, cudaStream_t stream) {
    if (max_n == 0) return PG_OK;
    const uint32_t blocks = min((max_n + 255u) / 256u, (uint32_t)(PG_SM_COUNT * 16));
    ranges_kernel<<<blocks, 256, 0, stream>>>(sorted_tile_keys, n_ptr, ranges);
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
