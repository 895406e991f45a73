// binning.cu — tile binning after the depth sort (re-designed):
//   emit_kernel      : fused inclusive scan of tiles_touched (decoupled look-back over 1024-Gaussian
//                      chunks, in DEPTH order) + key duplication: writes (tile id, Gaussian index)
//                      pairs and counts pairs per tile.  12 B read per Gaussian, 8 B written per pair.
//   tile_scan_kernel : exclusive scan of the per-tile counts -> ranges[tile] (identifyTileRanges
//                      without touching the sorted keys) + exclusive digit bases of both tile-sort passes.
//   export_keys_kernel (tests only): rebuilds the reference's 64-bit tile|depth keys.
#include "pg_common.cuh"

namespace pg {

constexpr uint32_t E_FLAG_AGG = 1u << 30;
constexpr uint32_t E_FLAG_INCL = 2u << 30;
constexpr uint32_t E_VAL_MASK = (1u << 30) - 1;

__global__ void __launch_bounds__(256)
emit_kernel(const uint32_t* __restrict__ sorted_dkey, const uint32_t* __restrict__ perm,
            const ushort4* __restrict__ rects, const GeomRec* __restrict__ recs, uint32_t P, uint32_t gx,
            int W, int H, uint32_t* __restrict__ tkeys,
            uint32_t* __restrict__ tvals, uint32_t R_cap, uint32_t* __restrict__ status,
            uint32_t* __restrict__ tile_count, uint32_t n_env, uint32_t* __restrict__ tile_obj_count,
            Counters* __restrict__ counters) {
    __shared__ uint32_t s_g[EMIT_CHUNK];
    __shared__ ushort4 s_rect[EMIT_CHUNK];
    __shared__ uint32_t s_off[EMIT_CHUNK];
    __shared__ float4 s_ga[EMIT_CHUNK];  // x, y, conic.x, conic.y
    __shared__ float2 s_gb[EMIT_CHUNK];  // conic.z, cut
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_chunk, s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_chunk = atomicAdd(&counters->tile_counter[4], 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    const uint32_t num_chunks = (P + EMIT_CHUNK - 1) / EMIT_CHUNK;
    if (chunk >= num_chunks) return;

    // blocked: thread owns 4 consecutive sorted positions
    uint32_t tt[4], local = 0;
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
        tt[j] = (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
        s_g[tid * 4 + j] = g;
        s_rect[tid * 4 + j] = r;
        local += tt[j];
    }
    // block exclusive scan of `local`
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
                counters->num_rendered = (uint32_t)min(R, (uint64_t)0xFFFFFFFFu);
                counters->sort_n = (uint32_t)min(R, (uint64_t)R_cap);
                if (R > R_cap) counters->overflow = 1;
            }
        }
    }
    __syncthreads();
    const uint32_t base = s_base;
    uint32_t o = base + excl;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_off[tid * 4 + j] = o;
        o += tt[j];
    }
    __syncthreads();
    // warp w expands Gaussians [w*128, w*128+128): lanes over the tiles of one rectangle
    for (int q = warp * 128; q < warp * 128 + 128; ++q) {
        const ushort4 r = s_rect[q];
        const uint32_t w_ = r.z - r.x, h_ = r.w - r.y;
        const uint32_t n = w_ * h_;
        if (n == 0) continue;
        const uint32_t g = s_g[q];
        const uint32_t off = s_off[q];
        const float4 ga = s_ga[q];
        const float2 gb = s_gb[q];
        const float inv = __frcp_rn((float)w_);
        for (uint32_t t = lane; t < n; t += 32) {
            uint32_t row = (uint32_t)__float2uint_rz(((float)t + 0.5f) * inv);
            int rem = (int)t - (int)(row * w_);
            if (rem < 0) { --row; rem += (int)w_; }
            else if (rem >= (int)w_) { ++row; rem -= (int)w_; }
            const uint32_t tyi = r.y + row, txi = r.x + (uint32_t)rem;
            const uint32_t tile = tyi * gx + txi;
            const uint32_t dst = off + t;
            if (dst < R_cap) {
                // can this Gaussian reach alpha >= 1/255 at any pixel centre of the tile?
                const float x0 = (float)(txi * PG_TILE), y0 = (float)(tyi * PG_TILE);
                const float x1 = (float)min((int)(txi * PG_TILE + PG_TILE - 1), W - 1);
                const float y1 = (float)min((int)(tyi * PG_TILE + PG_TILE - 1), H - 1);
                const bool culled = block_culled(ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, x0, x1, y0, y1);
                tkeys[dst] = tile;
                tvals[dst] = culled ? (g | PG_CULL_FLAG) : g;
                atomicAdd(&tile_count[tile], 1u);
                if (g >= n_env && !culled) atomicAdd(&tile_obj_count[tile], 1u);
            }
        }
    }
}

// ranges + digit bases from per-tile counts. One CTA, 1024 threads.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const uint32_t* __restrict__ tile_count, uint32_t tiles, int bits_lo, int bits_hi,
                 uint2* __restrict__ ranges, uint32_t* __restrict__ bins /*[2][256]*/) {
    __shared__ uint32_t s_lo[RADIX], s_hi[RADIX];
    __shared__ uint32_t s_scan[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < RADIX) { s_lo[tid] = 0; s_hi[tid] = 0; }
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const uint32_t mask_lo = (1u << bits_lo) - 1u;
    for (uint32_t b = 0; b < tiles; b += 1024) {
        uint32_t t = b + tid;
        uint32_t c = t < tiles ? tile_count[t] : 0u;
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[warp] = x;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            uint32_t v = s_scan[w];
            if (w < warp) wb += v;
            tot += v;
        }
        uint32_t start = s_carry + wb + x - c;
        if (t < tiles) {
            ranges[t] = c ? make_uint2(start, start + c) : make_uint2(0u, 0u);
            if (c) {
                atomicAdd(&s_lo[t & mask_lo], c);
                atomicAdd(&s_hi[t >> bits_lo], c);
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    // exclusive scans of the two 256-bin histograms: warps 0..7 -> lo, warps 8..15 -> hi
    if (tid < 2 * RADIX) {
        uint32_t* h = tid < RADIX ? s_lo : s_hi;
        int i = tid & (RADIX - 1);
        uint32_t v = h[i], x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[warp] = x;
        __syncwarp();
        // per-histogram warp totals live in s_scan[0..7] / s_scan[8..15]
        asm volatile("bar.sync 1, 512;");
        uint32_t wb = 0;
        int w0 = tid < RADIX ? 0 : 8;
        for (int w = w0; w < warp; ++w) wb += s_scan[w];
        bins[tid] = wb + x - v;
    }
    (void)bits_hi;
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

int launch_emit(const uint32_t* sorted_dkey, const uint32_t* perm, const ushort4* rects, const GeomRec* recs,
                uint32_t P, uint32_t gx, int W, int H, uint32_t* tkeys, uint32_t* tvals, uint32_t R_cap, uint32_t* status,
                uint32_t* tile_count, uint32_t n_env, uint32_t* tile_obj_count, Counters* counters,
                cudaStream_t stream) {
    uint32_t chunks = (P + EMIT_CHUNK - 1) / EMIT_CHUNK;
    if (chunks == 0) return PG_OK;
    emit_kernel<<<chunks, 256, 0, stream>>>(sorted_dkey, perm, rects, recs, P, gx, W, H, tkeys, tvals, R_cap, status,
                                            tile_count, n_env, tile_obj_count, counters);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_tile_scan(const uint32_t* tile_count, uint32_t tiles, int bits_lo, int bits_hi, uint2* ranges,
                     uint32_t* bins, cudaStream_t stream) {
    tile_scan_kernel<<<1, 1024, 0, stream>>>(tile_count, tiles, bits_lo, bits_hi, ranges, bins);
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
