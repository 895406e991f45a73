// composite.cu — per-tile front-to-back alpha compositing.
//
// One CTA (256 threads, 16x16 pixels; each warp owns an 8x4 pixel block) per tile.  The tile's sorted
// Gaussian indices are read coalesced, and each index's 48-byte record {xy, conic, opacity, depth,
// rgb, object id} is GATHERED into shared memory by its own TMA bulk copy (cp.async.bulk, 48 B,
// completion on an mbarrier) into a 2-stage ring: batch r+1 lands while batch r is composited, no
// registers or scoreboard slots are held by the loads.  All lanes then walk the staged batch with
// broadcast 16-byte shared loads.
//
// composite_kernel        : the reference's single pass -> color, depth (+ final_T, n_contrib).
// composite_masks_kernel  : the reference's K+3 passes in one walk -> RGB + depth, the objects-only
//                           flat-colour render (visible masks, sem-seg) and one transmittance chain
//                           per object (silhouettes).  alpha is evaluated once per (pixel, Gaussian).
//
// Compute-bound (FP32 pipe): ~12 FP32 ops to reject a pair, ~40 to blend one (exp is a 12-op FMA
// polynomial so that results are bit-reproducible on the CPU oracle).
#include <cstring>

#include "pg_common.cuh"

namespace pg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct CompArgs {
    const uint2* ranges;
    const uint32_t* point_list;
    const GeomRec* recs;
    int W, H, gx;
    const float* bg;
    float* out_color;
    float* out_depth;
    float* out_final_T;
    uint32_t* out_n_contrib;
    // masks
    uint32_t n_env;                 // Gaussian indices >= n_env belong to objects
    const uint32_t* tile_obj_count; // [tiles] number of object pairs per tile
    int num_objects, num_colors;
    float eff_color[PG_MAX_OBJECTS][3];  // colour the rasterizer produces for object k's flat SH
    float set_color[PG_MAX_COLORS][3];   // colour set the masks are tested against
    int color_index[PG_MAX_OBJECTS];
    float* seg_color;
    uint8_t* sem_seg;
    uint8_t* visible;
    uint8_t* silhouette;
    unsigned long long* stats;  // non-null: count pairs evaluated / exp'd / blended
};

// alpha of one (pixel, Gaussian) pair; returns false when the pair is skipped (A.7 `continue`s)
template <bool STATS>
__device__ __forceinline__ bool pair_alpha(const float4 A, const float4 B, float pfx, float pfy, float& alpha,
                                           uint32_t& n_eval, uint32_t& n_exp) {
    if (STATS) ++n_eval;
    float dx = sub(A.x, pfx), dy = sub(A.y, pfy);
    float u = mul(A.z, dx);
    float v = mul(B.x, dy);
    float w = mul(dy, v);
    float s = fma(dx, u, w);
    float bxy = mul(mul(A.w, dx), dy);
    float power = fma(s, -0.5f, -bxy);
    if (power > 0.0f) return false;
    if (power < B.w) return false;  // alpha < 1/255 guaranteed (B.w = -5.55 when opacity <= 1)
    if (STATS) ++n_exp;
    alpha = fminf(0.99f, mul(B.y, expf_exact(power)));
    return !(alpha < 1.0f / 255.0f);
}

__device__ __forceinline__ void flush_stats(unsigned long long* stats, uint32_t n_eval, uint32_t n_exp, uint32_t n_blend) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
        n_exp += __shfl_xor_sync(0xffffffffu, n_exp, o);
        n_blend += __shfl_xor_sync(0xffffffffu, n_blend, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], (unsigned long long)n_eval);
        atomicAdd(&stats[1], (unsigned long long)n_exp);
        atomicAdd(&stats[2], (unsigned long long)n_blend);
    }
}

template <bool STATS>
__global__ void __launch_bounds__(256) composite_kernel(const CompArgs a) {
    __shared__ GeomRec s_rec[2][256];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.y * a.gx + blockIdx.x;
    const int px = blockIdx.x * PG_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = blockIdx.y * PG_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pfx = (float)px, pfy = (float)py;
    const uint2 range = a.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int rounds = (n + 255) >> 8;

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int r) {
        const int cnt = min(256, n - (r << 8));
        uint64_t* bar = &s_bar[r & 1];
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)cnt * (uint32_t)sizeof(GeomRec));
        if (tid < cnt) {
            const uint32_t g = a.point_list[range.x + (r << 8) + tid];
            bulk_g2s(&s_rec[r & 1][tid], a.recs + g, sizeof(GeomRec), bar);
        }
    };

    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f;
    uint32_t contributor = 0, last = 0;
    uint32_t n_eval = 0, n_exp = 0, n_blend = 0;
    bool done = !inside;
    int issued = 0;
    if (rounds > 0) { issue(0); issued = 1; }
    int r = 0;
    for (; r < rounds; ++r) {
        if (r + 1 < rounds) { issue(r + 1); issued = r + 2; }
        mbar_wait(&s_bar[r & 1], (uint32_t)((r >> 1) & 1));
        const int cnt = min(256, n - (r << 8));
        const GeomRec* sr = s_rec[r & 1];
        for (int j = 0; j < cnt && !done; ++j) {
            contributor++;
            const float4 A = sr[j].a;
            const float4 B = sr[j].b;
            float alpha;
            if (!pair_alpha<STATS>(A, B, pfx, pfy, alpha, n_eval, n_exp)) continue;
            float test_T = mul(T, sub(1.0f, alpha));
            if (test_T < 0.0001f) { done = true; continue; }
            if (STATS) ++n_blend;
            const float4 Cc = sr[j].c;
            C0 = fma(mul(Cc.x, alpha), T, C0);
            C1 = fma(mul(Cc.y, alpha), T, C1);
            C2 = fma(mul(Cc.z, alpha), T, C2);
            D = fma(mul(B.z, alpha), T, D);
            T = test_T;
            last = contributor;
        }
        if (__syncthreads_and(done)) { ++r; break; }
    }
    // never leave with a bulk copy still in flight into our shared memory
    if (issued > r) mbar_wait(&s_bar[r & 1], (uint32_t)((r >> 1) & 1));

    if (inside) {
        const size_t HW = (size_t)a.W * a.H, pix = (size_t)py * a.W + px;
        a.out_color[pix] = fma(T, a.bg[0], C0);
        a.out_color[HW + pix] = fma(T, a.bg[1], C1);
        a.out_color[2 * HW + pix] = fma(T, a.bg[2], C2);
        a.out_depth[pix] = D;
        if (a.out_final_T) a.out_final_T[pix] = T;
        if (a.out_n_contrib) a.out_n_contrib[pix] = last;
    }
    if (STATS) flush_stats(a.stats, n_eval, n_exp, n_blend);
}

// ------------------------------------------------------------------------------------------------
// Fused K+3 passes.  Chains per pixel: main (all Gaussians), objects-only, and one per object.
// Phase 1 walks every entry until the main chain of all 256 pixels has terminated; phase 2 scans
// the remaining indices (4 B each), keeps object entries only and composites those.
// ------------------------------------------------------------------------------------------------
template <int KMAX, bool STATS>
__global__ void __launch_bounds__(256) composite_masks_kernel(const CompArgs a) {
    __shared__ GeomRec s_rec[2][256];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_ids[512];
    __shared__ uint32_t s_wcnt[8];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.y * a.gx + blockIdx.x;
    const int px = blockIdx.x * PG_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = blockIdx.y * PG_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pfx = (float)px, pfy = (float)py;
    const uint2 range = a.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int K = a.num_objects;
    int obj_left = (int)a.tile_obj_count[tile];  // object entries of this tile not yet walked

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f;
    float To = 1.0f, S0 = 0.0f, S1 = 0.0f, S2 = 0.0f;
    float Tk[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) Tk[k] = 1.0f;
    bool done_main = !inside, done_o = !inside;
    uint32_t n_eval = 0, n_exp = 0, n_blend = 0;
    uint32_t done_k = inside ? 0u : 0xFFFFFFFFu;

    // blend one staged record into every live chain of this pixel
    auto blend = [&](const GeomRec& rec, bool main_live) {
        const int obj = __float_as_int(rec.c.w);  // warp-uniform
        const bool k_live = obj > 0 && !((done_k >> (obj - 1)) & 1u);
        const bool o_live = obj > 0 && !done_o;
        if (!(main_live || k_live || o_live)) return;
        float alpha;
        if (!pair_alpha<STATS>(rec.a, rec.b, pfx, pfy, alpha, n_eval, n_exp)) return;
        if (STATS) ++n_blend;
        const float om = sub(1.0f, alpha);
        if (main_live) {
            float test_T = mul(T, om);
            if (test_T < 0.0001f) done_main = true;
            else {
                C0 = fma(mul(rec.c.x, alpha), T, C0);
                C1 = fma(mul(rec.c.y, alpha), T, C1);
                C2 = fma(mul(rec.c.z, alpha), T, C2);
                D = fma(mul(rec.b.z, alpha), T, D);
                T = test_T;
            }
        }
        if (obj > 0) {
            if (o_live) {
                float test_T = mul(To, om);
                if (test_T < 0.0001f) done_o = true;
                else {
                    S0 = fma(mul(a.eff_color[obj - 1][0], alpha), To, S0);
                    S1 = fma(mul(a.eff_color[obj - 1][1], alpha), To, S1);
                    S2 = fma(mul(a.eff_color[obj - 1][2], alpha), To, S2);
                    To = test_T;
                }
            }
            if (k_live) {
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k == obj - 1) {  // uniform across the warp
                        float test_T = mul(Tk[k], om);
                        if (test_T < 0.0001f) done_k |= 1u << k;
                        else Tk[k] = test_T;
                    }
                }
            }
        }
    };
    const uint32_t all_k = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);

    // ---------------- phase 1: every entry, until all main chains are done ----------------
    auto issue = [&](int r) {
        const int cnt = min(256, n - (r << 8));
        uint64_t* bar = &s_bar[r & 1];
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)cnt * (uint32_t)sizeof(GeomRec));
        if (tid < cnt) {
            const uint32_t g = a.point_list[range.x + (r << 8) + tid];
            bulk_g2s(&s_rec[r & 1][tid], a.recs + g, sizeof(GeomRec), bar);
        }
    };
    const int rounds = (n + 255) >> 8;
    int issued = 0, r = 0;
    uint32_t par0 = 0, par1 = 0;  // phase parity of the two barriers
    if (rounds > 0) { issue(0); issued = 1; }
    bool all_main_done = false;
    for (; r < rounds; ++r) {
        if (r + 1 < rounds) { issue(r + 1); issued = r + 2; }
        if (r & 1) { mbar_wait(&s_bar[1], par1); par1 ^= 1; } else { mbar_wait(&s_bar[0], par0); par0 ^= 1; }
        const int cnt = min(256, n - (r << 8));
        const GeomRec* sr = s_rec[r & 1];
        int seen_obj = 0;
        for (int j = 0; j < cnt; ++j) {
            seen_obj += __float_as_int(sr[j].c.w) > 0 ? 1 : 0;
            blend(sr[j], !done_main);
        }
        obj_left -= seen_obj;
        all_main_done = __syncthreads_and(done_main);
        if (all_main_done) { ++r; break; }
    }
    if (issued > r) {  // drain the prefetched batch we are not going to use in phase 1
        if (r & 1) { mbar_wait(&s_bar[1], par1); par1 ^= 1; } else { mbar_wait(&s_bar[0], par0); par0 ^= 1; }
    }
    __syncthreads();

    // ---------------- phase 2: object entries only ----------------
    int pos = r << 8;  // first entry not walked in phase 1
    if (pos < n && obj_left > 0 && K > 0) {
        int fill = 0;
        bool pix_done = done_o && (done_k & all_k) == all_k;
        while (true) {
            // scan ids until >= 256 object entries are buffered or the list ends
            while (fill < 256 && pos < n && fill < obj_left) {
                const int idx = pos + tid;
                uint32_t g = 0;
                bool is_obj = false;
                if (idx < n) { g = a.point_list[range.x + idx]; is_obj = g >= a.n_env; }
                const uint32_t bal = __ballot_sync(0xffffffffu, is_obj);
                if (lane == 0) s_wcnt[warp] = __popc(bal);
                __syncthreads();
                int wb = 0, tot = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { int c = (int)s_wcnt[w]; if (w < warp) wb += c; tot += c; }
                if (is_obj) s_ids[fill + wb + __popc(bal & ((1u << lane) - 1u))] = g;
                fill += tot;
                pos += 256;
                __syncthreads();
            }
            if (fill == 0) break;
            const int cnt = min(fill, 256);
            if (tid == 0) mbar_expect_tx(&s_bar[0], (uint32_t)cnt * (uint32_t)sizeof(GeomRec));
            if (tid < cnt) bulk_g2s(&s_rec[0][tid], a.recs + s_ids[tid], sizeof(GeomRec), &s_bar[0]);
            mbar_wait(&s_bar[0], par0); par0 ^= 1;
            if (!pix_done) {
                for (int j = 0; j < cnt; ++j) blend(s_rec[0][j], false);
                pix_done = done_o && (done_k & all_k) == all_k;
            }
            obj_left -= cnt;
            // shift the tail of the id buffer down
            const int rest = fill - cnt;
            uint32_t keep = 0;
            if (tid < rest) keep = s_ids[cnt + tid];
            const bool all_done = __syncthreads_and(pix_done);
            if (tid < rest) s_ids[tid] = keep;
            fill = rest;
            __syncthreads();
            if (all_done || (obj_left <= 0 && fill == 0)) break;
        }
    }

    if (inside) {
        const size_t HW = (size_t)a.W * a.H, pix = (size_t)py * a.W + px;
        const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
        a.out_color[pix] = fma(T, bg0, C0);
        a.out_color[HW + pix] = fma(T, bg1, C1);
        a.out_color[2 * HW + pix] = fma(T, bg2, C2);
        a.out_depth[pix] = D;
        if (a.out_final_T) a.out_final_T[pix] = T;
        const float s0 = fma(To, bg0, S0), s1 = fma(To, bg1, S1), s2 = fma(To, bg2, S2);
        if (a.seg_color) {
            a.seg_color[pix] = s0; a.seg_color[HW + pix] = s1; a.seg_color[2 * HW + pix] = s2;
        }
        if (a.sem_seg) {
            a.sem_seg[3 * pix] = (uint8_t)(int)mul(s0, 255.0f);
            a.sem_seg[3 * pix + 1] = (uint8_t)(int)mul(s1, 255.0f);
            a.sem_seg[3 * pix + 2] = (uint8_t)(int)mul(s2, 255.0f);
        }
        if (a.visible) {
            for (int c = 0; c < a.num_colors; ++c) {
                float d0 = sub(s0, a.set_color[c][0]), d1 = sub(s1, a.set_color[c][1]), d2 = sub(s2, a.set_color[c][2]);
                float dist = sqrt(add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)));
                a.visible[(size_t)c * HW + pix] = dist <= 0.1f ? 1 : 0;
            }
        }
        if (a.silhouette) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    const int ci = a.color_index[k];
                    const float w = sub(1.0f, Tk[k]);
                    float i0 = fma(Tk[k], bg0, mul(a.eff_color[k][0], w));
                    float i1 = fma(Tk[k], bg1, mul(a.eff_color[k][1], w));
                    float i2 = fma(Tk[k], bg2, mul(a.eff_color[k][2], w));
                    float d0 = sub(i0, a.set_color[ci][0]), d1 = sub(i1, a.set_color[ci][1]), d2 = sub(i2, a.set_color[ci][2]);
                    float dist = sqrt(add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)));
                    a.silhouette[(size_t)ci * HW + pix] = dist <= 0.1f ? 1 : 0;
                }
            }
        }
    }
    if (STATS) flush_stats(a.stats, n_eval, n_exp, n_blend);
}

int launch_composite(const CompArgs& a, int gy, bool masks, cudaStream_t stream) {
    dim3 grid(a.gx, gy), block(256);
    const bool st = a.stats != nullptr;
    if (!masks) {
        if (st) composite_kernel<true><<<grid, block, 0, stream>>>(a);
        else composite_kernel<false><<<grid, block, 0, stream>>>(a);
    } else if (a.num_objects <= 8) {
        if (st) composite_masks_kernel<8, true><<<grid, block, 0, stream>>>(a);
        else composite_masks_kernel<8, false><<<grid, block, 0, stream>>>(a);
    } else if (a.num_objects <= 16) {
        if (st) composite_masks_kernel<16, true><<<grid, block, 0, stream>>>(a);
        else composite_masks_kernel<16, false><<<grid, block, 0, stream>>>(a);
    } else {
        if (st) composite_masks_kernel<32, true><<<grid, block, 0, stream>>>(a);
        else composite_masks_kernel<32, false><<<grid, block, 0, stream>>>(a);
    }
    PG_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return PG_OK;
}

}  // namespace pg

namespace pg {

// Fills CompArgs from the ABI structs and launches the right kernel.
int launch_composite_from_abi(const uint2* ranges, const uint32_t* point_list, const GeomRec* recs, int W,
                              int H, const float* bg, const pg_raster_outputs* ro, const pg_frame_outputs* fo,
                              const pg_object_table* objs, uint32_t n_env, const uint32_t* tile_obj_count,
                              unsigned long long* stats, cudaStream_t stream) {
    CompArgs a;
    memset(&a, 0, sizeof(a));
    a.stats = stats;
    a.ranges = ranges; a.point_list = point_list; a.recs = recs;
    a.W = W; a.H = H; a.gx = (W + PG_TILE - 1) / PG_TILE;
    const int gy = (H + PG_TILE - 1) / PG_TILE;
    a.bg = bg;
    a.n_env = n_env;
    a.tile_obj_count = tile_obj_count;
    if (ro) {
        a.out_color = ro->color; a.out_depth = ro->depth; a.out_final_T = ro->final_T; a.out_n_contrib = ro->n_contrib;
        return launch_composite(a, gy, false, stream);
    }
    a.out_color = fo->color; a.out_depth = fo->depth; a.out_final_T = fo->final_T;
    a.seg_color = fo->seg_color; a.sem_seg = fo->sem_seg; a.visible = fo->visible; a.silhouette = fo->silhouette;
    a.num_objects = objs->num_objects; a.num_colors = objs->num_colors;
    const volatile float C0 = 0.28209479177387814f;
    for (int c = 0; c < objs->num_colors; ++c)
        for (int ch = 0; ch < 3; ++ch) a.set_color[c][ch] = objs->colors[c][ch];
    for (int k = 0; k < objs->num_objects; ++k) {
        a.color_index[k] = objs->color_index[k];
        for (int ch = 0; ch < 3; ++ch) {
            // dc = RGB2SH(colour) (GSP/utils/sh_utils.py:114), then the rasterizer's C0*dc + 0.5, clamp >= 0
            volatile float dc = (objs->colors[objs->color_index[k]][ch] - 0.5f) / C0;
            volatile float m = C0 * dc;
            volatile float e = m + 0.5f;
            a.eff_color[k][ch] = e < 0.0f ? 0.0f : e;
        }
    }
    return launch_composite(a, gy, true, stream);
}

}  // namespace pg
