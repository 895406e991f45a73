// composite_common.cuh — pieces shared by the compositing kernels (composite.cu, composite3.cu): mbarrier / TMA
// helpers, kernel arguments, the shared-memory ring, the producer warp, the bit-reproducible exp and the packed
// FP32 (f32x2) wrappers.
#pragma once
#include <cstdlib>
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking wait: try_wait with a suspend-time hint parks the warp in hardware (no issue slots burnt
// while other warps of the SM have work) and is re-armed until the phase completes.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
// WAITNS > 0: back off with nanosleep between polls instead of re-arming the suspended try_wait
// (a waiting warp then issues ~1 instruction per WAITNS ns instead of 3 per hardware time-out).
template <int WAITNS>
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity) {
    if (WAITNS == 0) mbar_wait(bar, parity);
    else
        while (!mbar_try_wait(bar, parity)) __nanosleep(WAITNS);
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
    const uint32_t* tile_order;  // launch order: CTA b composites tile tile_order[b]
    const uint32_t* point_list;  // bit 31 = PG_CULL_FLAG
    const GeomRec* recs;
    int W, H, gx;
    const float* bg;
    float* out_color;
    float* out_depth;
    float* out_final_T;
    uint32_t* out_n_contrib;
    // masks
    uint32_t n_env;                 // Gaussian indices >= n_env belong to objects
    const uint32_t* tile_obj_count; // [tiles][OBJ_SPREAD] partial counts of the un-culled object pairs per tile
    int num_objects, num_colors;
    float eff_color[PG_MAX_OBJECTS][3];  // colour the rasterizer produces for object k's flat SH
    float set_color[PG_MAX_COLORS][3];   // colour set the masks are tested against
    int color_index[PG_MAX_OBJECTS];
    float* seg_color;
    uint8_t* sem_seg;
    uint8_t* visible;
    uint8_t* silhouette;
    unsigned long long* stats;  // non-null: count pairs evaluated / exp'd / blended
    int fast;                   // PG_NUMERICS_FAST: MUFU exp, alpha * T formed once per hit (composite3_kernel)
};

__device__ __forceinline__ void flush_stats(unsigned long long* stats, uint32_t n_eval, uint32_t n_exp, uint32_t n_blend,
                                            uint32_t n_slots = 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
        n_exp += __shfl_xor_sync(0xffffffffu, n_exp, o);
        n_blend += __shfl_xor_sync(0xffffffffu, n_blend, o);
        n_slots += __shfl_xor_sync(0xffffffffu, n_slots, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], (unsigned long long)n_eval);
        atomicAdd(&stats[1], (unsigned long long)n_exp);
        atomicAdd(&stats[2], (unsigned long long)n_blend);
        atomicAdd(&stats[3], (unsigned long long)n_slots);  // pixel slots the walked hits occupied (hits x pixels per warp)
    }
}

// Warp-specialised: the CTA's last warp is the PRODUCER (TMA-prefetches the tile's ids in 1 KB chunks, drops
// culled / no-longer-needed entries by ballot compaction, gathers each surviving 48-byte record into
// a 4-stage shared-memory ring with cp.async whose completion arrives on the stage's `full`
// mbarrier); warps 0..3 are CONSUMERS, each owning an 8x8 pixel block (two pixels per lane): wait `full`, cull
// the batch lane-parallel against the block, walk the hits, arrive on `empty`.  No CTA-wide barrier in the steady
// state; a consumer whose pixels are finished keeps releasing stages, the producer stops when all have finished.
constexpr int COMP_BATCH = 128;

constexpr int COMP_IDCHUNK = 256;
constexpr int COMP_PEND = 512;
static_assert(COMP_BATCH - 1 + COMP_IDCHUNK <= COMP_PEND, "pending ring too small");

// POS: the 1-based list position of every staged entry travels with it (n_contrib of the plain pass); the fused
// frame has no use for it and its CTAs are 4 KB smaller, which is what lets a fifth CTA per SM in up to 13 objects.
template <int COMP_STAGES, bool POS>
struct CompSmemT {
    GeomRec rec[COMP_STAGES][COMP_BATCH];
    uint32_t pos[POS ? COMP_STAGES : 1][POS ? COMP_BATCH : 1];  // 1-based list position of each staged entry
    alignas(16) uint32_t ids[2][COMP_IDCHUNK];  // producer: raw id chunks, TMA double buffer (16-byte destinations)
    uint32_t pend[COMP_PEND];               // producer: compacted ids not yet staged (circular)
    uint32_t pendpos[POS ? COMP_PEND : 1];
    unsigned long long full[COMP_STAGES];
    unsigned long long empty[COMP_STAGES];
    unsigned long long idbar[2];
    int cnt[COMP_STAGES];
    GeomRec dummy;        // composite3_kernel: a record that never blends (completes an odd pair of hits)
    int warps_done;       // consumers whose pixels are completely finished
    int warps_main_done;  // consumers whose main chains are all finished
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// exp(x) for cut <= x <= 0 (cut >= -80.001): expf_exact without its x < -80 guard — identical results on
// that range, and below it alpha = opacity * 2^-115 is skipped by the alpha < 1/255 test either way.
__device__ __forceinline__ float expf_exact_nz(float x) {
    const float L2E = 1.44269502162933349609375f;
    const float MAGIC = 12582912.0f;
    float z = fma(x, L2E, MAGIC);
    float n = sub(z, MAGIC);
    float r = fma(n, -0.693145751953125f, x);
    r = fma(n, -1.428606765330187045e-06f, r);
    float p = 0x1.6b5016p-10f;
    p = fma(p, r, 0x1.126caep-7f);
    p = fma(p, r, 0x1.55578ep-5f);
    p = fma(p, r, 0x1.55540cp-3f);
    p = fma(p, r, 0x1.fffffcp-2f);
    p = fma(p, r, 1.0f);
    p = fma(p, r, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(z) << 23));  // MAGIC's low 9 bits are 0
}

// TUNING ONLY (PG_COMP_VARIANT 9): exp through MUFU ex2.approx — not reproducible on the CPU oracle, used
// to measure what the bit-exact polynomial costs; never the default.
template <bool FASTEXP>
__device__ __forceinline__ float exp_sel(float x) {
    if (FASTEXP) {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(mul(x, 1.44269502162933349609375f)));
        return y;
    }
    return expf_exact_nz(x);
}

// The producer warp of a compositing CTA (shared by the 1-pixel and the 2-pixel-per-thread kernels):
// TMA-prefetches the tile's ids in 1 KB chunks, drops culled / no-longer-needed entries by ballot
// compaction, gathers each surviving 48-byte record into the COMP_STAGES-deep shared-memory ring with
// cp.async whose completion arrives on the stage's `full` mbarrier.  NCW = number of consumer warps.
template <bool MASKS, int COMP_STAGES, int NCW, int WAITNS>
__device__ __forceinline__ void composite_producer(const CompArgs& a, CompSmemT<COMP_STAGES, !MASKS>& sm, const int tile,
                                                   const uint2 range, const int n, const int lane, const uint32_t lt) {
    // =========================== PRODUCER ===========================
    int obj_left = 0;  // un-culled object entries not yet compacted
    if (MASKS) {
        const uint4* oc = reinterpret_cast<const uint4*>(a.tile_obj_count + (size_t)tile * OBJ_SPREAD);
        static_assert(OBJ_SPREAD == 8, "two 16-byte loads");
        const uint4 c0 = oc[0], c1 = oc[1];
        obj_left = (int)(c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w);
    }
    // id chunks are fetched by TMA from a 16-byte aligned base; entries outside [range.x, range.y) are ignored
    const uint32_t a0 = range.x & ~3u;
    const int span = (int)(range.y - a0);
    const int nchunks = n > 0 ? (span + COMP_IDCHUNK - 1) / COMP_IDCHUNK : 0;
    auto fetch_ids = [&](int ci) {
        if (lane == 0) {
            const int left = span - ci * COMP_IDCHUNK;
            const uint32_t bytes = (uint32_t)((min(left, COMP_IDCHUNK) + 3) & ~3) * 4u;
            uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.idbar[ci & 1]);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(sm.ids[ci & 1], a.point_list + a0 + (size_t)ci * COMP_IDCHUNK, bytes, bar);
        }
    };
    if (nchunks > 0) fetch_ids(0);
    if (nchunks > 1) fetch_ids(1);
    int ci = 0, head = 0, fill = 0;
    for (int it = 0;; ++it) {
        const int s = it % COMP_STAGES;
        if (it >= COMP_STAGES)
            mbar_wait_t<(WAITNS < 0 ? -WAITNS : WAITNS)>(reinterpret_cast<uint64_t*>(&sm.empty[s]), (uint32_t)(((it / COMP_STAGES) - 1) & 1));
        const bool all_done = *(volatile int*)&sm.warps_done == NCW;
        const bool mode_all = !MASKS || *(volatile int*)&sm.warps_main_done < NCW;
        while (!all_done && fill < COMP_BATCH && ci < nchunks && (mode_all || obj_left > 0)) {
            mbar_wait(reinterpret_cast<uint64_t*>(&sm.idbar[ci & 1]), (uint32_t)((ci >> 1) & 1));
            const uint32_t* src = sm.ids[ci & 1];
            const uint32_t g0 = a0 + (uint32_t)ci * COMP_IDCHUNK;
#pragma unroll 4
            for (int u = 0; u < COMP_IDCHUNK / 32; ++u) {
                const uint32_t gi = g0 + u * 32 + lane;
                const uint32_t v = src[u * 32 + lane];
                const bool live = gi >= range.x && gi < range.y && !(v & PG_CULL_FLAG);
                const bool is_obj = MASKS && live && v >= a.n_env;
                const bool keep = live && (mode_all || is_obj);
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int slot = (head + fill + __popc(bal & lt)) & (COMP_PEND - 1);
                    sm.pend[slot] = v;
                    if (!MASKS) sm.pendpos[slot] = gi - range.x + 1u;
                }
                fill += __popc(bal);
                if (MASKS) obj_left -= __popc(__ballot_sync(0xffffffffu, is_obj));
            }
            __syncwarp();
            if (ci + 2 < nchunks) fetch_ids(ci + 2);
            ++ci;
        }
        __syncwarp();
        const int cnt = all_done ? 0 : min(fill, COMP_BATCH);
        if (lane == 0) sm.cnt[s] = cnt;
        if (!MASKS)
            for (int e = lane; e < cnt; e += 32) sm.pos[s][e] = sm.pendpos[(head + e) & (COMP_PEND - 1)];
        __syncwarp();
        // full[s] expects 33 arrivals: lane 0's release-arrive (publishes cnt / pos) + one per lane that
        // fires when that lane's cp.async gathers have landed
        if (lane == 0) mbar_arrive(&sm.full[s]);
        if (cnt == 0) {
            mbar_arrive(&sm.full[s]);  // end marker: plain arrivals
            break;
        }
        // gather: 3 x 16-byte cp.async per record (per-lane addresses issue as ordinary SIMT instructions;
        // a per-record TMA bulk copy needs uniform operands and serialises over the 32 lanes)
        for (int e = lane; e < cnt; e += 32) {
            const char* src = reinterpret_cast<const char*>(a.recs + sm.pend[(head + e) & (COMP_PEND - 1)]);
            const uint32_t dst = smem_u32(&sm.rec[s][e]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 16) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 32), "l"(src + 32) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sm.full[s])) : "memory");
        head = (head + cnt) & (COMP_PEND - 1);
        fill -= cnt;
        __syncwarp();
    }
    // drain id chunks that were fetched but never consumed (never exit with a copy in flight)
    for (int c2 = ci; c2 < min(nchunks, ci + 2); ++c2)
        mbar_wait(reinterpret_cast<uint64_t*>(&sm.idbar[c2 & 1]), (uint32_t)((c2 >> 1) & 1));
}


typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 bc2(float s) { return pk2(s, s); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// expf_exact_nz on both halves (same operation sequence per half)
__device__ __forceinline__ void exp2_exact_nz(f32x2 x, float& e0, float& e1) {
    const f32x2 z = fma2(x, bc2(1.44269502162933349609375f), bc2(12582912.0f));
    const f32x2 n = add2(z, bc2(-12582912.0f));
    f32x2 r = fma2(n, bc2(-0.693145751953125f), x);
    r = fma2(n, bc2(-1.428606765330187045e-06f), r);
    f32x2 p = bc2(0x1.6b5016p-10f);
    p = fma2(p, r, bc2(0x1.126caep-7f));
    p = fma2(p, r, bc2(0x1.55578ep-5f));
    p = fma2(p, r, bc2(0x1.55540cp-3f));
    p = fma2(p, r, bc2(0x1.fffffcp-2f));
    p = fma2(p, r, bc2(1.0f));
    p = fma2(p, r, bc2(1.0f));
    float p0, p1, z0, z1;
    unpk2(p, p0, p1);
    unpk2(z, z0, z1);
    e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(z0) << 23));
    e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(z1) << 23));
}

// A.7's three skips as ONE predicate chain: !(power > 0) && !(power < cut) && !(alpha < 1/255) (NaNs pass,
// as in the scalar kernel).  Written in PTX so that it stays 3 FSETP instead of a chain of selects.
__device__ __forceinline__ bool blends(float power, float cut, float alpha) {
    uint32_t r;
    asm("{\n.reg .pred q;\n"
        "setp.leu.f32 q, %1, 0f00000000;\n"
        "setp.geu.and.f32 q, %1, %2, q;\n"
        "setp.geu.and.f32 q, %3, 0f3B808081, q;\n"
        "selp.u32 %0, 1, 0, q;\n}\n"
        : "=r"(r) : "f"(power), "f"(cut), "f"(alpha));
    return r != 0;
}

constexpr int COMP2_CW = 4;                       // consumer warps: 8x8 pixel blocks of the 16x16 tile
constexpr int COMP2_THREADS = (COMP2_CW + 1) * 32;

}  // namespace pg
