// composite3.cu — the default compositing kernel (SURVEY Appendix A.7 + the fused K+3 passes).  Same CTA shape as
// composite2_kernel (composite.cu): one CTA per 16x16 tile, a producer warp staging the tile's sorted list into a
// shared-memory ring (TMA bulk copies of the id list, cp.async gathers of the 48-byte records), 4 consumer warps of
// 8x8 pixels, two pixels per lane in packed FP32.  What differs is how a consumer warp spends its instructions.
//
// Measured on composite2_kernel (round 2, C2 workload, debug-bit-1 counters): 7.3 M warp-hits per frame, of which
// 4.9 M are environment entries, 0.8 M object entries met while a main (RGB) chain is alive and 1.6 M object entries
// met AFTER every main chain of the warp has terminated (the objects-only chains of the fused mask passes keep
// walking).  That last phase found 1.6 hits per 32-entry cull pass, so most pairs of hits ran half empty, each pass
// paid its 65-instruction cull, and every hit dragged the dead main chain along: 22 % of the hits cost 44 % of the
// kernel's instructions.  Here, per staged batch of up to 128 entries:
//   1. CULL: all four 32-entry chunks are tested lane-parallel against the warp's pixel block back to back (the
//      four tests interleave) and the hits are appended to a per-warp queue in shared memory, so hits pair up across
//      chunk boundaries (one half-empty pair per batch at most);
//   2. WALK: one of two loops, chosen once per batch: the MAIN loop (RGB + depth chain, object chains on object hits)
//      while a main chain of the warp is alive, else the OBJECT loop, which evaluates alpha and updates only the
//      objects-only chains (no colour record, no main-chain arithmetic).
// A finished main chain is a per-pixel threshold folded into the validity test (alpha := 0 makes every update an
// exact no-op); records whose opacity is <= 0.99 skip min(0.99, .), which cannot bind for them (PG_REC_GENERAL).
// A silhouette chain is also finished as soon as its mask bit is decided: the bit is ||c_k (1 - T_k) + T_k bg - c_k||
// <= 0.1, T_k only decreases, so once T_k is below 0.0999 / ||c_k - bg|| the bit is 1 whatever follows.
// Every float operation of the default (exact) instantiation is the same individually rounded operation in the same
// order as in composite2_kernel, so images, final_T and n_contrib stay bit-identical to the CPU oracle.
//
// FAST instantiation (pg_launch_opts.numerics = PG_NUMERICS_FAST): exp through MUFU ex2.approx and the blend weight
// alpha * T formed once per hit (C += c * (alpha * T) instead of (c * alpha) * T).  Not bit-reproducible on a CPU;
// within the reference's own tolerance (its expf is MUFU-based too) — tests compare it with the north_star bounds.
#include "composite_common.cuh"

namespace pg {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// alpha where this half blends, else 0.  A.7's three skips — !(power > 0), !(power < cut), !(alpha < thr) — as one
// predicate chain (NaNs pass, as in the reference); thr is 1/255 while the pixel's main chain is alive and +inf once
// it has terminated, so "chain alive" costs no instruction.  PTX, because nvcc turns the C++ form into a cascade
// of selects.
__device__ __forceinline__ float valid_alpha(float power, float cut, float alpha, float thr) {
    float r;
    asm("{\n.reg .pred q;\n"
        "setp.leu.f32 q, %1, 0f00000000;\n"
        "setp.geu.and.f32 q, %1, %2, q;\n"
        "setp.geu.and.f32 q, %3, %4, q;\n"
        "selp.f32 %0, %3, 0f00000000, q;\n}\n"
        : "=f"(r) : "f"(power), "f"(cut), "f"(alpha), "f"(thr));
    return r;
}
constexpr float THR_ALIVE = 1.0f / 255.0f;          // 0x3B808081, the reference's alpha threshold
#define THR_DEAD __int_as_float(0x7f800000)         // +inf: no finite alpha passes

// block_culled (pg_common.cuh) with MUFU reciprocals: the 1-ulp error of the two clamped minimisers changes the
// evaluated minimum by a relative 1e-14 (second order), far inside the test's 1e-4 + 1e-3 margin.
__device__ __forceinline__ bool block_culled_fast(float gx, float gy, float qa, float qb, float qc, float cut,
                                                  float x0, float x1, float y0, float y1) {
    const float xl = x0 - gx, xh = x1 - gx, yl = y0 - gy, yh = y1 - gy;
    const float cx = fminf(fmaxf(0.0f, xl), xh), cy = fminf(fmaxf(0.0f, yl), yh);
    const float dy1 = fminf(fmaxf(-qb * cx * rcp_approx(qc), yl), yh);
    const float dx2 = fminf(fmaxf(-qb * cy * rcp_approx(qa), xl), xh);
    const float q1 = 0.5f * (qa * cx * cx + qc * dy1 * dy1) + qb * cx * dy1;
    const float q2 = 0.5f * (qa * dx2 * dx2 + qc * cy * cy) + qb * dx2 * cy;
    const float qmin = fminf(q1, q2);
    // denormal / huge conics: the approximate reciprocal may flush or overflow; such Gaussians are never culled
    const bool pd = qa > 1e-30f && qc > 1e-30f && qa < 1e30f && qc < 1e30f && qa * qc - qb * qb > 0.0f;
    return pd && (qmin * 0.9999f - 1e-3f > -cut);
}


constexpr int QSTRIDE = COMP_BATCH + 16;  // bytes of one warp's hit queue (entry indices; 0xFF = no second hit)

// Dynamic shared memory: CompSmem | hit queues [COMP2_CW][QSTRIDE] | (MASKS) per-warp, per-object boxes
// [COMP2_CW][PG_MAX_OBJECTS] float4 | (MASKS) eff[PG_MAX_OBJECTS] float4 {colour the
// rasterizer produces for object k's flat SH, silhouette decision threshold} | Tk2[K][128] float2 — the standalone
// transmittance of object k at the two pixels of each lane (slot = warp * 32 + lane).
template <bool MASKS, bool NCONTRIB, bool FAST, int COMP_STAGES, int MINB>
__global__ void __launch_bounds__(COMP2_THREADS, MINB) composite3_kernel(const CompArgs a) {
    using CompSmem = CompSmemT<COMP_STAGES, !MASKS>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CompSmem& sm = *reinterpret_cast<CompSmem*>(smem_raw);
    constexpr int kQueueOff = (int)((sizeof(CompSmem) + 15) / 16 * 16);
    constexpr int kBoxOff = kQueueOff + COMP2_CW * QSTRIDE;
    constexpr int kEffOff = kBoxOff + (MASKS ? COMP2_CW * PG_MAX_OBJECTS * (int)sizeof(float4) : 0);
    float4* sm_eff = reinterpret_cast<float4*>(smem_raw + kEffOff);
    float2* sm_tk = reinterpret_cast<float2*>(smem_raw + kEffOff + PG_MAX_OBJECTS * sizeof(float4));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = (int)a.tile_order[blockIdx.x];
    const int tile_x = tile % a.gx, tile_y = tile / a.gx;
    const uint2 range = a.ranges[tile];
    const int n = (int)(range.y - range.x);
    const uint32_t lt = (1u << lane) - 1u;

    if (tid == 0) {
        for (int s = 0; s < COMP_STAGES; ++s) {
            mbar_init(reinterpret_cast<uint64_t*>(&sm.full[s]), 33);
            mbar_init(reinterpret_cast<uint64_t*>(&sm.empty[s]), COMP2_CW);
        }
        mbar_init(reinterpret_cast<uint64_t*>(&sm.idbar[0]), 1);
        mbar_init(reinterpret_cast<uint64_t*>(&sm.idbar[1]), 1);
        sm.warps_done = 0;
        sm.warps_main_done = 0;
        sm.dummy.a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        sm.dummy.b = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(__float_as_int(-80.0f) & ~PG_REC_FLAGS));  // opacity 0: never blends
        sm.dummy.c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (MASKS && tid < a.num_objects) {
        // silhouette of object k: || c_k (1 - T_k) + T_k bg - c_k || = T_k || bg - c_k || <= 0.1 (src/gs/render.py:60-63);
        // below thr_k = max(1e-4, 0.0999 / || bg - c_k ||) the bit is decided (the 1e-3 margin dwarfs the rounding of
        // the final test and the 1-ulp gap between the colour set and the rasterized flat colour)
        const float e0 = a.eff_color[tid][0], e1 = a.eff_color[tid][1], e2 = a.eff_color[tid][2];
        const float d0 = a.bg[0] - e0, d1 = a.bg[1] - e1, d2 = a.bg[2] - e2;
        const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        const float tau = nrm > 0.0999f ? 0.0999f / nrm : 2.0f;  // 2: decided from the start (T_k <= 1)
        sm_eff[tid] = make_float4(e0, e1, e2, fmaxf(0.0001f, tau));
    }
    __syncthreads();

    if (warp == COMP2_CW) {
        composite_producer<MASKS, COMP_STAGES, COMP2_CW, 0>(a, sm, tile, range, n, lane, lt);
        return;
    }

    // =========================== CONSUMERS ===========================
    const int wx0 = tile_x * PG_TILE + (warp & 1) * 8, wy0 = tile_y * PG_TILE + (warp >> 1) * 8;
    const int px = wx0 + (lane & 7), py0 = wy0 + (lane >> 3) * 2, py1 = py0 + 1;
    const bool in0 = px < a.W && py0 < a.H, in1 = px < a.W && py1 < a.H;
    float pfx = (float)px;
    asm volatile("" : "+f"(pfx));
    const f32x2 npfy = pk2(-(float)py0, -(float)py1);
    const int K = MASKS ? a.num_objects : 0;
    const uint32_t all_k = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);
    float2* my_tk = sm_tk + (warp * 32 + lane);  // object k's silhouette chains at my_tk[k * 128]
    uint8_t* const queue = smem_raw + kQueueOff + warp * QSTRIDE;
    float4* const my_box = reinterpret_cast<float4*>(smem_raw + kBoxOff) + warp * PG_MAX_OBJECTS;  // MASKS: per object

    // Main chain: T stays the transmittance (frozen once the chain has terminated); thr0 / thr1 say whether it is alive
    // (1/255: alive, +inf: finished — see valid_alpha).  Object chains (To, Tk): a finished chain holds -|T|.
    f32x2 T = bc2(1.0f);
    float thr0 = in0 ? THR_ALIVE : THR_DEAD, thr1 = in1 ? THR_ALIVE : THR_DEAD;
    f32x2 To = pk2((in0 && MASKS) ? 1.0f : -1.0f, (in1 && MASKS) ? 1.0f : -1.0f);
    f32x2 C0 = bc2(0.0f), C1 = C0, C2 = C0, D = C0, S0 = C0, S1 = C0, S2 = C0;
    uint32_t dk0 = (in0 && MASKS) ? 0u : 0xFFFFFFFFu, dk1 = (in1 && MASKS) ? 0u : 0xFFFFFFFFu;
    if (MASKS)
        for (int k = 0; k < K; ++k) {
            // a chain whose bit is decided from the start (colour within 0.1 of the background) never runs
            const bool decided = sm_eff[k].w > 1.0f;
            my_tk[k * 128] = make_float2((in0 && !decided) ? 1.0f : -1.0f, (in1 && !decided) ? 1.0f : -1.0f);
            if (decided) { dk0 |= 1u << k; dk1 |= 1u << k; }
        }
    uint32_t last0 = 0, last1 = 0;
    bool w_main_done = false, w_done = false;
    const f32x2 ONE = bc2(1.0f), MONE = bc2(-1.0f);

    // power of one Gaussian at the lane's two pixels (A.7 operation order per half)
    auto power_of = [&](const float4& A, const float4& B) -> f32x2 {
        const float dx = sub(A.x, pfx);
        const float t1 = mul(A.z, dx);     // conic.x * dx
        const float nt2 = mul(-A.w, dx);   // -(conic.y * dx)
        const f32x2 dy = add2(bc2(A.y), npfy);
        const f32x2 w = mul2(dy, mul2(bc2(B.x), dy));
        const f32x2 sq = fma2(bc2(dx), bc2(t1), w);
        const f32x2 nbxy = mul2(bc2(nt2), dy);
        return fma2(sq, bc2(-0.5f), nbxy);
    };
    // opacity * exp(power) on both halves, before the min(0.99, .)
    auto raw_alpha = [&](f32x2 pw, float op, float& a0, float& a1) {
        if (FAST) {
            float q0, q1;
            unpk2(mul2(pw, bc2(1.44269502162933349609375f)), q0, q1);
            unpk2(mul2(pk2(ex2_approx(q0), ex2_approx(q1)), bc2(op)), a0, a1);
        } else {
            float e0, e1;
            exp2_exact_nz(pw, e0, e1);
            unpk2(mul2(pk2(e0, e1), bc2(op)), a0, a1);
        }
    };
    // C += c * am * Told for the four accumulated channels
    auto accumulate = [&](const float4& Cc, float depth, f32x2 am, f32x2 Told) {
        if (FAST) {
            const f32x2 w = mul2(am, Told);
            C0 = fma2(bc2(Cc.x), w, C0);
            C1 = fma2(bc2(Cc.y), w, C1);
            C2 = fma2(bc2(Cc.z), w, C2);
            D = fma2(bc2(depth), w, D);
        } else {
            C0 = fma2(mul2(bc2(Cc.x), am), Told, C0);
            C1 = fma2(mul2(bc2(Cc.y), am), Told, C1);
            C2 = fma2(mul2(bc2(Cc.z), am), Told, C2);
            D = fma2(mul2(bc2(depth), am), Told, D);
        }
    };
    // the objects-only chains of one object hit: av = alpha where the half blends at all (A.7's three skips,
    // independent of the main chain), else 0
    auto object_chains = [&](int obj, float av0, float av1) {
        const f32x2 om = fma2(pk2(av0, av1), MONE, ONE);
        const float4 ec = sm_eff[obj - 1];
        {   // objects-only render (visible masks, sem-seg)
            const f32x2 tT = mul2(To, om);
            float n0, n1, o0, o1;
            unpk2(tT, n0, n1);
            unpk2(To, o0, o1);
            const bool d0 = n0 < 0.0001f, d1 = n1 < 0.0001f;  // also true for a finished (negative) chain
            const f32x2 am = pk2(d0 ? 0.0f : av0, d1 ? 0.0f : av1);
            const f32x2 Told = To;
            To = pk2(d0 ? -fabsf(o0) : n0, d1 ? -fabsf(o1) : n1);
            // a finished chain has Told < 0 and am == 0: the product is -0, the sum unchanged
            S0 = fma2(mul2(bc2(ec.x), am), Told, S0);
            S1 = fma2(mul2(bc2(ec.y), am), Told, S1);
            S2 = fma2(mul2(bc2(ec.z), am), Told, S2);
        }
        // silhouette chain of this object: terminated (T (1 - alpha) < 1e-4: keeps T, as the reference's pass would) or
        // decided (below the object's threshold: keeps the new value); both are stored negative = finished
        const uint32_t kbit = 1u << ((uint32_t)(obj - 1) & 31u);
        const float2 tk = my_tk[(obj - 1) * 128];
        const f32x2 tT = mul2(pk2(tk.x, tk.y), om);
        float n0, n1;
        unpk2(tT, n0, n1);
        const bool d0 = n0 < ec.w, d1 = n1 < ec.w;          // finished now (or before: n <= 0)
        if (d0) dk0 |= kbit;
        if (d1) dk1 |= kbit;
        const float f0 = n0 < 0.0001f ? -fabsf(tk.x) : -n0, f1 = n1 < 0.0001f ? -fabsf(tk.y) : -n1;
        my_tk[(obj - 1) * 128] = make_float2(d0 ? f0 : n0, d1 ? f1 : n1);
    };

    for (int it = 0;; ++it) {
        const int s = it % COMP_STAGES;
        mbar_wait(reinterpret_cast<uint64_t*>(&sm.full[s]), (uint32_t)((it / COMP_STAGES) & 1));
        const int cnt = *(volatile int*)&sm.cnt[s];
        if (cnt == 0) break;
        if (!w_done) {
            const GeomRec* sr = sm.rec[s];
            const bool wm = !w_main_done;  // a main chain of this warp is alive: every entry is wanted
            // objects some pixel of this warp still needs (bit k-1): every object while an objects-only render chain is
            // alive, else those whose silhouette chain is alive somewhere
            uint32_t need = 0;
            if (MASKS && !wm) {
                float to0, to1;
                unpk2(To, to0, to1);
                need = __any_sync(0xffffffffu, to0 > 0.0f || to1 > 0.0f)
                           ? all_k : (__reduce_or_sync(0xffffffffu, ~(dk0 & dk1)) & all_k);
            }
            // ---- 1. cull the batch against the bounding box of this warp's pixels it can still change: pixels saturate
            // one by one, the box shrinks, and a Gaussian that only touches finished pixels is a no-op.  An environment
            // entry can change pixels whose main chain is alive; an entry of object k also those whose objects-only
            // render chain or whose silhouette chain k is alive (one box per object, in shared memory).
            float ax0, ax1, ay0, ay1;
            {
                const bool m0 = thr0 == THR_ALIVE, m1 = thr1 == THR_ALIVE;
                const int big = 1 << 30;
                auto box = [&](bool l0, bool l1, float& x0, float& x1, float& y0, float& y1) {
                    x0 = (float)__reduce_min_sync(0xffffffffu, (l0 || l1) ? px : big);
                    x1 = (float)__reduce_max_sync(0xffffffffu, (l0 || l1) ? px : -big);
                    y0 = (float)__reduce_min_sync(0xffffffffu, l0 ? py0 : (l1 ? py1 : big));
                    y1 = (float)__reduce_max_sync(0xffffffffu, l1 ? py1 : (l0 ? py0 : -big));
                };
                box(m0, m1, ax0, ax1, ay0, ay1);
                if (MASKS) {
                    float to0, to1;
                    unpk2(To, to0, to1);
                    const bool o0 = m0 || to0 > 0.0f, o1 = m1 || to1 > 0.0f;
                    for (int k = 0; k < K; ++k) {
                        float x0, x1, y0, y1;
                        box(o0 || !((dk0 >> k) & 1u), o1 || !((dk1 >> k) & 1u), x0, x1, y0, y1);
                        if (lane == 0) my_box[k] = make_float4(x0, x1, y0, y1);
                    }
                    __syncwarp();
                }
            }
            int nq = 0;
            if (wm || need != 0) {
#pragma unroll
                for (int c = 0; c < COMP_BATCH / 32; ++c) {
                    const int e = c * 32 + lane;
                    bool hit = false;
                    if (e < cnt) {
                        const float4 A = sr[e].a;
                        const float4 B = sr[e].b;
                        const int eo = MASKS ? (__float_as_int(B.w) & 63) : 0;
                        const bool wanted = wm || (MASKS && eo > 0 && ((need >> ((uint32_t)(eo - 1) & 31u)) & 1u));
                        float4 bb = make_float4(ax0, ax1, ay0, ay1);
                        if (MASKS && eo > 0) bb = my_box[eo - 1];
                        hit = wanted && !block_culled_fast(A.x, A.y, A.z, A.w, B.x, B.w, bb.x, bb.y, bb.z, bb.w);
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, hit);
                    if (hit) queue[nq + __popc(m & lt)] = (uint8_t)e;
                    nq += __popc(m);
                }
                if (lane == 0) queue[nq] = 0xFF;  // completes an odd number of hits
                __syncwarp();
            }
            // ---- 2. walk the hits two at a time
            int i_obj = wm ? nq : 0;  // where the objects-only loop takes over
            if (wm) {
#pragma unroll 1
                for (int i = 0; i < nq; i += 2) {
                    const uint32_t pr = *reinterpret_cast<const uint16_t*>(queue + i);
                    const uint32_t e1 = pr & 255u, e2 = pr >> 8;
                    const GeomRec* r1 = sr + e1;
                    const GeomRec* r2 = e2 == 255u ? &sm.dummy : sr + e2;  // a record that never blends (opacity 0)
                    const float4 A1 = r1->a, B1 = r1->b, A2 = r2->a, B2 = r2->b;
                    const float4 Cc1 = r1->c, Cc2 = r2->c;
                    const f32x2 p1 = power_of(A1, B1), p2 = power_of(A2, B2);
                    const int fl = (__float_as_int(B1.w) | __float_as_int(B2.w)) & PG_REC_FLAGS;  // warp-uniform
                    float a10, a11, a20, a21, q10, q11, q20, q21;
                    raw_alpha(p1, B1.y, a10, a11);
                    raw_alpha(p2, B2.y, a20, a21);
                    unpk2(p1, q10, q11);
                    unpk2(p2, q20, q21);
                    if (fl & PG_REC_GENERAL) {
                        // A.7's min(0.99, .): cannot bind for the other records (opacity <= 0.99, exp(power) <= 1)
                        a10 = fminf(0.99f, a10); a11 = fminf(0.99f, a11);
                        a20 = fminf(0.99f, a20); a21 = fminf(0.99f, a21);
                    }
                    if (MASKS && (fl & 63)) {
                        const int o1 = __float_as_int(B1.w) & 63, o2 = __float_as_int(B2.w) & 63;
                        if (o1) object_chains(o1, valid_alpha(q10, B1.w, a10, THR_ALIVE), valid_alpha(q11, B1.w, a11, THR_ALIVE));
                        if (o2) object_chains(o2, valid_alpha(q20, B2.w, a20, THR_ALIVE), valid_alpha(q21, B2.w, a21, THR_ALIVE));
                    }
                    // main chain (RGB + depth): every hit tests its own termination — a chain that would drop below 1e-4
                    // does not blend the hit and is dead from then on (a dead chain has alpha 0, so tT == T >= 1e-4)
                    float n0, n1, o0, o1;
                    float av10 = valid_alpha(q10, B1.w, a10, thr0), av11 = valid_alpha(q11, B1.w, a11, thr1);
                    unpk2(mul2(T, fma2(pk2(av10, av11), MONE, ONE)), n0, n1);
                    unpk2(T, o0, o1);
                    const bool d10 = n0 < 0.0001f, d11 = n1 < 0.0001f;
                    av10 = d10 ? 0.0f : av10; thr0 = d10 ? THR_DEAD : thr0;
                    av11 = d11 ? 0.0f : av11; thr1 = d11 ? THR_DEAD : thr1;
                    const f32x2 T1 = pk2(d10 ? o0 : n0, d11 ? o1 : n1);
                    float av20 = valid_alpha(q20, B2.w, a20, thr0), av21 = valid_alpha(q21, B2.w, a21, thr1);
                    unpk2(mul2(T1, fma2(pk2(av20, av21), MONE, ONE)), n0, n1);
                    unpk2(T1, o0, o1);
                    const bool d20 = n0 < 0.0001f, d21 = n1 < 0.0001f;
                    av20 = d20 ? 0.0f : av20; thr0 = d20 ? THR_DEAD : thr0;
                    av21 = d21 ? 0.0f : av21; thr1 = d21 ? THR_DEAD : thr1;
                    const f32x2 T2 = pk2(d20 ? o0 : n0, d21 ? o1 : n1);
                    accumulate(Cc1, B1.z, pk2(av10, av11), T);
                    accumulate(Cc2, B2.z, pk2(av20, av21), T1);
                    T = T2;
                    if (NCONTRIB) {
                        const uint32_t pos1 = sm.pos[s][e1], pos2 = e2 == 255u ? 0u : sm.pos[s][e2];
                        if (av10 > 0.0f) last0 = pos1;
                        if (av11 > 0.0f) last1 = pos1;
                        if (av20 > 0.0f) last0 = pos2;
                        if (av21 > 0.0f) last1 = pos2;
                    }
                    // every 4th pair: once all main chains of the warp are dead the rest of the batch is no-ops
                    // (object entries still have to reach their chains, so only without masks)
                    if ((i & 6) == 6 && !__any_sync(0xffffffffu, thr0 == THR_ALIVE || thr1 == THR_ALIVE)) {
                        i_obj = i + 2;  // the rest of the batch can only matter to the objects-only chains
                        break;
                    }
                }
                if (!__any_sync(0xffffffffu, thr0 == THR_ALIVE || thr1 == THR_ALIVE)) {
                    w_main_done = true;
                    if (lane == 0) atomicAdd(&sm.warps_main_done, 1);
                }
            }
            if (MASKS && i_obj < nq) {
                // every main chain of the warp has terminated: only the objects-only chains run, on object entries
                // (a batch culled while a main chain was alive still has environment entries queued: skipped)
#pragma unroll 1
                for (int i = i_obj; i < nq; i += 2) {
                    const uint32_t pr = *reinterpret_cast<const uint16_t*>(queue + i);
                    const uint32_t e1 = pr & 255u, e2 = pr >> 8;
                    const GeomRec* r1 = sr + e1;
                    const GeomRec* r2 = e2 == 255u ? &sm.dummy : sr + e2;
                    const float4 B1 = r1->b, B2 = r2->b;
                    const int fl = (__float_as_int(B1.w) | __float_as_int(B2.w)) & PG_REC_FLAGS;
                    if (!(fl & 63)) continue;
                    const float4 A1 = r1->a, A2 = r2->a;
                    const f32x2 p1 = power_of(A1, B1), p2 = power_of(A2, B2);
                    float a10, a11, a20, a21, q10, q11, q20, q21;
                    raw_alpha(p1, B1.y, a10, a11);
                    raw_alpha(p2, B2.y, a20, a21);
                    unpk2(p1, q10, q11);
                    unpk2(p2, q20, q21);
                    if (fl & PG_REC_GENERAL) {
                        a10 = fminf(0.99f, a10); a11 = fminf(0.99f, a11);
                        a20 = fminf(0.99f, a20); a21 = fminf(0.99f, a21);
                    }
                    const int o1 = __float_as_int(B1.w) & 63, o2 = __float_as_int(B2.w) & 63;
                    if (o1) object_chains(o1, valid_alpha(q10, B1.w, a10, THR_ALIVE), valid_alpha(q11, B1.w, a11, THR_ALIVE));
                    if (o2) object_chains(o2, valid_alpha(q20, B2.w, a20, THR_ALIVE), valid_alpha(q21, B2.w, a21, THR_ALIVE));
                }
            }
            // ---- report progress to the producer
            if (w_main_done) {
                float o0, o1;
                unpk2(To, o0, o1);
                const bool pix_done = !MASKS || (o0 < 0.0f && o1 < 0.0f && (dk0 & dk1 & all_k) == all_k);
                if (__all_sync(0xffffffffu, pix_done)) {
                    w_done = true;
                    if (lane == 0) atomicAdd(&sm.warps_done, 1);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }

    const size_t HW = (size_t)a.W * a.H;
    const float bg0 = a.bg[0], bg1 = a.bg[1], bg2 = a.bg[2];
    float Th[2], Toh[2], Ch[3][2], Dh[2], Sh[3][2];
    unpk2(T, Th[0], Th[1]); unpk2(To, Toh[0], Toh[1]);
    unpk2(C0, Ch[0][0], Ch[0][1]); unpk2(C1, Ch[1][0], Ch[1][1]); unpk2(C2, Ch[2][0], Ch[2][1]);
    unpk2(D, Dh[0], Dh[1]);
    unpk2(S0, Sh[0][0], Sh[0][1]); unpk2(S1, Sh[1][0], Sh[1][1]); unpk2(S2, Sh[2][0], Sh[2][1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (!(h == 0 ? in0 : in1)) continue;
        const float Tf = Th[h], Tof = fabsf(Toh[h]);
        const size_t pix = (size_t)(h == 0 ? py0 : py1) * a.W + px;
        a.out_color[pix] = fma(Tf, bg0, Ch[0][h]);
        a.out_color[HW + pix] = fma(Tf, bg1, Ch[1][h]);
        a.out_color[2 * HW + pix] = fma(Tf, bg2, Ch[2][h]);
        a.out_depth[pix] = Dh[h];
        if (a.out_final_T) a.out_final_T[pix] = Tf;
        if (NCONTRIB && a.out_n_contrib) a.out_n_contrib[pix] = h == 0 ? last0 : last1;
        if (MASKS) {
            const float s0 = fma(Tof, bg0, Sh[0][h]), s1 = fma(Tof, bg1, Sh[1][h]), s2 = fma(Tof, bg2, Sh[2][h]);
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
                for (int kk = 0; kk < K; ++kk) {
                    const int ci = a.color_index[kk];
                    const float2 t2 = my_tk[kk * 128];
                    const float tk = fabsf(h == 0 ? t2.x : t2.y);
                    const float4 ec = sm_eff[kk];
                    const float w = sub(1.0f, tk);
                    float i0 = fma(tk, bg0, mul(ec.x, w));
                    float i1 = fma(tk, bg1, mul(ec.y, w));
                    float i2 = fma(tk, bg2, mul(ec.z, w));
                    float d0 = sub(i0, a.set_color[ci][0]), d1 = sub(i1, a.set_color[ci][1]), d2 = sub(i2, a.set_color[ci][2]);
                    float dist = sqrt(add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)));
                    a.silhouette[(size_t)ci * HW + pix] = dist <= 0.1f ? 1 : 0;
                }
            }
        }
    }
}

template <bool MASKS, bool NCONTRIB, bool FAST, int STAGES, int MINB>
static int launch_three(const CompArgs& a, dim3 grid, cudaStream_t stream) {
    const int smem = (int)((sizeof(CompSmemT<STAGES, !MASKS>) + 15) / 16 * 16) + COMP2_CW * QSTRIDE +
                     (MASKS ? (int)((COMP2_CW + 1) * PG_MAX_OBJECTS * sizeof(float4)) + a.num_objects * 256 * (int)sizeof(float) : 0);
    PG_CUDA_CHECK(ensure_dynamic_smem(composite3_kernel<MASKS, NCONTRIB, FAST, STAGES, MINB>, smem, true));
    composite3_kernel<MASKS, NCONTRIB, FAST, STAGES, MINB><<<grid, COMP2_THREADS, smem, stream>>>(a);
    return PG_OK;
}

// masks: fused K+3 passes; fast: PG_NUMERICS_FAST
// CTAs per SM the register allocation is bounded for.  With masks the kernel wants 87 registers; bounded for 5 CTAs
// per SM (72 registers, 4 of them spilled) it runs 0.899 instead of 0.948 ms (fast; exact 1.008 vs 1.057): 20 instead
// of 16 consumer warps per SM hide more of the dependent chains.  6 CTAs (64 registers, 12 spilled) 0.956,
// 7 CTAs 1.173; __maxnreg__(80), spill-free, 0.923.  The plain kernel needs 51-64 registers: unaffected.
#ifndef PG_COMP3_MINB
#define PG_COMP3_MINB 5
#endif
#ifndef PG_COMP3_MINB_PLAIN
#define PG_COMP3_MINB_PLAIN 4
#endif

int launch_composite3(const CompArgs& a, dim3 grid, bool masks, bool fast, int variant, cudaStream_t stream) {
    (void)variant;
    constexpr int MB = PG_COMP3_MINB, MBP = PG_COMP3_MINB_PLAIN;
    if (!masks) {
        const bool nc = a.out_n_contrib != nullptr;
        if (fast) return nc ? launch_three<false, true, true, 4, MBP>(a, grid, stream) : launch_three<false, false, true, 4, MBP>(a, grid, stream);
        return nc ? launch_three<false, true, false, 4, MBP>(a, grid, stream) : launch_three<false, false, false, 4, MBP>(a, grid, stream);
    }
    // the tighter register bound only pays when shared memory lets the extra CTA in (227 KB per SM, 1 KB per CTA
    // reserved by the runtime: up to 13 objects); frames with more objects keep the spill-free 4-CTA build
    const int smem = (int)((sizeof(CompSmemT<4, false>) + 15) / 16 * 16) + COMP2_CW * QSTRIDE +
                     (int)((COMP2_CW + 1) * PG_MAX_OBJECTS * sizeof(float4)) + a.num_objects * 256 * (int)sizeof(float);
    if (MB > 4 && MB * (smem + 1024) > 227 * 1024)
        return fast ? launch_three<true, false, true, 4, 4>(a, grid, stream) : launch_three<true, false, false, 4, 4>(a, grid, stream);
    return fast ? launch_three<true, false, true, 4, MB>(a, grid, stream) : launch_three<true, false, false, 4, MB>(a, grid, stream);
}

}  // namespace pg
