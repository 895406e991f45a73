// composite3.cu — the default compositing kernel (SURVEY Appendix A.7 + the fused K+3 passes): same CTA shape as
// composite2_kernel (composite.cu: one CTA per 16x16 tile, a producer warp staging the sorted list into a
// shared-memory ring, 4 consumer warps of 8x8 pixels, two pixels per lane in packed FP32), with a leaner hit loop.
//
// What the hit loop of composite2_kernel spent per pair of hits (ncu, round 1): 68 packed FP32 instructions and ~60
// predicate / select / min / integer ones; both the FMA and the ALU pipe ran at 50 % with 4 consumer warps per
// scheduler.  Here:
//   * termination is tested once per PAIR of hits: T only decreases, so if T after both hits is still >= 1e-4 at every
//     pixel of the warp (one min, one compare, one vote) neither hit terminated any chain and both are blended without
//     per-hit compares / selects; otherwise the pair is redone hit by hit (at most once per pixel of the warp);
//   * a finished main chain is a predicate folded into the validity test (alpha = 0 makes every update an exact
//     no-op) instead of a sign trick that costs two selects per chain and hit;
//   * records whose opacity is <= 0.99 (flagged by preprocess, PG_REC_GENERAL clear) skip min(0.99, .): exp(power) <= 1
//     for power <= 0, so the minimum cannot bind;
//   * hits of object Gaussians, flagged records and pairs in which a chain terminates take the general path, which
//     is composite2_kernel's arithmetic hit by hit.
// Every float operation of the default (exact) instantiation is the same individually rounded operation in the same
// order as before, so images, final_T and n_contrib stay bit-identical to the CPU oracle.
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

template <bool MASKS, bool NCONTRIB, bool FAST, int COMP_STAGES, int MINB, bool BRANCHY>
__global__ void __launch_bounds__(COMP2_THREADS, MINB) composite3_kernel(const CompArgs a) {
    using CompSmem = CompSmemT<COMP_STAGES>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CompSmem& sm = *reinterpret_cast<CompSmem*>(smem_raw);
    float4* sm_eff = reinterpret_cast<float4*>(smem_raw + sizeof(CompSmem));
    float2* sm_tk = reinterpret_cast<float2*>(smem_raw + sizeof(CompSmem) + PG_MAX_OBJECTS * sizeof(float4));

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
    if (MASKS && tid < a.num_objects)
        sm_eff[tid] = make_float4(a.eff_color[tid][0], a.eff_color[tid][1], a.eff_color[tid][2], 0.0f);
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
    const float bx0 = (float)wx0, bx1 = (float)min(wx0 + 7, a.W - 1);
    const float by0 = (float)wy0, by1 = (float)min(wy0 + 7, a.H - 1);
    const int K = MASKS ? a.num_objects : 0;
    const uint32_t all_k = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);
    float2* my_tk = sm_tk + (warp * 32 + lane);  // object k's silhouette chains at my_tk[k * 128]
    if (MASKS)
        for (int k = 0; k < K; ++k) my_tk[k * 128] = make_float2(in0 ? 1.0f : -1.0f, in1 ? 1.0f : -1.0f);

    // Main chain: T stays the transmittance (frozen once the chain has terminated); thr0 / thr1 say whether it is alive.
    // Object chains (To, Tk) keep composite2_kernel's convention: a finished chain holds -|T|.
    f32x2 T = bc2(1.0f);
    float thr0 = in0 ? THR_ALIVE : THR_DEAD, thr1 = in1 ? THR_ALIVE : THR_DEAD;  // main chain alive <=> thr == 1/255
    f32x2 To = pk2((in0 && MASKS) ? 1.0f : -1.0f, (in1 && MASKS) ? 1.0f : -1.0f);
    f32x2 C0 = bc2(0.0f), C1 = C0, C2 = C0, D = C0, S0 = C0, S1 = C0, S2 = C0;
    uint32_t dk0 = (in0 && MASKS) ? 0u : 0xFFFFFFFFu, dk1 = (in1 && MASKS) ? 0u : 0xFFFFFFFFu;
    uint32_t last0 = 0, last1 = 0;
    bool w_main_done = false, w_done = false;
    const f32x2 ONE = bc2(1.0f), MONE = bc2(-1.0f);

    for (int it = 0;; ++it) {
        const int s = it % COMP_STAGES;
        mbar_wait(reinterpret_cast<uint64_t*>(&sm.full[s]), (uint32_t)((it / COMP_STAGES) & 1));
        const int cnt = *(volatile int*)&sm.cnt[s];
        if (cnt == 0) break;
        if (!w_done) {
            const GeomRec* sr = sm.rec[s];
            bool wm = !w_main_done;
#pragma unroll 1
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int e = c0 + lane;
                uint32_t need = 0;
                if (MASKS && !wm) {
                    float to0, to1;
                    unpk2(To, to0, to1);
                    need = __any_sync(0xffffffffu, to0 > 0.0f || to1 > 0.0f)
                               ? all_k : (__reduce_or_sync(0xffffffffu, ~(dk0 & dk1)) & all_k);
                    if (need == 0) break;
                }
                bool hit = false;
                if (e < cnt) {
                    const float4 A = sr[e].a;
                    const float4 B = sr[e].b;
                    const int eo = MASKS ? (__float_as_int(B.w) & 63) : 0;
                    const bool wanted = wm || (MASKS && eo > 0 && ((need >> ((uint32_t)(eo - 1) & 31u)) & 1u));
                    hit = wanted && !block_culled_fast(A.x, A.y, A.z, A.w, B.x, B.w, bx0, bx1, by0, by1);
                }
                uint32_t mm = __ballot_sync(0xffffffffu, hit);

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
                // the object-only chains of one object hit (composite2_kernel's arithmetic): av = alpha where the half
                // blends at all (A.7's three skips, independent of the main chain), else 0
                auto object_chains = [&](int obj, float av0, float av1) {
                    const f32x2 om = fma2(pk2(av0, av1), MONE, ONE);
                    {   // objects-only render (visible masks, sem-seg)
                        const f32x2 tT = mul2(To, om);
                        float n0, n1, o0, o1;
                        unpk2(tT, n0, n1);
                        unpk2(To, o0, o1);
                        const bool d0 = n0 < 0.0001f, d1 = n1 < 0.0001f;  // also true for a finished (negative) chain
                        const f32x2 am = pk2(d0 ? 0.0f : av0, d1 ? 0.0f : av1);
                        const f32x2 Told = To;
                        To = pk2(d0 ? -fabsf(o0) : n0, d1 ? -fabsf(o1) : n1);
                        const float4 ec = sm_eff[obj - 1];
                        // a finished chain has Told < 0 and am == 0: the product is -0, the sum unchanged
                        S0 = fma2(mul2(bc2(ec.x), am), Told, S0);
                        S1 = fma2(mul2(bc2(ec.y), am), Told, S1);
                        S2 = fma2(mul2(bc2(ec.z), am), Told, S2);
                    }
                    const uint32_t kbit = 1u << ((uint32_t)(obj - 1) & 31u);
                    const float2 tk = my_tk[(obj - 1) * 128];
                    const f32x2 tT = mul2(pk2(tk.x, tk.y), om);
                    float n0, n1;
                    unpk2(tT, n0, n1);
                    const bool d0 = n0 < 0.0001f, d1 = n1 < 0.0001f;
                    if (d0) dk0 |= kbit;
                    if (d1) dk1 |= kbit;
                    my_tk[(obj - 1) * 128] = make_float2(d0 ? -fabsf(tk.x) : n0, d1 ? -fabsf(tk.y) : n1);
                };
#pragma unroll 1
                while (mm) {
                    const GeomRec* r1 = sr + (c0 + __ffs(mm) - 1);
                    mm &= mm - 1;
                    // without a second hit the pair is completed by a record that never blends (opacity 0)
                    const GeomRec* r2 = mm ? sr + (c0 + __ffs(mm) - 1) : &sm.dummy;
                    mm &= mm - 1;
                    const float4 A1 = r1->a, B1 = r1->b, A2 = r2->a, B2 = r2->b;
                    const float4 Cc1 = r1->c, Cc2 = r2->c;  // loaded with the rest: a late load makes ptxas shuffle the alphas
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
                    // ---- main chain (RGB + depth), both hits
                    // T only decreases: if T after both hits is >= 1e-4 at every pixel of the warp, no chain terminates
                    // at either hit (one min, one compare, one vote for the pair).  Otherwise — at most once per pixel —
                    // the chains that terminate are marked dead from the terminating hit on (a chain that terminates
                    // at a hit does not blend it) and the pair's alphas are formed again; that evaluation passes.
                    f32x2 T1, T2;
                    float av10, av11, av20, av21;
                    if (BRANCHY) {
                    av10 = valid_alpha(q10, B1.w, a10, thr0), av11 = valid_alpha(q11, B1.w, a11, thr1);
                    av20 = valid_alpha(q20, B2.w, a20, thr0), av21 = valid_alpha(q21, B2.w, a21, thr1);
                    T1 = mul2(T, fma2(pk2(av10, av11), MONE, ONE));
                    T2 = mul2(T1, fma2(pk2(av20, av21), MONE, ONE));
                    {
                        float t20, t21;
                        unpk2(T2, t20, t21);
                        if (__builtin_expect(__any_sync(0xffffffffu, fminf(t20, t21) < 0.0001f), 0)) {
                            float t10, t11;
                            unpk2(T1, t10, t11);
                            // hit 1: a chain with T1 < 1e-4 terminates there (no blend, dead for hit 2 as well)
                            const bool d10 = t10 < 0.0001f, d11 = t11 < 0.0001f;
                            av10 = d10 ? 0.0f : av10; av20 = d10 ? 0.0f : av20; thr0 = d10 ? THR_DEAD : thr0;
                            av11 = d11 ? 0.0f : av11; av21 = d11 ? 0.0f : av21; thr1 = d11 ? THR_DEAD : thr1;
                            T1 = mul2(T, fma2(pk2(av10, av11), MONE, ONE));
                            // hit 2: chains still alive whose T would drop below 1e-4 terminate there
                            unpk2(mul2(T1, fma2(pk2(av20, av21), MONE, ONE)), t20, t21);
                            const bool d20 = t20 < 0.0001f, d21 = t21 < 0.0001f;
                            av20 = d20 ? 0.0f : av20; thr0 = d20 ? THR_DEAD : thr0;
                            av21 = d21 ? 0.0f : av21; thr1 = d21 ? THR_DEAD : thr1;
                            T2 = mul2(T1, fma2(pk2(av20, av21), MONE, ONE));
                        }
                    }
                    } else {
                        // branch-free: every hit tests its own termination (a dead chain has alpha 0, so tT == T >= 1e-4)
                        float n0, n1, o0, o1;
                        av10 = valid_alpha(q10, B1.w, a10, thr0), av11 = valid_alpha(q11, B1.w, a11, thr1);
                        unpk2(mul2(T, fma2(pk2(av10, av11), MONE, ONE)), n0, n1);
                        unpk2(T, o0, o1);
                        const bool d10 = n0 < 0.0001f, d11 = n1 < 0.0001f;
                        av10 = d10 ? 0.0f : av10; thr0 = d10 ? THR_DEAD : thr0;
                        av11 = d11 ? 0.0f : av11; thr1 = d11 ? THR_DEAD : thr1;
                        T1 = pk2(d10 ? o0 : n0, d11 ? o1 : n1);
                        av20 = valid_alpha(q20, B2.w, a20, thr0), av21 = valid_alpha(q21, B2.w, a21, thr1);
                        unpk2(mul2(T1, fma2(pk2(av20, av21), MONE, ONE)), n0, n1);
                        unpk2(T1, o0, o1);
                        const bool d20 = n0 < 0.0001f, d21 = n1 < 0.0001f;
                        av20 = d20 ? 0.0f : av20; thr0 = d20 ? THR_DEAD : thr0;
                        av21 = d21 ? 0.0f : av21; thr1 = d21 ? THR_DEAD : thr1;
                        T2 = pk2(d20 ? o0 : n0, d21 ? o1 : n1);
                    }
                    accumulate(Cc1, B1.z, pk2(av10, av11), T);
                    accumulate(Cc2, B2.z, pk2(av20, av21), T1);
                    T = T2;
                    if (NCONTRIB) {
                        const uint32_t pos1 = sm.pos[s][r1 - sr], pos2 = r2 != &sm.dummy ? sm.pos[s][r2 - sr] : 0u;
                        if (av10 > 0.0f) last0 = pos1;
                        if (av11 > 0.0f) last1 = pos1;
                        if (av20 > 0.0f) last0 = pos2;
                        if (av21 > 0.0f) last1 = pos2;
                    }
                }
                if (wm && !__any_sync(0xffffffffu, thr0 == THR_ALIVE || thr1 == THR_ALIVE)) {
                    wm = false;
                    if (!MASKS) break;
                }
            }
            if (!w_main_done && !wm) {
                w_main_done = true;
                if (lane == 0) atomicAdd(&sm.warps_main_done, 1);
            }
            float o0, o1;
            unpk2(To, o0, o1);
            const bool pix_done = thr0 != THR_ALIVE && thr1 != THR_ALIVE && (!MASKS || (o0 < 0.0f && o1 < 0.0f && (dk0 & dk1 & all_k) == all_k));
            if (__all_sync(0xffffffffu, pix_done)) {
                w_done = true;
                if (lane == 0) atomicAdd(&sm.warps_done, 1);
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

template <bool MASKS, bool NCONTRIB, bool FAST, int STAGES, int MINB, bool BRANCHY = false>
static int launch_three(const CompArgs& a, dim3 grid, cudaStream_t stream) {
    const int smem = (int)sizeof(CompSmemT<STAGES>) +
                     (MASKS ? (int)(PG_MAX_OBJECTS * sizeof(float4)) + a.num_objects * 256 * (int)sizeof(float) : 0);
    PG_CUDA_CHECK(ensure_dynamic_smem(composite3_kernel<MASKS, NCONTRIB, FAST, STAGES, MINB, BRANCHY>, smem, true));
    composite3_kernel<MASKS, NCONTRIB, FAST, STAGES, MINB, BRANCHY><<<grid, COMP2_THREADS, smem, stream>>>(a);
    return PG_OK;
}

// masks: fused K+3 passes; fast: PG_NUMERICS_FAST; variant (tuning): 31 = 5 CTAs / SM
int launch_composite3(const CompArgs& a, dim3 grid, bool masks, bool fast, int variant, cudaStream_t stream) {
    if (!masks) {
        const bool nc = a.out_n_contrib != nullptr;
        if (fast) return nc ? launch_three<false, true, true, 4, 4>(a, grid, stream) : launch_three<false, false, true, 4, 4>(a, grid, stream);
        return nc ? launch_three<false, true, false, 4, 4>(a, grid, stream) : launch_three<false, false, false, 4, 4>(a, grid, stream);
    }
    if (variant == 31) return fast ? launch_three<true, false, true, 4, 5>(a, grid, stream) : launch_three<true, false, false, 4, 5>(a, grid, stream);
    if (variant == 32) return fast ? launch_three<true, false, true, 4, 4, true>(a, grid, stream) : launch_three<true, false, false, 4, 4, true>(a, grid, stream);
    return fast ? launch_three<true, false, true, 4, 4>(a, grid, stream) : launch_three<true, false, false, 4, 4>(a, grid, stream);
}

}  // namespace pg
