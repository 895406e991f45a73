// composite.cu — per-tile front-to-back alpha compositing: the ABI-side argument set-up and
// launch selection, plus composite2_kernel, the round-1 kernel (one CTA per 16x16 tile, warp-specialised producer /
// consumers, two pixels per lane in packed FP32, exact arithmetic only).  The default kernel is composite3_kernel
// (composite3.cu); composite2_kernel stays for two jobs:
//   * the counting runs (settings.debug bit 1 -> pg_read_stats): its STATS instantiation counts pair evaluations,
//     exps, blends and warp-level hits of the same walk in exact arithmetic;
//   * PG_COMP_VARIANT=20: a second implementation of the same arithmetic for A/B measurements and cross-checks.
//
// Walk of a tile's sorted list, per batch of up to COMP_BATCH entries (shared with composite3, composite_common.cuh):
//   1. ids are prefetched by TMA bulk copies (1 KB chunks, double-buffered); entries the binning stage proved
//      invisible for the whole tile (PG_CULL_FLAG) and, once every main chain has terminated, environment
//      entries are dropped by a ballot compaction — they are never fetched;
//   2. each surviving id's 48-byte record {xy, conic, opacity, depth, cut|object, rgb} is GATHERED into a
//      shared-memory ring by 16-byte cp.async whose completion arrives on the stage's `full` mbarrier
//      (4-stage ring: later batches land while the current one is composited);
//   3. every consumer warp tests the batch against its pixel block LANE-PARALLEL (lane l tests entry 32c+l:
//      exact minimum of the conic's quadratic form over the block vs. the entry's alpha<1/255 cut) and
//      ballots the hits;
//   4. the warp walks only its hits, in list order, with broadcast 16-byte shared loads, then releases the
//      stage on its `empty` mbarrier.  No CTA-wide barrier in the steady state.
// Culling is conservative (a margin covers float rounding), so every (pixel, Gaussian) pair that
// the reference would blend is blended with the reference's operation order: results are unchanged.
//
// MASKS == false : the reference's single pass -> color, depth (+ final_T, n_contrib).
// MASKS == true  : the reference's K+3 passes in one walk -> RGB + depth, the objects-only flat-colour
//             render (visible masks, sem-seg) and one transmittance chain per object (silhouettes);
//             alpha is evaluated once per (pixel, Gaussian).
//
// exp is a 12-op FMA polynomial so that results are bit-reproducible on the CPU oracle.
#include "composite_common.cuh"

namespace pg {

// Dynamic shared memory after CompSmem (MASKS only): eff[PG_MAX_OBJECTS] float4 (colour the
// rasterizer produces for object k's flat SH), then Tk[K][256] — the standalone transmittance of
// object k at each of the tile's 256 pixels (silhouette chains), slot = warp * 32 + lane.
// COMP_STAGES: depth of the record ring (how far fast warps may run ahead of the slowest one);
// ILP: hits evaluated together (2: geometry + exp of two hits interleave, blending stays in list order);
// MINB: CTAs per SM the register allocation is bounded for.


// =================================================================================================
// 2 pixels per thread: packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2 — PTX fma/mul/add.rn.f32x2).
//
// A 1-pixel-per-lane kernel is issue-bound (83 % issue-active, FMA pipe 41 %, ALU pipe 39 %, profiles/r1ab_*): every
// instruction of its hit loop serves 32 pixels.  Here a warp owns an 8x8 pixel block, a lane owns the two
// vertically adjacent pixels (x, y0) and (x, y0 + 1), and the per-pixel arithmetic of both runs in ONE packed
// instruction per operation: each half is an individually rounded IEEE binary32 op in the same order as the
// scalar kernel (the per-Gaussian operands ride along as broadcast scalars, SASS `R.F32`), so results stay
// bit-identical to the oracle.  Record loads, hit-list scanning, address arithmetic and the lane-parallel
// cull (4 consumer warps per tile instead of 8) are shared by the two pixels.
//
// Blending is branch-free: a half that does not blend (outside [cut, 0], alpha < 1/255, chain finished)
// gets alpha = 0, which makes every update an exact no-op (1 - 0 = 1, T * 1 = T, C + 0 * T = C); a finished
// chain keeps -|T| as in the scalar kernel.  The silhouette chains keep the same sign convention in
// shared memory (float2 per lane and object).
//
// ptxas contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 (observed with 12.9, unlike the scalar
// .rn forms), so no packed add here consumes a packed mul: 1 - alpha is fma(alpha, -1, 1), the polynomial
// and the accumulations are fma by definition, negations ride on scalar operands.
// =================================================================================================

// Dynamic shared memory after CompSmem (MASKS only): eff[PG_MAX_OBJECTS] float4, then Tk2[K][128] float2 —
// the standalone transmittance of object k at the two pixels of each lane (slot = warp * 32 + lane).
template <bool MASKS, bool STATS, int COMP_STAGES, int ILP, int MINB>
__global__ void __launch_bounds__(COMP2_THREADS, MINB) composite2_kernel(const CompArgs a) {
    using CompSmem = CompSmemT<COMP_STAGES, !MASKS>;
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
    float2* my_tk = sm_tk + (warp * 32 + lane);  // object k's chains at my_tk[k * 128]
    if (MASKS)
        for (int k = 0; k < K; ++k) my_tk[k * 128] = make_float2(in0 ? 1.0f : -1.0f, in1 ? 1.0f : -1.0f);

    // finished chains keep -|T| (see the scalar kernel)
    f32x2 T = pk2(in0 ? 1.0f : -1.0f, in1 ? 1.0f : -1.0f);
    f32x2 To = pk2((in0 && MASKS) ? 1.0f : -1.0f, (in1 && MASKS) ? 1.0f : -1.0f);
    f32x2 C0 = bc2(0.0f), C1 = C0, C2 = C0, D = C0, S0 = C0, S1 = C0, S2 = C0;
    uint32_t dk0 = (in0 && MASKS) ? 0u : 0xFFFFFFFFu, dk1 = (in1 && MASKS) ? 0u : 0xFFFFFFFFu;
    uint32_t last0 = 0, last1 = 0;
    uint32_t n_eval = 0, n_exp = 0, n_blend = 0, n_slots = 0;
    uint32_t h_env = 0, h_obj_main = 0, h_obj_after = 0, c_after = 0;  // STATS, warp-uniform
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
                if (STATS && !wm) ++c_after;
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
                    hit = wanted && !block_culled(A.x, A.y, A.z, A.w, B.x, B.w, bx0, bx1, by0, by1);
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
                // alpha of both halves; v = the half blends at all (A.7's three skips)
                auto alpha_of = [&](f32x2 pw, const float4& B, float& a0, float& a1, bool& v0, bool& v1) {
                    float e0, e1, p0, p1;
                    exp2_exact_nz(pw, e0, e1);
                    unpk2(pw, p0, p1);
                    a0 = fminf(0.99f, mul(B.y, e0));
                    a1 = fminf(0.99f, mul(B.y, e1));
                    v0 = blends(p0, B.w, a0);
                    v1 = blends(p1, B.w, a1);
                    if (STATS) {
                        float t0, t1, o0, o1;
                        unpk2(T, t0, t1);
                        unpk2(To, o0, o1);
                        const int obj = MASKS ? (__float_as_int(B.w) & 63) : 0;
                        const bool l0 = t0 > 0.0f || (MASKS && obj > 0 && (o0 > 0.0f || !((dk0 >> (obj - 1)) & 1u)));
                        const bool l1 = t1 > 0.0f || (MASKS && obj > 0 && (o1 > 0.0f || !((dk1 >> (obj - 1)) & 1u)));
                        n_slots += 2;
                        n_eval += (l0 ? 1 : 0) + (l1 ? 1 : 0);
                        n_exp += ((l0 && !(p0 > 0.0f) && !(p0 < B.w)) ? 1 : 0) + ((l1 && !(p1 > 0.0f) && !(p1 < B.w)) ? 1 : 0);
                        n_blend += ((l0 && v0) ? 1 : 0) + ((l1 && v1) ? 1 : 0);
                    }
                };
                // one transmittance chain, both halves: returns the chain's alpha (0 where it does not blend)
                auto chain = [&](f32x2& Tc, f32x2 om, float av0, float av1, float& am0, float& am1) {
                    const f32x2 tT = mul2(Tc, om);
                    float n0, n1, o0, o1;
                    unpk2(tT, n0, n1);
                    unpk2(Tc, o0, o1);
                    const bool d0 = n0 < 0.0001f, d1 = n1 < 0.0001f;
                    am0 = d0 ? 0.0f : av0;
                    am1 = d1 ? 0.0f : av1;
                    Tc = pk2(d0 ? -fabsf(o0) : n0, d1 ? -fabsf(o1) : n1);
                };
                auto blend = [&](const GeomRec* r, const float4& B, float a0, float a1, bool v0, bool v1) {
                    const float av0 = v0 ? a0 : 0.0f, av1 = v1 ? a1 : 0.0f;
                    const f32x2 om = fma2(pk2(av0, av1), MONE, ONE);  // 1 - alpha
                    {
                        const f32x2 Told = T;
                        float am0, am1;
                        chain(T, om, av0, av1, am0, am1);
                        const f32x2 am = pk2(am0, am1);
                        const float4 Cc = r->c;
                        C0 = fma2(mul2(bc2(Cc.x), am), Told, C0);
                        C1 = fma2(mul2(bc2(Cc.y), am), Told, C1);
                        C2 = fma2(mul2(bc2(Cc.z), am), Told, C2);
                        D = fma2(mul2(bc2(B.z), am), Told, D);
                        if (!MASKS) {
                            const uint32_t pos = sm.pos[s][r - sr];
                            if (am0 > 0.0f) last0 = pos;
                            if (am1 > 0.0f) last1 = pos;
                        }
                    }
                    const int obj = MASKS ? (__float_as_int(B.w) & 63) : 0;  // warp-uniform
                    if (MASKS && obj > 0) {
                        {
                            const f32x2 Told = To;
                            float am0, am1;
                            chain(To, om, av0, av1, am0, am1);
                            const f32x2 am = pk2(am0, am1);
                            const float4 ec = sm_eff[obj - 1];
                            S0 = fma2(mul2(bc2(ec.x), am), Told, S0);
                            S1 = fma2(mul2(bc2(ec.y), am), Told, S1);
                            S2 = fma2(mul2(bc2(ec.z), am), Told, S2);
                        }
                        const uint32_t kbit = 1u << ((uint32_t)(obj - 1) & 31u);
                        const float2 tk = my_tk[(obj - 1) * 128];
                        const f32x2 tT = mul2(pk2(tk.x, tk.y), om);
                        float n0, n1;
                        unpk2(tT, n0, n1);
                        const bool d0 = n0 < 0.0001f, d1 = n1 < 0.0001f;  // also true for a finished (negative) chain
                        if (d0) dk0 |= kbit;
                        if (d1) dk1 |= kbit;
                        my_tk[(obj - 1) * 128] = make_float2(d0 ? -fabsf(tk.x) : n0, d1 ? -fabsf(tk.y) : n1);
                    }
                };
#pragma unroll 1
                while (mm) {
                    const GeomRec* r1 = sr + (c0 + __ffs(mm) - 1);
                    mm &= mm - 1;
                    if (ILP == 2) {
                        const bool two = mm != 0;  // warp-uniform
                        const GeomRec* r2 = two ? sr + (c0 + __ffs(mm) - 1) : r1;
                        mm &= mm - 1;
                        const float4 A1 = r1->a, B1 = r1->b, A2 = r2->a, B2 = r2->b;
                        const f32x2 p1 = power_of(A1, B1), p2 = power_of(A2, B2);
                        float a10, a11, a20, a21;
                        bool v10, v11, v20, v21;
                        alpha_of(p1, B1, a10, a11, v10, v11);
                        alpha_of(p2, B2, a20, a21, v20, v21);
                        // without a second hit r2 == r1 blends with alpha 0: an exact no-op, cheaper than a branch
                        // whose two sides keep the accumulators in different registers
                        blend(r1, B1, a10, a11, v10, v11);
                        blend(r2, B2, a20, a21, v20 && two, v21 && two);
                    } else {
                        const float4 A1 = r1->a, B1 = r1->b;
                        if (STATS) {
                            if ((__float_as_int(B1.w) & 63) == 0) ++h_env;
                            else if (wm) ++h_obj_main;
                            else ++h_obj_after;
                        }
                        const f32x2 p1 = power_of(A1, B1);
                        float a10, a11;
                        bool v10, v11;
                        alpha_of(p1, B1, a10, a11, v10, v11);
                        blend(r1, B1, a10, a11, v10, v11);
                    }
                }
                if (wm) {
                    float t0, t1;
                    unpk2(T, t0, t1);
                    if (__all_sync(0xffffffffu, t0 < 0.0f && t1 < 0.0f)) {
                        wm = false;
                        if (!MASKS) break;
                    }
                }
            }
            if (!w_main_done && !wm) {
                w_main_done = true;
                if (lane == 0) atomicAdd(&sm.warps_main_done, 1);
            }
            float t0, t1, o0, o1;
            unpk2(T, t0, t1);
            unpk2(To, o0, o1);
            const bool pix_done = t0 < 0.0f && t1 < 0.0f &&
                                  (!MASKS || (o0 < 0.0f && o1 < 0.0f && (dk0 & dk1 & all_k) == all_k));
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
        const float Tf = fabsf(Th[h]), Tof = fabsf(Toh[h]);
        const size_t pix = (size_t)(h == 0 ? py0 : py1) * a.W + px;
        a.out_color[pix] = fma(Tf, bg0, Ch[0][h]);
        a.out_color[HW + pix] = fma(Tf, bg1, Ch[1][h]);
        a.out_color[2 * HW + pix] = fma(Tf, bg2, Ch[2][h]);
        a.out_depth[pix] = Dh[h];
        if (a.out_final_T) a.out_final_T[pix] = Tf;
        if (!MASKS && a.out_n_contrib) a.out_n_contrib[pix] = h == 0 ? last0 : last1;
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
    if (STATS) {
        flush_stats(a.stats, n_eval, n_exp, n_blend, n_slots);
        if (lane == 0) {
            atomicAdd(&a.stats[4], (unsigned long long)h_env);
            atomicAdd(&a.stats[5], (unsigned long long)h_obj_main);
            atomicAdd(&a.stats[6], (unsigned long long)h_obj_after);
            atomicAdd(&a.stats[7], (unsigned long long)c_after);
        }
    }
}

// PG_COMP_SMEM_PAD (tuning only): extra dynamic shared memory per compositing CTA, i.e. a cap on the CTAs
// per SM, which leaves registers / shared memory for another frame's binning kernels to co-run.
static int comp_smem_pad() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PG_COMP_SMEM_PAD");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <bool MASKS, bool STATS, int STAGES, int ILP, int MINB>
static int launch_two(const CompArgs& a, dim3 grid, cudaStream_t stream) {
    const int smem = (int)sizeof(CompSmemT<STAGES, !MASKS>) + comp_smem_pad() +
                     (MASKS ? (int)(PG_MAX_OBJECTS * sizeof(float4)) + a.num_objects * 256 * (int)sizeof(float) : 0);
    // residency is bounded by shared memory (~39 KB per CTA): ask for the largest carve-out
    PG_CUDA_CHECK(ensure_dynamic_smem(composite2_kernel<MASKS, STATS, STAGES, ILP, MINB>, smem, true));
    composite2_kernel<MASKS, STATS, STAGES, ILP, MINB><<<grid, COMP2_THREADS, smem, stream>>>(a);
    return PG_OK;
}

// PG_COMP_VARIANT (tuning only, read once): 30 (default) composite3_kernel; 20 the round-1 composite2_kernel
// (1.05 ms on the C2 workload where composite3 takes 1.08 exact / 0.95 fast; profiles/README.md).
static int comp_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PG_COMP_VARIANT");
        v = e ? atoi(e) : 30;
    }
    return v;
}

int launch_composite3(const CompArgs& a, dim3 grid, bool masks, bool fast, int variant, cudaStream_t stream);

int launch_composite(const CompArgs& a, int gy, bool masks, cudaStream_t stream) {
    dim3 grid(a.gx * gy);
    const bool st = a.stats != nullptr;
    const int var = comp_variant();
    int rc;
    if (var >= 30 && !st) {
        // default: composite3_kernel (composite3.cu).  Statistics runs (debug bit 1) use composite2_kernel below,
        // whose exact arithmetic walks the same hits.
        rc = launch_composite3(a, grid, masks, a.fast != 0, var, stream);
    } else if (!masks) {
        rc = st ? launch_two<false, true, 4, 1, 4>(a, grid, stream) : launch_two<false, false, 4, 2, 4>(a, grid, stream);
    } else {
        rc = st ? launch_two<true, true, 4, 1, 4>(a, grid, stream) : launch_two<true, false, 4, 2, 4>(a, grid, stream);
    }
    if (rc) return rc;
    PG_CUDA_CHECK(cudaGetLastError());
    count_launch(1);
    return PG_OK;
}

// Fills CompArgs from the ABI structs and launches the right kernel.
int launch_composite_from_abi(const uint2* ranges, const uint32_t* tile_order, const uint32_t* point_list, const GeomRec* recs, int W,
                              int H, const float* bg, const pg_raster_outputs* ro, const pg_frame_outputs* fo,
                              const pg_object_table* objs, uint32_t n_env, const uint32_t* tile_obj_count,
                              unsigned long long* stats, bool fast, cudaStream_t stream) {
    CompArgs a;
    memset(&a, 0, sizeof(a));
    a.stats = stats;
    a.fast = fast ? 1 : 0;
    a.ranges = ranges; a.tile_order = tile_order; a.point_list = point_list; a.recs = recs;
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
    if (a.num_objects == 0 && !(fo->seg_color || fo->sem_seg || fo->visible || fo->silhouette))
        return launch_composite(a, gy, false, stream);
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
