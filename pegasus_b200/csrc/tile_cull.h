// tile_cull.h — which tiles of one tile ROW can a Gaussian contribute to?  (host + device)
//
// A Gaussian with pixel centre (gx, gy) and conic (qa, qb, qc) contributes to pixel (px, py) only if
//     power = -0.5 (qa u^2 + qc v^2) - qb u v >= cut,   u = px - gx, v = py - gy          (A.7)
// because below its per-Gaussian `cut` alpha < 1/255 is certain (preprocess.cu).  With
// Q(u,v) = 0.5 qa u^2 + qb u v + 0.5 qc v^2 and thr = -cut this is the ellipse Q <= thr.
// For the pixel rows [y0, y1] of a tile row, v ranges over [vl, vh] and the set of u with
//     g(u) = min_{v in [vl,vh]} Q(u, v) <= thr
// is an interval (projection of a convex set).  g(u) = Q(u, clamp(-qb u / qc, vl, vh)), hence
//     {g <= thr} = Jl  U  Jh  U  I0
//     Jv = {u : Q(u, v) <= thr}                       chord of the ellipse on the line v (v = vl, vh)
//     I0 = {|u| <= u_max} ^ {u : -qb u / qc in [vl, vh]},  u_max = sqrt(2 thr qc / det)
// (dropping the side conditions of Jl, Jh keeps the set exact because g <= Q(., v) everywhere).
// Every quantity is evaluated so that rounding can only ENLARGE the interval: thr is inflated by
// 1e-4 relative + 2e-3 absolute (the same slack block_culled uses), det is replaced by a lower bound
// that covers its cancellation error, and the pixel interval is widened by 0.02 px + 1e-5 relative.
// All operations are IEEE (division, sqrt, fma), so the host build used by the CPU tests computes
// exactly what the kernel computes.  Indefinite / NaN conics are never culled.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD static inline
#endif

namespace pg {

struct CullGauss {      // per-Gaussian constants of the row test
    float gx, gy, qa, qb, qc;
    float thr2qa;       // 2 thr qa                 (disc of a chord = thr2qa - det_lo v^2)
    float det_lo;       // lower bound of qa qc - qb^2
    float inv_qa;       // 1 / qa
    float umax;         // upper bound of sqrt(2 thr qc / det)
    float k;            // -qc / qb  (u of the unconstrained minimiser line at height v is k v); 0 if qb == 0
    int ok;             // 0: never cull (indefinite conic, NaN, ill-conditioned)
};

PG_HD float cg_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
PG_HD float cg_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
PG_HD float cg_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
PG_HD float cg_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
PG_HD float cg_sqrt(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

PG_HD CullGauss cull_setup(float gx, float gy, float qa, float qb, float qc, float cut) {
    CullGauss c;
    c.gx = gx; c.gy = gy; c.qa = qa; c.qb = qb; c.qc = qc;
    const float thr = cg_fma(-cut, 1.0001f, 2e-3f);             // inflated -cut (> 0)
    const float qq = cg_mul(qa, qc);
    const float det = cg_fma(qa, qc, -cg_mul(qb, qb));          // error <= ulp(qb^2) + ulp(qa qc)
    c.det_lo = cg_fma(qq, -4e-7f, det);
    c.ok = (qa > 0.0f) && (qc > 0.0f) && (c.det_lo > 0.0f) && (thr > 0.0f) && (thr < 1e4f);
    c.thr2qa = cg_mul(cg_mul(2.0f, thr), qa);
    c.inv_qa = cg_div(1.0f, qa);
    // u_max^2 = 2 thr qc / det; 1.00001 covers the roundings of the three operations
    c.umax = cg_mul(cg_sqrt(cg_div(cg_mul(cg_mul(2.0f, thr), qc), c.det_lo)), 1.00001f);
    c.k = (qb != 0.0f) ? cg_div(-qc, qb) : 0.0f;
    if (!(c.umax == c.umax) || !(c.k == c.k)) c.ok = 0;
    return c;
}

// Tile columns [*ta, *tb) of tile row `ty` (clipped to the Gaussian's rectangle [rx0, rx1)) that can
// contain a contributing pixel.  Returns the count tb - ta (0: nothing in this row).
PG_HD int cull_row_run(const CullGauss& c, int ty, int rx0, int rx1, int W, int H, int* ta, int* tb) {
    if (!c.ok) { *ta = rx0; *tb = rx1; return rx1 - rx0; }
    const int y0 = ty * 16;
    const int y1 = (y0 + 15 < H - 1) ? y0 + 15 : H - 1;
    const float vl = cg_add((float)y0, -c.gy), vh = cg_add((float)y1, -c.gy);
    float lo = 3.0e38f, hi = -3.0e38f;
    // chords on the two boundary lines
    {
        const float disc = cg_fma(-c.det_lo, cg_mul(vl, vl), c.thr2qa);
        if (disc >= 0.0f) {
            const float s = cg_sqrt(disc), m = -cg_mul(c.qb, vl);
            lo = fminf(lo, cg_mul(cg_add(m, -s), c.inv_qa));
            hi = fmaxf(hi, cg_mul(cg_add(m, s), c.inv_qa));
        }
    }
    {
        const float disc = cg_fma(-c.det_lo, cg_mul(vh, vh), c.thr2qa);
        if (disc >= 0.0f) {
            const float s = cg_sqrt(disc), m = -cg_mul(c.qb, vh);
            lo = fminf(lo, cg_mul(cg_add(m, -s), c.inv_qa));
            hi = fmaxf(hi, cg_mul(cg_add(m, s), c.inv_qa));
        }
    }
    // interior piece: the unconstrained minimiser v*(u) = -qb u / qc lies inside [vl, vh]
    {
        float a, b;
        if (c.qb != 0.0f) {
            const float s1 = cg_mul(c.k, vl), s2 = cg_mul(c.k, vh);
            a = fminf(s1, s2); b = fmaxf(s1, s2);
            // widen by the rounding of k and the products
            const float w = cg_fma(fmaxf(fabsf(a), fabsf(b)), 4e-7f, 1e-6f);
            a = cg_add(a, -w); b = cg_add(b, w);
        } else {
            if (vl <= 0.0f && vh >= 0.0f) { a = -3.0e38f; b = 3.0e38f; }
            else { a = 1.0f; b = -1.0f; }
        }
        a = fmaxf(a, -c.umax); b = fminf(b, c.umax);
        if (a <= b) { lo = fminf(lo, a); hi = fmaxf(hi, b); }
    }
    if (!(lo <= hi)) { *ta = rx0; *tb = rx0; return 0; }
    // pixel columns that may contribute, widened; then tile columns
    const float wl = cg_fma(fabsf(lo), 1e-5f, 0.02f), wh = cg_fma(fabsf(hi), 1e-5f, 0.02f);
    const float fl = cg_add(cg_add(c.gx, lo), -wl), fh = cg_add(cg_add(c.gx, hi), wh);
    // clamp in float before the int conversion (values can be huge)
    const float flc = fminf(fmaxf(fl, -1.0f), (float)W), fhc = fminf(fmaxf(fh, -1.0f), (float)W);
    int pa = (int)ceilf(flc), pb = (int)floorf(fhc);
    if (pa < 0) pa = 0;
    if (pb > W - 1) pb = W - 1;
    if (pa > pb) { *ta = rx0; *tb = rx0; return 0; }
    int a = pa >> 4, b = (pb >> 4) + 1;
    if (a < rx0) a = rx0;
    if (b > rx1) b = rx1;
    if (a >= b) { *ta = rx0; *tb = rx0; return 0; }
    *ta = a; *tb = b;
    return b - a;
}

}  // namespace pg
