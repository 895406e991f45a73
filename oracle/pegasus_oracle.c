/*
 * pegasus_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (pegasus_b200/, diff_gaussian_rasterization/)
 * never imports, links or executes anything under oracle/.
 *
 * PARITY UNPINNED.  This file restates, in plain C, the forward pass of the rasterizer the
 * reference binds as `diff_gaussian_rasterization`
 *   (call sites: /root/reference/submodules/gaussian-splatting-pegasus/gaussian_renderer/__init__.py:14,38-53,87-95)
 * whose source — meyerls/depth-diff-gaussian-rasterization @ 0062df9757ed57330b49b7ad181d97f5c34d547e —
 * is an un-vendored submodule that is ABSENT from /root/reference (see /root/reference/.SUBMODULES.json).
 * The reference ships no golden vectors, known-answer tests or fixtures for this path, so the
 * rasterizer arithmetic below follows the published algorithm of graphdeco-inria/diff-gaussian-rasterization
 * (forward path) plus the depth accumulation of the "depth" fork, as written down in SURVEY.md Appendix A.
 * What IS pinned against the reference's in-tree Python (tests/golden/, tools/make_golden.py):
 *   - SH basis / constants      : GSP/utils/sh_utils.py:26-43,57-112 (eval_sh)
 *   - quaternion -> rotation    : GSP/utils/general_utils.py:78-99 (build_rotation)
 *   - covariance R S S^T R^T    : GSP/utils/general_utils.py:101-110, src/gs/gaussian_model.py:38-42
 *   - view / projection matrices: GSP/utils/graphics_utils.py:38-71, GSP/scene/cameras.py:54-57
 * Cross-checks that stand in for the missing rasterizer vectors (they narrow "unpinned", they do not lift
 * it): tests/test_oracle_analytic.py (an independent float64 numpy restatement of the published equations:
 * radii, tile rectangles and pair counts exactly, images within float32 round-off; a closed-form
 * single-Gaussian known answer) and baseline/upstream_style.cu (a second GPU implementation in the
 * upstream's expression form, tests/test_gpu_baseline.py).
 *
 * NUMERICAL SPEC.  Every floating-point operation below is an individually rounded IEEE-754
 * binary32 operation in the stated order; fused multiply-adds appear only where FMA() is written.
 * The order is the one nvcc (-fmad=true, the upstream default) contracts the upstream expressions
 * into (a*x + b*y + c*z + d  ->  t=b*y; t=fma(a,x,t); t=fma(c,z,t); t=d+t), verified on nvcc 12.9.
 * The CUDA kernels in pegasus_b200/csrc use the same sequence with explicit __f*_rn intrinsics, so
 * the integer outputs (radii, tile rectangles, 64-bit keys, sorted order, ranges) are bit-exact
 * CPU <-> GPU, and so are the images because exp() is a software polynomial (pg_expf below)
 * rather than the GPU's MUFU approximation.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC  (oracle/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FMA(a, b, c) fmaf((a), (b), (c))
#define BLOCK_X 16
#define BLOCK_Y 16

static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * exp(x) for x <= 0 (0 below -80): Cody-Waite reduction + degree-6 polynomial, <= 1.3 ulp.
 * Stands in for the upstream kernel's expf; built from IEEE ops only so
 * that CPU and GPU agree bit for bit.
 * ------------------------------------------------------------------------------------------ */
float orc_expf(float x) {
    const float L2E = 1.44269502162933349609375f;
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23, bits 0x4B400000 */
    const float LN2_HI = 0.693145751953125f;
    const float LN2_LO = 1.428606765330187045e-06f;
    if (x < -80.0f) return 0.0f;
    float z = FMA(x, L2E, MAGIC);
    float n = z - MAGIC;
    float r = FMA(n, -LN2_HI, x);
    r = FMA(n, -LN2_LO, r);
    float p = 0x1.6b5016p-10f;
    p = FMA(p, r, 0x1.126caep-7f);
    p = FMA(p, r, 0x1.55578ep-5f);
    p = FMA(p, r, 0x1.55540cp-3f);
    p = FMA(p, r, 0x1.fffffcp-2f);
    p = FMA(p, r, 1.0f);
    p = FMA(p, r, 1.0f);
    int32_t ni = (int32_t)(as_uint(z) - 0x4B400000u);
    return as_float((uint32_t)((int32_t)as_uint(p) + ni * (1 << 23)));
}

/* SH constants: GSP/utils/sh_utils.py:26-43 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* p_view / p_hom rows: t = m[4+r]*y; t = fma(m[r],x,t); t = fma(m[8+r],z,t); t = m[12+r] + t
 * (matrix flat index [4*col + row]; GSP/scene/cameras.py:54-56 stores the transpose). */
static inline float xform_row(const float* m, int r, float x, float y, float z) {
    float t = m[4 + r] * y;
    t = FMA(m[r], x, t);
    t = FMA(m[8 + r], z, t);
    return m[12 + r] + t;
}

/* Appendix A.3: cov3D from scale and (already unit) quaternion (r,x,y,z);
 * rotation convention GSP/utils/general_utils.py:78-99. */
void orc_cov3d(const float* scale, float mod, const float* q, float* cov6) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3];
    float t;
    t = z * z; t = FMA(y, y, t); R[0][0] = 1.0f - (t + t);
    t = FMA(x, y, -(r * z));     R[0][1] = t + t;
    t = r * y; t = FMA(x, z, t); R[0][2] = t + t;
    t = r * z; t = FMA(x, y, t); R[1][0] = t + t;
    t = z * z; t = FMA(x, x, t); R[1][1] = 1.0f - (t + t);
    t = FMA(y, z, -(r * x));     R[1][2] = t + t;
    t = FMA(x, z, -(r * y));     R[2][0] = t + t;
    t = r * x; t = FMA(y, z, t); R[2][1] = t + t;
    t = y * y; t = FMA(x, x, t); R[2][2] = 1.0f - (t + t);
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    float M[3][3]; /* M[i][k] = s_k * R[i][k] */
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) M[i][k] = s[k] * R[i][k];
    /* Sigma[i][j] = M[i][0]*M[j][0] + M[i][1]*M[j][1] + M[i][2]*M[j][2] */
    int idx = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            float a = M[i][1] * M[j][1];
            a = FMA(M[i][0], M[j][0], a);
            a = FMA(M[i][2], M[j][2], a);
            cov6[idx++] = a;
        }
}

/* Appendix A.4: EWA 2D covariance (a, b, c) with the 0.3 dilation added. */
static void cov2d(const float* pv, float fx, float fy, float tanx, float tany, const float* c3,
                  const float* V, float* out3) {
    float tz = pv[2];
    float limx = 1.3f * tanx, limy = 1.3f * tany;
    float txtz = pv[0] / tz, tytz = pv[1] / tz;
    float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    float J00 = fx / tz, J11 = fy / tz;
    float tz2 = tz * tz;
    float J02 = -(fx * tx) / tz2;
    float J12 = -(fy * ty) / tz2;
    /* T0[k] = fma(V[4k+2], J02, V[4k]*J00); T1[k] = fma(V[4k+2], J12, V[4k+1]*J11) */
    float T0[3], T1[3];
    for (int k = 0; k < 3; ++k) {
        T0[k] = FMA(V[4 * k + 2], J02, V[4 * k] * J00);
        T1[k] = FMA(V[4 * k + 2], J12, V[4 * k + 1] * J11);
    }
    float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    /* A_i[k] = T_i[0]*Vrk[0][k] + T_i[1]*Vrk[1][k] + T_i[2]*Vrk[2][k] */
    float A0[3], A1[3];
    for (int k = 0; k < 3; ++k) {
        float a = T0[1] * Vrk[1][k];
        a = FMA(T0[0], Vrk[0][k], a);
        A0[k] = FMA(T0[2], Vrk[2][k], a);
        float b = T1[1] * Vrk[1][k];
        b = FMA(T1[0], Vrk[0][k], b);
        A1[k] = FMA(T1[2], Vrk[2][k], b);
    }
    float c00 = A0[1] * T0[1]; c00 = FMA(A0[0], T0[0], c00); c00 = FMA(A0[2], T0[2], c00);
    float c01 = A1[1] * T0[1]; c01 = FMA(A1[0], T0[0], c01); c01 = FMA(A1[2], T0[2], c01);
    float c11 = A1[1] * T1[1]; c11 = FMA(A1[0], T1[0], c11); c11 = FMA(A1[2], T1[2], c11);
    out3[0] = c00 + 0.3f;
    out3[1] = c01;
    out3[2] = c11 + 0.3f;
}

/* Appendix A.5 colour: SH basis exactly GSP/utils/sh_utils.py:74-100, +0.5, clamp >= 0. */
static void sh_to_rgb(int deg, const float* p, const float* campos, const float* sh /*[16][3]*/,
                      float* rgb) {
    float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
    float l2 = dy * dy; l2 = FMA(dx, dx, l2); l2 = FMA(dz, dz, l2);
    float len = sqrtf(l2);
    float x = dx / len, y = dy / len, z = dz / len;
    float b[16];
    int n = 1;
    b[0] = SH_C0;
    if (deg > 0) {
        b[1] = -(SH_C1 * y); b[2] = SH_C1 * z; b[3] = -(SH_C1 * x);
        n = 4;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = SH_C2[0] * xy;
            b[5] = SH_C2[1] * yz;
            b[6] = SH_C2[2] * ((2.0f * zz - xx) - yy);
            b[7] = SH_C2[3] * xz;
            b[8] = SH_C2[4] * (xx - yy);
            n = 9;
            if (deg > 2) {
                b[9] = (SH_C3[0] * y) * (3.0f * xx - yy);
                b[10] = (SH_C3[1] * xy) * z;
                b[11] = (SH_C3[2] * y) * ((4.0f * zz - xx) - yy);
                b[12] = (SH_C3[3] * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
                b[13] = (SH_C3[4] * x) * ((4.0f * zz - xx) - yy);
                b[14] = (SH_C3[5] * z) * (xx - yy);
                b[15] = (SH_C3[6] * x) * (xx - 3.0f * yy);
                n = 16;
            }
        }
    }
    for (int c = 0; c < 3; ++c) {
        float acc = b[0] * sh[c];
        for (int k = 1; k < n; ++k) acc = FMA(b[k], sh[3 * k + c], acc);
        acc = acc + 0.5f;
        rgb[c] = acc < 0.0f ? 0.0f : acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * Appendix A.1-A.5 preprocess.  All outputs are caller-allocated, zero-initialised here.
 *   radii[P] i32, xy[P][2], depth[P], cov3d[P][6], conic_opacity[P][4], rgb[P][3],
 *   tiles_touched[P] u32, rect[P][4] i32 = (min.x, min.y, max.x, max.y)
 * shs: (P,16,3) or NULL with colors_precomp (P,3); scales+rots or cov3d_precomp (P,6).
 * ------------------------------------------------------------------------------------------ */
void orc_preprocess(int P, int deg, const float* means, const float* scales, float scale_mod,
                    const float* rots, const float* opac, const float* shs,
                    const float* cov3d_precomp, const float* colors_precomp, const float* V,
                    const float* M, const float* campos, int W, int H, float tanx, float tany,
                    int32_t* radii, float* xy, float* depth, float* cov3d, float* conic_opacity,
                    float* rgb, uint32_t* tiles_touched, int32_t* rect) {
    const float fx = (float)W / (2.0f * tanx);
    const float fy = (float)H / (2.0f * tany);
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        xy[2 * i] = xy[2 * i + 1] = 0.0f;
        depth[i] = 0.0f;
        for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = 0.0f;
        for (int k = 0; k < 4; ++k) conic_opacity[4 * i + k] = 0.0f;
        for (int k = 0; k < 3; ++k) rgb[3 * i + k] = 0.0f;
        for (int k = 0; k < 4; ++k) rect[4 * i + k] = 0;
        const float* p = means + 3 * i;
        float pv[3];
        pv[0] = xform_row(V, 0, p[0], p[1], p[2]);
        pv[1] = xform_row(V, 1, p[0], p[1], p[2]);
        pv[2] = xform_row(V, 2, p[0], p[1], p[2]);
        if (pv[2] <= 0.2f) continue; /* A.1 */
        float hx = xform_row(M, 0, p[0], p[1], p[2]);
        float hy = xform_row(M, 1, p[0], p[1], p[2]);
        float hw = xform_row(M, 3, p[0], p[1], p[2]);
        float pw = 1.0f / (hw + 0.0000001f);
        float ppx = hx * pw, ppy = hy * pw;
        float c3[6];
        if (cov3d_precomp) {
            for (int k = 0; k < 6; ++k) c3[k] = cov3d_precomp[6 * i + k];
        } else {
            orc_cov3d(scales + 3 * i, scale_mod, rots + 4 * i, c3);
        }
        float cov[3];
        cov2d(pv, fx, fy, tanx, tany, c3, V, cov);
        float det = FMA(cov[0], cov[2], -(cov[1] * cov[1]));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float disc = sqrtf(fmaxf(0.1f, FMA(mid, mid, -det)));
        float l1 = mid + disc, l2 = mid - disc;
        float my_radius = ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
        /* ndc2Pix is evaluated in double upstream (double literals), then rounded to float */
        float px = (float)((((double)ppx + 1.0) * (double)W - 1.0) * 0.5);
        float py = (float)((((double)ppy + 1.0) * (double)H - 1.0) * 0.5);
        int ir = (int)my_radius;
        float fr = (float)ir;
        int rminx = (int)((px - fr) / 16.0f), rminy = (int)((py - fr) / 16.0f);
        int rmaxx = (int)(((px + fr) + 15.0f) / 16.0f), rmaxy = (int)(((py + fr) + 15.0f) / 16.0f);
        rminx = rminx < 0 ? 0 : rminx; rminx = rminx > gx ? gx : rminx;
        rminy = rminy < 0 ? 0 : rminy; rminy = rminy > gy ? gy : rminy;
        rmaxx = rmaxx < 0 ? 0 : rmaxx; rmaxx = rmaxx > gx ? gx : rmaxx;
        rmaxy = rmaxy < 0 ? 0 : rmaxy; rmaxy = rmaxy > gy ? gy : rmaxy;
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue;
        if (colors_precomp) {
            for (int k = 0; k < 3; ++k) rgb[3 * i + k] = colors_precomp[3 * i + k];
        } else {
            sh_to_rgb(deg, p, campos, shs + 48 * (size_t)i, rgb + 3 * i);
        }
        for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = c3[k];
        depth[i] = pv[2];
        radii[i] = ir;
        xy[2 * i] = px; xy[2 * i + 1] = py;
        conic_opacity[4 * i] = conic[0]; conic_opacity[4 * i + 1] = conic[1];
        conic_opacity[4 * i + 2] = conic[2]; conic_opacity[4 * i + 3] = opac[i];
        tiles_touched[i] = (uint32_t)((rmaxx - rminx) * (rmaxy - rminy));
        rect[4 * i] = rminx; rect[4 * i + 1] = rminy; rect[4 * i + 2] = rmaxx; rect[4 * i + 3] = rmaxy;
    }
}

/* Stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits (what CUB's
 * DeviceRadixSort::SortPairs yields: any stable sort gives the same permutation). */
static void radix_sort_pairs(uint64_t n, uint64_t* keys, uint32_t* vals, int bits) {
    if (n == 0) return;
    uint64_t* k2 = (uint64_t*)malloc(n * 8);
    uint32_t* v2 = (uint32_t*)malloc(n * 4);
    uint64_t *ka = keys, *kb = k2;
    uint32_t *va = vals, *vb = v2;
#ifdef _OPENMP
    int nt = omp_get_max_threads();
#else
    int nt = 1;
#endif
    uint64_t* hist = (uint64_t*)malloc((size_t)nt * 256 * 8);
    for (int shift = 0; shift < bits; shift += 8) {
        memset(hist, 0, (size_t)nt * 256 * 8);
#pragma omp parallel num_threads(nt)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            uint64_t lo = n * (uint64_t)t / nt, hi = n * (uint64_t)(t + 1) / nt;
            uint64_t* h = hist + (size_t)t * 256;
            for (uint64_t i = lo; i < hi; ++i) h[(ka[i] >> shift) & 255]++;
#pragma omp barrier
#pragma omp single
            {
                uint64_t run = 0;
                for (int d = 0; d < 256; ++d)
                    for (int tt = 0; tt < nt; ++tt) {
                        uint64_t c = hist[(size_t)tt * 256 + d];
                        hist[(size_t)tt * 256 + d] = run;
                        run += c;
                    }
            }
            for (uint64_t i = lo; i < hi; ++i) {
                uint64_t pos = h[(ka[i] >> shift) & 255]++;
                kb[pos] = ka[i];
                vb[pos] = va[i];
            }
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys) {
        memcpy(keys, ka, n * 8);
        memcpy(vals, va, n * 4);
    }
    free(k2); free(v2); free(hist);
}

/* Appendix A.6.  Returns R.  keys/vals must hold sum(tiles_touched) entries; ranges[gx*gy][2]. */
uint64_t orc_binning(int P, int W, int H, const int32_t* radii, const float* depth,
                     const uint32_t* tiles_touched, const int32_t* rect, uint64_t* keys,
                     uint32_t* vals, uint32_t* ranges) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    uint64_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        const int32_t* r = rect + 4 * i;
        for (int y = r[1]; y < r[3]; ++y)
            for (int x = r[0]; x < r[2]; ++x) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key = (key << 32) | as_uint(depth[i]);
                keys[off] = key;
                vals[off] = (uint32_t)i;
                ++off;
            }
    }
    (void)tiles_touched;
    uint32_t tiles = (uint32_t)(gx * gy);
    int msb = 0;
    while ((tiles >> msb) != 0) ++msb;
    radix_sort_pairs(off, keys, vals, 32 + msb);
    memset(ranges, 0, (size_t)tiles * 8);
    for (uint64_t i = 0; i < off; ++i) {
        uint32_t t = (uint32_t)(keys[i] >> 32);
        if (i == 0) ranges[2 * t] = 0;
        else {
            uint32_t pt = (uint32_t)(keys[i - 1] >> 32);
            if (t != pt) { ranges[2 * pt + 1] = (uint32_t)i; ranges[2 * t] = (uint32_t)i; }
        }
        if (i == off - 1) ranges[2 * t + 1] = (uint32_t)off;
    }
    return off;
}

uint64_t orc_count_pairs(int P, const uint32_t* tiles_touched) {
    uint64_t s = 0;
    for (int i = 0; i < P; ++i) s += tiles_touched[i];
    return s;
}

/* Appendix A.7 compositing.  out_color[3][H][W], out_depth[H][W], final_T[H][W], n_contrib[H][W]. */
void orc_composite(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                   const float* xy, const float* rgb, const float* depth,
                   const float* conic_opacity, const float* bg, float* out_color,
                   float* out_depth, float* final_T, uint32_t* n_contrib) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)W * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                float pfx = (float)px, pfy = (float)py;
                float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, D = 0.0f;
                uint32_t contributor = 0, last = 0;
                for (uint32_t e = beg; e < end; ++e) {
                    uint32_t g = point_list[e];
                    contributor++;
                    float dx = xy[2 * g] - pfx, dy = xy[2 * g + 1] - pfy;
                    const float* co = conic_opacity + 4 * (size_t)g;
                    float u = co[0] * dx;
                    float v = co[2] * dy;
                    float w = dy * v;
                    float s = FMA(dx, u, w);
                    float bxy = (co[1] * dx) * dy;
                    float power = FMA(s, -0.5f, -bxy);
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * orc_expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) break;
                    const float* c = rgb + 3 * (size_t)g;
                    C0 = FMA(c[0] * alpha, T, C0);
                    C1 = FMA(c[1] * alpha, T, C1);
                    C2 = FMA(c[2] * alpha, T, C2);
                    D = FMA(depth[g] * alpha, T, D);
                    T = test_T;
                    last = contributor;
                }
                size_t pix = (size_t)py * W + px;
                out_color[pix] = FMA(T, bg[0], C0);
                out_color[HW + pix] = FMA(T, bg[1], C1);
                out_color[2 * HW + pix] = FMA(T, bg[2], C2);
                out_depth[pix] = D;
                final_T[pix] = T;
                n_contrib[pix] = last;
            }
    }
}
