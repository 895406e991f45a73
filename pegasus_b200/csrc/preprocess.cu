// preprocess.cu — per-Gaussian stage: frustum test, cov3D from
// scale/quaternion, EWA projection with 0.3 dilation, conic, radius, pixel centre, tile rectangle,
// SH -> RGB; writes the 48-byte record (compositing + binning) and the depth sort key.
//
// HBM-bound: 236 B read per visible Gaussian (12 B when culled), 48+8+4+4 B written.  One thread
// per Gaussian; SH rows (192 B, 64-B aligned) are fetched with 12 independent 16-byte loads issued
// before first use so each warp keeps 12 x 32 requests in flight.
#include "pg_common.cuh"

namespace pg {

struct PreArgs {
    int P, deg, M;
    const float* means;
    const float* shs;
    const float* colors_precomp;
    const float* opac;
    const float* scales;
    const float* rots;
    const float* cov3d_precomp;
    const float* view;
    const float* proj;
    const float* campos;
    int W, H, gx, gy;
    float tanx, tany, fx, fy, scale_mod;
    int num_objects;
    int first[PG_MAX_OBJECTS + 1];
    // outputs
    int32_t* radii;
    GeomRec* recs;
    uint32_t* dkey;
    Counters* counters;
};

__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
    float t = mul(m[4 + r], y);
    t = fma(m[r], x, t);
    t = fma(m[8 + r], z, t);
    return add(m[12 + r], t);
}

__device__ __forceinline__ void cov3d_from_scale_rot(const float* s_in, float mod, const float* q, float* c) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3];
    float t;
    t = mul(z, z); t = fma(y, y, t); R[0][0] = sub(1.0f, add(t, t));
    t = fma(x, y, -mul(r, z));       R[0][1] = add(t, t);
    t = mul(r, y); t = fma(x, z, t); R[0][2] = add(t, t);
    t = mul(r, z); t = fma(x, y, t); R[1][0] = add(t, t);
    t = mul(z, z); t = fma(x, x, t); R[1][1] = sub(1.0f, add(t, t));
    t = fma(y, z, -mul(r, x));       R[1][2] = add(t, t);
    t = fma(x, z, -mul(r, y));       R[2][0] = add(t, t);
    t = mul(r, x); t = fma(y, z, t); R[2][1] = add(t, t);
    t = mul(y, y); t = fma(x, x, t); R[2][2] = sub(1.0f, add(t, t));
    float s[3] = {mul(mod, s_in[0]), mul(mod, s_in[1]), mul(mod, s_in[2])};
    float Mm[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) Mm[i][k] = mul(s[k], R[i][k]);
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
            float a = mul(Mm[i][1], Mm[j][1]);
            a = fma(Mm[i][0], Mm[j][0], a);
            a = fma(Mm[i][2], Mm[j][2], a);
            c[idx++] = a;
        }
}

__constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};

__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
    const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
    b[0] = C0;
    if (deg > 0) {
        b[1] = -mul(C1, y); b[2] = mul(C1, z); b[3] = -mul(C1, x);
        if (deg > 1) {
            float xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
            float xy = mul(x, y), yz = mul(y, z), xz = mul(x, z);
            b[4] = mul(1.0925484305920792f, xy);
            b[5] = mul(-1.0925484305920792f, yz);
            b[6] = mul(0.31539156525252005f, sub(sub(mul(2.0f, zz), xx), yy));
            b[7] = mul(-1.0925484305920792f, xz);
            b[8] = mul(0.5462742152960396f, sub(xx, yy));
            if (deg > 2) {
                b[9] = mul(mul(-0.5900435899266435f, y), sub(mul(3.0f, xx), yy));
                b[10] = mul(mul(2.890611442640554f, xy), z);
                b[11] = mul(mul(-0.4570457994644658f, y), sub(sub(mul(4.0f, zz), xx), yy));
                b[12] = mul(mul(0.3731763325901154f, z), sub(sub(mul(2.0f, zz), mul(3.0f, xx)), mul(3.0f, yy)));
                b[13] = mul(mul(-0.4570457994644658f, x), sub(sub(mul(4.0f, zz), xx), yy));
                b[14] = mul(mul(1.445305721320277f, z), sub(xx, yy));
                b[15] = mul(mul(-0.5900435899266435f, x), sub(xx, mul(3.0f, yy)));
            }
        }
    }
}

// bounded for 6 CTAs per SM (<= 40 registers): 0.143 vs 0.145 ms unbounded (48 registers), 0.161 at 8
__global__ void __launch_bounds__(256, 6) preprocess_kernel(const PreArgs a) {
    __shared__ float s_view[16], s_proj[16], s_cam[3];
    if (threadIdx.x < 16) {
        s_view[threadIdx.x] = a.view[threadIdx.x];
        s_proj[threadIdx.x] = a.proj[threadIdx.x];
    }
    if (threadIdx.x < 3) s_cam[threadIdx.x] = a.campos[threadIdx.x];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = idx < a.P;  // no early return: the warp reductions at the end want all 32 lanes

    int radius_out = 0;
    uint32_t area = 0;
    uint32_t key = 0xFFFFFFFFu;
    bool visible = false;

    const float px3 = in_range ? a.means[3 * (size_t)idx] : 0.0f, py3 = in_range ? a.means[3 * (size_t)idx + 1] : 0.0f,
                pz3 = in_range ? a.means[3 * (size_t)idx + 2] : 0.0f;
    float pv[3];
    pv[0] = xform_row(s_view, 0, px3, py3, pz3);
    pv[1] = xform_row(s_view, 1, px3, py3, pz3);
    pv[2] = xform_row(s_view, 2, px3, py3, pz3);
    if (in_range && pv[2] > 0.2f) {
        float hx = xform_row(s_proj, 0, px3, py3, pz3);
        float hy = xform_row(s_proj, 1, px3, py3, pz3);
        float hw = xform_row(s_proj, 3, px3, py3, pz3);
        float pw = div(1.0f, add(hw, 0.0000001f));
        float ppx = mul(hx, pw), ppy = mul(hy, pw);
        float c3[6];
        if (a.cov3d_precomp) {
#pragma unroll
            for (int k = 0; k < 6; ++k) c3[k] = a.cov3d_precomp[6 * (size_t)idx + k];
        } else {
            float s[3] = {a.scales[3 * (size_t)idx], a.scales[3 * (size_t)idx + 1], a.scales[3 * (size_t)idx + 2]};
            float4 q4 = *reinterpret_cast<const float4*>(a.rots + 4 * (size_t)idx);
            float q[4] = {q4.x, q4.y, q4.z, q4.w};
            cov3d_from_scale_rot(s, a.scale_mod, q, c3);
        }
        // --- EWA 2D covariance (A.4)
        float tz = pv[2];
        float limx = mul(1.3f, a.tanx), limy = mul(1.3f, a.tany);
        float txtz = div(pv[0], tz), tytz = div(pv[1], tz);
        float tx = mul(fminf(limx, fmaxf(-limx, txtz)), tz);
        float ty = mul(fminf(limy, fmaxf(-limy, tytz)), tz);
        float J00 = div(a.fx, tz), J11 = div(a.fy, tz);
        float tz2 = mul(tz, tz);
        float J02 = div(-mul(a.fx, tx), tz2);
        float J12 = div(-mul(a.fy, ty), tz2);
        float T0[3], T1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T0[k] = fma(s_view[4 * k + 2], J02, mul(s_view[4 * k], J00));
            T1[k] = fma(s_view[4 * k + 2], J12, mul(s_view[4 * k + 1], J11));
        }
        float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        float A0[3], A1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float u = mul(T0[1], Vrk[1][k]);
            u = fma(T0[0], Vrk[0][k], u);
            A0[k] = fma(T0[2], Vrk[2][k], u);
            float v = mul(T1[1], Vrk[1][k]);
            v = fma(T1[0], Vrk[0][k], v);
            A1[k] = fma(T1[2], Vrk[2][k], v);
        }
        float c00 = mul(A0[1], T0[1]); c00 = fma(A0[0], T0[0], c00); c00 = fma(A0[2], T0[2], c00);
        float c01 = mul(A1[1], T0[1]); c01 = fma(A1[0], T0[0], c01); c01 = fma(A1[2], T0[2], c01);
        float c11 = mul(A1[1], T1[1]); c11 = fma(A1[0], T1[0], c11); c11 = fma(A1[2], T1[2], c11);
        float cva = add(c00, 0.3f), cvb = c01, cvc = add(c11, 0.3f);
        float det = fma(cva, cvc, -mul(cvb, cvb));
        if (det != 0.0f) {
            float det_inv = div(1.0f, det);
            float conx = mul(cvc, det_inv), cony = mul(-cvb, det_inv), conz = mul(cva, det_inv);
            float mid = mul(0.5f, add(cva, cvc));
            float disc = sqrt(fmaxf(0.1f, fma(mid, mid, -det)));
            float l1 = add(mid, disc), l2 = sub(mid, disc);
            float my_radius = ceilf(mul(3.0f, sqrt(fmaxf(l1, l2))));
            float pixx = (float)__dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)ppx, 1.0), (double)a.W), 1.0), 0.5);
            float pixy = (float)__dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)ppy, 1.0), (double)a.H), 1.0), 0.5);
            int ir = (int)my_radius;
            const ushort4 tr = tile_rect(pixx, pixy, ir, a.gx, a.gy);
            const int rminx = tr.x, rminy = tr.y, rmaxx = tr.z, rmaxy = tr.w;
            if ((rmaxx - rminx) * (rmaxy - rminy) != 0) {
                // --- colour (A.5)
                float rgb[3];
                if (a.colors_precomp) {
                    rgb[0] = a.colors_precomp[3 * (size_t)idx];
                    rgb[1] = a.colors_precomp[3 * (size_t)idx + 1];
                    rgb[2] = a.colors_precomp[3 * (size_t)idx + 2];
                } else {
                    float dx = sub(px3, s_cam[0]), dy = sub(py3, s_cam[1]), dz = sub(pz3, s_cam[2]);
                    float l2n = mul(dy, dy); l2n = fma(dx, dx, l2n); l2n = fma(dz, dz, l2n);
                    float len = sqrt(l2n);
                    float b[16];
                    sh_basis(a.deg, div(dx, len), div(dy, len), div(dz, len), b);
                    const int n = (a.deg + 1) * (a.deg + 1);
                    const float* shp = a.shs + (size_t)idx * a.M * 3;
                    float acc0, acc1, acc2;
                    if (a.M == 16 && a.deg == 3) {
                        const float4* sh4 = reinterpret_cast<const float4*>(shp);
                        float4 v[12];
#pragma unroll
                        for (int k = 0; k < 12; ++k) v[k] = __ldg(sh4 + k);
                        const float* f = reinterpret_cast<const float*>(v);
                        acc0 = mul(b[0], f[0]); acc1 = mul(b[0], f[1]); acc2 = mul(b[0], f[2]);
#pragma unroll
                        for (int k = 1; k < 16; ++k) {
                            acc0 = fma(b[k], f[3 * k], acc0);
                            acc1 = fma(b[k], f[3 * k + 1], acc1);
                            acc2 = fma(b[k], f[3 * k + 2], acc2);
                        }
                    } else {
                        acc0 = mul(b[0], shp[0]); acc1 = mul(b[0], shp[1]); acc2 = mul(b[0], shp[2]);
                        for (int k = 1; k < n; ++k) {
                            acc0 = fma(b[k], shp[3 * k], acc0);
                            acc1 = fma(b[k], shp[3 * k + 1], acc1);
                            acc2 = fma(b[k], shp[3 * k + 2], acc2);
                        }
                    }
                    acc0 = add(acc0, 0.5f); acc1 = add(acc1, 0.5f); acc2 = add(acc2, 0.5f);
                    rgb[0] = acc0 < 0.0f ? 0.0f : acc0;
                    rgb[1] = acc1 < 0.0f ? 0.0f : acc1;
                    rgb[2] = acc2 < 0.0f ? 0.0f : acc2;
                }
                const float op = a.opac[idx];
                int obj = 0;
                for (int k = 0; k < a.num_objects; ++k)
                    if (idx >= a.first[k] && idx < a.first[k + 1]) obj = k + 1;
                GeomRec rec;
                rec.a = make_float4(pixx, pixy, conx, cony);
                // cut: alpha = min(.99, op*exp(power)) < 1/255 is certain for power < -(ln(255*op) + 0.01)
                // (the 1 % margin dwarfs the error of logf and of the exp polynomial); never above 0.
                // __logf is the one approximate operation here (abs. error < 1e-5 on this range): it only moves the
                // cut inside the 0.01 margin.  The low 7 mantissa bits carry the object id (bits 0-5) and the
                // "general path" flag (bit 6: opacity > 0.99, so min(0.99, .) can bind, or a non-finite record);
                // clearing them makes the cut at most 127 ulp (< 2e-3 at -80, < 1e-4 typically) LESS negative,
                // which the margin covers as well.
                float cut = -80.0f;
                if (op > 0.0f) cut = fminf(-(__logf(255.0f * op) + 0.01f), -1e-6f);
                if (!(cut > -80.0f)) cut = -80.0f;
                const bool plain = op <= 0.99f && isfinite(pixx) && isfinite(pixy) && isfinite(conx) && isfinite(cony) &&
                                   isfinite(conz) && isfinite(pv[2]) && isfinite(rgb[0]) && isfinite(rgb[1]) && isfinite(rgb[2]);
                cut = __int_as_float((__float_as_int(cut) & ~127) | obj | (plain ? 0 : PG_REC_GENERAL));
                rec.b = make_float4(conz, op, pv[2], cut);
                rec.c = make_float4(rgb[0], rgb[1], rgb[2], __int_as_float(ir));
                a.recs[idx] = rec;
                radius_out = ir;
                key = __float_as_uint(pv[2]);
                visible = true;
                area = (uint32_t)(rmaxx - rminx) * (uint32_t)(rmaxy - rminy);
            }
        }
    }
    if (in_range) {
        a.radii[idx] = radius_out;
        a.dkey[idx] = key;
    }
    // visible count and the reference's R = sum of all rectangle areas (pg_status.num_rendered): reduced per CTA — one
    // atomic per warp on the same two words serialises in L2 (94 k warps: ~45 us of a 145 us kernel)
    const unsigned vis = __ballot_sync(0xffffffffu, visible);
    unsigned long long full = area;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) full += __shfl_xor_sync(0xffffffffu, full, o);
    __shared__ unsigned long long s_full[8];
    __shared__ uint32_t s_vis[8];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_full[warp] = full; s_vis[warp] = __popc(vis); }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long f = 0;
        uint32_t v = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { f += s_full[w]; v += s_vis[w]; }
        if (v) {
            atomicAdd(&a.counters->num_visible, v);
            atomicAdd(&a.counters->rendered_full, f);
        }
    }
}

__global__ void mark_visible_kernel(int P, const float* means, const float* view, uint8_t* present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float z = xform_row(view, 2, means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]);
    present[idx] = z > 0.2f ? 1 : 0;
}

int launch_preprocess(const pg_raster_settings* s, const pg_gaussians* g, const pg_object_table* objs,
                      int32_t* radii, GeomRec* recs, uint32_t* dkey, Counters* counters, cudaStream_t stream) {
    PreArgs a;
    a.P = g->P; a.deg = s->sh_degree; a.M = g->sh_coeffs;
    a.means = g->means3D; a.shs = g->shs; a.colors_precomp = g->colors_precomp; a.opac = g->opacities;
    a.scales = g->scales; a.rots = g->rotations; a.cov3d_precomp = g->cov3D_precomp;
    a.view = s->viewmatrix; a.proj = s->projmatrix; a.campos = s->campos;
    a.W = s->image_width; a.H = s->image_height;
    a.gx = (a.W + PG_TILE - 1) / PG_TILE; a.gy = (a.H + PG_TILE - 1) / PG_TILE;
    a.tanx = s->tanfovx; a.tany = s->tanfovy;
    a.fx = (float)a.W / (2.0f * s->tanfovx);
    a.fy = (float)a.H / (2.0f * s->tanfovy);
    a.scale_mod = s->scale_modifier;
    a.num_objects = objs ? objs->num_objects : 0;
    for (int k = 0; k <= PG_MAX_OBJECTS; ++k) a.first[k] = (objs && k <= objs->num_objects) ? objs->first[k] : 0;
    a.radii = radii; a.recs = recs; a.dkey = dkey; a.counters = counters;
    if (a.P == 0) return PG_OK;
    preprocess_kernel<<<(a.P + 255) / 256, 256, 0, stream>>>(a);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_mark_visible(int P, const float* means, const float* view, uint8_t* present, cudaStream_t stream) {
    if (P == 0) return PG_OK;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means, view, present);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

}  // namespace pg
