// upstream_style.cu — BASELINE, not product: a plain restatement of how the public 3DGS forward
// rasterizer (graphdeco-inria/diff-gaussian-rasterization + the depth accumulation of the "depth" fork
// the reference pins as an un-vendored submodule) is organised on a GPU:
//
//   preprocess (one thread per Gaussian)  ->  cub::DeviceScan::InclusiveSum of tiles_touched
//   ->  blocking D2H copy of the pair count  ->  duplicateWithKeys (64-bit tile|depth keys)
//   ->  cub::DeviceRadixSort::SortPairs  ->  identifyTileRanges  ->  render (one 16x16 CTA per tile,
//   256-entry cooperative fetches, __syncthreads_count early exit).
//
// The pinned source is not in /root/reference and cannot be fetched offline, so this file
// is written from SURVEY Appendix A; natural float expressions, nvcc default FMA contraction, IEEE
// expf — no Blackwell features, CUB for scan/sort exactly as upstream uses it.  It exists so that
// bench.py can put a number on north_star's ">= 1.5x the reference CUDA rasterizer" and so that the
// GPU tests have a second, independently written implementation to compare images with.
// Nothing under pegasus_b200/ links or loads it.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE (BLOCK_X * BLOCK_Y)

namespace {

__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f

struct Buffers {
    // geometry state (per Gaussian)
    float* depths = nullptr; float2* xy = nullptr; float* cov3D = nullptr; float4* conic_opacity = nullptr;
    float* rgb = nullptr; uint32_t* tiles_touched = nullptr; uint32_t* point_offsets = nullptr; bool* clamped = nullptr;
    void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0; size_t capP = 0;
    // binning state (per pair)
    uint64_t* keys_unsorted = nullptr; uint64_t* keys = nullptr; uint32_t* vals_unsorted = nullptr; uint32_t* vals = nullptr;
    void* sort_tmp = nullptr; size_t sort_tmp_bytes = 0; size_t capR = 0;
    // image state
    uint2* ranges = nullptr; uint32_t* n_contrib = nullptr; float* accum_alpha = nullptr; size_t capPix = 0;
};

__device__ inline float3 xform4x3(const float3& p, const float* m) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ inline float4 xform4x4(const float3& p, const float* m) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}
__device__ inline float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

__device__ void tile_rect(float2 p, int r, uint2& lo, uint2& hi, dim3 grid) {
    lo = make_uint2(min((int)grid.x, max(0, (int)((p.x - r) / BLOCK_X))), min((int)grid.y, max(0, (int)((p.y - r) / BLOCK_Y))));
    hi = make_uint2(min((int)grid.x, max(0, (int)((p.x + r + BLOCK_X - 1) / BLOCK_X))),
                    min((int)grid.y, max(0, (int)((p.y + r + BLOCK_Y - 1) / BLOCK_Y))));
}

__device__ float3 sh_to_rgb(int idx, int deg, int M, const float3* means, float3 campos, const float* shs, bool* clamped) {
    float3 pos = means[idx];
    float3 dir = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x /= len; dir.y /= len; dir.z /= len;
    const float3* sh = reinterpret_cast<const float3*>(shs) + (size_t)idx * M;
    auto ax = [](float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); };
    auto ad = [](float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); };
    float3 res = ax(sh[0], SH_C0);
    if (deg > 0) {
        float x = dir.x, y = dir.y, z = dir.z;
        res = ad(ad(ad(res, ax(sh[1], -SH_C1 * y)), ax(sh[2], SH_C1 * z)), ax(sh[3], -SH_C1 * x));
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            res = ad(res, ax(sh[4], SH_C2[0] * xy));
            res = ad(res, ax(sh[5], SH_C2[1] * yz));
            res = ad(res, ax(sh[6], SH_C2[2] * (2.0f * zz - xx - yy)));
            res = ad(res, ax(sh[7], SH_C2[3] * xz));
            res = ad(res, ax(sh[8], SH_C2[4] * (xx - yy)));
            if (deg > 2) {
                res = ad(res, ax(sh[9], SH_C3[0] * y * (3.0f * xx - yy)));
                res = ad(res, ax(sh[10], SH_C3[1] * xy * z));
                res = ad(res, ax(sh[11], SH_C3[2] * y * (4.0f * zz - xx - yy)));
                res = ad(res, ax(sh[12], SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy)));
                res = ad(res, ax(sh[13], SH_C3[4] * x * (4.0f * zz - xx - yy)));
                res = ad(res, ax(sh[14], SH_C3[5] * z * (xx - yy)));
                res = ad(res, ax(sh[15], SH_C3[6] * x * (xx - 3.0f * yy)));
            }
        }
    }
    res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
    clamped[3 * idx + 0] = res.x < 0; clamped[3 * idx + 1] = res.y < 0; clamped[3 * idx + 2] = res.z < 0;
    return make_float3(fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f));
}

// Sigma = (S R)^T (S R) with R from the (unnormalised-in-kernel) quaternion; upper triangle
__device__ void make_cov3D(float3 s, float mod, float4 q, float* c) {
    float r = q.x, x = q.y, y = q.z, z = q.w;
    float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                     {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                     {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    float sc[3] = {mod * s.x, mod * s.y, mod * s.z};
    float Mm[3][3];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) Mm[i][k] = R[i][k] * sc[k];
    int n = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) c[n++] = Mm[i][0] * Mm[j][0] + Mm[i][1] * Mm[j][1] + Mm[i][2] * Mm[j][2];
}

__device__ float3 make_cov2D(const float3& mean, float fx, float fy, float tanx, float tany, const float* c3, const float* V) {
    float3 t = xform4x3(mean, V);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    t.x = fminf(limx, fmaxf(-limx, t.x / t.z)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, t.y / t.z)) * t.z;
    // rows of T = J * W (2 x 3)
    float J00 = fx / t.z, J02 = -(fx * t.x) / (t.z * t.z), J11 = fy / t.z, J12 = -(fy * t.y) / (t.z * t.z);
    float T0[3], T1[3];
    for (int k = 0; k < 3; ++k) {
        T0[k] = J00 * V[4 * k] + J02 * V[4 * k + 2];
        T1[k] = J11 * V[4 * k + 1] + J12 * V[4 * k + 2];
    }
    float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float A0[3], A1[3];
    for (int k = 0; k < 3; ++k) {
        A0[k] = T0[0] * S[0][k] + T0[1] * S[1][k] + T0[2] * S[2][k];
        A1[k] = T1[0] * S[0][k] + T1[1] * S[1][k] + T1[2] * S[2][k];
    }
    float c00 = A0[0] * T0[0] + A0[1] * T0[1] + A0[2] * T0[2];
    float c01 = A1[0] * T0[0] + A1[1] * T0[1] + A1[2] * T0[2];
    float c11 = A1[0] * T1[0] + A1[1] * T1[1] + A1[2] * T1[2];
    return make_float3(c00 + 0.3f, c01, c11 + 0.3f);
}

__global__ void preprocess(int P, int D, int M, const float* means3D, const float3* scales, float mod, const float4* rots,
                           const float* opac, const float* shs, bool* clamped, const float* V, const float* PM,
                           const float3* campos, int W, int H, float tanx, float tany, float fx, float fy, int* radii,
                           float2* xy, float* depths, float* cov3Ds, float* rgb, float4* conic_opacity, dim3 grid,
                           uint32_t* tiles_touched) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    radii[idx] = 0;
    tiles_touched[idx] = 0;
    float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    float3 pv = xform4x3(p, V);
    if (pv.z <= 0.2f) return;
    float4 ph = xform4x4(p, PM);
    float pw = 1.0f / (ph.w + 0.0000001f);
    float3 pp = make_float3(ph.x * pw, ph.y * pw, ph.z * pw);
    make_cov3D(scales[idx], mod, rots[idx], cov3Ds + 6 * (size_t)idx);
    float3 cov = make_cov2D(p, fx, fy, tanx, tany, cov3Ds + 6 * (size_t)idx, V);
    float det = cov.x * cov.z - cov.y * cov.y;
    if (det == 0.0f) return;
    float det_inv = 1.f / det;
    float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);
    float mid = 0.5f * (cov.x + cov.z);
    float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det)), l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
    float2 pix = make_float2(ndc2pix(pp.x, W), ndc2pix(pp.y, H));
    uint2 lo, hi;
    tile_rect(pix, (int)my_radius, lo, hi, grid);
    if ((hi.x - lo.x) * (hi.y - lo.y) == 0) return;
    float3 c = sh_to_rgb(idx, D, M, reinterpret_cast<const float3*>(means3D), *campos, shs, clamped);
    rgb[3 * idx] = c.x; rgb[3 * idx + 1] = c.y; rgb[3 * idx + 2] = c.z;
    depths[idx] = pv.z;
    radii[idx] = (int)my_radius;
    xy[idx] = pix;
    conic_opacity[idx] = make_float4(conic.x, conic.y, conic.z, opac[idx]);
    tiles_touched[idx] = (hi.y - lo.y) * (hi.x - lo.x);
}

__global__ void duplicate_with_keys(int P, const float2* xy, const float* depths, const uint32_t* offsets, uint64_t* keys,
                                    uint32_t* vals, const int* radii, dim3 grid) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || radii[idx] <= 0) return;
    uint32_t off = idx == 0 ? 0 : offsets[idx - 1];
    uint2 lo, hi;
    tile_rect(xy[idx], radii[idx], lo, hi, grid);
    for (uint32_t y = lo.y; y < hi.y; ++y)
        for (uint32_t x = lo.x; x < hi.x; ++x) {
            uint64_t key = y * grid.x + x;
            key <<= 32;
            key |= *reinterpret_cast<const uint32_t*>(&depths[idx]);
            keys[off] = key;
            vals[off] = idx;
            ++off;
        }
}

__global__ void identify_tile_ranges(int L, const uint64_t* keys, uint2* ranges) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    uint32_t tile = keys[idx] >> 32;
    if (idx == 0) ranges[tile].x = 0;
    else {
        uint32_t prev = keys[idx - 1] >> 32;
        if (tile != prev) { ranges[prev].y = idx; ranges[tile].x = idx; }
    }
    if (idx == L - 1) ranges[tile].y = L;
}

__global__ void __launch_bounds__(BLOCK_SIZE)
render(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
       const float2* __restrict__ xy, const float* __restrict__ features, const float* __restrict__ depths,
       const float4* __restrict__ conic_opacity, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
       const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth) {
    const uint32_t hb = (W + BLOCK_X - 1) / BLOCK_X;
    uint2 pmin = {blockIdx.x * BLOCK_X, blockIdx.y * BLOCK_Y};
    uint2 pix = {pmin.x + threadIdx.x, pmin.y + threadIdx.y};
    uint32_t pix_id = W * pix.y + pix.x;
    float2 pixf = {(float)pix.x, (float)pix.y};
    bool inside = pix.x < W && pix.y < H;
    bool done = !inside;
    uint2 range = ranges[blockIdx.y * hb + blockIdx.x];
    const int rounds = (range.y - range.x + BLOCK_SIZE - 1) / BLOCK_SIZE;
    int todo = range.y - range.x;
    __shared__ int s_id[BLOCK_SIZE];
    __shared__ float2 s_xy[BLOCK_SIZE];
    __shared__ float4 s_co[BLOCK_SIZE];
    __shared__ float s_depth[BLOCK_SIZE];
    float T = 1.0f;
    uint32_t contributor = 0, last = 0;
    float C[3] = {0, 0, 0}, Dp = 0;
    const int rank = threadIdx.y * BLOCK_X + threadIdx.x;
    for (int i = 0; i < rounds; ++i, todo -= BLOCK_SIZE) {
        if (__syncthreads_count(done) == BLOCK_SIZE) break;
        int progress = i * BLOCK_SIZE + rank;
        if (range.x + progress < range.y) {
            int id = point_list[range.x + progress];
            s_id[rank] = id; s_xy[rank] = xy[id]; s_co[rank] = conic_opacity[id]; s_depth[rank] = depths[id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK_SIZE, todo); ++j) {
            ++contributor;
            float2 d = {s_xy[j].x - pixf.x, s_xy[j].y - pixf.y};
            float4 co = s_co[j];
            float power = -0.5f * (co.x * d.x * d.x + co.z * d.y * d.y) - co.y * d.x * d.y;
            if (power > 0.0f) continue;
            float alpha = fminf(0.99f, co.w * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            for (int ch = 0; ch < 3; ++ch) C[ch] += features[s_id[j] * 3 + ch] * alpha * T;
            Dp += s_depth[j] * alpha * T;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last;
        for (int ch = 0; ch < 3; ++ch) out_color[ch * H * W + pix_id] = C[ch] + T * bg[ch];
        out_depth[pix_id] = Dp;
    }
}

uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) ++msb;
    return msb;
}

template <typename T>
bool grow(T*& p, size_t n) {
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc(&p, n * sizeof(T)) == cudaSuccess;
}

}  // namespace

extern "C" {

void* base_create(void) { return new Buffers(); }

void base_destroy(void* h) {
    Buffers* b = static_cast<Buffers*>(h);
    if (!b) return;
    void* ptrs[] = {b->depths, b->xy, b->cov3D, b->conic_opacity, b->rgb, b->tiles_touched, b->point_offsets, b->clamped,
                    b->scan_tmp, b->keys_unsorted, b->keys, b->vals_unsorted, b->vals, b->sort_tmp, b->ranges, b->n_contrib,
                    b->accum_alpha};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete b;
}

// One forward pass, upstream style.  Returns the number of (tile, Gaussian) pairs, or -1 on a CUDA error.
// Synchronises with the host once (the pair count), as upstream does.
long long base_forward(void* h, int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                       const float* shs, const float* opacities, const float* scales, float scale_modifier,
                       const float* rotations, const float* viewmatrix, const float* projmatrix, const float* campos,
                       float tanfovx, float tanfovy, float* out_color, float* out_depth, int* radii, void* stream_) {
    Buffers& b = *static_cast<Buffers*>(h);
    cudaStream_t st = (cudaStream_t)stream_;
    const float fy = H / (2.0f * tanfovy), fx = W / (2.0f * tanfovx);
    dim3 grid((W + BLOCK_X - 1) / BLOCK_X, (H + BLOCK_Y - 1) / BLOCK_Y, 1), block(BLOCK_X, BLOCK_Y, 1);
    if ((size_t)P > b.capP) {
        bool ok = grow(b.depths, P) && grow(b.xy, P) && grow(b.cov3D, (size_t)P * 6) && grow(b.conic_opacity, P) &&
                  grow(b.rgb, (size_t)P * 3) && grow(b.tiles_touched, P) && grow(b.point_offsets, P) && grow(b.clamped, (size_t)P * 3);
        if (!ok) return -1;
        cub::DeviceScan::InclusiveSum(nullptr, b.scan_tmp_bytes, b.tiles_touched, b.point_offsets, P);
        if (b.scan_tmp) cudaFree(b.scan_tmp);
        if (cudaMalloc(&b.scan_tmp, b.scan_tmp_bytes) != cudaSuccess) return -1;
        b.capP = P;
    }
    if ((size_t)W * H > b.capPix) {
        if (!(grow(b.ranges, (size_t)W * H) && grow(b.n_contrib, (size_t)W * H) && grow(b.accum_alpha, (size_t)W * H))) return -1;
        b.capPix = (size_t)W * H;
    }
    if (P == 0) return 0;
    preprocess<<<(P + 255) / 256, 256, 0, st>>>(P, D, M, means3D, (const float3*)scales, scale_modifier, (const float4*)rotations,
                                                opacities, shs, b.clamped, viewmatrix, projmatrix, (const float3*)campos, W, H,
                                                tanfovx, tanfovy, fx, fy, radii, b.xy, b.depths, b.cov3D, b.rgb,
                                                b.conic_opacity, grid, b.tiles_touched);
    cub::DeviceScan::InclusiveSum(b.scan_tmp, b.scan_tmp_bytes, b.tiles_touched, b.point_offsets, P, st);
    uint32_t R = 0;
    if (cudaMemcpyAsync(&R, b.point_offsets + P - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if ((size_t)R > b.capR) {
        size_t n = (size_t)R + R / 4 + 1024;
        if (!(grow(b.keys_unsorted, n) && grow(b.keys, n) && grow(b.vals_unsorted, n) && grow(b.vals, n))) return -1;
        cub::DeviceRadixSort::SortPairs(nullptr, b.sort_tmp_bytes, b.keys_unsorted, b.keys, b.vals_unsorted, b.vals, (int)n);
        if (b.sort_tmp) cudaFree(b.sort_tmp);
        if (cudaMalloc(&b.sort_tmp, b.sort_tmp_bytes) != cudaSuccess) return -1;
        b.capR = n;
    }
    cudaMemsetAsync(b.ranges, 0, (size_t)grid.x * grid.y * sizeof(uint2), st);
    if (R > 0) {
        duplicate_with_keys<<<(P + 255) / 256, 256, 0, st>>>(P, b.xy, b.depths, b.point_offsets, b.keys_unsorted, b.vals_unsorted, radii, grid);
        const int bit = (int)higher_msb(grid.x * grid.y);
        cub::DeviceRadixSort::SortPairs(b.sort_tmp, b.sort_tmp_bytes, b.keys_unsorted, b.keys, b.vals_unsorted, b.vals, (int)R, 0, 32 + bit, st);
        identify_tile_ranges<<<(R + 255) / 256, 256, 0, st>>>((int)R, b.keys, b.ranges);
    }
    render<<<grid, block, 0, st>>>(b.ranges, b.vals, W, H, b.xy, b.rgb, b.depths, b.conic_opacity, b.accum_alpha, b.n_contrib, bg,
                                   out_color, out_depth);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (long long)R;
}

// tests: sorted keys / values of the last forward
int base_export(void* h, long long R, uint64_t* keys, uint32_t* vals, uint32_t* ranges, int tiles, void* stream_) {
    Buffers& b = *static_cast<Buffers*>(h);
    cudaStream_t st = (cudaStream_t)stream_;
    if (R > 0) {
        cudaMemcpyAsync(keys, b.keys, (size_t)R * 8, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(vals, b.vals, (size_t)R * 4, cudaMemcpyDeviceToDevice, st);
    }
    cudaMemcpyAsync(ranges, b.ranges, (size_t)tiles * 8, cudaMemcpyDeviceToDevice, st);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
