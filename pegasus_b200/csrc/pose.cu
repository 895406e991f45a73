// pose.cu — per-object 6DoF pose transform written straight into the composed scene
// (replaces GaussianModel.apply_transformation + merge_gaussians: src/gs/gaussian_model.py:482-591).
//
//   means     x' = R (x - pivot) + pivot + t                      (apply_transformation_on_xyz)
//   rotations q' = q_R (x) normalize(q)   (w,x,y,z)               (apply_rotation_on_splats)
//   SH        c_l' = D_l c_l, l = 1..3, per colour channel        (apply_rotation_on_sh)
//
// HBM-bound: 208 B read + 208 B written per object Gaussian.  grid = (ceil(n_max/256), K); a CTA
// handles 256 Gaussians of ONE object: their 256 x 180 B of SH coefficients are one contiguous
// 46 080-byte block that a single TMA bulk copy (cp.async.bulk) stages into shared memory; threads
// then read their 45 coefficients at a stride of 45 words (conflict-free), rotate in registers and
// the CTA stores the block back coalesced into the scene's (P,16,3) SH rows.
#include "pg_common.cuh"

namespace pg {

struct PoseDev {  // device copy of pg_pose without padding surprises
    float R[9], t[3], pivot[3], q[4], D1[9], D2[25], D3[49];
    int rotate_sh;
};

struct PoseArgs {
    int K;
    int first[PG_MAX_OBJECTS + 1];
    const PoseDev* poses;  // device [K]
    const float* xyz;
    const float* rot;
    const float* rest;
    int scene_offset;
    float* means;
    float* rots;
    float* shs;
};

__device__ __forceinline__ uint32_t smem_u32p(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) pose_kernel(const PoseArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_sh = reinterpret_cast<float*>(smem_raw);               // 256*45 floats
    PoseDev* s_pose = reinterpret_cast<PoseDev*>(s_sh + 256 * 45);  // 103 words
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + 256 * 45 * 4 + 512);

    const int k = blockIdx.y;
    const int n_k = a.first[k + 1] - a.first[k];
    const int i0 = blockIdx.x * 256;
    if (i0 >= n_k) return;
    const int cnt = min(256, n_k - i0);
    const int tid = threadIdx.x;
    const size_t src0 = (size_t)a.first[k] + i0;                   // first canonical Gaussian of this CTA
    const size_t dst0 = (size_t)a.scene_offset + a.first[k] + i0;  // its slot in the composed scene

    const float* src_sh = a.rest + src0 * 45;
    const uint32_t bytes = (uint32_t)cnt * 180u;
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src_sh) & 15) == 0) && (bytes % 16 == 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32p(s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (bulk_ok) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32p(s_bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32p(s_sh)), "l"(src_sh), "r"(bytes), "r"(smem_u32p(s_bar)) : "memory");
        }
    }
    // pose packet -> shared
    {
        const float* pp = reinterpret_cast<const float*>(a.poses + k);
        float* sp = reinterpret_cast<float*>(s_pose);
        for (int i = tid; i < (int)(sizeof(PoseDev) / 4); i += 256) sp[i] = pp[i];
    }
    if (!bulk_ok) {
        for (int i = tid; i < cnt * 45; i += 256) s_sh[i] = src_sh[i];
    }
    __syncthreads();

    // ---- means + quaternion (thread per Gaussian) while the SH block is in flight ----
    if (tid < cnt) {
        const PoseDev& p = *s_pose;
        const size_t s = src0 + tid, d = dst0 + tid;
        float x = a.xyz[3 * s], y = a.xyz[3 * s + 1], z = a.xyz[3 * s + 2];
        float d0 = sub(x, p.pivot[0]), d1 = sub(y, p.pivot[1]), d2 = sub(z, p.pivot[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float v = mul(p.R[3 * i], d0);
            v = fma(p.R[3 * i + 1], d1, v);
            v = fma(p.R[3 * i + 2], d2, v);
            a.means[3 * d + i] = add(add(v, p.pivot[i]), p.t[i]);
        }
        float4 q = *reinterpret_cast<const float4*>(a.rot + 4 * s);
        float nrm = sqrt(fma(q.w, q.w, fma(q.z, q.z, fma(q.y, q.y, mul(q.x, q.x)))));
        float iw = div(q.x, nrm), ix = div(q.y, nrm), iy = div(q.z, nrm), iz = div(q.w, nrm);
        const float aw = p.q[0], ax = p.q[1], ay = p.q[2], az = p.q[3];
        float4 o;
        o.x = sub(sub(sub(mul(aw, iw), mul(ax, ix)), mul(ay, iy)), mul(az, iz));
        o.y = sub(add(add(mul(aw, ix), mul(ax, iw)), mul(ay, iz)), mul(az, iy));
        o.z = add(add(sub(mul(aw, iy), mul(ax, iz)), mul(ay, iw)), mul(az, ix));
        o.w = add(sub(add(mul(aw, iz), mul(ax, iy)), mul(ay, ix)), mul(az, iw));
        *reinterpret_cast<float4*>(a.rots + 4 * d) = o;
    }

    if (bulk_ok) {
        asm volatile(
            "{\n.reg .pred p;\nPWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra PDONE;\nbra PWAIT;\nPDONE:\n}\n" ::"r"(
                smem_u32p(s_bar))
            : "memory");
    }

    // ---- SH bands ----
    if (tid < cnt && s_pose->rotate_sh) {
        const PoseDev& p = *s_pose;
        float* c = s_sh + tid * 45;  // [15][3]
        float in[15][3];
#pragma unroll
        for (int j = 0; j < 15; ++j)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) in[j][ch] = c[3 * j + ch];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float acc = mul(p.D1[3 * i], in[0][ch]);
#pragma unroll
                for (int j = 1; j < 3; ++j) acc = fma(p.D1[3 * i + j], in[j][ch], acc);
                c[3 * i + ch] = acc;
            }
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float acc = mul(p.D2[5 * i], in[3][ch]);
#pragma unroll
                for (int j = 1; j < 5; ++j) acc = fma(p.D2[5 * i + j], in[3 + j][ch], acc);
                c[3 * (3 + i) + ch] = acc;
            }
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float acc = mul(p.D3[7 * i], in[8][ch]);
#pragma unroll
                for (int j = 1; j < 7; ++j) acc = fma(p.D3[7 * i + j], in[8 + j][ch], acc);
                c[3 * (8 + i) + ch] = acc;
            }
    }
    __syncthreads();
    // coalesced store into rows of 48 floats, skipping the 3 DC floats at the head of each row
    float* dst = a.shs + dst0 * 48;
    for (int i = tid; i < cnt * 45; i += 256) {
        int r = i / 45, e = i - r * 45;
        dst[(size_t)r * 48 + 3 + e] = s_sh[i];
    }
}

int launch_pose(int K, const int32_t* first, const PoseDev* poses_dev, const pg_canonical* canon,
                int scene_offset, const pg_scene* scene, cudaStream_t stream) {
    if (K == 0) return PG_OK;
    PoseArgs a;
    a.K = K;
    int n_max = 0;
    for (int k = 0; k <= PG_MAX_OBJECTS; ++k) a.first[k] = k <= K ? first[k] : 0;
    for (int k = 0; k < K; ++k) n_max = max(n_max, first[k + 1] - first[k]);
    if (n_max == 0) return PG_OK;
    a.poses = poses_dev;
    a.xyz = canon->xyz; a.rot = canon->rotation; a.rest = canon->features_rest;
    a.scene_offset = scene_offset;
    a.means = scene->means3D; a.rots = scene->rotations; a.shs = scene->shs;
    const int smem = 256 * 45 * 4 + 512 + 16;
    PG_CUDA_CHECK(ensure_dynamic_smem(pose_kernel, smem));
    dim3 grid((n_max + 255) / 256, K);
    pose_kernel<<<grid, 256, smem, stream>>>(a);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

// ---- output packing (pegasus.py:340-358 host conversions, done before the D2H copy) ------------
__global__ void pack_kernel(int W, int H, const float* __restrict__ color, const float* __restrict__ depth,
                            uint8_t* __restrict__ rgb_u8, uint16_t* __restrict__ depth_u16) {
    const size_t HW = (size_t)W * H;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    if (rgb_u8) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = mul(color[c * HW + i], 255.0f);
            // numpy's float -> uint8 cast on the host truncates; for values in [0, 256) this is it.  DELIBERATE
            // DEVIATION outside that range: the reference's (rgb * 255).astype("uint8") wraps modulo 256 (a pixel
            // brighter than 1.0 — SH colours are clamped only below — comes out dark), depth beyond 65.535 m wraps
            // modulo 65536; here both saturate.  Pinned by tests/test_gpu_generate.py.
            v = fminf(fmaxf(v, 0.0f), 255.0f);
            rgb_u8[3 * i + c] = (uint8_t)(int)v;
        }
    }
    if (depth_u16) {
        float v = mul(depth[i], 1000.0f);
        v = fminf(fmaxf(v, 0.0f), 65535.0f);
        depth_u16[i] = (uint16_t)(int)v;
    }
}

// one thread per output byte: 8 mask pixels of one row -> 1 byte, least-significant bit first
__global__ void pack_masks_kernel(int W, int H, int n_planes, const uint8_t* __restrict__ masks, uint8_t* __restrict__ bits) {
    const int Wb = (W + 7) / 8;
    const size_t total = (size_t)n_planes * H * Wb;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int xb = (int)(i % Wb);
    const size_t row = i / Wb;  // plane * H + y
    const uint8_t* src = masks + row * W + 8 * (size_t)xb;
    uint32_t b = 0;
    if (8 * xb + 8 <= W && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
        const uint2 v = *reinterpret_cast<const uint2*>(src);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            b |= (((v.x >> (8 * k)) & 255u) ? 1u : 0u) << k;
            b |= (((v.y >> (8 * k)) & 255u) ? 1u : 0u) << (4 + k);
        }
    } else {
        for (int k = 0; k < 8 && 8 * xb + k < W; ++k) b |= (src[k] ? 1u : 0u) << k;
    }
    bits[i] = (uint8_t)b;
}

int launch_pack_masks(int W, int H, int n_planes, const uint8_t* masks, uint8_t* bits, cudaStream_t stream) {
    const size_t total = (size_t)n_planes * H * ((W + 7) / 8);
    if (total == 0) return PG_OK;
    pack_masks_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(W, H, n_planes, masks, bits);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

int launch_pack(int W, int H, const float* color, const float* depth, uint8_t* rgb_u8, uint16_t* depth_u16,
                cudaStream_t stream) {
    size_t HW = (size_t)W * H;
    if (HW == 0) return PG_OK;
    pack_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(W, H, color, depth, rgb_u8, depth_u16);
    count_launch(1);
    PG_CUDA_CHECK(cudaGetLastError());
    return PG_OK;
}

}  // namespace pg
