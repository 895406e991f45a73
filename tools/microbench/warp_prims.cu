// Latency / throughput of the warp primitives a radix-sort ranking step can be built from (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_prims warp_prims.cu && ./warp_prims
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512

template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed, long long* cyc) {
    __shared__ uint32_t sh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = (MODE & 1) ? (lane * 7u + seed) & 255u : (seed & 255u);  // odd modes: 32 distinct digits
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
        if (MODE / 2 == 0) {          // match.any, dependent chain
            uint32_t m = __match_any_sync(0xffffffffu, v);
            acc += m; v = (v + (m & 1u)) & 255u;
        } else if (MODE / 2 == 1) {   // 8 explicit votes, dependent chain
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                uint32_t bal, bit = (v >> b) & 1u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\nvote.sync.ballot.b32 %0, p, 0xffffffff;\n}\n" : "=r"(bal) : "r"(bit));
                peers &= bal ^ (bit - 1u);
            }
            acc += peers; v = (v + (peers & 1u)) & 255u;
        } else if (MODE / 2 == 2) {   // shared atomicAdd with return (warp-private row), dependent chain
            uint32_t r = atomicAdd(&sh[warp][v], 1u);
            acc += r; v = (v + (r & 1u)) & 255u;
        } else if (MODE / 2 == 3) {   // LDS -> STS chain
            uint32_t r = sh[warp][v];
            __syncwarp();
            sh[warp][v] = r + 1;
            __syncwarp();
            acc += r; v = (v + (r & 1u)) & 255u;
        } else if (MODE / 2 == 4) {   // shfl chain
            uint32_t r = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
            acc += r; v = (r + 1) & 255u;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
    uint32_t* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    k<MODE><<<1, threads>>>(out, 3, cyc);
    k<MODE><<<1, threads>>>(out, 3, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s threads=%4d  %.1f cycles/iter (warp 0)\n", name, threads, (double)h / ITERS);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {32, 256, 1024}) {
        run<0>("match.any uniform", th);
        run<1>("match.any 32 distinct", th);
        run<2>("8x vote uniform", th);
        run<3>("8x vote 32 distinct", th);
        run<4>("smem atomicAdd(ret) same addr", th);
        run<5>("smem atomicAdd(ret) 32 distinct", th);
        run<6>("LDS->STS same addr", th);
        run<7>("LDS->STS 32 distinct", th);
        run<8>("shfl chain", th);
    }
    return 0;
}
