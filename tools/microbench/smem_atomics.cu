// Throughput of shared-memory reductions as the binning / radix kernels use them (B200): one atomicAdd per lane to
// consecutive / random / identical counters, 8 warps per SMSP, result never read back inside the loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t h[8][256];  // one row per warp pair: spreads the traffic like per-warp histograms do
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* row = h[warp & 7];
    uint32_t x = seed + threadIdx.x * 2654435761u;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        uint32_t a;
        if (MODE == 0) a = (it + lane) & 127;                 // consecutive counters (a run of tiles)
        if (MODE == 1) { x = x * 1664525u + 1013904223u; a = x >> 25; }  // random 7-bit digit
        if (MODE == 2) a = it & 127;                          // all lanes the same counter
        if (MODE == 3) a = ((it + lane) * 2) & 127;           // 2-way bank spread
        if (MODE <= 3) atomicAdd(&row[a], 1u);
        if (MODE == 4) { a = (it + lane) & 127; asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&row[a])) : "memory"); }
        if (MODE == 5) { a = (it + lane) & 127; row[a] += 1; }  // plain LDS + STS (racy: cost reference only)
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = row[lane] + x;
}
template <int MODE>
void run(const char* name, int threads) {
    uint32_t* out; long long* cyc; long long hcy;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<MODE><<<148, threads>>>(out, cyc, 1);
    k<MODE><<<148, threads>>>(out, cyc, 2);
    cudaMemcpy(&hcy, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps = threads / 32.0;
    printf("%-44s threads=%4d  cycles per warp-wide atomic per SM = %.2f\n", name, threads, (double)hcy / (ITERS * warps));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {256, 1024}) {
        run<0>("atomicAdd, 32 consecutive counters", th);
        run<1>("atomicAdd, random 7-bit digit", th);
        run<2>("atomicAdd, one counter (REDUX-aggregated?)", th);
        run<3>("atomicAdd, stride-2 counters", th);
        run<4>("red.shared.add, 32 consecutive counters", th);
        run<5>("plain LDS+STS increment (racy reference)", th);
    }
    return 0;
}
