// Throughput of packed FP32 (FFMA2, sm_100) against scalar FFMA on B200, and what it takes to feed a packed
// instruction with a broadcast scalar.  Decides whether a 2-pixels-per-thread compositing loop pays.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

#define ITERS 2048
// MODE 0: scalar FFMA, 8 independent chains / thread      (8 FMA-lane-ops per iteration-step)
// MODE 1: FFMA2, 8 independent packed chains / thread      (16 FMA-lane-ops)
// MODE 2: FFMA2 with a per-iteration broadcast scalar operand (pack of (s,s) rebuilt every step)
// MODE 3: scalar FFMA interleaved 1:1 with ALU ops (LOP3)  (dual-pipe issue)
// MODE 4: FFMA2 interleaved 1:1 with ALU ops
template <int MODE>
__global__ void k(float* out, float a0, float b0, long long* cyc) {
    float x[8];
    unsigned long long X[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; X[i] = pk(x[i], x[i] + 0.5f); u[i] = threadIdx.x + i; }
    float a = a0, b = b0;
    unsigned long long A = pk(a, a), B = pk(b, b);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = __fmaf_rn(x[i], a, b);
            if (MODE == 1) X[i] = fma2(X[i], A, B);
            if (MODE == 2) { float s = __int_as_float(__float_as_int(a) + (it & 1)); asm volatile("" : "+f"(s)); X[i] = fma2(X[i], pk(s, s), B); }
            if (MODE == 3) { x[i] = __fmaf_rn(x[i], a, b); u[i] = (u[i] ^ (uint32_t)it) & 0x7fffffffu; asm volatile("" : "+r"(u[i])); }
            if (MODE == 4) { X[i] = fma2(X[i], A, B); u[i] = (u[i] ^ (uint32_t)it) & 0x7fffffffu; asm volatile("" : "+r"(u[i])); }
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i] + __uint_as_float((uint32_t)X[i]) + __uint_as_float((uint32_t)(X[i] >> 32)) + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int lane_ops_per_step) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<MODE><<<148, threads>>>(out, 0.999f, 0.001f, cyc);
    k<MODE><<<148, threads>>>(out, 0.999f, 0.001f, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double steps = (double)ITERS * 8;                       // instruction groups per thread
    double warps = threads / 32.0;
    double per_smsp = steps * warps / 4.0 / (double)h;      // warp-level steps per cycle per SMSP
    printf("%-44s threads=%4d  %8lld cyc  steps/clk/SMSP=%.3f  FMA lane-ops/clk/SM=%.1f\n", name, threads, h, per_smsp,
           per_smsp * 4 * 32 * lane_ops_per_step);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {128, 512, 1024}) {
        run<0>("FFMA scalar", th, 1);
        run<1>("FFMA2 packed", th, 2);
        run<2>("FFMA2 packed + (s,s) broadcast each step", th, 2);
        run<3>("FFMA scalar + 1 ALU op", th, 1);
        run<4>("FFMA2 packed + 1 ALU op", th, 2);
    }
    return 0;
}
