// Issue cost (cycles per warp instruction per SM sub-partition) of the instruction kinds the compositing hit loop
// is made of, alone and mixed, on B200.  Each thread runs 8 independent chains so that latency is hidden with
// 8 warps per SMSP; the loop overhead is < 2 %.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

#define ITERS 1024
#define NCH 8

enum { FFMA, FFMA2, FMUL2, FADD2, FSETP_SEL, FMNMX, LOP3, LEA, IMAD, MUFU, FFMA2_FSEL, FFMA2_FMNMX, FFMA2_2ALU, FFMA2_MUFU,
       FFMA_FSEL, FFMA2_BCAST, FFMA2_LDS, NMODES };
static const char* kNames[NMODES] = {"FFMA", "FFMA2", "FMUL2", "FADD2", "FSETP+FSEL (2 instr)", "FMNMX", "LOP3", "LEA", "IMAD", "MUFU.EX2",
                                     "FFMA2 + FSETP+FSEL", "FFMA2 + FMNMX", "FFMA2 + 2 FMNMX", "FFMA2 + MUFU.EX2", "FFMA + FSETP+FSEL",
                                     "FFMA2 with scalar-broadcast operand", "FFMA2 + LDS.128 (broadcast)"};
static const int kInstr[NMODES] = {1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 3, 2, 3, 2, 3, 1, 2};

template <int MODE>
__global__ void k(float* out, float a0, float b0, int i0, long long* cyc) {
    __shared__ float4 sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = make_float4(a0, b0, a0, b0);
    __syncthreads();
    float x[NCH], y[NCH];
    u64 X[NCH];
    uint32_t u[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = 0.5f * i; X[i] = pk(x[i], x[i] + 0.5f); u[i] = threadIdx.x + i + i0; }
    float a = a0, b = b0;
    u64 A = pk(a, a + 1e-3f), B = pk(b, b + 1e-3f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (MODE == FFMA2 || MODE == FFMA2_FSEL || MODE == FFMA2_FMNMX || MODE == FFMA2_2ALU || MODE == FFMA2_MUFU || MODE == FFMA2_LDS)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(A), "l"(B));
            if (MODE == FFMA2_BCAST) asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nfma.rn.f32x2 %0, %0, t, %2;\n}" : "+l"(X[i]) : "f"(a), "l"(B));
            if (MODE == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(A));
            if (MODE == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(B));
            if (MODE == FFMA_FSEL) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (MODE == FSETP_SEL || MODE == FFMA2_FSEL || MODE == FFMA_FSEL)
                asm volatile("{\n.reg .pred p;\nsetp.lt.f32 p, %0, %1;\nselp.f32 %0, %0, %2, p;\n}" : "+f"(y[i]) : "f"(a), "f"(b));
            if (MODE == FMNMX || MODE == FFMA2_FMNMX || MODE == FFMA2_2ALU) asm volatile("min.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(a));
            if (MODE == FFMA2_2ALU) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (MODE == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(it), "r"(i0));
            if (MODE == LEA) asm volatile("{\n.reg .b32 t;\nshl.b32 t, %0, 23;\nadd.s32 %0, t, %1;\n}" : "+r"(u[i]) : "r"(it));
            if (MODE == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(i0), "r"(it));
            if (MODE == MUFU || MODE == FFMA2_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y[i]));
            if (MODE == FFMA2_LDS) { float4 v = sh[(it + i) & 63]; asm volatile("" ::"f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)); }
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc += x[i] + y[i] + __uint_as_float((uint32_t)X[i]) + __uint_as_float((uint32_t)(X[i] >> 32)) + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(int threads) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<MODE><<<148, threads>>>(out, 0.999f, 0.001f, 3, cyc);
    k<MODE><<<148, threads>>>(out, 0.999f, 0.001f, 3, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double steps = (double)ITERS * NCH, warps_per_smsp = threads / 32.0 / 4.0;
    const double cyc_per_step = (double)h / (steps * warps_per_smsp);  // cycles one SMSP spends per warp-level step
    printf("%-40s threads=%4d  cycles/step/SMSP=%6.3f  (%d instr/step -> %.3f cycles/instr)\n", kNames[MODE], threads, cyc_per_step,
           kInstr[MODE], cyc_per_step / kInstr[MODE]);
    cudaFree(out); cudaFree(cyc);
}

template <int M>
void run_all(int th) {
    run<M>(th);
    if constexpr (M + 1 < NMODES) run_all<M + 1>(th);
}

int main() {
    for (int th : {512, 1024}) run_all<0>(th);
    return 0;
}
