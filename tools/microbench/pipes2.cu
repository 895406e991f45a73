// Operand-form costs of packed / scalar FP32 and of the predicate / select ops next to them (B200).
// Same harness as pipes.cu: 8 independent chains per thread, 8 warps per SMSP, cycles per warp instruction per SMSP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
#define ITERS 1024
#define NCH 8
enum { F2_PPP, F2_PPI, F2_PBP, F2_PBI, F2_BPP_ACC, M2_PP, M2_PB, A2_PB, F_RRR, F_RRI, F_RIR, FMUL_RR, FSETP1, FSEL1, FSETP_AND, PLOP, F2PPI_FSETP, F2PBP_FSEL,
       F2PPI_F_RRR, SEL_F2, LEA1, MOV64, NMODES };
static const char* kNames[NMODES] = {
    "FFMA2 d=d*p+p (3 packed regs)", "FFMA2 d=d*p+imm (Horner)", "FFMA2 d=d*bcast+p", "FFMA2 d=d*bcast+imm", "FFMA2 acc=bcast*p+acc",
    "FMUL2 d=d*p", "FMUL2 d=d*bcast", "FADD2 d=d+bcast", "FFMA d=d*r+r", "FFMA d=d*r+imm", "FFMA d=d*imm+r", "FMUL d=d*r",
    "FSETP + predicated IADD (2 instr)", "ISETP + FSEL (2 instr)", "FSETP.AND chain of 3 + FSEL (4 instr)", "2 ISETP + PLOP3 + pred IADD (4 instr)",
    "FFMA2 Horner + FSETP+FSEL (3 instr)", "FFMA2 d=d*bcast+p + FSETP+FSEL (3 instr)", "FFMA2 Horner + FFMA rrr (2 instr)",
    "ISETP + 2 FSEL + FFMA2 on the selected pair (4 instr)", "LEA (shl+add)", "MOV64 pack (2 MOV)"};
static const int kInstr[NMODES] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 4, 4, 3, 3, 2, 4, 1, 2};

template <int MODE>
__global__ void k(float* out, float a0, float b0, int i0, long long* cyc) {
    float x[NCH], y[NCH];
    u64 X[NCH];
    uint32_t u[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = 0.5f * i; X[i] = pk(x[i], x[i] + 0.5f); u[i] = threadIdx.x + i + i0; }
    float a = a0 + (float)(threadIdx.x >> 10), b = b0 + (float)(threadIdx.x >> 11);  // not provably uniform: stays in R registers
    u64 A = pk(a, a + 1e-3f), B = pk(b, b + 1e-3f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == F2_PPP) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(A), "l"(B));
            if (MODE == F2_PPI || MODE == F2PPI_FSETP || MODE == F2PPI_F_RRR)
                asm volatile("{\n.reg .b64 t;\nmov.b64 t, {0f3F000000, 0f3F000000};\nfma.rn.f32x2 %0, %0, %1, t;\n}" : "+l"(X[i]) : "l"(A));
            if (MODE == F2_PBP || MODE == F2PBP_FSEL) asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nfma.rn.f32x2 %0, %0, t, %2;\n}" : "+l"(X[i]) : "f"(a), "l"(B));
            if (MODE == F2_PBI) asm volatile("{\n.reg .b64 t, c;\nmov.b64 t, {%1, %1};\nmov.b64 c, {0f3F000000, 0f3F000000};\nfma.rn.f32x2 %0, %0, t, c;\n}" : "+l"(X[i]) : "f"(a));
            if (MODE == F2_BPP_ACC) asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nfma.rn.f32x2 %0, t, %2, %0;\n}" : "+l"(X[i]) : "f"(a), "l"(B));
            if (MODE == M2_PP) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(A));
            if (MODE == M2_PB) asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nmul.rn.f32x2 %0, %0, t;\n}" : "+l"(X[i]) : "f"(a));
            if (MODE == A2_PB) asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nadd.rn.f32x2 %0, %0, t;\n}" : "+l"(X[i]) : "f"(b));
            if (MODE == F_RRR || MODE == F2PPI_F_RRR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (MODE == F_RRI) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(x[i]) : "f"(a));
            if (MODE == F_RIR) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, %1;" : "+f"(x[i]) : "f"(b));
            if (MODE == FMUL_RR) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
            if (MODE == FSETP1) asm volatile("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n@p add.s32 %0, %0, 1;\n}" : "+r"(u[i]) : "f"(y[i]), "f"(a));
            if (MODE == FSEL1) asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f32 %0, %0, %2, p;\n}" : "+f"(y[i]) : "f"(a), "f"(b), "r"(u[i]));
            if (MODE == FSETP_AND) asm volatile("{\n.reg .pred p;\nsetp.leu.f32 p, %0, 0f00000000;\nsetp.geu.and.f32 p, %0, %1, p;\nsetp.geu.and.f32 p, %2, 0f3B808081, p;\nselp.f32 %0, %0, %2, p;\n}" : "+f"(y[i]) : "f"(a), "f"(b));
            if (MODE == PLOP) asm volatile("{\n.reg .pred p, q;\nsetp.ne.s32 p, %1, 0;\nsetp.ne.s32 q, %0, 0;\nand.pred p, p, q;\n@p add.s32 %0, %0, 1;\n}" : "+r"(u[i]) : "r"(i0));
            if (MODE == F2PPI_FSETP || MODE == F2PBP_FSEL) asm volatile("{\n.reg .pred p;\nsetp.lt.f32 p, %0, %1;\nselp.f32 %0, %0, %2, p;\n}" : "+f"(y[i]) : "f"(a), "f"(b));
            if (MODE == SEL_F2) {
                float s0, s1;
                asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %4, 0;\nselp.f32 %0, %2, 0f00000000, p;\nselp.f32 %1, %3, 0f00000000, p;\n}" : "=f"(s0), "=f"(s1) : "f"(x[i]), "f"(y[i]), "r"(u[i]));
                asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %2};\nfma.rn.f32x2 %0, t, %3, %0;\n}" : "+l"(X[i]) : "f"(s0), "f"(s1), "l"(B));
            }
            if (MODE == LEA1) asm volatile("{\n.reg .b32 t;\nshl.b32 t, %0, 23;\nadd.s32 %0, t, %1;\n}" : "+r"(u[i]) : "r"(it));
            if (MODE == MOV64) { asm volatile("mov.b64 %0, {%1, %2};" : "=l"(X[i]) : "f"(x[i]), "f"(y[i])); asm volatile("" : "+l"(X[i])); }
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
    const double steps = (double)ITERS * NCH, wps = threads / 32.0 / 4.0;
    const double c = (double)h / (steps * wps);
    printf("%-56s thr=%4d  cycles/step/SMSP=%6.3f  (%d instr nominal -> %.3f/instr)\n", kNames[MODE], threads, c, kInstr[MODE], c / kInstr[MODE]);
    cudaFree(out); cudaFree(cyc);
}
template <int M> void run_all(int th) { run<M>(th); if constexpr (M + 1 < NMODES) run_all<M + 1>(th); }
int main() { for (int th : {1024}) run_all<0>(th); return 0; }
