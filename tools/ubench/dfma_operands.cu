// Micro-benchmark: does the FP64 pipe of sm_100a sustain one DFMA per 2 cycles per sub-partition whatever the operand
// forms, or do three distinct register-pair operands cost register-file bandwidth?  Decides whether the 21 FP64
// instructions per path-step of the C2 loop (most of them with three register pairs) can run at the pipe's nominal rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu && ./dfma_operands
// 4 warps per sub-partition, 8 independent chains per thread; per iteration 8 DFMA of one operand form:
//   RRR  d = fma(a, b, c)  three distinct register pairs (a = chain value, b and c per-chain invariants)
//   RRd  d = fma(a, b, d)  accumulate: two distinct pairs + the destination
//   RdR  d = fma(d, b, c)
//   Rdd  d = fma(d, b, d)  two distinct pairs
//   ddd  d = fma(d, d, d)  one pair
//   RCR  d = fma(d, c[..], c)  constant-bank operand + two pairs
//   RIR  d = fma(d, 1.5, c)    immediate operand + two pairs
//   RCd  d = fma(d, c[..], d)  constant-bank operand + one pair
//   RSd  d = fma(a, B, d)      three pairs, B the SAME register in all 8 consecutive instructions (operand reuse cache?)
//   RSd+ the same with a 3-register LOP3 between the DFMAs (does another pipe's operand traffic disturb it?)
//   RRd+ three distinct pairs + a LOP3 each (is register bandwidth shared between the FP64 and the ALU pipe?)
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double kc[8] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996};

template <int FORM>
__global__ void k(double* out, int iters, double s, long long* cyc) {
    double d[8], b[8], c[8], a[8];
    unsigned y[8], z[8], m = (unsigned)iters * 2654435761u;
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[i] = threadIdx.x * 7 + i; z[i] = threadIdx.x * 13 + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = 1.0 + 1e-9 * (threadIdx.x + i); b[i] = 1.0 + s * (i + 1); c[i] = s * (i + 2); a[i] = 1.0 - s * (i + 3); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (FORM == 0) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d[i]) : "d"(a[i]), "d"(b[i]), "d"(c[i]));
            if (FORM == 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[i]));
            if (FORM == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(b[i]), "d"(c[i]));
            if (FORM == 3) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(b[i]));
            if (FORM == 4) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(d[i]));
            if (FORM == 5) d[i] = fma(d[i], kc[i], c[i]);
            if (FORM == 6) d[i] = fma(d[i], 1.5, c[i]);
            if (FORM == 7) d[i] = fma(d[i], kc[i], d[i]);
            if (FORM == 8 || FORM == 9) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[0]));
            if (FORM == 10) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[i]));
            if (FORM == 9 || FORM == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m), "r"(z[i]));
        }
    }
    long long t1 = clock64();
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += d[i] + a[i] + b[i] + c[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 24);
    cudaMallocManaged(&cyc, 8);
    const int iters = 4000, warps = 4, threads = 128 * warps;
    const char* names[] = {"RRR d=fma(a,b,c)", "RRd d=fma(a,b,d)", "RdR d=fma(d,b,c)", "Rdd d=fma(d,b,d)", "ddd d=fma(d,d,d)", "RCR d=fma(d,c[],c)", "RIR d=fma(d,1.5,c)", "RCd d=fma(d,c[],d)",
                           "RSd d=fma(a,B,d) B shared", "RSd + LOP3", "RRd + LOP3"};
#define RUN(F)                                                                                          \
    k<F><<<148, threads>>>(out, iters, 1e-12, cyc);                                                    \
    cudaDeviceSynchronize();                                                                            \
    printf("%-22s %6.2f cycles per DFMA per sub-partition (4 warps, 8 chains each)\n", names[F], (double)*cyc / iters / warps / 8);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    return 0;
}
