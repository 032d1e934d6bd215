// Micro-benchmark: do half-rate instructions of different pipes (FP64, ALU) overlap, or does each hold the dispatch
// port for two cycles?  Decides whether the C2 step loop (21 FP64 + ~10 half-rate integer + ~10 other instructions per
// path-step) is bound by the FP64 pipe (42 cycles) or by dispatch (the sum).  B200, sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NL, int NS, int NM>
__global__ void k_mix(double* out, int iters, double a, double b, unsigned m, long long* cyc) {
    double x[8];
    unsigned y[8], z[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a + i + threadIdx.x; y[i] = threadIdx.x * 7 + i; z[i] = threadIdx.x * 13 + i; w[i] = threadIdx.x + 3 * i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NF) x[i] = fma(x[i], a, b);
            if (i < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m), "r"(z[i]));
            if (i < NS) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(w[i]), "r"(m));
            if (i < NM) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(m), "r"(m));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 24);
    cudaMallocManaged(&cyc, 8);
    const int iters = 4000;
    const int warps = 4, threads = 128 * warps;              // 4 warps on each of the 4 sub-partitions
#define RUN(NF, NL, NS, NM)                                                                                             \
    k_mix<NF, NL, NS, NM><<<148, threads>>>(out, iters, 1.0000001, 1e-9, 0x5bd1e995u, cyc);                            \
    cudaDeviceSynchronize();                                                                                            \
    printf("per iteration and warp: %d DFMA + %d LOP3 + %d SHF + %d IMAD : %6.2f cycles of the sub-partition (%d instructions)\n", NF, NL, NS, NM, \
           (double)*cyc / iters / warps, NF + NL + NS + NM);
    RUN(8, 0, 0, 0) RUN(0, 8, 0, 0) RUN(0, 0, 8, 0) RUN(0, 0, 0, 8)
    RUN(8, 4, 0, 0) RUN(8, 8, 0, 0) RUN(8, 0, 8, 0) RUN(8, 4, 4, 0) RUN(8, 0, 0, 8) RUN(8, 4, 0, 4) RUN(8, 8, 8, 0) RUN(8, 8, 8, 8)
    RUN(0, 8, 8, 0) RUN(0, 8, 0, 8)
    return 0;
}
