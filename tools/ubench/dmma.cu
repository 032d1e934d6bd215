// Micro-benchmark: FP64 tensor-core (DMMA, mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64) throughput against the DFMA
// peak on B200 (sm_100a).  Decides whether the 64 x 64 correlation product of wide models (C4) belongs on DMMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
    double c[ILP][4], av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(c[i], av, bv);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
    double c[ILP][4], av[8], bv[4] = {b, b + 1, b + 2, b + 3};
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(c[i], av, bv);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA and DFMA interleaved: do the two share the FP64 pipe, or do they add up?
template <int ILP>
__global__ void k_mix(double* out, int iters, double a, double b) {
    double c[ILP][2], x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = c[i][1] = threadIdx.x + i; x[i] = a + i; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) { dmma884(c[i], a, b); x[i] = fma(x[i], a, b); }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 512, blocks = sms * 2, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    const double warps = (double)blocks * threads / 32;
    printf("%s, %d SMs, %d blocks x %d threads, %d iterations\n", p.name, sms, blocks, threads, iters);
    {
        float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DFMA            ILP 8 : %8.3f ms  %7.2f T FMA/s\n", ms, warps * 32.0 * 8 * iters / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_dmma884<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m8n8k4     ILP 4 : %8.3f ms  %7.2f T FMA/s\n", ms, warps * 256.0 * 4 * iters / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m8n8k4     ILP 8 : %8.3f ms  %7.2f T FMA/s\n", ms, warps * 256.0 * 8 * iters / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_dmma1688<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m16n8k8    ILP 4 : %8.3f ms  %7.2f T FMA/s\n", ms, warps * 1024.0 * 4 * iters / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_dmma16816<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m16n8k16   ILP 4 : %8.3f ms  %7.2f T FMA/s\n", ms, warps * 2048.0 * 4 * iters / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_mix<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m8n8k4 + DFMA 1:1: %8.3f ms  %7.2f T FMA/s (DMMA part %7.2f)\n", ms, warps * (256.0 + 32.0) * 4 * iters / ms * 1e-9,
               warps * 256.0 * 4 * iters / ms * 1e-9);
    }
    cudaFree(out);
    return 0;
}
