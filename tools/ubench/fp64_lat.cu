// Micro-benchmarks behind the scheduling model in DESIGN.md: dependent-issue latency and throughput of the
// FP64 pipe, MUFU.RSQ64H / RCP64H, FLO and LDS on one SM sub-partition (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// DFMA chains interleaved with independent integer work: does the scheduler fill the FP64 pipe's idle issue slot?
template <int ILP, int NINT>
__global__ void k_mix(double* out, int iters, double a, double b, unsigned m, long long* cyc) {
    double x[ILP];
    unsigned y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a + i + threadIdx.x; y[i] = threadIdx.x * 7 + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                x[i] = fma(x[i], a, b);
#pragma unroll
                for (int q = 0; q < NINT; ++q) y[i] = (y[i] ^ m) + (y[i] >> 3);
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_rsq(double* out, int iters, double a, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x[i]));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_lds(double* out, int iters, long long* cyc) {
    __shared__ unsigned tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (threadIdx.x + i * 32) & 1023;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = tab[x[i]];
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 24);
    cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    auto report = [&](const char* name, int ilp, int warps_per_smsp, double ops_per_iter) {
        cudaDeviceSynchronize();
        printf("%-28s ILP %d  warps/SMSP %d : %.2f cycles per op per warp, %.3f warp-ops/cycle/SMSP\n", name, ilp, warps_per_smsp,
               (double)*cyc / (iters * ops_per_iter / ilp) / ilp * 1.0, warps_per_smsp * iters * ops_per_iter / (double)*cyc);
    };
    for (int w : {1, 2, 4, 8}) {
        int threads = 128 * w;   // w warps on each of the 4 sub-partitions
        k_dfma<1><<<148, threads>>>(out, iters, 1.0000001, 1e-9, cyc); report("DFMA dependent", 1, w, 8);
        k_dfma<2><<<148, threads>>>(out, iters, 1.0000001, 1e-9, cyc); report("DFMA", 2, w, 16);
        k_dfma<4><<<148, threads>>>(out, iters, 1.0000001, 1e-9, cyc); report("DFMA", 4, w, 32);
        k_dfma<8><<<148, threads>>>(out, iters, 1.0000001, 1e-9, cyc); report("DFMA", 8, w, 64);
    }
    for (int w : {1, 2, 4}) {
        int threads = 128 * w;
        k_mix<4, 1><<<148, threads>>>(out, iters, 1.0000001, 1e-9, 0x5bd1e995u, cyc); report("DFMA + 3 INT each", 4, w, 32);
        k_mix<4, 2><<<148, threads>>>(out, iters, 1.0000001, 1e-9, 0x5bd1e995u, cyc); report("DFMA + 6 INT each", 4, w, 32);
        k_mix<8, 1><<<148, threads>>>(out, iters, 1.0000001, 1e-9, 0x5bd1e995u, cyc); report("DFMA + 3 INT each", 8, w, 64);
    }
    for (int w : {1, 4}) {
        int threads = 128 * w;
        k_rsq<1><<<148, threads>>>(out, iters, 1.5, cyc); report("MUFU.RSQ64H dependent", 1, w, 8);
        k_rsq<4><<<148, threads>>>(out, iters, 1.5, cyc); report("MUFU.RSQ64H", 4, w, 32);
        k_lds<1><<<148, threads>>>(out, iters, cyc); report("LDS dependent", 1, w, 8);
        k_lds<4><<<148, threads>>>(out, iters, cyc); report("LDS", 4, w, 32);
    }
    return 0;
}
