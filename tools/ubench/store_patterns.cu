// Store-pattern ceilings for full-path output [N][T] f64 with T = 253 (row = 2024 B, not a multiple of 32 B):
//   A  each lane owns one row and writes one aligned 32-byte sector per step group (st.global.v4.f64)  — rows 4 apart per warp
//   B  each lane stages a 128-byte line in shared memory and hands it to the TMA (cp.async.bulk shared -> global) — rows 16 apart
//   C  4 lanes write one 128-byte line of one row with st.global.v4.f64 (needs a transpose in a real kernel)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cstdlib>
#include <string>

constexpr int T = 253;

__global__ void __launch_bounds__(1024, 1) k_sector(double* out, long long n_rows, int iters_dummy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long n_items = n_rows / 32;
    for (long long item = (long long)blockIdx.x * nw + warp; item < n_items; item += (long long)gridDim.x * nw) {
        const long long s = ((item >> 2) << 7) + (item & 3) + 4 * lane;
        double* row = out + s * T;
        const int gamma = (int)((4 - ((s * T + 1) & 3)) & 3);
        double v = (double)lane;
        for (int t = gamma; t + 4 <= T - 1; t += 4) {
            double* dst = row + t + 1;
            v += 1.0;
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(v), "d"(v), "d"(v), "d"(v) : "memory");
        }
    }
}

// D: as A, but a lane collects NS sectors (NS*4 steps) and writes them back to back: complete 64- / 128-byte pieces reach L2 at once
template <int NS>
__global__ void __launch_bounds__(1024, 1) k_sector_burst(double* out, long long n_rows, int iters_dummy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long n_items = n_rows / 32;
    for (long long item = (long long)blockIdx.x * nw + warp; item < n_items; item += (long long)gridDim.x * nw) {
        const long long s = ((item >> 2) << 7) + (item & 3) + 4 * lane;
        double* row = out + s * T;
        const int gamma = (int)((4 - ((s * T + 1) & 3)) & 3);
        double v = (double)lane;
        for (int t = gamma; t + 4 * NS <= T - 1; t += 4 * NS) {
            double* dst = row + t + 1;
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                v += 1.0;
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "d"(v), "d"(v), "d"(v), "d"(v) : "memory");
            }
        }
    }
}

__global__ void __launch_bounds__(1024, 1) k_tma(double* out, long long n_rows, int dummy) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* stage = reinterpret_cast<double*>(smem) + (size_t)(warp * 32 + lane) * 2 * 18;   // 2 buffers x 144 B
    const unsigned s_addr = (unsigned)__cvta_generic_to_shared(stage);
    const long long n_items = n_rows / 32;
    int buf = 0;
    for (long long item = (long long)blockIdx.x * nw + warp; item < n_items; item += (long long)gridDim.x * nw) {
        const long long s = ((item >> 4) << 9) + (item & 15) + 16 * lane;
        const long long e0 = s * T;
        const int phi = (int)(e0 & 15);
        double* line = out + (e0 - phi);                       // 128-byte aligned
        double v = (double)lane;
        // full lines only (the partial first / last line is noise for this measurement)
        for (int L = 1; (L + 1) * 16 <= phi + T; ++L) {
            double* st = stage + buf * 18;
#pragma unroll
            for (int q = 0; q < 8; ++q) { v += 1.0; *reinterpret_cast<double2*>(st + 2 * q) = make_double2(v, v); }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(line + 16 * L), "r"(s_addr + buf * 144) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            buf ^= 1;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(1024, 1) k_line4(double* out, long long n_rows, int dummy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long n_items = n_rows / 32;
    for (long long item = (long long)blockIdx.x * nw + warp; item < n_items; item += (long long)gridDim.x * nw) {
        // 32 rows 16 apart; instruction k writes one line of rows 8k .. 8k+7 (4 lanes per line)
        const long long s0 = ((item >> 4) << 9) + (item & 15);
        double v = (double)lane;
        for (int L = 1; L < 15; ++L) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long s = s0 + 16 * (8 * k + (lane >> 2));
                const long long e0 = s * T;
                double* line = out + (e0 - (e0 & 15)) + 16 * L + 4 * (lane & 3);
                v += 1.0;
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(line), "d"(v), "d"(v), "d"(v), "d"(v) : "memory");
            }
        }
    }
}

// usage: store_patterns                      the bandwidth table (all patterns x threads per SM)
//        store_patterns <pattern> <threads> <seconds>   one pattern in a loop for power sampling (nvidia-smi alongside):
//        patterns: A D2 D4 D8 B C S   (S = sequential: consecutive warps write consecutive 1 KB pieces)
__global__ void __launch_bounds__(1024, 1) k_seq(double* out, long long n_rows, int dummy) {
    const long long total4 = n_rows * T / 4;                              // 32-byte pieces
    double v = (double)threadIdx.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        v += 1.0;
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out + 4 * i), "d"(v), "d"(v), "d"(v), "d"(v) : "memory");
    }
}

int main(int argc, char** argv) {
    const long long n_rows = 1ll << 24;
    double* out;
    if (cudaMalloc(&out, (size_t)n_rows * T * 8 + 4096) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (argc >= 4) {
        const std::string pat = argv[1];
        const int threads = atoi(argv[2]);
        const double seconds = atof(argv[3]);
        double bytes = 0;
        auto launch = [&] {
            if (pat == "A") { k_sector<<<148, threads>>>(out, n_rows, 0); bytes = (double)n_rows * 62 * 32; }
            else if (pat == "D2") { k_sector_burst<2><<<148, threads>>>(out, n_rows, 0); bytes = (double)n_rows * 31 * 64; }
            else if (pat == "D4") { k_sector_burst<4><<<148, threads>>>(out, n_rows, 0); bytes = (double)n_rows * 15 * 128; }
            else if (pat == "D8") { k_sector_burst<8><<<148, threads>>>(out, n_rows, 0); bytes = (double)n_rows * 7 * 256; }
            else if (pat == "C") { k_line4<<<148, threads>>>(out, n_rows, 0); bytes = (double)n_rows * 14 * 128; }
            else if (pat == "S") { k_seq<<<148, threads>>>(out, n_rows, 0); bytes = (double)(n_rows * T / 4) * 32; }
            else if (pat == "B") {
                cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, threads * 288 + 128);
                k_tma<<<148, threads, threads * 288 + 128>>>(out, n_rows, 0); bytes = (double)n_rows * 14 * 128;
            }
        };
        launch(); launch();
        cudaDeviceSynchronize();
        float ms1;
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        const int n = (int)(seconds * 1e3 / ms1) + 1;
        cudaEventRecord(e0);
        for (int i = 0; i < n; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s threads %d: %d launches, %.3f ms each, %.0f GB/s (%s)\n", pat.c_str(), threads, n, ms / n, bytes / (ms / n) * 1e-6,
               cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    auto run = [&](const char* name, auto launch, double bytes) {
        launch(); launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %.3f ms  %.0f GB/s  (%s)\n", name, ms / 5, bytes / (ms / 5) * 1e-6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int threads : {256, 512, 768, 1024}) {
        printf("threads/SM %d\n", threads);
        run("A sector per lane (rows 4 apart)", [&] { k_sector<<<148, threads>>>(out, n_rows, 0); }, (double)n_rows * 62 * 32);
        run("D2 two sectors back to back", [&] { k_sector_burst<2><<<148, threads>>>(out, n_rows, 0); }, (double)n_rows * 31 * 64);
        run("D4 four sectors back to back", [&] { k_sector_burst<4><<<148, threads>>>(out, n_rows, 0); }, (double)n_rows * 15 * 128);
        run("D8 eight sectors back to back", [&] { k_sector_burst<8><<<148, threads>>>(out, n_rows, 0); }, (double)n_rows * 7 * 256);
        cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, threads * 288 + 128);
        run("B TMA bulk 128 B per lane (rows 16 apart)", [&] { k_tma<<<148, threads, threads * 288 + 128>>>(out, n_rows, 0); }, (double)n_rows * 14 * 128);
        run("C 4 lanes per 128 B line", [&] { k_line4<<<148, threads>>>(out, n_rows, 0); }, (double)n_rows * 14 * 128);
    }
    return 0;
}
