// engine.cpp — see engine.h.
#include "engine.h"

#include <chrono>
#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

extern "C" {
extern const char sde_blob_util_cubin_begin[], sde_blob_util_cubin_end[];
extern const char sde_blob_joe_kuo_begin[], sde_blob_joe_kuo_end[];
}

namespace sde {

// --------------------------------------------------------------------------------------
// Joe–Kuo table and host-side Sobol set-up
// --------------------------------------------------------------------------------------
namespace {
struct JoeKuo {
    uint32_t ndims = 0, stride = 0;
    const uint32_t* poly = nullptr;
    const uint32_t* minit = nullptr;
};
const JoeKuo& joe_kuo() {
    static JoeKuo jk = [] {
        JoeKuo t;
        const char* b = sde_blob_joe_kuo_begin;
        size_t len = (size_t)(sde_blob_joe_kuo_end - sde_blob_joe_kuo_begin);
        if (len < 16 || std::memcmp(b, "SDEJK601", 8) != 0) throw CudaError{"embedded Joe-Kuo table is corrupt"};
        std::memcpy(&t.ndims, b + 8, 4);
        std::memcpy(&t.stride, b + 12, 4);
        if (len < 16 + (size_t)t.ndims * 4 * (1 + t.stride)) throw CudaError{"embedded Joe-Kuo table is truncated"};
        t.poly = reinterpret_cast<const uint32_t*>(b + 16);
        t.minit = t.poly + t.ndims;
        return t;
    }();
    return jk;
}
}  // namespace

void joe_kuo_params(uint32_t dims, uint32_t* poly, uint32_t* minit) {
    const JoeKuo& jk = joe_kuo();
    if (dims > jk.ndims) throw ExprError{"Sobol dimension " + std::to_string(dims) + " exceeds the Joe-Kuo table (21201)"};
    std::memcpy(poly, jk.poly, (size_t)dims * 4);
    for (uint32_t d = 0; d < dims; ++d)
        for (uint32_t i = 0; i < 18; ++i) minit[(size_t)d * 18 + i] = i < jk.stride ? jk.minit[(size_t)d * jk.stride + i] : 0;
}

void sobol_tables(uint32_t dims, std::vector<uint32_t>& V, std::vector<uint32_t>& lane, std::vector<uint32_t>* nib, uint32_t lane_stride) {
    const JoeKuo& jk = joe_kuo();
    if (dims > jk.ndims) throw ExprError{"Sobol dimension " + std::to_string(dims) + " exceeds the Joe-Kuo table (21201)"};
    V.assign((size_t)dims * 32, 0);
    lane.assign((size_t)dims * 32, 0);
    if (nib) nib->assign((size_t)dims * 128, 0);
    for (uint32_t d = 0; d < dims; ++d) {
        uint32_t m[32];
        if (d == 0) {
            for (int i = 0; i < 32; ++i) m[i] = 1;          // van der Corput
        } else {
            const uint32_t p = jk.poly[d];
            int s = 31 - __builtin_clz(p);                   // degree of the primitive polynomial
            for (int i = 0; i < s && i < 32; ++i) m[i] = jk.minit[(size_t)d * jk.stride + i];
            for (int i = s; i < 32; ++i) {                   // m_i = 2 a_1 m_{i-1} ^ ... ^ 2^s m_{i-s} ^ m_{i-s}
                uint32_t v = m[i - s] ^ (m[i - s] << s);
                for (int k = 1; k < s; ++k) if ((p >> (s - k)) & 1u) v ^= m[i - k] << k;
                m[i] = v;
            }
        }
        uint32_t* Vd = &V[(size_t)d * 32];
        for (int i = 0; i < 32; ++i) Vd[i] = m[i] << (31 - i);   // top 32 bits of m_i << (63 - i)
        uint32_t* Ld = &lane[(size_t)d * 32];
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t idx = l * lane_stride;
            uint32_t g = idx ^ (idx >> 1), x = 0;
            for (int b = 0; b < 32; ++b) if ((g >> b) & 1u) x ^= Vd[b];
            Ld[l] = x;
        }
        if (nib) {
            uint32_t* Nd = &(*nib)[(size_t)d * 128];
            for (int i = 0; i < 8; ++i)
                for (uint32_t v = 0; v < 16; ++v) {
                    uint32_t x = 0;
                    for (int b = 0; b < 4; ++b) if ((v >> b) & 1u) x ^= Vd[4 * i + b];
                    Nd[i * 16 + v] = x;
                }
        }
    }
}

namespace {
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
void chacha8_block_host(const uint32_t key[8], uint64_t ctr, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                       key[4], key[5], key[6], key[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t x[16];
    std::memcpy(x, in, sizeof x);
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int r = 0; r < 4; ++r) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}
}  // namespace

void chacha8_u64_host(uint64_t seed, size_t n, uint64_t* out) {
    uint32_t key[8];
    uint64_t st = seed;
    for (int i = 0; i < 8; ++i) {                            // rand_core seed_from_u64 (PCG32)
        st = st * 6364136223846793005ull + 11634580027462260723ull;
        uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27), rot = (uint32_t)(st >> 59);
        key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    uint32_t buf[16];
    for (size_t i = 0; i < n; ++i) {
        if ((i & 7) == 0) chacha8_block_host(key, i >> 3, buf);
        out[i] = (uint64_t)buf[2 * (i & 7)] | ((uint64_t)buf[2 * (i & 7) + 1] << 32);
    }
}

// --------------------------------------------------------------------------------------
// Plan
// --------------------------------------------------------------------------------------
namespace {
struct SdeParamsHost {   // must mirror SdeParams in csrc/kernels/sde_sim_kernel.cuh
    uint64_t n_paths, scen_offset, n_base, seed;
    int32_t n_steps, reserved;
    CUdeviceptr times, dts, sqrt_dts, x0, sobol_nib, sobol_lane, xor_masks, inject, out, partials;
};
bool uses_sobol(int rng) { return rng == RNG_SOBOL_CP || rng == RNG_SOBOL_XOR || rng == RNG_SOBOL_RAW; }
}  // namespace

namespace {
// SDE_B200_TRACE=1: phase times of plan creation on stderr (lowering, cubin, module load, tables)
struct PhaseTrace {
    bool on = std::getenv("SDE_B200_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[sde_b200] plan: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

Plan::Plan(const Universe& u, const PlanOptions& opt) : u_(u), opt_(opt) {
    PhaseTrace trace;
    low_ = lower_model(u_, opt_.lower);
    trace.mark("lowering");
    const int S = u_.T() - 1, K = u_.K();
    const size_t dims = (size_t)S * K;
    if (uses_sobol(opt_.lower.rng) && dims > kMaxSobolDims)
        throw ExprError{"sobol needs (T-1)*K = " + std::to_string(dims) + " dimensions; the Joe-Kuo table has 21201 (src/rng/sobol.rs:16)"};
    use_device(opt_.device);
    const DriverApi& d = driver();
    trace.mark("device context");
    std::string log;
    std::vector<char> cubin = nvrtc_compile(low_.source, "sde_plan.cu", &log);
    prelowered_ = log.rfind("(cached on disk", 0) == 0;       // the cubin shipped in the ahead-of-time cache (build/jit_cache)
    trace.mark(prelowered_ ? "cubin (disk cache)" : "cubin (NVRTC)");
    cu_check(d.cuModuleLoadData(&mod_, cubin.data()), "cuModuleLoadData(plan)");
    trace.mark("cuModuleLoadData");
    cu_check(d.cuModuleGetFunction(&fn_sim_, mod_, "sde_sim_kernel"), "cuModuleGetFunction(sde_sim_kernel)");
    cu_check(d.cuModuleGetFunction(&fn_fin_, mod_, "sde_moments_finalize"), "cuModuleGetFunction(sde_moments_finalize)");
    if (low_.smem_bytes > 48 * 1024)
        cu_check(d.cuFuncSetAttribute(fn_sim_, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)low_.smem_bytes), "cuFuncSetAttribute(smem)");
    // time grid: dts / sqrt_dts exactly as the incrementors precompute them (increment.rs:38-41,75-79)
    std::vector<double> dts(S), sq(S);
    for (int t = 0; t < S; ++t) { dts[t] = u_.times[t + 1] - u_.times[t]; sq[t] = std::sqrt(dts[t]); }
    d_times_.upload(u_.times.data(), u_.times.size() * 8);
    d_dts_.upload(dts.data(), dts.size() * 8);
    d_sqrt_dts_.upload(sq.data(), sq.size() * 8);
    d_x0_.alloc((size_t)u_.P() * 8);
    if (uses_sobol(opt_.lower.rng) && dims > 0) {
        std::vector<uint32_t> V, lane, nib;
        sobol_tables((uint32_t)dims, V, lane, &nib, low_.direct ? 4u : 1u);
        if (low_.wide) {
            // sde_sim_wide.cuh reads the lane part in A-fragment order: [S][NKK][MT][32], entry (t, kk, m, l) =
            // x_d(8m + (l >> 2)) for d = t K + 4kk + (l & 3) (zero for the pad factors of the last factor step)
            const size_t K = (size_t)u_.K(), NKK = (size_t)low_.wide_nkk, MT = (size_t)low_.wide_mt;
            std::vector<uint32_t> lt((size_t)S * NKK * MT * 32, 0u);
            for (size_t t = 0; t < (size_t)S; ++t)
                for (size_t kk = 0; kk < NKK; ++kk)
                    for (size_t m = 0; m < MT; ++m)
                        for (size_t l = 0; l < 32; ++l) {
                            const size_t k = 4 * kk + (l & 3);
                            if (k < K) lt[((t * NKK + kk) * MT + m) * 32 + l] = lane[(t * K + k) * 32 + 8 * m + (l >> 2)];
                        }
            lane.swap(lt);
        }
        if (low_.resident || low_.wide) {
            // sde_sim_resident.cuh reads the nibble table dimension-fastest: [8][16][ld], ld = dims rounded up to 32
            const size_t ld = (dims + 31) & ~(size_t)31;
            std::vector<uint32_t> nt(128 * ld, 0u);
            for (size_t dd = 0; dd < dims; ++dd)
                for (size_t qv = 0; qv < 128; ++qv) nt[qv * ld + dd] = nib[dd * 128 + qv];
            nib.swap(nt);
        }
        d_nib_.upload(nib.data(), nib.size() * 4);
        if (low_.lane_global) {
            // sde_sim_resident.cuh with SDE_RES_LANE_GLOBAL reads the table as the CTA prologue would have built it in shared
            // memory; it depends on the digital-shift masks, i.e. on the seed (ensure_masks); without masks it is final now
            lane_host_ = lane;
            const size_t nq = (dims + 6) / 4;
            d_lane_.alloc(4 * nq * 128 * 4);
            if (opt_.lower.rng != RNG_SOBOL_XOR) upload_prepared_lane_table(nullptr);
        } else {
            d_lane_.upload(lane.data(), lane.size() * 4);
        }
        if (opt_.lower.rng == RNG_SOBOL_XOR) d_masks_.alloc(dims * 4);
    }
    trace.mark("tables (host build + upload)");
    cu_check(d.cuStreamCreate(&own_stream_, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    cu_check(d.cuStreamCreate(&copy_stream_, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    cu_check(d.cuEventCreate(&ev_a_, CU_EVENT_DEFAULT), "cuEventCreate");
    cu_check(d.cuEventCreate(&ev_b_, CU_EVENT_DEFAULT), "cuEventCreate");
    for (int i = 0; i < 2; ++i) {
        cu_check(d.cuEventCreate(&ev_done_[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        cu_check(d.cuEventCreate(&ev_copied_[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    }
}

Plan::~Plan() {
    std::string why;
    if (!driver_available(&why)) return;
    const DriverApi& d = driver();
    try { use_device(opt_.device); } catch (...) { return; }
    if (own_stream_) d.cuStreamSynchronize(own_stream_);
    if (copy_stream_) d.cuStreamSynchronize(copy_stream_);
    for (int i = 0; i < 2; ++i) { if (ev_done_[i]) d.cuEventDestroy(ev_done_[i]); if (ev_copied_[i]) d.cuEventDestroy(ev_copied_[i]); }
    if (ev_a_) d.cuEventDestroy(ev_a_);
    if (ev_b_) d.cuEventDestroy(ev_b_);
    if (own_stream_) d.cuStreamDestroy(own_stream_);
    if (copy_stream_) d.cuStreamDestroy(copy_stream_);
    if (mod_) d.cuModuleUnload(mod_);
}

size_t Plan::output_elems(uint64_t n) const {
    const size_t P = u_.P(), T = u_.T();
    switch (opt_.lower.out) {
        case OUT_PATHS_NTP: case OUT_PATHS_TPN: return (size_t)n * T * P;
        case OUT_TERMINAL: return (size_t)n * P;
        default: return P * 3;
    }
}

void Plan::set_initial_values(const std::vector<std::pair<std::string, double>>& init, CUstream stream) {
    // ScenarioFiltration::new (filtration.rs:42-50): unknown names ignored, missing -> 0.0
    std::vector<double> x0(u_.P(), 0.0);
    for (auto& kv : init) {
        auto it = u_.process_registry.find(kv.first);
        if (it != u_.process_registry.end()) x0[it->second] = kv.second;
    }
    if (x0 == x0_host_) return;
    x0_host_ = x0;
    const DriverApi& d = driver();
    // stream-ordered so a previous launch that still reads the old row is not disturbed
    cu_check(d.cuStreamSynchronize(stream), "cuStreamSynchronize");
    cu_check(d.cuMemcpyHtoD(d_x0_.ptr(), x0_host_.data(), x0_host_.size() * 8), "cuMemcpyHtoD(x0)");
}

void Plan::upload_prepared_lane_table(const uint32_t* masks) {
    const size_t dims = lane_host_.size() / 32, nq = (dims + 6) / 4;
    std::vector<uint32_t> t(4 * nq * 128, 0u);
    for (size_t off = 0; off < 4; ++off)
        for (size_t dd0 = 0; dd0 < dims; ++dd0) {
            const size_t dd = dd0 + off;
            const uint32_t m = masks ? masks[dd0] : 0u;
            for (size_t l = 0; l < 32; ++l) {
                uint32_t v = lane_host_[dd0 * 32 + l] ^ m;
                if (low_.res_fold) v ^= (uint32_t)((int32_t)v >> 31) & 0x7fffffffu;     // sign-folded form (SDE_RES_FOLD)
                t[off * nq * 128 + ((dd >> 2) * 32 + l) * 4 + (dd & 3)] = v;
            }
        }
    cu_check(driver().cuMemcpyHtoD(d_lane_.ptr(), t.data(), t.size() * 4), "cuMemcpyHtoD(prepared lane table)");
}

void Plan::ensure_masks(uint64_t seed, CUstream stream) {
    if (opt_.lower.rng != RNG_SOBOL_XOR || d_masks_.bytes() == 0) return;
    if (masks_valid_ && masks_seed_ == seed) return;
    // digital-shift mask of dimension d = top 32 bits of u64 #d of ChaCha8Rng::seed_from_u64(seed)
    std::vector<uint64_t> m64(d_masks_.bytes() / 4);
    chacha8_u64_host(seed, m64.size(), m64.data());
    std::vector<uint32_t> m(m64.size());
    for (size_t i = 0; i < m.size(); ++i) m[i] = (uint32_t)(m64[i] >> 32);
    const DriverApi& d = driver();
    cu_check(d.cuStreamSynchronize(stream), "cuStreamSynchronize");
    cu_check(d.cuMemcpyHtoD(d_masks_.ptr(), m.data(), m.size() * 4), "cuMemcpyHtoD(masks)");
    if (low_.lane_global) upload_prepared_lane_table(m.data());
    masks_valid_ = true;
    masks_seed_ = seed;
}

void Plan::launch(uint64_t n, uint64_t seed, uint64_t scenario_offset, double* d_out, const double* d_inject, CUstream stream, int* n_launches) {
    const DriverApi& d = driver();
    const uint64_t block = (uint64_t)low_.block;
    const uint64_t first_n = scenario_offset + 5;            // Sobol::new(..).skip(5)  (sobol.rs:17)
    if (uses_sobol(opt_.lower.rng) && first_n + n > (1ull << 32))
        throw ExprError{"sobol point index exceeds 2^32 (the reference's scenario index is i32, src/sim/mod.rs:47)"};
    if (low_.wide && opt_.lower.out == OUT_PATHS_NTP && ((uintptr_t)d_out & 15u))
        throw ExprError{"full-path output buffer must be 16-byte aligned (128-bit row stores of the tensor-core kernel)"};
    if (low_.tma && ((uintptr_t)d_out & 15u))
        throw ExprError{"full-path output buffer must be 16-byte aligned (bulk copies of row segments)"};
    if (low_.direct && ((uintptr_t)d_out & 31u))
        throw ExprError{"full-path output buffer must be 32-byte aligned (256-bit sector stores)"};
    SdeParamsHost prm{};
    prm.n_paths = n;
    prm.scen_offset = scenario_offset;
    prm.n_base = (first_n / block) * block;
    prm.seed = seed;
    prm.n_steps = u_.T() - 1;
    prm.times = d_times_.ptr(); prm.dts = d_dts_.ptr(); prm.sqrt_dts = d_sqrt_dts_.ptr(); prm.x0 = d_x0_.ptr();
    prm.sobol_nib = d_nib_.ptr(); prm.sobol_lane = d_lane_.ptr(); prm.xor_masks = d_masks_.ptr();
    prm.inject = (CUdeviceptr)d_inject;
    prm.out = (CUdeviceptr)d_out;
    uint64_t grid = (first_n + n - prm.n_base + block - 1) / block;
    if (low_.wide) {
        // persistent warps: one work item = 8 wide_mt paths, see sde_sim_wide.cuh
        const uint64_t wp = 8ull * (uint64_t)low_.wide_mt;
        const uint64_t items = (first_n + n - (first_n & ~(wp - 1)) + wp - 1) / wp;
        const uint64_t warps = block / 32;
        grid = std::min<uint64_t>((items + warps - 1) / warps, (uint64_t)sm_count(opt_.device));
    }
    if (low_.resident) {
        // persistent warps: one work item = 32 paths (lane stride 4 inside a 128-path block), see sde_sim_resident.cuh
        const uint64_t base128 = first_n & ~127ull;
        const uint64_t items = 4 * ((first_n + n - base128 + 127) / 128);
        const uint64_t warps = block / 32;
        grid = std::min<uint64_t>((items + warps - 1) / warps, (uint64_t)sm_count(opt_.device) * (uint64_t)low_.min_blocks);
        // a CTA serves one item class (item mod 4 = blockIdx mod 4: one step shift per CTA): whole groups of 4 CTAs
        grid = std::max<uint64_t>(4, grid & ~3ull);
    }
    if (grid == 0 || n == 0) return;
    if (low_.resident) {
        // the <= 5 leading and <= 127 trailing pad lanes of a launch store their (meaningless) rows here: one scratch row
        size_t need = ((size_t)u_.T() * u_.P() + 8) * 8;
        if (d_partials_.bytes() < need) { cu_check(d.cuStreamSynchronize(stream), "cuStreamSynchronize"); d_partials_.alloc(need); }
        prm.partials = d_partials_.ptr();
    }
    if (grid > 0x7fffffffull) throw ExprError{"too many scenarios for one launch"};
    const uint64_t n_partials = low_.wide ? grid * (block / 32) : grid;      // the wide kernel leaves one partial per warp
    if (opt_.lower.out == OUT_MOMENTS) {
        size_t need = (size_t)n_partials * u_.P() * 3 * 8;
        if (d_partials_.bytes() < need) { cu_check(d.cuStreamSynchronize(stream), "cuStreamSynchronize"); d_partials_.alloc(need); }
        prm.partials = d_partials_.ptr();
    }
    alignas(64) CUtensorMap tmap;
    void* args[] = {&prm, &tmap};
    if (low_.tma2) {
        // the output as a 2-D tensor [n rows][T P doubles]; a box is 32 rows x 128 bytes, written from shared memory in the
        // 128-byte swizzle pattern; rows / columns outside the tensor are clipped by the copy engine
        if (!d.cuTensorMapEncodeTiled) throw ExprError{"this CUDA driver has no cuTensorMapEncodeTiled (needed by ntp_direct=5)"};
        if (n >= (1ull << 31)) throw ExprError{"tensor-map stores address paths with 32-bit coordinates: at most 2^31 - 1 paths per launch"};
        const cuuint64_t TP = (cuuint64_t)u_.T() * (cuuint64_t)u_.P();
        const cuuint64_t dims[2] = {TP, (cuuint64_t)n};
        const cuuint64_t strides[1] = {TP * 8};
        const cuuint32_t box[2] = {16, 32};
        const cuuint32_t estr[2] = {1, 1};
        cu_check(d.cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d_out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
                 "cuTensorMapEncodeTiled");
    }
    cu_check(d.cuLaunchKernel(fn_sim_, (unsigned)grid, 1, 1, (unsigned)low_.block, 1, 1, (unsigned)low_.smem_bytes, stream, args, nullptr),
             "cuLaunchKernel(sde_sim_kernel)");
    if (n_launches) ++*n_launches;
    if (opt_.lower.out == OUT_MOMENTS) {
        CUdeviceptr parts = d_partials_.ptr();
        uint64_t np = n_partials;
        CUdeviceptr outp = (CUdeviceptr)d_out;
        void* fargs[] = {&parts, &np, &outp};
        cu_check(d.cuLaunchKernel(fn_fin_, 1, 1, 1, 256, 1, 1, 0, stream, fargs, nullptr), "cuLaunchKernel(sde_moments_finalize)");
        if (n_launches) ++*n_launches;
    }
}

void Plan::run_device(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed,
                      uint64_t scenario_offset, double* d_out, const double* d_inject, CUstream stream, int* n_launches) {
    // asynchronous on `stream`; nullptr is the legacy default stream (what torch.cuda.current_stream().cuda_stream is
    // unless the caller switched streams), so the launch is ordered with the caller's pending work on it
    use_device(opt_.device);
    if (opt_.lower.rng == RNG_INJECT && !d_inject) throw ExprError{"injected-draw plan needs options.inject"};
    set_initial_values(init, stream);
    ensure_masks(seed, stream);
    launch(n, seed, scenario_offset, d_out, d_inject, stream, n_launches);
}

void Plan::run_timed(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed,
                     uint64_t scenario_offset, double* d_out, const double* d_inject, CUstream stream, int* n_launches) {
    // synchronous: launches on `stream` (nullptr = the plan's own stream), waits, records the kernel time
    use_device(opt_.device);
    if (opt_.lower.rng == RNG_INJECT && !d_inject) throw ExprError{"injected-draw plan needs options.inject"};
    CUstream s = stream ? stream : own_stream_;
    set_initial_values(init, s);
    ensure_masks(seed, s);
    const DriverApi& d = driver();
    cu_check(d.cuEventRecord(ev_a_, s), "cuEventRecord");
    launch(n, seed, scenario_offset, d_out, d_inject, s, n_launches);
    cu_check(d.cuEventRecord(ev_b_, s), "cuEventRecord");
    cu_check(d.cuStreamSynchronize(s), "cuStreamSynchronize");
    float ms = 0;
    d.cuEventElapsedTime(&ms, ev_a_, ev_b_);
    last_ms_ = ms;
}

void Plan::run_host(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed,
                    uint64_t scenario_offset, double* h_out, int* n_launches) {
    use_device(opt_.device);
    if (opt_.lower.rng == RNG_INJECT) throw ExprError{"injected-draw plans run through sde_plan_run_device"};
    const DriverApi& d = driver();
    set_initial_values(init, own_stream_);
    ensure_masks(seed, own_stream_);
    const int out = opt_.lower.out;
    if (out == OUT_MOMENTS || out == OUT_PATHS_TPN) {
        // single shot: moments are tiny; the transposed layout is not scenario-contiguous
        size_t bytes = output_elems(n) * (out == OUT_MOMENTS ? 8 : elem_bytes());
        DeviceBuffer tmp(bytes);
        launch(n, seed, scenario_offset, tmp.as<double>(), nullptr, own_stream_, n_launches);
        cu_check(d.cuMemcpyDtoHAsync(h_out, tmp.ptr(), bytes, own_stream_), "cuMemcpyDtoHAsync");
        cu_check(d.cuStreamSynchronize(own_stream_), "cuStreamSynchronize");
        return;
    }
    // scenario-chunked, double-buffered: chunk i simulates while chunk i-1 drains over PCIe
    const size_t row_elems = out == OUT_TERMINAL ? (size_t)u_.P() : (size_t)u_.T() * u_.P();
    const size_t target_bytes = (size_t)512 << 20;
    const size_t eb = elem_bytes();
    uint64_t chunk = std::max<uint64_t>(1, target_bytes / (row_elems * eb));
    chunk = std::min<uint64_t>(n, std::max<uint64_t>(chunk, (uint64_t)low_.block));
    for (int b = 0; b < 2; ++b)
        if (d_chunk_[b].bytes() < chunk * row_elems * eb) d_chunk_[b].alloc(chunk * row_elems * eb);
    uint64_t done = 0;
    int i = 0;
    bool used[2] = {false, false};
    while (done < n) {
        const uint64_t m = std::min<uint64_t>(chunk, n - done);
        const int b = i & 1;
        if (used[b]) cu_check(d.cuStreamWaitEvent(own_stream_, ev_copied_[b], 0), "cuStreamWaitEvent");
        launch(m, seed, scenario_offset + done, d_chunk_[b].as<double>(), nullptr, own_stream_, n_launches);
        cu_check(d.cuEventRecord(ev_done_[b], own_stream_), "cuEventRecord");
        cu_check(d.cuStreamWaitEvent(copy_stream_, ev_done_[b], 0), "cuStreamWaitEvent");
        cu_check(d.cuMemcpyDtoHAsync(reinterpret_cast<unsigned char*>(h_out) + done * row_elems * eb, d_chunk_[b].ptr(), m * row_elems * eb, copy_stream_), "cuMemcpyDtoHAsync");
        cu_check(d.cuEventRecord(ev_copied_[b], copy_stream_), "cuEventRecord");
        used[b] = true;
        done += m;
        ++i;
    }
    cu_check(d.cuStreamSynchronize(copy_stream_), "cuStreamSynchronize");
    cu_check(d.cuStreamSynchronize(own_stream_), "cuStreamSynchronize");
}

// --------------------------------------------------------------------------------------
// Stand-alone kernels (AOT cubin)
// --------------------------------------------------------------------------------------
namespace {
struct UtilModule {
    CUmodule mod = nullptr;
    CUfunction sobol = nullptr, chacha = nullptr, icdf = nullptr, icdf_wide = nullptr, poisson = nullptr, fill = nullptr, dfma = nullptr, ffma = nullptr,
               merge = nullptr, cp_uniforms = nullptr;
};
UtilModule& util_module(int device) {
    static std::mutex mu;
    static UtilModule mods[64];
    std::lock_guard<std::mutex> lock(mu);
    use_device(device);
    UtilModule& m = mods[device];
    if (!m.mod) {
        const DriverApi& d = driver();
        cu_check(d.cuModuleLoadData(&m.mod, sde_blob_util_cubin_begin), "cuModuleLoadData(util)");
        cu_check(d.cuModuleGetFunction(&m.sobol, m.mod, "sde_k_sobol_points"), "sde_k_sobol_points");
        cu_check(d.cuModuleGetFunction(&m.chacha, m.mod, "sde_k_chacha8_u64"), "sde_k_chacha8_u64");
        cu_check(d.cuModuleGetFunction(&m.icdf, m.mod, "sde_k_icdf_normal"), "sde_k_icdf_normal");
        cu_check(d.cuModuleGetFunction(&m.icdf_wide, m.mod, "sde_k_icdf_normal_wide"), "sde_k_icdf_normal_wide");
        cu_check(d.cuFuncSetAttribute(m.icdf_wide, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 1024 * 2 * 8 * 8), "cuFuncSetAttribute(wide)");
        cu_check(d.cuModuleGetFunction(&m.poisson, m.mod, "sde_k_icdf_poisson"), "sde_k_icdf_poisson");
        cu_check(d.cuModuleGetFunction(&m.fill, m.mod, "sde_k_fill"), "sde_k_fill");
        cu_check(d.cuModuleGetFunction(&m.dfma, m.mod, "sde_k_dfma"), "sde_k_dfma");
        cu_check(d.cuModuleGetFunction(&m.ffma, m.mod, "sde_k_ffma"), "sde_k_ffma");
        cu_check(d.cuModuleGetFunction(&m.merge, m.mod, "sde_k_moments_merge"), "sde_k_moments_merge");
        cu_check(d.cuModuleGetFunction(&m.cp_uniforms, m.mod, "sde_k_sobol_cp_uniforms"), "sde_k_sobol_cp_uniforms");
    }
    return m;
}
void launch1d(CUfunction f, uint64_t grid, unsigned block, unsigned smem, void** args) {
    cu_check(driver().cuLaunchKernel(f, (unsigned)grid, 1, 1, block, 1, 1, smem, nullptr, args, nullptr), "cuLaunchKernel(util)");
}
}  // namespace

void util_sobol_points(int device, uint32_t dims, uint64_t first, uint64_t count, uint64_t* h_out) {
    if (dims == 0 || count == 0) return;
    if (first + count > (1ull << 32)) throw ExprError{"sobol point index exceeds 2^32"};
    UtilModule& m = util_module(device);
    const DriverApi& d = driver();
    std::vector<uint32_t> V, lane;
    sobol_tables(dims, V, lane);
    DeviceBuffer dV, dL, dout(count * dims * 8);
    dV.upload(V.data(), V.size() * 4);
    dL.upload(lane.data(), lane.size() * 4);
    uint64_t n_base = (first / 256) * 256;
    uint64_t grid = (first + count - n_base + 255) / 256;
    CUdeviceptr pV = dV.ptr(), pL = dL.ptr(), po = dout.ptr();
    void* args[] = {&pV, &pL, &dims, &n_base, &first, &count, &po};
    launch1d(m.sobol, grid, 256, 256 * 8 * 4, args);
    cu_check(d.cuMemcpyDtoH(h_out, dout.ptr(), count * dims * 8), "cuMemcpyDtoH");
}

void util_sobol_cp_uniforms(int device, uint32_t dims, uint64_t seed, uint64_t first_scenario, uint64_t count, double* h_out) {
    if (dims == 0 || count == 0) return;
    if (first_scenario + 5 + count > (1ull << 32)) throw ExprError{"sobol point index exceeds 2^32"};
    UtilModule& m = util_module(device);
    std::vector<uint32_t> V, lane;
    sobol_tables(dims, V, lane);
    DeviceBuffer dV, dout(count * dims * 8);
    dV.upload(V.data(), V.size() * 4);
    CUdeviceptr pV = dV.ptr(), po = dout.ptr();
    void* args[] = {&pV, &dims, &seed, &first_scenario, &count, &po};
    launch1d(m.cp_uniforms, (count + 127) / 128, 128, 0, args);
    cu_check(driver().cuMemcpyDtoH(h_out, dout.ptr(), count * dims * 8), "cuMemcpyDtoH");
}

void util_chacha8_u64(int device, uint64_t seed, size_t n, uint64_t* h_out) {
    if (!n) return;
    UtilModule& m = util_module(device);
    DeviceBuffer dout(n * 8);
    uint64_t nn = n;
    CUdeviceptr po = dout.ptr();
    void* args[] = {&seed, &nn, &po};
    launch1d(m.chacha, ((n + 7) / 8 + 127) / 128, 128, 0, args);
    cu_check(driver().cuMemcpyDtoH(h_out, dout.ptr(), n * 8), "cuMemcpyDtoH");
}

void util_icdf_normal(int device, int mode, const double* h_p, size_t n, double* h_out) {
    if (!n) return;
    UtilModule& m = util_module(device);
    DeviceBuffer din(n * 8), dout(n * 8);
    din.upload(h_p, n * 8);
    uint64_t nn = n;
    CUdeviceptr pi = din.ptr(), po = dout.ptr();
    if (mode == 5 || mode == 6) {                            // 32-bit front end, 1024-entry log table (6: FP32-unit seeds)
        int f32seed = mode == 6 ? 1 : 0;
        void* wargs[] = {&pi, &nn, &po, &f32seed};
        launch1d(m.icdf_wide, std::min<uint64_t>((n + 255) / 256, 1184), 256, 1024 * 2 * 8 * 8, wargs);
    } else {
        void* args[] = {&pi, &nn, &mode, &po};
        launch1d(m.icdf, (n + 255) / 256, 256, 0, args);
    }
    cu_check(driver().cuMemcpyDtoH(h_out, dout.ptr(), n * 8), "cuMemcpyDtoH");
}

void util_icdf_poisson(int device, const double* h_u, const double* h_lambda, size_t n, double* h_out) {
    if (!n) return;
    UtilModule& m = util_module(device);
    DeviceBuffer du(n * 8), dl(n * 8), dout(n * 8);
    du.upload(h_u, n * 8);
    dl.upload(h_lambda, n * 8);
    uint64_t nn = n;
    CUdeviceptr pu = du.ptr(), pl = dl.ptr(), po = dout.ptr();
    void* args[] = {&pu, &pl, &nn, &po};
    launch1d(m.poisson, (n + 127) / 128, 128, 0, args);
    cu_check(driver().cuMemcpyDtoH(h_out, dout.ptr(), n * 8), "cuMemcpyDtoH");
}

void util_measure_peaks(int device, double* fill_gbs, double* dfma_tflops, double* ffma_tflops) {
    UtilModule& m = util_module(device);
    const DriverApi& d = driver();
    const int sms = sm_count(device);
    CUevent e0, e1;
    cu_check(d.cuEventCreate(&e0, CU_EVENT_DEFAULT), "cuEventCreate");
    cu_check(d.cuEventCreate(&e1, CU_EVENT_DEFAULT), "cuEventCreate");
    auto timed = [&](auto&& body, int reps) {
        body();                                             // warm-up
        body();
        body();
        float best = 1e30f;
        for (int r = 0; r < reps; ++r) {
            d.cuEventRecord(e0, nullptr);
            body();
            d.cuEventRecord(e1, nullptr);
            d.cuEventSynchronize(e1);
            float ms = 0;
            d.cuEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        return (double)best * 1e-3;
    };
    {
        const size_t bytes = (size_t)8 << 30;               // 8 GiB >> L2
        DeviceBuffer buf(bytes);
        uint64_t nvec = bytes / 16;
        CUdeviceptr p = buf.ptr();
        double v = 1.0;
        void* args[] = {&p, &nvec, &v};
        double s = timed([&] { launch1d(m.fill, (uint64_t)sms * 16, 256, 0, args); }, 10);
        if (fill_gbs) *fill_gbs = (double)bytes / s * 1e-9;
    }
    {
        const uint64_t grid = (uint64_t)sms * 8;
        DeviceBuffer out(grid * 256 * 8);
        CUdeviceptr p = out.ptr();
        int iters = 1 << 15;
        double a = 1.0000001, b = 1e-9;
        void* args[] = {&p, &iters, &a, &b};
        double s = timed([&] { launch1d(m.dfma, grid, 256, 0, args); }, 5);
        if (dfma_tflops) *dfma_tflops = (double)grid * 256 * iters * 8 * 2 / s * 1e-12;
        float af = 1.0000001f, bf = 1e-9f;
        void* fargs[] = {&p, &iters, &af, &bf};
        s = timed([&] { launch1d(m.ffma, grid, 256, 0, fargs); }, 5);
        if (ffma_tflops) *ffma_tflops = (double)grid * 256 * iters * 8 * 2 / s * 1e-12;
    }
    d.cuEventDestroy(e0);
    d.cuEventDestroy(e1);
}

void moments_merge_device(int device, const double* d_shards, size_t n_shards, size_t P, double* d_out, CUstream stream) {
    UtilModule& m = util_module(device);
    CUdeviceptr ps = (CUdeviceptr)d_shards, po = (CUdeviceptr)d_out;
    uint64_t ns = n_shards;
    int p = (int)P;
    void* args[] = {&ps, &ns, &p, &po};
    cu_check(driver().cuLaunchKernel(m.merge, (unsigned)((P + 127) / 128), 1, 1, 128, 1, 1, 0, stream, args, nullptr), "cuLaunchKernel(sde_k_moments_merge)");
}

void shard_range(uint64_t n, size_t part, size_t parts, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = n / parts, rem = n % parts, i = part;
    const uint64_t l = i * base + std::min<uint64_t>(i, rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (i < rem ? 1 : 0);
}

DevicePlans::DevicePlans(const Universe& u, const PlanOptions& opt, const std::vector<int>& devices) : devices_(devices) {
    if (devices_.empty()) throw ExprError{"device list is empty"};
    const int n_dev = device_count();
    bool repeated = false;                                   // a device listed twice (several shards on one GPU): fine, but not for NCCL
    for (size_t i = 0; i < devices_.size(); ++i) {
        if (devices_[i] < 0 || devices_[i] >= n_dev) throw ExprError{"device ordinal " + std::to_string(devices_[i]) + " out of range (" + std::to_string(n_dev) + " visible)"};
        for (size_t j = 0; j < i; ++j) if (devices_[j] == devices_[i]) repeated = true;
    }
    for (int dev : devices_) {
        PlanOptions po = opt;
        po.device = dev;
        plans_.push_back(std::make_shared<Plan>(u, po));    // same lowering: the cubin comes out of the in-memory cache after the first
    }
    const size_t G = devices_.size(), P = (size_t)u.P();
    if (opt.lower.out == OUT_MOMENTS && G > 1) {
        const DriverApi& d = driver();
        local_.resize(G);
        gathered_.resize(G);
        ev_local_.assign(G, nullptr);
        for (size_t i = 0; i < G; ++i) {
            use_device(devices_[i]);
            local_[i].alloc(P * 3 * 8);
            gathered_[i].alloc(G * P * 3 * 8);
            cu_check(d.cuEventCreate(&ev_local_[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        }
        use_device(devices_[0]);
        cu_check(d.cuEventCreate(&ev_c0_, CU_EVENT_DEFAULT), "cuEventCreate");
        cu_check(d.cuEventCreate(&ev_c1_, CU_EVENT_DEFAULT), "cuEventCreate");
        collective_ = 2;
        if (!repeated && !std::getenv("SDE_B200_NO_NCCL")) {
            if (const NcclApi* nc = nccl()) {
                comms_.assign(G, nullptr);
                const int rc = nc->ncclCommInitAll(comms_.data(), (int)G, devices_.data());
                if (rc == 0) collective_ = 1;
                else comms_.clear();                         // e.g. a NCCL build without these GPUs: peer copies still work
            }
        }
        if (collective_ == 2) {
            for (size_t i = 0; i < G; ++i)
                for (size_t j = 0; j < G; ++j) {
                    if (devices_[i] == devices_[j]) continue;
                    use_device(devices_[i]);
                    const CUresult r = d.cuCtxEnablePeerAccess(device_context(devices_[j]), 0);
                    (void)r;                                 // already enabled / not supported: cuMemcpyPeerAsync stages through the host then
                }
        }
    }
}

DevicePlans::~DevicePlans() {
    std::string why;
    if (!driver_available(&why)) return;
    const DriverApi& d = driver();
    try {
        for (size_t i = 0; i < plans_.size(); ++i) {
            use_device(devices_[i]);
            d.cuStreamSynchronize(plans_[i]->stream());
            if (i < ev_local_.size() && ev_local_[i]) d.cuEventDestroy(ev_local_[i]);
            if (i < local_.size()) { local_[i].release(); gathered_[i].release(); }
        }
        if (ev_c0_) { use_device(devices_[0]); d.cuEventDestroy(ev_c0_); d.cuEventDestroy(ev_c1_); }
        if (!comms_.empty()) if (const NcclApi* nc = nccl()) for (void* c : comms_) if (c) nc->ncclCommDestroy(c);
    } catch (...) {}
}

void DevicePlans::run(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed, uint64_t scenario_offset,
                      double* const* d_out, int* n_launches, double* collective_ms) {
    const DriverApi& d = driver();
    const size_t G = plans_.size(), P = (size_t)plans_[0]->universe().P();
    const bool moments = plans_[0]->options().lower.out == OUT_MOMENTS;
    if (!d_out) throw ExprError{"d_out must not be NULL"};
    if (collective_ms) *collective_ms = 0.0;
    // 1. every device's shard is launched on its plan's stream; nothing is synchronised in this loop once the plans
    //    hold their tables (first call: initial row / digital-shift masks are uploaded)
    for (size_t i = 0; i < G; ++i) {
        uint64_t lo, hi;
        shard_range(n, i, G, &lo, &hi);
        use_device(devices_[i]);
        double* dst = (moments && G > 1) ? local_[i].as<double>() : d_out[i];
        if (hi > lo) {
            if (!dst) throw ExprError{"d_out[" + std::to_string(i) + "] must not be NULL"};
            plans_[i]->run_device(init, hi - lo, seed, scenario_offset + lo, dst, nullptr, plans_[i]->stream(), n_launches);
        } else if (moments && G > 1) {
            cu_check(d.cuMemsetD8Async(local_[i].ptr(), 0, P * 3 * 8, plans_[i]->stream()), "cuMemsetD8Async");   // count 0: skipped by the merge
        }
        if (moments && G > 1 && collective_ == 2) cu_check(d.cuEventRecord(ev_local_[i], plans_[i]->stream()), "cuEventRecord");
    }
    // 2. moments: all-gather of the 3 P doubles per device + merge kernel, on the compute streams
    if (moments && G > 1) {
        use_device(devices_[0]);
        cu_check(d.cuEventRecord(ev_c0_, plans_[0]->stream()), "cuEventRecord");
        if (collective_ == 1) {
            const NcclApi* nc = nccl();
            auto nccl_check = [&](int rc, const char* what) { if (rc != 0) throw CudaError{std::string(what) + ": " + nc->ncclGetErrorString(rc)}; };
            nccl_check(nc->ncclGroupStart(), "ncclGroupStart");
            for (size_t i = 0; i < G; ++i)
                nccl_check(nc->ncclAllGather(local_[i].as<double>(), gathered_[i].as<double>(), P * 3, /*ncclDouble*/ 8, comms_[i], plans_[i]->stream()), "ncclAllGather");
            nccl_check(nc->ncclGroupEnd(), "ncclGroupEnd");
        } else {
            for (size_t j = 0; j < G; ++j) {                 // destination device j pulls every shard's triple
                use_device(devices_[j]);
                for (size_t i = 0; i < G; ++i) {
                    if (i != j) cu_check(d.cuStreamWaitEvent(plans_[j]->stream(), ev_local_[i], 0), "cuStreamWaitEvent");
                    cu_check(d.cuMemcpyPeerAsync(gathered_[j].ptr() + i * P * 3 * 8, device_context(devices_[j]), local_[i].ptr(),
                                                 device_context(devices_[i]), P * 3 * 8, plans_[j]->stream()), "cuMemcpyPeerAsync");
                }
            }
        }
        for (size_t i = 0; i < G; ++i) {
            use_device(devices_[i]);
            if (d_out[i]) moments_merge_device(devices_[i], gathered_[i].as<double>(), G, P, d_out[i], plans_[i]->stream());
            if (n_launches && d_out[i]) ++*n_launches;
        }
        use_device(devices_[0]);
        cu_check(d.cuEventRecord(ev_c1_, plans_[0]->stream()), "cuEventRecord");
    }
    // 3. wait for everything
    for (size_t i = 0; i < G; ++i) {
        use_device(devices_[i]);
        cu_check(d.cuStreamSynchronize(plans_[i]->stream()), "cuStreamSynchronize");
    }
    if (moments && G > 1 && collective_ms) {
        float ms = 0;
        d.cuEventElapsedTime(&ms, ev_c0_, ev_c1_);
        *collective_ms = ms;
    }
}

void moments_merge(const double* shards, size_t n_shards, size_t P, double* out) {
    for (size_t p = 0; p < P; ++p) {
        double n = 0, mean = 0, m2 = 0;
        for (size_t s = 0; s < n_shards; ++s) {
            const double* b = shards + (s * P + p) * 3;
            const double nb = b[0];
            if (nb == 0) continue;
            const double nn = n + nb, dlt = b[1] - mean, f = nb / nn;
            m2 = m2 + b[2] + dlt * dlt * n * f;
            mean = mean + dlt * f;
            n = nn;
        }
        out[p * 3] = n; out[p * 3 + 1] = mean; out[p * 3 + 2] = m2;
    }
}

}  // namespace sde
