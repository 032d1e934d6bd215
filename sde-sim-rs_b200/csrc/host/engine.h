// engine.h — plans and launches: the host side of sim::simulate (src/sim/mod.rs:20-92).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "cuda_rt.h"
#include "lower.h"
#include "universe.h"

namespace sde {

// ---- Joe–Kuo / Sobol host tables (replaces sobol::params::JoeKuoD6::extended, src/rng/sobol.rs:7,16)
constexpr uint32_t kMaxSobolDims = 21201;
void joe_kuo_params(uint32_t dims, uint32_t* poly, uint32_t* minit /*[dims][18]*/);
// V[d][b], b < 32: top 32 bits of the 64-bit direction numbers.  lane[d][l] = x_d(l), l < 32.
// nib[d][i][v] (optional), i < 8, v < 16: XOR of V[d][4i+b] over the set bits b of v, so that
// x_d(n) = XOR_i nib[d][i][(gray(n) >> 4i) & 15] — 8 independent loads instead of a bit loop.
void sobol_tables(uint32_t dims, std::vector<uint32_t>& V, std::vector<uint32_t>& lane, std::vector<uint32_t>* nib = nullptr,
                  uint32_t lane_stride = 1);   // lane[d][l] = x_d(lane_stride * l)
// u64 #i of ChaCha8Rng::seed_from_u64(seed) (host copy, used for the XOR digital-shift masks)
void chacha8_u64_host(uint64_t seed, size_t n, uint64_t* out);

struct PlanOptions {
    int device = 0;
    LowerOptions lower;
};

class Plan {
  public:
    Plan(const Universe& u, const PlanOptions& opt);
    ~Plan();
    Plan(const Plan&) = delete;
    Plan& operator=(const Plan&) = delete;

    const Universe& universe() const { return u_; }
    const Lowered& lowered() const { return low_; }
    const PlanOptions& options() const { return opt_; }
    bool prelowered() const { return prelowered_; }
    size_t output_elems(uint64_t n) const;
    // bytes per element of paths / terminal output (moments are always f64)
    size_t elem_bytes() const { return opt_.lower.f32 ? 4 : 8; }
    size_t output_bytes(uint64_t n) const { return output_elems(n) * (opt_.lower.out == OUT_MOMENTS ? 8 : elem_bytes()); }

    // Asynchronous on `stream` (nullptr = the legacy default stream).
    void run_device(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed,
                    uint64_t scenario_offset, double* d_out, const double* d_inject, CUstream stream, int* n_launches);
    // Synchronous on `stream` (nullptr = the plan's own stream): waits for the result and records last_kernel_ms().
    void run_timed(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed,
                   uint64_t scenario_offset, double* d_out, const double* d_inject, CUstream stream, int* n_launches);
    void run_host(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed,
                  uint64_t scenario_offset, double* h_out, int* n_launches);
    double last_kernel_ms() const { return last_ms_; }
    int device() const { return opt_.device; }
    CUstream stream() const { return own_stream_; }          // the plan's own (non-blocking) stream

  private:
    Universe u_;
    PlanOptions opt_;
    Lowered low_;
    bool prelowered_ = false;
    CUmodule mod_ = nullptr;
    CUfunction fn_sim_ = nullptr, fn_fin_ = nullptr;
    DeviceBuffer d_times_, d_dts_, d_sqrt_dts_, d_x0_, d_nib_, d_lane_, d_masks_, d_partials_;
    std::vector<uint32_t> lane_host_;     // persistent kernel with the lane table in global memory: x_d(4 lane), [dims][32]
    void upload_prepared_lane_table(const uint32_t* masks);   // [4 offsets][quads][32 lanes][4]: mask and sign fold applied
    std::vector<double> x0_host_;
    bool masks_valid_ = false;
    uint64_t masks_seed_ = 0;
    CUstream own_stream_ = nullptr, copy_stream_ = nullptr;
    CUevent ev_a_ = nullptr, ev_b_ = nullptr;
    DeviceBuffer d_chunk_[2];
    CUevent ev_done_[2] = {nullptr, nullptr}, ev_copied_[2] = {nullptr, nullptr};
    double last_ms_ = 0.0;

    void set_initial_values(const std::vector<std::pair<std::string, double>>& init, CUstream stream);
    void ensure_masks(uint64_t seed, CUstream stream);
    void launch(uint64_t n, uint64_t seed, uint64_t scenario_offset, double* d_out, const double* d_inject, CUstream stream, int* n_launches);
};

// ---- several GPUs of one box driven by ONE host thread: replaces the rayon par_iter over scenarios inside
// sim::simulate (src/sim/mod.rs:41-43).  Device i of G simulates scenarios shard_range(N, i, G) (disjoint Sobol index
// ranges / ChaCha keys: the union is bit-identical to a one-device run).  Every launch is issued before any device is
// synchronised.  Moments: each device reduces its shard, the 3 P doubles per device are all-gathered over NCCL
// (ncclAllGather inside one group, on the compute streams) and Chan-merged by a device kernel right behind it, so every
// device ends up with the merged [P][3]; without libnccl the same gather runs as peer-to-peer copies.
void shard_range(uint64_t n, size_t part, size_t parts, uint64_t* lo, uint64_t* hi);
class DevicePlans {
  public:
    DevicePlans(const Universe& u, const PlanOptions& opt, const std::vector<int>& devices);
    ~DevicePlans();
    DevicePlans(const DevicePlans&) = delete;
    DevicePlans& operator=(const DevicePlans&) = delete;
    size_t size() const { return plans_.size(); }
    const std::shared_ptr<Plan>& plan(size_t i) const { return plans_[i]; }
    int device(size_t i) const { return devices_[i]; }
    int collective() const { return collective_; }           // 0 none, 1 NCCL, 2 peer-to-peer copies
    // d_out[i]: device memory on device i for that shard's rows / terminal values ([n_i][T][P] / [n_i][P]) or, for moments,
    // the merged [P][3] (every device receives it).  Synchronous: returns when every device has finished.
    void run(const std::vector<std::pair<std::string, double>>& init, uint64_t n, uint64_t seed, uint64_t scenario_offset,
             double* const* d_out, int* n_launches, double* collective_ms);

  private:
    std::vector<std::shared_ptr<Plan>> plans_;
    std::vector<int> devices_;
    int collective_ = 0;
    std::vector<void*> comms_;                                // ncclComm_t per device
    std::vector<DeviceBuffer> local_, gathered_;              // [P][3] / [G][P][3] per device
    std::vector<CUevent> ev_local_;                           // shard moments ready (peer-copy path)
    CUevent ev_c0_ = nullptr, ev_c1_ = nullptr;               // collective + merge on device 0
};
// Chan merge of [n_shards][P][3] on the device (util cubin), asynchronous on `stream` of `device`
void moments_merge_device(int device, const double* d_shards, size_t n_shards, size_t P, double* d_out, CUstream stream);

// ---- stand-alone building blocks (ahead-of-time cubin)
void util_sobol_points(int device, uint32_t dims, uint64_t first, uint64_t count, uint64_t* h_out);
void util_sobol_cp_uniforms(int device, uint32_t dims, uint64_t seed, uint64_t first_scenario, uint64_t count, double* h_out);
void util_chacha8_u64(int device, uint64_t seed, size_t n, uint64_t* h_out);
void util_icdf_normal(int device, int mode, const double* h_p, size_t n, double* h_out);
void util_icdf_poisson(int device, const double* h_u, const double* h_lambda, size_t n, double* h_out);
void util_measure_peaks(int device, double* fill_gbs, double* dfma_tflops, double* ffma_tflops);
void moments_merge(const double* shards, size_t n_shards, size_t P, double* out);

}  // namespace sde
