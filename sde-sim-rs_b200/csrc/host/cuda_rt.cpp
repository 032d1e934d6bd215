// cuda_rt.cpp — see cuda_rt.h.
#include "cuda_rt.h"

#include <dlfcn.h>

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#define SDE_STR2(x) #x
#define SDE_STR(x) SDE_STR2(x)

extern "C" {
extern const char sde_blob_sim_kernel_cuh_begin[], sde_blob_sim_kernel_cuh_end[];
extern const char sde_blob_sim_common_cuh_begin[], sde_blob_sim_common_cuh_end[];
extern const char sde_blob_sim_resident_cuh_begin[], sde_blob_sim_resident_cuh_end[];
extern const char sde_blob_sim_wide_cuh_begin[], sde_blob_sim_wide_cuh_end[];
extern const char sde_blob_device_rng_cuh_begin[], sde_blob_device_rng_cuh_end[];
extern const char sde_blob_device_icdf_cuh_begin[], sde_blob_device_icdf_cuh_end[];
extern const char sde_blob_icdf_tables_cuh_begin[], sde_blob_icdf_tables_cuh_end[];
extern const char sde_blob_expr_helpers_cuh_begin[], sde_blob_expr_helpers_cuh_end[];
}

namespace sde {
namespace {

void* open_first(const char* const* names, std::string* tried) {
    for (int i = 0; names[i]; ++i) {
        if (void* h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL)) return h;
        if (tried) { *tried += names[i]; *tried += " "; }
    }
    return nullptr;
}

struct DriverState {
    DriverApi api{};
    bool ok = false;
    std::string why;
};

DriverState& driver_state() {
    static DriverState st;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libcuda.so.1", "libcuda.so", nullptr};
        std::string tried;
        void* h = open_first(names, &tried);
        if (!h) { st.why = "CUDA driver library not found (tried " + tried + "): no GPU, and this library has no CPU fallback"; return; }
        bool missing = false;
        std::string miss;
#define SDE_LOAD(field)                                                                 \
        st.api.field = reinterpret_cast<decltype(st.api.field)>(dlsym(h, SDE_STR(field))); \
        if (!st.api.field) { missing = true; miss += SDE_STR(field) " "; }
        SDE_LOAD(cuInit) SDE_LOAD(cuDeviceGet) SDE_LOAD(cuDeviceGetCount) SDE_LOAD(cuDeviceGetAttribute)
        SDE_LOAD(cuDevicePrimaryCtxRetain) SDE_LOAD(cuCtxSetCurrent) SDE_LOAD(cuCtxGetCurrent)
        SDE_LOAD(cuModuleLoadData) SDE_LOAD(cuModuleUnload) SDE_LOAD(cuModuleGetFunction)
        SDE_LOAD(cuFuncSetAttribute) SDE_LOAD(cuFuncGetAttribute) SDE_LOAD(cuLaunchKernel)
        SDE_LOAD(cuMemAlloc) SDE_LOAD(cuMemFree) SDE_LOAD(cuMemcpyHtoD) SDE_LOAD(cuMemcpyDtoH)
        SDE_LOAD(cuMemcpyHtoDAsync) SDE_LOAD(cuMemcpyDtoHAsync) SDE_LOAD(cuMemsetD8Async)
        SDE_LOAD(cuMemHostAlloc) SDE_LOAD(cuMemFreeHost) SDE_LOAD(cuMemHostRegister) SDE_LOAD(cuMemHostUnregister)
        SDE_LOAD(cuMemGetInfo) SDE_LOAD(cuStreamCreate) SDE_LOAD(cuStreamDestroy) SDE_LOAD(cuStreamSynchronize)
        SDE_LOAD(cuStreamWaitEvent) SDE_LOAD(cuEventCreate) SDE_LOAD(cuEventDestroy) SDE_LOAD(cuEventRecord)
        SDE_LOAD(cuEventSynchronize) SDE_LOAD(cuEventElapsedTime) SDE_LOAD(cuGetErrorString)
        SDE_LOAD(cuOccupancyMaxActiveBlocksPerMultiprocessor)
        SDE_LOAD(cuMemcpyDtoDAsync) SDE_LOAD(cuMemcpyPeerAsync) SDE_LOAD(cuCtxEnablePeerAccess) SDE_LOAD(cuDeviceCanAccessPeer)
#undef SDE_LOAD
        st.api.cuTensorMapEncodeTiled = reinterpret_cast<decltype(st.api.cuTensorMapEncodeTiled)>(dlsym(h, "cuTensorMapEncodeTiled"));
        if (missing) { st.why = "CUDA driver is missing symbols: " + miss; return; }
        CUresult r = st.api.cuInit(0);
        if (r != CUDA_SUCCESS) { st.why = "cuInit failed (" + std::to_string((int)r) + "): no usable GPU"; return; }
        int n = 0;
        if (st.api.cuDeviceGetCount(&n) != CUDA_SUCCESS || n <= 0) { st.why = "no CUDA device visible"; return; }
        st.ok = true;
    });
    return st;
}

struct NvrtcState {
    NvrtcApi api{};
    int major = 0, minor = 0;
    bool ok = false;
    std::string why;
};

// Where the toolkit's NVRTC would be loaded from (see nvrtc_state): the two path candidates.
void nvrtc_path_candidates(std::string* p1, std::string* p2) {
    const std::string cuda_home = std::getenv("CUDA_HOME") ? std::getenv("CUDA_HOME") : "/usr/local/cuda";
    *p1 = cuda_home + "/lib64/libnvrtc.so.12";
    *p2 = cuda_home + "/targets/x86_64-linux/lib/libnvrtc.so.12";
}

// NVRTC's version WITHOUT loading it: the real file behind the soname is libnvrtc.so.<major>.<minor>.<patch>.  Loading the
// library costs ~90 ms (measured: the whole of a first plan creation that hits the cubin cache), and a cache hit needs
// nothing else from it.  False when the name does not parse (then the caller loads the library and asks it).
bool nvrtc_version_from_filename(int* major, int* minor) {
    std::string c[2];
    nvrtc_path_candidates(&c[0], &c[1]);
    for (const std::string& p : c) {
        char buf[4096];
        if (!realpath(p.c_str(), buf)) continue;
        const std::string real(buf);
        const size_t at = real.rfind("libnvrtc.so.");
        int a = 0, b = 0;
        if (at != std::string::npos && std::sscanf(real.c_str() + at, "libnvrtc.so.%d.%d", &a, &b) == 2 && a >= 11) {
            *major = a; *minor = b;
            return true;
        }
        return false;                                        // this is the file that would be loaded, and its name says nothing
    }
    return false;
}

NvrtcState& nvrtc_state() {
    static NvrtcState st;
    static std::once_flag once;
    std::call_once(once, [] {
        // Prefer the toolkit's NVRTC by path: a process that imported PyTorch already holds torch's bundled
        // libnvrtc.so.12 (CUDA 12.8), which a bare soname lookup would return and whose ptxas predates
        // 256-bit vector stores.
        std::string p1, p2;
        nvrtc_path_candidates(&p1, &p2);
        const char* names[] = {p1.c_str(), p2.c_str(), "libnvrtc.so.12", "libnvrtc.so", nullptr};
        std::string tried;
        void* h = open_first(names, &tried);
        if (!h) { st.why = "NVRTC not found (tried " + tried + ")"; return; }
        bool missing = false;
#define SDE_LOAD(field)                                                                 \
        st.api.field = reinterpret_cast<decltype(st.api.field)>(dlsym(h, SDE_STR(field))); \
        if (!st.api.field) missing = true;
        SDE_LOAD(nvrtcCreateProgram) SDE_LOAD(nvrtcDestroyProgram) SDE_LOAD(nvrtcCompileProgram)
        SDE_LOAD(nvrtcGetCUBINSize) SDE_LOAD(nvrtcGetCUBIN) SDE_LOAD(nvrtcGetProgramLogSize)
        SDE_LOAD(nvrtcGetProgramLog) SDE_LOAD(nvrtcGetErrorString) SDE_LOAD(nvrtcVersion)
#undef SDE_LOAD
        if (missing) { st.why = "NVRTC is missing symbols"; return; }
        st.api.nvrtcVersion(&st.major, &st.minor);
        st.ok = true;
    });
    return st;
}

}  // namespace

const DriverApi& driver() {
    DriverState& st = driver_state();
    if (!st.ok) throw CudaError{st.why};
    return st.api;
}

bool driver_available(std::string* why) {
    DriverState& st = driver_state();
    if (!st.ok && why) *why = st.why;
    return st.ok;
}

const NvrtcApi& nvrtc() {
    NvrtcState& st = nvrtc_state();
    if (!st.ok) throw CudaError{st.why};
    return st.api;
}

void cu_check(CUresult r, const char* what) {
    if (r == CUDA_SUCCESS) return;
    const char* s = nullptr;
    DriverState& st = driver_state();
    if (st.api.cuGetErrorString) st.api.cuGetErrorString(r, &s);
    throw CudaError{std::string(what) + ": " + (s ? s : "unknown CUDA error") + " (" + std::to_string((int)r) + ")"};
}

CUcontext device_context(int device) {
    const DriverApi& d = driver();
    static std::mutex mu;
    static CUcontext ctxs[64] = {};
    if (device < 0 || device >= 64) throw CudaError{"bad device ordinal"};
    std::lock_guard<std::mutex> lock(mu);
    if (!ctxs[device]) {
        CUdevice dev;
        cu_check(d.cuDeviceGet(&dev, device), "cuDeviceGet");
        cu_check(d.cuDevicePrimaryCtxRetain(&ctxs[device], dev), "cuDevicePrimaryCtxRetain");
    }
    return ctxs[device];
}

void use_device(int device) {
    cu_check(driver().cuCtxSetCurrent(device_context(device)), "cuCtxSetCurrent");
}

int device_count() {
    int n = 0;
    cu_check(driver().cuDeviceGetCount(&n), "cuDeviceGetCount");
    return n;
}

const NcclApi* nccl(std::string* why) {
    static NcclApi api{};
    static bool ok = false;
    static std::string err;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        std::string tried;
        void* h = open_first(names, &tried);
        if (!h) { err = "NCCL not found (tried " + tried + ")"; return; }
        bool missing = false;
#define SDE_LOAD(field)                                                            \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, SDE_STR(field))); \
        if (!api.field) missing = true;
        SDE_LOAD(ncclCommInitAll) SDE_LOAD(ncclCommDestroy) SDE_LOAD(ncclGroupStart) SDE_LOAD(ncclGroupEnd)
        SDE_LOAD(ncclAllGather) SDE_LOAD(ncclGetErrorString)
#undef SDE_LOAD
        if (missing) { err = "libnccl is missing symbols"; return; }
        ok = true;
    });
    if (!ok && why) *why = err;
    return ok ? &api : nullptr;
}

int sm_count(int device) {
    const DriverApi& d = driver();
    CUdevice dev;
    cu_check(d.cuDeviceGet(&dev, device), "cuDeviceGet");
    int n = 0;
    cu_check(d.cuDeviceGetAttribute(&n, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev), "cuDeviceGetAttribute");
    return n;
}

namespace {
// ---- compiled-kernel cache: the same model lowered with the same options compiles once per process and,
//      when $SDE_B200_CACHE names a directory, once per machine (cubins keyed by source + headers + NVRTC version).
uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
std::mutex g_cache_mu;
std::unordered_map<uint64_t, std::vector<char>> g_cubin_cache;

std::string cache_dir() {
    // $SDE_B200_CACHE when the caller names a directory (the Python package points it inside its own build/ tree)
    if (std::getenv("SDE_B200_NO_DISK_CACHE")) return "";
    if (const char* d = std::getenv("SDE_B200_CACHE")) return d;
    // default for callers that did not choose a place (the pure-C path): $XDG_CACHE_HOME/sde_b200 or ~/.cache/sde_b200
    std::string base;
    if (const char* x = std::getenv("XDG_CACHE_HOME")) base = x;
    else if (const char* h = std::getenv("HOME")) { base = std::string(h) + "/.cache"; ::mkdir(base.c_str(), 0755); }
    if (base.empty()) return "";
    return base + "/sde_b200";
}
bool read_file(const std::string& path, std::vector<char>* out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bool ok = n > 0;
    if (ok) { out->resize((size_t)n); ok = std::fread(out->data(), 1, (size_t)n, f) == (size_t)n; }
    std::fclose(f);
    return ok;
}
void write_file_atomic(const std::string& dir, const std::string& path, const std::vector<char>& data) {
    ::mkdir(dir.c_str(), 0755);                              // best effort; parents are expected to exist
    std::string tmp = path + ".tmp." + std::to_string((long)::getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return;
    bool ok = std::fwrite(data.data(), 1, data.size(), f) == data.size();
    std::fclose(f);
    if (ok) std::rename(tmp.c_str(), path.c_str()); else std::remove(tmp.c_str());
}
}  // namespace

std::vector<char> nvrtc_compile(const std::string& source, const std::string& name, std::string* log) {
    struct H { const char* name; const char* b; const char* e; };
    const H hs[] = {
        {"sde_sim_kernel.cuh", sde_blob_sim_kernel_cuh_begin, sde_blob_sim_kernel_cuh_end},
        {"sde_sim_common.cuh", sde_blob_sim_common_cuh_begin, sde_blob_sim_common_cuh_end},
        {"sde_sim_resident.cuh", sde_blob_sim_resident_cuh_begin, sde_blob_sim_resident_cuh_end},
        {"sde_sim_wide.cuh", sde_blob_sim_wide_cuh_begin, sde_blob_sim_wide_cuh_end},
        {"sde_device_rng.cuh", sde_blob_device_rng_cuh_begin, sde_blob_device_rng_cuh_end},
        {"sde_device_icdf.cuh", sde_blob_device_icdf_cuh_begin, sde_blob_device_icdf_cuh_end},
        {"sde_icdf_tables.cuh", sde_blob_icdf_tables_cuh_begin, sde_blob_icdf_tables_cuh_end},
        {"sde_expr_helpers.cuh", sde_blob_expr_helpers_cuh_begin, sde_blob_expr_helpers_cuh_end},
    };
    // The cache key covers the NVRTC version — read from the library's file name, so that a cache hit never loads NVRTC
    int vmaj = 0, vmin = 0;
    if (!nvrtc_version_from_filename(&vmaj, &vmin)) { NvrtcState& q = nvrtc_state(); vmaj = q.major; vmin = q.minor; }
    // st.global.v4.f64 (256-bit) needs the CUDA 12.9 ptxas; older NVRTC falls back to two 128-bit stores
    const bool st256_key = vmaj > 12 || (vmaj == 12 && vmin >= 9);

    uint64_t key = fnv1a(source.data(), source.size());
    for (const H& h : hs) key = fnv1a(h.b, (size_t)(h.e - h.b), key);
    const int ver[3] = {vmaj, vmin, st256_key ? 1 : 0};
    key = fnv1a(ver, sizeof ver, key);
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        auto it = g_cubin_cache.find(key);
        if (it != g_cubin_cache.end()) { if (log) *log = "(cached in memory)"; return it->second; }
    }
    const std::string dir = cache_dir();
    char hex[32];
    std::snprintf(hex, sizeof hex, "%016llx", (unsigned long long)key);
    const std::string path = dir.empty() ? "" : dir + "/" + hex + ".cubin";
    if (!path.empty()) {
        std::vector<char> cached;
        if (read_file(path, &cached) && cached.size() > 64 && std::memcmp(cached.data(), "\x7f" "ELF", 4) == 0) {
            std::lock_guard<std::mutex> lock(g_cache_mu);
            g_cubin_cache[key] = cached;
            if (log) *log = "(cached on disk: " + path + ")";
            return cached;
        }
    }

    const NvrtcApi& n = nvrtc();                             // a miss: now the compiler is needed
    NvrtcState& nst = nvrtc_state();
    const bool st256 = nst.major > 12 || (nst.major == 12 && nst.minor >= 9);
    std::vector<std::string> bodies;
    std::vector<const char*> names, ptrs;
    for (const H& h : hs) bodies.emplace_back(h.b, h.e);
    for (size_t i = 0; i < bodies.size(); ++i) { names.push_back(hs[i].name); ptrs.push_back(bodies[i].c_str()); }
    nvrtcProgram prog;
    nvrtcResult r = n.nvrtcCreateProgram(&prog, source.c_str(), name.c_str(), (int)names.size(), ptrs.data(), names.data());
    if (r != NVRTC_SUCCESS) throw CudaError{std::string("nvrtcCreateProgram: ") + n.nvrtcGetErrorString(r)};
    // -default-device: the generic lambda that instantiates the persistent kernel's item loop per step shift has no execution
    // space annotation of its own (nvcc infers it from the enclosing __global__ function, NVRTC asks for the flag)
    const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", st256 ? "-DSDE_ST256=1" : "-DSDE_ST256=0", "-default-device"};
    r = n.nvrtcCompileProgram(prog, 5, opts);
    size_t ls = 0;
    n.nvrtcGetProgramLogSize(prog, &ls);
    std::string lg(ls, '\0');
    if (ls) n.nvrtcGetProgramLog(prog, lg.data());
    if (log) *log = lg;
    if (r != NVRTC_SUCCESS) {
        n.nvrtcDestroyProgram(&prog);
        throw CudaError{std::string("NVRTC compilation failed: ") + n.nvrtcGetErrorString(r) + "\n" + lg};
    }
    size_t cs = 0;
    n.nvrtcGetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    n.nvrtcGetCUBIN(prog, cubin.data());
    n.nvrtcDestroyProgram(&prog);
    if (cs == 0) throw CudaError{"NVRTC produced an empty cubin"};
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        g_cubin_cache[key] = cubin;
    }
    if (!path.empty()) write_file_atomic(dir, path, cubin);
    return cubin;
}

void DeviceBuffer::alloc(size_t bytes) {
    release();
    if (bytes == 0) return;
    cu_check(driver().cuMemAlloc(&ptr_, bytes), "cuMemAlloc");
    bytes_ = bytes;
}
void DeviceBuffer::release() {
    if (ptr_) { driver_state().api.cuMemFree(ptr_); ptr_ = 0; bytes_ = 0; }
}
void DeviceBuffer::upload(const void* src, size_t bytes) {
    if (bytes > bytes_) alloc(bytes);
    if (bytes) cu_check(driver().cuMemcpyHtoD(ptr_, src, bytes), "cuMemcpyHtoD");
}

}  // namespace sde
