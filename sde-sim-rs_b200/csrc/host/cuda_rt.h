// cuda_rt.h — thin runtime over the CUDA driver API and NVRTC, both dlopen'ed lazily so
// that libsde_b200.so loads (and exports its symbols) on a machine without a GPU.
// There is no CPU fallback: every call fails with a descriptive error when CUDA is absent.
#pragma once
#include <cuda.h>
#include <nvrtc.h>

#include <string>
#include <vector>

namespace sde {

struct CudaError {
    std::string msg;
};

struct DriverApi {
    CUresult (*cuInit)(unsigned);
    CUresult (*cuDeviceGet)(CUdevice*, int);
    CUresult (*cuDeviceGetCount)(int*);
    CUresult (*cuDeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
    CUresult (*cuDevicePrimaryCtxRetain)(CUcontext*, CUdevice);
    CUresult (*cuCtxSetCurrent)(CUcontext);
    CUresult (*cuCtxGetCurrent)(CUcontext*);
    CUresult (*cuModuleLoadData)(CUmodule*, const void*);
    CUresult (*cuModuleUnload)(CUmodule);
    CUresult (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*cuFuncGetAttribute)(int*, CUfunction_attribute, CUfunction);
    CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
    CUresult (*cuMemAlloc)(CUdeviceptr*, size_t);
    CUresult (*cuMemFree)(CUdeviceptr);
    CUresult (*cuMemcpyHtoD)(CUdeviceptr, const void*, size_t);
    CUresult (*cuMemcpyDtoH)(void*, CUdeviceptr, size_t);
    CUresult (*cuMemcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream);
    CUresult (*cuMemcpyDtoHAsync)(void*, CUdeviceptr, size_t, CUstream);
    CUresult (*cuMemsetD8Async)(CUdeviceptr, unsigned char, size_t, CUstream);
    CUresult (*cuMemHostAlloc)(void**, size_t, unsigned);
    CUresult (*cuMemFreeHost)(void*);
    CUresult (*cuMemHostRegister)(void*, size_t, unsigned);
    CUresult (*cuMemHostUnregister)(void*);
    CUresult (*cuMemGetInfo)(size_t*, size_t*);
    CUresult (*cuStreamCreate)(CUstream*, unsigned);
    CUresult (*cuStreamDestroy)(CUstream);
    CUresult (*cuStreamSynchronize)(CUstream);
    CUresult (*cuStreamWaitEvent)(CUstream, CUevent, unsigned);
    CUresult (*cuEventCreate)(CUevent*, unsigned);
    CUresult (*cuEventDestroy)(CUevent);
    CUresult (*cuEventRecord)(CUevent, CUstream);
    CUresult (*cuEventSynchronize)(CUevent);
    CUresult (*cuEventElapsedTime)(float*, CUevent, CUevent);
    CUresult (*cuGetErrorString)(CUresult, const char**);
    CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t);
    CUresult (*cuMemcpyDtoDAsync)(CUdeviceptr, CUdeviceptr, size_t, CUstream);
    CUresult (*cuMemcpyPeerAsync)(CUdeviceptr, CUcontext, CUdeviceptr, CUcontext, size_t, CUstream);
    CUresult (*cuCtxEnablePeerAccess)(CUcontext, unsigned);
    CUresult (*cuDeviceCanAccessPeer)(int*, CUdevice, CUdevice);
    // optional (CUDA >= 12.0 drivers): tensor maps for the per-warp 2-D bulk stores of the time-tiled kernel; may be null
    CUresult (*cuTensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
};

// NCCL, dlopen'ed like the driver (libnccl.so.2: the copy PyTorch already mapped when running under it, else the
// system's).  Only what the moment all-gather needs (3 P doubles per rank).
struct NcclApi {
    int (*ncclCommInitAll)(void** comms, int ndev, const int* devlist);
    int (*ncclCommDestroy)(void* comm);
    int (*ncclGroupStart)();
    int (*ncclGroupEnd)();
    int (*ncclAllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, CUstream stream);
    const char* (*ncclGetErrorString)(int);
};
// nullptr when libnccl cannot be loaded (the callers then fall back to peer-to-peer copies)
const NcclApi* nccl(std::string* why = nullptr);

struct NvrtcApi {
    nvrtcResult (*nvrtcCreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*nvrtcDestroyProgram)(nvrtcProgram*);
    nvrtcResult (*nvrtcCompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*nvrtcGetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*nvrtcGetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*nvrtcGetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*nvrtcGetProgramLog)(nvrtcProgram, char*);
    const char* (*nvrtcGetErrorString)(nvrtcResult);
    nvrtcResult (*nvrtcVersion)(int*, int*);
};

// Throws CudaError when the library cannot be loaded.
const DriverApi& driver();
const NvrtcApi& nvrtc();
bool driver_available(std::string* why);

void cu_check(CUresult r, const char* what);

// Makes the primary context of `device` current on this thread (shared with the CUDA
// runtime / PyTorch, so their device pointers are valid here).
void use_device(int device);
CUcontext device_context(int device);
int sm_count(int device);
int device_count();

struct EmbeddedHeader { const char* name; const char* begin; const char* end; };
// Compile `source` for sm_100a with the embedded kernel headers; returns the cubin.
std::vector<char> nvrtc_compile(const std::string& source, const std::string& name, std::string* log);

// RAII device buffer
class DeviceBuffer {
  public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t bytes) { alloc(bytes); }
    ~DeviceBuffer() { release(); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : ptr_(o.ptr_), bytes_(o.bytes_) { o.ptr_ = 0; o.bytes_ = 0; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept { if (this != &o) { release(); ptr_ = o.ptr_; bytes_ = o.bytes_; o.ptr_ = 0; o.bytes_ = 0; } return *this; }
    void alloc(size_t bytes);
    void release();
    void upload(const void* src, size_t bytes);
    CUdeviceptr ptr() const { return ptr_; }
    size_t bytes() const { return bytes_; }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr_); }

  private:
    CUdeviceptr ptr_ = 0;
    size_t bytes_ = 0;
};

}  // namespace sde
