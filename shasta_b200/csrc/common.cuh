// Shared host-side helpers for the CUDA library: error propagation to the C ABI, RAII device
// buffers, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <algorithm>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/shasta_b200.h"

namespace shb {

struct Error : public std::runtime_error {
    shb_status status;
    Error(shb_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};

void setLastError(const std::string& message);

#define SHB_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t shbErr_ = (call);                                                        \
        if(shbErr_ != cudaSuccess) {                                                         \
            throw ::shb::Error(shbErr_ == cudaErrorMemoryAllocation ? SHB_ERR_OOM : SHB_ERR_CUDA, \
                std::string(#call) + " failed at " + __FILE__ + ":" + std::to_string(__LINE__) + ": " + \
                cudaGetErrorString(shbErr_));                                                \
        }                                                                                    \
    } while(0)

#define SHB_CHECK_LAUNCH() SHB_CUDA(cudaGetLastError())

#define SHB_REQUIRE(cond, status, message)                                                   \
    do { if(!(cond)) throw ::shb::Error(status, message); } while(0)

// Counts kernels launched through SHB_LAUNCH (bench.py reports it as gpu_launches).
extern thread_local uint64_t g_launchCount;

#define SHB_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        SHB_CHECK_LAUNCH();                                                                  \
        ++::shb::g_launchCount;                                                              \
    } while(0)

// Grow-only typed device buffer.
template<class T> class DeviceBuffer {
public:
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }

    T* get() const { return ptr_; }
    uint64_t capacity() const { return capacity_; }

    // Ensure room for n elements. Contents are NOT preserved unless keep is true.
    void reserve(uint64_t n, bool keep = false, cudaStream_t stream = 0)
    {
        if(n <= capacity_) return;
        // Grow geometrically to amortise repeated appends.
        uint64_t newCap = n;
        if(keep && capacity_) newCap = (n > capacity_ + capacity_/2) ? n : capacity_ + capacity_/2;
        if(!keep) release();            // nothing to preserve: return the old block first (the buffers can be tens of GB)
        T* p = nullptr;
        SHB_CUDA(cudaMalloc(&p, (newCap ? newCap : 1) * sizeof(T)));
        if(keep && ptr_ && capacity_) {
            SHB_CUDA(cudaMemcpyAsync(p, ptr_, capacity_ * sizeof(T), cudaMemcpyDeviceToDevice, stream));
            SHB_CUDA(cudaStreamSynchronize(stream));
        }
        release();
        ptr_ = p;
        capacity_ = newCap;
    }
    void release()
    {
        if(ptr_) cudaFree(ptr_);
        ptr_ = nullptr;
        capacity_ = 0;
    }
    void swap(DeviceBuffer& o) { std::swap(ptr_, o.ptr_); std::swap(capacity_, o.capacity_); }

private:
    T* ptr_ = nullptr;
    uint64_t capacity_ = 0;
};

inline uint32_t ceilDiv(uint64_t a, uint64_t b) { return uint32_t((a + b - 1) / b); }

} // namespace shb
