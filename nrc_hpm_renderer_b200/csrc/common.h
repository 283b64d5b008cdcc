// Shared host-side plumbing of libnrchpm_b200: error convention of the C ABI (include/nrc_hpm_b200.h), CUDA error
// checks, kernel-launch accounting.  The reference reports errors by throwing std::runtime_error
// (src/Log.cpp:16-21, include/engine/cuda_common.hpp:14); nothing may be thrown across a C ABI, so every entry point
// is wrapped in guard() which turns exceptions into return codes + nrchpm_last_error().
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>
#include "../../include/nrc_hpm_b200.h"

namespace nrchpm {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define NRCHPM_CUDA(expr)                                                                                             \
    do {                                                                                                              \
        cudaError_t e__ = (expr);                                                                                     \
        if (e__ != cudaSuccess)                                                                                       \
            throw ::nrchpm::Error(NRCHPM_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(e__));       \
    } while (0)

#define NRCHPM_REQUIRE(cond, msg)                                                 \
    do {                                                                          \
        if (!(cond)) throw ::nrchpm::Error(NRCHPM_ERR_INVALID, std::string(msg)); \
    } while (0)

template <class F>
int guard(F&& f) {
    try {
        f();
        return NRCHPM_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return NRCHPM_ERR_INVALID;
    } catch (...) {
        set_last_error("unknown error");
        return NRCHPM_ERR_INVALID;
    }
}

// check the launch that was just issued
inline void check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(NRCHPM_ERR_CUDA, std::string(what) + " launch failed: " + cudaGetErrorString(e));
    count_launch();
}

template <class T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }
    void allocate(size_t n) {
        release();
        if (n == 0) return;
        NRCHPM_CUDA(cudaMalloc((void**)&ptr, n * sizeof(T)));
        count = n;
    }
    void ensure(size_t n) { if (n > count) allocate(n); }
    void zero(cudaStream_t s = nullptr) { if (ptr) NRCHPM_CUDA(cudaMemsetAsync(ptr, 0, count * sizeof(T), s)); }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; count = 0; }
    size_t bytes() const { return count * sizeof(T); }
};

}  // namespace nrchpm
