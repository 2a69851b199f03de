// Shared helpers for libgpde_b200 (sm_100a).  Internal header, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/gpde_b200.h"

namespace gpde {

extern thread_local char g_last_error[512];

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

#define GPDE_CUDA_OK(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return gpde::fail(GPDE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

// RAII device switch: every entry point runs on the plan's device and restores the caller's.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
inline cudaError_t upload(T **dst, const std::vector<T> &src) {
    size_t bytes = (src.size() ? src.size() : 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void **)dst, bytes);
    if (e != cudaSuccess) return e;
    if (src.size()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

inline int sm_count(int device) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    return v;
}

}  // namespace gpde
