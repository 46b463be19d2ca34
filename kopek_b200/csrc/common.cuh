// Shared host-side helpers of libkopek_cuda.so: thread-local error text and
// CUDA call checking.  No exception ever crosses the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kopek_cuda.h"

namespace kopek {

char *err_buf();                      // thread-local, 512 bytes
kopek_status fail(kopek_status code, const char *fmt, ...);

#define KOPEK_CUDA(call_)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = (call_);                                                           \
        if (e_ != cudaSuccess)                                                              \
            return ::kopek::fail(KOPEK_ERR_CUDA, "%s failed: %s (%s:%d)", #call_,           \
                                 cudaGetErrorString(e_), __FILE__, __LINE__);               \
    } while (0)

// RAII device switch: restores the caller's current device on scope exit.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace kopek
