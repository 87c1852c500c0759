// Host-side helpers shared by the C-ABI translation units (api.cu: one shard; group.cu: a row-sharded group of shards).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <shared_mutex>

#include "../../include/fr_index.h"

namespace fr {

// Records the thread-local message fr_last_error() returns and hands back `code` (api.cu).
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
// FR_OK when `device` is an sm_100 device (asks the runtime once per device); optionally its SM count.
int check_device(int device, int *sm_count);

// CUDA-graph work (stream capture, instantiation, graph launch) and host waits on events of other callers never run at
// the same time.  Measured on B200 / driver 580 (scripts/stress_concurrent.py): seven threads searching one shard, each
// holding the shard lock only while it enqueues, crash inside cuGraphLaunch within seconds when one thread replays or
// captures a search graph while others sit in cudaEventSynchronize on events of the same stream; with the waits
// serialised, or without graphs, nothing happens.  So graph work takes this lock exclusively -- by try_lock only: when
// somebody is waiting the call simply runs its kernels eagerly, graphs are a latency optimisation for the lone caller
// -- and host waits outside an object's own lock take it shared.
std::shared_mutex &graph_wait_mutex();

// cudaGetDeviceProperties costs milliseconds; the answer never changes, so ask once per device.
struct DevInfo {
    int state = 0;  // 0 unknown, 1 ok, -1 not sm_100
    int sm_count = 0, major = 0, minor = 0;
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// grow-only device / pinned buffers
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t need(size_t n) {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = n + n / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = n;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t need(size_t n) {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        // portable: a group's staging block is read by every device of the group
        cudaError_t e = cudaHostAlloc(&p, n + n / 4, cudaHostAllocPortable);
        if (e == cudaSuccess) bytes = n + n / 4;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

}  // namespace fr

#define FR_CUDA(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess)                                                                                     \
            return fr::fail(_e == cudaErrorMemoryAllocation ? FR_ENOMEM : FR_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                                           \
    } while (0)
