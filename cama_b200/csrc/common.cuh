// Host-side plumbing shared by the translation units of libcama_b200: error reporting, the
// context object, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/cama_b200.h"

namespace cama {

extern thread_local char g_last_error[512];

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CAMA_CUDA_TRY(expr)                                                                       \
    do {                                                                                          \
        cudaError_t cama_e_ = (expr);                                                             \
        if (cama_e_ != cudaSuccess)                                                               \
            return ::cama::fail(CAMA_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(cama_e_)); \
    } while (0)

#define CAMA_REQUIRE(cond, ...)                                        \
    do {                                                               \
        if (!(cond)) return ::cama::fail(CAMA_E_INVALID, __VA_ARGS__); \
    } while (0)

// counts the launch and surfaces launch-configuration errors immediately
#define CAMA_LAUNCHED(ctx)                                                                            \
    do {                                                                                              \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                                      \
        cudaError_t cama_e_ = cudaPeekAtLastError();                                                  \
        if (cama_e_ != cudaSuccess)                                                                   \
            return ::cama::fail(CAMA_E_CUDA, "kernel launch failed (%s:%d): %s", __FILE__, __LINE__,  \
                                cudaGetErrorString(cama_e_));                                         \
    } while (0)

// Makes the context's device current for the duration of a call and restores the caller's.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace cama

struct cama_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    std::atomic<uint64_t> launches{0};
    // phase profiling (cama_ctx_profile_*): (CAMA_CLIP_PHASES + 1) events per recorded call
    std::vector<cudaEvent_t> prof_events;
    int prof_capacity = 0;
    int prof_calls = 0;
    // frame-group pipeline of cama_clip_render (created on first use): the sort and raster lanes, and the
    // events that order the lanes (2 per group + 2)
    cudaStream_t pipe_streams[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> pipe_events;
    std::vector<cudaEvent_t> fetch_events;     // cama_overlay_fetch_apply: one per record slice
};
