// Common definitions for the ddl (dedalus-b200) CUDA library.
//
// The kernels are written as "block bodies": a sequence of data-parallel phases over work
// items separated by barriers.  Under nvcc a phase is a blockDim-strided loop and the barrier
// is __syncthreads(); under DDL_HOST_EMUL (plain g++, used ONLY by tests/host/ to check index
// logic in the GPU-less build container) one host thread walks every item of every phase in
// order and the barrier is a no-op.  The emulation build is test infrastructure: the Python
// package never loads it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "fft_core.cuh"

#if defined(__CUDACC__) && !defined(DDL_HOST_EMUL)
#define DDL_DEVICE_BUILD 1
#define DDL_FOR_ITEMS(i, count) for (int i = threadIdx.x; i < (count); i += blockDim.x)
#define DDL_SYNC() __syncthreads()
#define DDL_BODY __device__ __forceinline__
typedef cudaStream_t ddl_stream_t;
#else
#define DDL_DEVICE_BUILD 0
#define DDL_FOR_ITEMS(i, count) for (int i = 0; i < (count); ++i)
#define DDL_SYNC() ((void)0)
#define DDL_BODY inline
typedef void* ddl_stream_t;
#endif

#define DDL_MAXF 20           // max fields (pencil groups) a fused tile carries
#define DDL_MAXC 8            // max state components per stage-update launch

namespace ddl {

// thread-local last error (include/ddl.h: ddl_last_error)
void set_error(const char* fmt, ...);

#if DDL_DEVICE_BUILD
#define DDL_CUDA_CHECK(expr)                                                               \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ddl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -2;                                                                     \
        }                                                                                  \
    } while (0)
#endif

#if DDL_DEVICE_BUILD
// One-time, per-device set-up of a kernel instantiation (function attributes, occupancy queries): a function-local static
// of this type replaces a plain `static bool done`, which is neither thread-safe nor valid for a second device.
#define DDL_MAXDEV 64
struct DeviceOnce {
    std::mutex mu;
    int val[DDL_MAXDEV] = {};
    // value cached for the current device (> 0), else init() is run under the lock: it returns the value (> 0) or an error (< 0)
    template <class F> int get(F&& init) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DDL_MAXDEV) { set_error("DeviceOnce: no current device"); return -2; }
        std::lock_guard<std::mutex> guard(mu);
        if (val[dev] <= 0) val[dev] = init();
        return val[dev];
    }
};
#endif

// launch accounting + optional per-launch CUDA-event timing (ddl_profile_* in include/ddl.h)
void prof_begin(const char* name, ddl_stream_t stream);
void prof_end(ddl_stream_t stream);

// small device buffers owned by a plan (twiddles, index tables)
void* dev_alloc(size_t bytes);
void dev_free(void* p);
int dev_upload(void* dst, const void* src, size_t bytes);

}  // namespace ddl
