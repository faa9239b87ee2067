// Shared helpers for the libmfk_b200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mfk.h"

namespace mfk {

void set_error(const char *fmt, ...);

#define MFK_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            mfk::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,          \
                           cudaGetErrorString(e__));                                     \
            return MFK_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define MFK_REQUIRE(cond, ...)        \
    do {                              \
        if (!(cond)) {                \
            mfk::set_error(__VA_ARGS__); \
            return MFK_ERR_ARG;       \
        }                             \
    } while (0)

#define MFK_LAUNCH_CHECK() MFK_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

struct DeviceProps {
    int device = -1;
    int sm_count = 0;
    int cc = 0;
    size_t smem_optin = 0;
};
// cached per device; returns non-zero status on failure
int device_props(DeviceProps *out);

// Device memory of the plans and their scratch comes from a private, RETAINING memory pool: cudaMalloc / cudaFree of a
// plan's gigabyte of buffers cost 40 ms .. 900 ms from call to call (map / unmap), which made every host-buffer fit
// (mfk_*_host) vary by a second.  Same synchronisation semantics as cudaMalloc / cudaFree: pool_free waits for the
// device before the block goes back to the pool.  MFK_POOL=0 falls back to cudaMalloc / cudaFree.
cudaError_t pool_malloc(void **p, size_t bytes);
template <typename T>
static inline cudaError_t pool_malloc(T **p, size_t bytes) {
    return pool_malloc(reinterpret_cast<void **>(p), bytes);
}
cudaError_t pool_free(void *p);

// ---- device-side primitives ---------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Strong (L2-coherent, L1-bypassing) accesses for rows that migrate between workers/SMs.
__device__ __forceinline__ float4 ld_strong_f4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ float ld_strong_f(const float *p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_strong_i(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_i(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// streaming (read-once) loads of the schedule records
__device__ __forceinline__ int ld_stream_i(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream_f(const float *p) { return __ldcs(p); }

// kernels.py:6-18 (sigmoid) pieces shared by the update and prediction paths
__device__ __forceinline__ float kmf_predict_from_dot(int kernel, float mu, float bu, float bi,
                                                       float acc, float gamma, float a, float c) {
    if (kernel == MFK_KERNEL_LINEAR) return mu + bi + bu + acc;            // kernels.py:41-44
    if (kernel == MFK_KERNEL_SIGMOID) {                                     // kernels.py:72-77
        float x = mu + bu + bi + acc;
        return a + c * (1.0f / (1.0f + expf(-x)));
    }
    return a + c * expf(-gamma * acc);                                      // kernels.py:102-104 (acc = |p-q|^2)
}

}  // namespace mfk
