// Shared device/host helpers for the rgbd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define RGBD_OK 0
#define RGBD_ERR_ARG 1
#define RGBD_ERR_CUDA 2
#define RGBD_ERR_UNSUPPORTED 3

void rgbd_set_error(const char* fmt, ...);

#define RGBD_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            rgbd_set_error(__VA_ARGS__);          \
            return RGBD_ERR_ARG;                  \
        }                                         \
    } while (0)

#define RGBD_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            rgbd_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return RGBD_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define RGBD_CHECK_LAUNCH() RGBD_CHECK_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Per-device launch state.  SM count, opt-in shared memory and cudaFuncSetAttribute are properties of the CURRENT device:
// a process that drives several GPUs (one module per device, tests touching cuda:1 after cuda:0) must see each device's own.
#define RGBD_MAX_DEVICES 64
struct RgbdDeviceInfo { int device; int num_sms; int max_smem; };
int rgbd_device_info(RgbdDeviceInfo* info);      // api.cu: current device, cached per device, thread-safe

#ifdef __cplusplus
#include <mutex>
// Runs `body` once per device (per call site), under a lock; `body` may `return` an error code.
#define RGBD_ONCE_PER_DEVICE(dev, body)                                   \
    do {                                                                  \
        static std::mutex _rgbd_m;                                        \
        static bool _rgbd_done[RGBD_MAX_DEVICES] = {};                    \
        std::lock_guard<std::mutex> _rgbd_g(_rgbd_m);                     \
        if (!_rgbd_done[(dev)]) {                                         \
            body;                                                         \
            _rgbd_done[(dev)] = true;                                     \
        }                                                                 \
    } while (0)
#endif

#ifdef __CUDACC__
// Order-preserving float <-> uint32 encoding (for atomicMin/atomicMax on floats of any sign).
__device__ __forceinline__ uint32_t f32_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
#endif
