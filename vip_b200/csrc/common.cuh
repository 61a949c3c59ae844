// Shared helpers for the vip_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>

namespace vb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// last-error string returned by vb_last_error() (C ABI, include/vip_b200.h)
void set_error(const char* fmt, ...);

#define VB_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            vb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,           \
                          cudaGetErrorString(_e));                                       \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

#define VB_CHECK_LAUNCH()                                                                \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            vb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                          cudaGetErrorString(_e));                                       \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

#define VB_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            vb::set_error(__VA_ARGS__);                                                  \
            return -2;                                                                   \
        }                                                                                \
    } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit global load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace vb
