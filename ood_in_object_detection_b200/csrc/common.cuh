// Shared helpers for the oodb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/oodb200.h"

namespace oodb200 {

// thread-local error message behind oodb200_last_error()
char* error_buffer();
inline void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
}
inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return OODB200_ERR_CUDA;
    }
    return OODB200_OK;
}
#define OODB200_REQUIRE(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            ::oodb200::set_error(__VA_ARGS__);     \
            return OODB200_ERR_INVALID;            \
        }                                          \
    } while (0)

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// read-only global load; feature maps are read (at most) a few times, centroids many times
__device__ __forceinline__ float ldg_f32(const float* p) { return __ldg(p); }

// Block-wide sum for blockDim.x == 32 * NW threads; `scratch` holds NW floats. All threads get the result.
// Fixed combination order -> deterministic.
template <int NW>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();   // scratch may still be read from a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += scratch[i];
    return t;
}

}  // namespace oodb200
