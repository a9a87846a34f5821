// Shared helpers for the sm_100a kernels (error reporting, RNG hash, vector loads).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "igcn_b200.h"

namespace igcn {

void set_error(const char *fmt, ...);

#define IGCN_CHECK_ARG(cond, msg)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            igcn::set_error("%s: %s", __func__, msg);               \
            return -1;                                              \
        }                                                           \
    } while (0)

#define IGCN_CHECK_LAUNCH()                                                        \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            igcn::set_error("%s: %s", __func__, cudaGetErrorString(e__));          \
            return (int)e__;                                                       \
        }                                                                          \
    } while (0)

// splitmix64 finaliser: counter-based RNG for dropout masks and the triple sampler.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// 32 uniform bits for edge (row r, column-node c) under `seed`.
__device__ __forceinline__ uint32_t edge_hash(uint64_t seed, uint32_t r, uint32_t c) {
    return (uint32_t)(mix64(seed ^ (((uint64_t)r << 32) | c)) >> 32);
}

// keep threshold: keep iff hash >= thresh  (P[keep] = 1 - p)
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
    double t = (double)p * 4294967296.0;
    if (t < 0.0) t = 0.0;
    if (t > 4294967295.0) t = 4294967295.0;
    return (uint32_t)t;
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void fma4(float4 &a, float s, const float4 &x) {
    a.x = fmaf(s, x.x, a.x); a.y = fmaf(s, x.y, a.y); a.z = fmaf(s, x.z, a.z); a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void add4(float4 &a, const float4 &x) {
    a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
}
__device__ __forceinline__ float4 scale4(const float4 &x, float s) {
    return make_float4(x.x * s, x.y * s, x.z * s, x.w * s);
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace igcn
