// Shared device/host helpers for the scanerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SNRF_API extern "C" __attribute__((visibility("default")))

// Thread-local last error text, returned by snrf_last_error().
void snrf_set_error(const char* fmt, ...);

#define SNRF_CHECK_ARG(cond, ...)                     \
    do {                                              \
        if (!(cond)) {                                \
            snrf_set_error(__VA_ARGS__);              \
            return (int)cudaErrorInvalidValue;        \
        }                                             \
    } while (0)

#define SNRF_RETURN_LAUNCH(name)                                              \
    do {                                                                      \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) {                                             \
            snrf_set_error("%s: %s", name, cudaGetErrorString(e__));          \
            return (int)e__;                                                  \
        }                                                                     \
        return 0;                                                             \
    } while (0)

// Stream-ordered scratch from a private pool that keeps its memory (core.cu); snrf_scratch_release() trims it.
cudaError_t snrf_scratch_alloc(void** ptr, size_t bytes, cudaStream_t s);
int snrf_scratch_release();

// Number of SMs of the current device (cached).
int snrf_sm_count();

static inline int snrf_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// small float3 helpers (own implementation; semantics noted where the reference's
// cutil_math.h conventions matter for bit-exactness)
// ---------------------------------------------------------------------------
struct f3 { float x, y, z; };
struct i3 { int x, y, z; };

__host__ __device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__host__ __device__ __forceinline__ void st3(float* p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__host__ __device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ f3 cross3(f3 a, f3 b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// reference convention: divide-by-zero yields 1e8 (cutil_math.h:924-926), sign(0)=+1 (:913-916)
__host__ __device__ __forceinline__ float safe_div(float a, float b) { return b != 0.0f ? a / b : 100000000.0f; }
__host__ __device__ __forceinline__ int sign_pos0(float a) { return a >= 0.0f ? 1 : -1; }

// Ray / axis-aligned box slab test with the reference's conventions
// (cuda/include/cuda_utils.h:564-613): near clamps at 0, far starts at 1e5,
// miss is (-1,-1).  half = half extents.
__host__ __device__ __forceinline__ float2 ray_aabb(f3 o, f3 d, f3 c, f3 half)
{
    float lo_acc = 0.0f, hi_acc = 100000.0f;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    const float cc[3] = {c.x, c.y, c.z}, hh[3] = {half.x, half.y, half.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float inv = safe_div(1.0f, dd[a]);
        float lo = (cc[a] - hh[a] - oo[a]) * inv;
        float hi = (cc[a] + hh[a] - oo[a]) * inv;
        if (hi < lo) { float t = lo; lo = hi; hi = t; }
        if (hi < lo_acc) return make_float2(-1.0f, -1.0f);
        if (lo > hi_acc) return make_float2(-1.0f, -1.0f);
        lo_acc = lo > lo_acc ? lo : lo_acc;
        hi_acc = hi < hi_acc ? hi : hi_acc;
        if (lo_acc > hi_acc) return make_float2(-1.0f, -1.0f);
    }
    return make_float2(lo_acc, hi_acc);
}
