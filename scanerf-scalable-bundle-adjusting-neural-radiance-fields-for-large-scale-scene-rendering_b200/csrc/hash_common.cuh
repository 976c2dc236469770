// Device helpers shared by the hash-grid encode kernels (hash_encode.cu: the reference-shaped
// operators; field_encode.cu: the fused sample -> contraction -> encode training path).
#pragma once
#include "common.cuh"

namespace hashgrid {

__device__ __forceinline__ uint32_t hash3(int x, int y, int z, uint32_t mask)
{
    // three-prime spatial hash, uint32 wrap-around (hashgrid_bg_kernel.cu:14-24)
    return (((uint32_t)x) ^ ((uint32_t)y * 2654435761u) ^ ((uint32_t)z * 805459861u)) & mask;
}

struct Cell {
    int ix, iy, iz;      // bottom-left vertex
    float ox, oy, oz;    // trilinear offsets
    float sx, sy, sz;    // d(offset)/d(point)
};

// contracted-space variant: points already in [-2,2]^3
__device__ __forceinline__ Cell locate_bg(f3 p, const int* __restrict__ res)
{
    Cell c;
    const float rx = (float)(res[0] - 1), ry = (float)(res[1] - 1), rz = (float)(res[2] - 1);
    // (p + 2) / 4 in the reference is add, then multiply by the exact reciprocal 0.25
    const float vx = ((p.x + 2.0f) * 0.25f) * rx;
    const float vy = ((p.y + 2.0f) * 0.25f) * ry;
    const float vz = ((p.z + 2.0f) * 0.25f) * rz;
    c.ix = (int)vx; c.iy = (int)vy; c.iz = (int)vz;
    c.ox = vx - (float)c.ix; c.oy = vy - (float)c.iy; c.oz = vz - (float)c.iz;
    c.sx = rx * 0.25f; c.sy = ry * 0.25f; c.sz = rz * 0.25f;
    return c;
}

// world-space variant: clamp into the box, IEEE divides as in hashgrid_kernel.cu:126-141
__device__ __forceinline__ Cell locate_bbox(f3 p, const int* __restrict__ res, f3 corner, f3 size)
{
    Cell c;
    const float px = fmaxf(corner.x, fminf(p.x, corner.x + size.x));
    const float py = fmaxf(corner.y, fminf(p.y, corner.y + size.y));
    const float pz = fmaxf(corner.z, fminf(p.z, corner.z + size.z));
    const float gx = size.x / (float)(res[0] - 1);
    const float gy = size.y / (float)(res[1] - 1);
    const float gz = size.z / (float)(res[2] - 1);
    c.ix = (int)((px - corner.x) / gx);
    c.iy = (int)((py - corner.y) / gy);
    c.iz = (int)((pz - corner.z) / gz);
    c.ox = (px - ((float)c.ix * gx + corner.x)) / gx;
    c.oy = (py - ((float)c.iy * gy + corner.y)) / gy;
    c.oz = (pz - ((float)c.iz * gz + corner.z)) / gz;
    c.sx = 1.0f / gx; c.sy = 1.0f / gy; c.sz = 1.0f / gz;
    return c;
}

template <bool BBOX>
__device__ __forceinline__ Cell locate(f3 p, const int* __restrict__ res, f3 corner, f3 size)
{
    if (BBOX) return locate_bbox(p, res, corner, size);
    return locate_bg(p, res);
}

// corner order c = 4*dx + 2*dy + dz (hashgrid_bg_kernel.cu:79-90)
__device__ __forceinline__ void corner_idx(uint32_t idx[8], const Cell& c, uint32_t mask)
{
#pragma unroll
    for (int k = 0; k < 8; ++k)
        idx[k] = hash3(c.ix + ((k >> 2) & 1), c.iy + ((k >> 1) & 1), c.iz + (k & 1), mask);
}

__device__ __forceinline__ void corner_w(float w[8], const Cell& c)
{
    const float ax = 1.0f - c.ox, ay = 1.0f - c.oy, az = 1.0f - c.oz;
    w[0] = ax * ay * az;     w[1] = ax * ay * c.oz;
    w[2] = ax * c.oy * az;   w[3] = ax * c.oy * c.oz;
    w[4] = c.ox * ay * az;   w[5] = c.ox * ay * c.oz;
    w[6] = c.ox * c.oy * az; w[7] = c.ox * c.oy * c.oz;
}

__device__ __forceinline__ float2 ldg2(const float2* p) { return __ldg(p); }


// Segmented warp sum over runs of equal keys in consecutive lanes.  `seg` is the
// run id of the lane (monotone), returns the run total in the run's first lane.
__device__ __forceinline__ float seg_sum(float v, int seg, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float o = __shfl_down_sync(0xffffffffu, v, off);
        const int so = __shfl_down_sync(0xffffffffu, seg, off);
        if (lane + off < 32 && so == seg) v += o;
    }
    return v;
}

}  // namespace hashgrid
