// Cell walk (3-D DDA) over an occupancy grid, shared by the training sampler (rays.cu) and the
// multi-tile render sampler (render.cu).
#pragma once
#include "common.cuh"

// Cell walk over a (2^lx, 2^ly, 2^lz) grid with the reference's conventions
// (cuda/include/dda.h:206-268): start cell clamped into the grid, tie rule
// x if (tx<ty && tx<=tz), y if (ty<tz && ty<=tx), else z.
struct Walk {
    int cx, cy, cz, sx, sy, sz, mx, my, mz;
    float tmx, tmy, tmz, tdx, tdy, tdz, t0, t1;
    int nx, ny, nz;

    __device__ __forceinline__ void init(f3 o, f3 d, float2 tb, int rx, int ry, int rz, f3 cell)
    {
        nx = rx; ny = ry; nz = rz;
        o = o + tb.x * d;
        cx = min(max((int)(o.x / cell.x), 0), nx - 1);
        cy = min(max((int)(o.y / cell.y), 0), ny - 1);
        cz = min(max((int)(o.z / cell.z), 0), nz - 1);
        sx = sign_pos0(d.x); sy = sign_pos0(d.y); sz = sign_pos0(d.z);
        float bx = (float)(cx + sx) * cell.x, by = (float)(cy + sy) * cell.y, bz = (float)(cz + sz) * cell.z;
        if (sx < 0) bx += cell.x;
        if (sy < 0) by += cell.y;
        if (sz < 0) bz += cell.z;
        t0 = tb.x; t1 = tb.y;
        tmx = fmaxf(safe_div(bx - o.x, d.x), 0.0f) + t0;
        tmy = fmaxf(safe_div(by - o.y, d.y), 0.0f) + t0;
        tmz = fmaxf(safe_div(bz - o.z, d.z), 0.0f) + t0;
        tdx = fabsf(safe_div(cell.x, d.x));
        tdy = fabsf(safe_div(cell.y, d.y));
        tdz = fabsf(safe_div(cell.z, d.z));
    }
    __device__ __forceinline__ void next()
    {
        mx = (tmx < tmy) & (tmx <= tmz);
        my = (tmy < tmz) & (tmy <= tmx);
        mz = !(mx | my);
        t1 = mx ? tmx : (my ? tmy : tmz);
    }
    __device__ __forceinline__ void step()
    {
        t0 = t1;
        tmx += (float)mx * tdx; tmy += (float)my * tdy; tmz += (float)mz * tdz;
        cx += mx * sx; cy += my * sy; cz += mz * sz;
    }
    __device__ __forceinline__ bool done() const
    {
        return cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz ||
               (tmx <= 0 && tmy <= 0 && tmz <= 0);
    }
};

