// Multi-resolution hash-grid trilinear encode, forward and backward, sm_100a.
//
// Replaces (behaviour, not code) the reference operators
//   embedding_bg_forward_cuda / embedding_bg_backward_cuda   hashgrid/src/hashgrid_bg_kernel.cu:229-275
//   embedding_forward_cuda    / embedding_backward_cuda      hashgrid/src/hashgrid_kernel.cu:246-300
//
// Design (B200): the op is a pure HBM gather/scatter -- 8 random 8-byte reads
// per (point, level) and, in the backward, 8 vector reductions.  One thread
// owns one POINT and walks all levels, so
//   * the coordinate prologue is done once per point instead of once per
//     (point, level) thread as in the reference's grid.y = level launch,
//   * each thread writes / reads its own full 128-byte output row (every
//     sector fully used; the reference's 8-byte stores strided by 128 B are gone),
//   * grad_points is accumulated in registers -- no atomics on it at all
//     (the reference issues 48 atomics per point),
//   * the 32 lanes of a warp are 32 consecutive samples, which along a ray fall
//     in the same coarse cell: the backward merges equal-cell lanes with a
//     segmented warp reduction and issues ONE red.global.add.v2.f32 per corner
//     per run instead of 2 scalar atomics per lane.
// Levels are processed two at a time, keeping 16 independent gathers in flight
// per thread.  Grids are sized in whole waves of the 148 SMs by the launcher.
#include "hash_common.cuh"
#include <cuda_bf16.h>
using namespace hashgrid;

namespace {

constexpr int kThreads = 256;

// ------------------------------- forward -----------------------------------
// OUT_MODE 0: fp32 out[B, L, 2] (the reference operator layout)
// OUT_MODE 1: bf16 out[B, 2L]    (row-major A operand of the tensor-core MLP)
template <bool BBOX, int OUT_MODE>
__global__ void __launch_bounds__(kThreads)
hash_fwd_kernel(const float* __restrict__ points, const float2* __restrict__ table,
                const int* __restrict__ res, const float* __restrict__ corner_p,
                const float* __restrict__ size_p, void* __restrict__ out_v,
                uint32_t* __restrict__ idx_out, int B, int L, uint32_t T, int lpb)
{
    const uint32_t mask = T - 1u;
    f3 corner = mk3(0, 0, 0), size = mk3(1, 1, 1);
    if (BBOX) { corner = ld3(corner_p); size = ld3(size_p); }
    // blockIdx.y selects a group of `lpb` consecutive levels: CTAs are dispatched
    // x-fastest, so the whole GPU works on one level group at a time and the live
    // table footprint (128 MiB per level at T=2^24) stays within TLB / L2 reach.
    const int l_begin = blockIdx.y * lpb;
    const int l_end = min(L, l_begin + lpb);

    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const f3 p = ld3(points + 3 * (size_t)b);
        int l = l_begin;
        // two levels per trip: 16 independent 8-byte gathers in flight
        for (; l + 1 < l_end; l += 2) {
            const Cell c0 = locate<BBOX>(p, res + 3 * l, corner, size);
            const Cell c1 = locate<BBOX>(p, res + 3 * (l + 1), corner, size);
            uint32_t i0[8], i1[8];
            corner_idx(i0, c0, mask);
            corner_idx(i1, c1, mask);
            const float2* t0 = table + (size_t)l * T;
            const float2* t1 = table + (size_t)(l + 1) * T;
            float2 f0[8], f1[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { f0[k] = ldg2(t0 + i0[k]); f1[k] = ldg2(t1 + i1[k]); }
            float w0[8], w1[8];
            corner_w(w0, c0);
            corner_w(w1, c1);
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a0.x += w0[k] * f0[k].x; a0.y += w0[k] * f0[k].y;
                a1.x += w1[k] * f1[k].x; a1.y += w1[k] * f1[k].y;
            }
            if (OUT_MODE == 0) {
                float* o = (float*)out_v + ((size_t)b * L + l) * 2;
                if (((L | l) & 1) == 0) {
                    *reinterpret_cast<float4*>(o) = make_float4(a0.x, a0.y, a1.x, a1.y);
                } else {
                    o[0] = a0.x; o[1] = a0.y; o[2] = a1.x; o[3] = a1.y;
                }
            } else {
                __nv_bfloat162* o = (__nv_bfloat162*)out_v + ((size_t)b * L + l);
                o[0] = __floats2bfloat162_rn(a0.x, a0.y);
                o[1] = __floats2bfloat162_rn(a1.x, a1.y);
            }
            if (idx_out) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    idx_out[((size_t)b * L + l) * 8 + k] = i0[k];
                    idx_out[((size_t)b * L + l + 1) * 8 + k] = i1[k];
                }
            }
        }
        if (l < l_end) {  // odd level count tail
            const Cell c0 = locate<BBOX>(p, res + 3 * l, corner, size);
            uint32_t i0[8]; float w0[8];
            corner_idx(i0, c0, mask);
            corner_w(w0, c0);
            const float2* t0 = table + (size_t)l * T;
            float2 a0 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 f = ldg2(t0 + i0[k]);
                a0.x += w0[k] * f.x; a0.y += w0[k] * f.y;
            }
            if (OUT_MODE == 0) {
                float* o = (float*)out_v + ((size_t)b * L + l) * 2;
                o[0] = a0.x; o[1] = a0.y;
            } else {
                ((__nv_bfloat162*)out_v)[(size_t)b * L + l] = __floats2bfloat162_rn(a0.x, a0.y);
            }
            if (idx_out)
                for (int k = 0; k < 8; ++k) idx_out[((size_t)b * L + l) * 8 + k] = i0[k];
        }
    }
}

// ------------------------------- backward ----------------------------------
template <bool BBOX, bool NEED_DX>
__global__ void __launch_bounds__(kThreads)
hash_bwd_kernel(const float* __restrict__ points, const float2* __restrict__ grad_in,
                const float2* __restrict__ table, const int* __restrict__ res,
                const float* __restrict__ corner_p, const float* __restrict__ size_p,
                float* __restrict__ grad_points, float2* __restrict__ grad_table,
                int B, int L, uint32_t T, int aggregate_levels, int lpb)
{
    const uint32_t mask = T - 1u;
    const int lane = threadIdx.x & 31;
    const int l_begin = blockIdx.y * lpb;
    const int l_end = min(L, l_begin + lpb);
    const bool sole_writer = (lpb >= L);   // one CTA row owns every level of a point
    f3 corner = mk3(0, 0, 0), size = mk3(1, 1, 1);
    if (BBOX) { corner = ld3(corner_p); size = ld3(size_p); }

    // whole warps iterate together so the shuffles below are always convergent
    const int warp_base0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
    for (int wb = warp_base0; wb < B; wb += gridDim.x * blockDim.x) {
        const int b = wb + lane;
        const bool live = b < B;
        const f3 p = live ? ld3(points + 3 * (size_t)b) : mk3(0, 0, 0);
        float gpx = 0.f, gpy = 0.f, gpz = 0.f;

        for (int l = l_begin; l < l_end; ++l) {
            const Cell c = locate<BBOX>(p, res + 3 * l, corner, size);
            uint32_t idx[8]; float w[8];
            corner_idx(idx, c, mask);
            corner_w(w, c);
            const float2 g = live ? __ldg(grad_in + (size_t)b * L + l) : make_float2(0.f, 0.f);
            const float2* tl = table + (size_t)l * T;
            float2* gl = grad_table + (size_t)l * T;

            if (NEED_DX && live) {
                float2 f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = ldg2(tl + idx[k]);
                const float ax = 1.0f - c.ox, ay = 1.0f - c.oy, az = 1.0f - c.oz;
                // d(out)/d(offset) = sum_c F_c * dw_c/d(offset)  (hashgrid_bg_kernel.cu:40-77,208-218)
                const float dxw[8] = {-ay * az, -ay * c.oz, -c.oy * az, -c.oy * c.oz, ay * az, ay * c.oz, c.oy * az, c.oy * c.oz};
                const float dyw[8] = {-ax * az, -ax * c.oz, ax * az, ax * c.oz, -c.ox * az, -c.ox * c.oz, c.ox * az, c.ox * c.oz};
                const float dzw[8] = {-ax * ay, ax * ay, -ax * c.oy, ax * c.oy, -c.ox * ay, c.ox * ay, -c.ox * c.oy, c.ox * c.oy};
                float2 dx = make_float2(0.f, 0.f), dy = dx, dz = dx;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    dx.x += f[k].x * dxw[k]; dx.y += f[k].y * dxw[k];
                    dy.x += f[k].x * dyw[k]; dy.y += f[k].y * dyw[k];
                    dz.x += f[k].x * dzw[k]; dz.y += f[k].y * dzw[k];
                }
                gpx += c.sx * (g.x * dx.x + g.y * dx.y);
                gpy += c.sy * (g.x * dy.x + g.y * dy.y);
                gpz += c.sz * (g.x * dz.x + g.y * dz.y);
            }

            // gradient scatter.  Coarse levels: consecutive samples share a cell, so
            // merge runs of equal cells inside the warp first (warp-aggregated reduction).
            bool done = false;
            if (l < aggregate_levels) {
                // cell identity: 3 x 21 bits is ample for any resolution ladder in use
                const unsigned long long key = live
                    ? (((unsigned long long)(uint32_t)c.ix & 0x1fffffull) << 42) |
                      (((unsigned long long)(uint32_t)c.iy & 0x1fffffull) << 21) |
                      ((unsigned long long)(uint32_t)c.iz & 0x1fffffull)
                    : ~0ull;
                const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
                const bool head = (lane == 0) || (prev != key);
                const unsigned heads = __ballot_sync(0xffffffffu, head);
                if (__popc(heads) <= 12) {      // warp-uniform: worth aggregating
                    const int seg = __popc(heads & ((2u << lane) - 1u));
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float sx = seg_sum(w[k] * g.x, seg, lane);
                        const float sy = seg_sum(w[k] * g.y, seg, lane);
                        if (head && live) atomicAdd(gl + idx[k], make_float2(sx, sy));
                    }
                    done = true;
                }
            }
            if (!done && live) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(gl + idx[k], make_float2(w[k] * g.x, w[k] * g.y));
            }
        }
        if (NEED_DX && live) {
            float* gp = grad_points + 3 * (size_t)b;
            if (sole_writer) {   // single writer per point: accumulate semantics, no atomics
                gp[0] += gpx; gp[1] += gpy; gp[2] += gpz;
            } else {             // one reduction per level group (the reference: one per level)
                atomicAdd(gp + 0, gpx); atomicAdd(gp + 1, gpy); atomicAdd(gp + 2, gpz);
            }
        }
    }
}

inline int grid_for(int B)
{
    // whole waves of the SMs; 8 resident 256-thread CTAs per SM is the cap
    const int sms = snrf_sm_count();
    const int want = snrf_div_up(B, kThreads);
    const int wave = sms * 8;
    if (want <= wave) return want > 0 ? want : 1;
    return wave * ((want + wave - 1) / wave > 4 ? 4 : (want + wave - 1) / wave);
}

// Levels per CTA row (blockIdx.y = level group).  Measured on B200 (tools/microbench.py
// encode --sweep): walking ONE level at a time is fastest once a level's table slice
// is large (T=2^24: fwd 3.2 ms vs 5.1 ms all-levels-per-thread; the 2 MiB-page TLB
// reaches 256 MiB and a 128 MiB level fits L2), while small tables prefer a few
// levels per thread (T=2^19: 4).  Rule: as many levels as fit 16 MiB, power of two.
int g_lpb_override = 0;
inline int pick_lpb(int L, int T)
{
    if (g_lpb_override > 0) return g_lpb_override > L ? L : g_lpb_override;
    const long long level_bytes = (long long)T * 8;
    int lpb = 1;
    while (lpb * 2 <= L && (long long)lpb * 2 * level_bytes <= (16ll << 20)) lpb *= 2;
    return lpb;
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
SNRF_API void snrf_hash_set_levels_per_block(int lpb) { g_lpb_override = lpb; }

SNRF_API int snrf_hash_fwd(const float* points, const float* table, const int* res,
                           const float* corner, const float* size, void* out, unsigned* idx_out,
                           int B, int L, int T, int out_bf16, void* stream)
{
    SNRF_CHECK_ARG(B >= 0 && L > 0 && T > 0 && (T & (T - 1)) == 0, "snrf_hash_fwd: T must be a power of two (got B=%d L=%d T=%d)", B, L, T);
    SNRF_CHECK_ARG((corner == nullptr) == (size == nullptr), "snrf_hash_fwd: corner and size must be given together");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int lpb = pick_lpb(L, T);
    const dim3 grid(grid_for(B), snrf_div_up(L, lpb));
    const float2* tb = (const float2*)table;
    if (corner) {
        if (out_bf16) hash_fwd_kernel<true, 1><<<grid, kThreads, 0, s>>>(points, tb, res, corner, size, out, idx_out, B, L, (uint32_t)T, lpb);
        else          hash_fwd_kernel<true, 0><<<grid, kThreads, 0, s>>>(points, tb, res, corner, size, out, idx_out, B, L, (uint32_t)T, lpb);
    } else {
        if (out_bf16) hash_fwd_kernel<false, 1><<<grid, kThreads, 0, s>>>(points, tb, res, corner, size, out, idx_out, B, L, (uint32_t)T, lpb);
        else          hash_fwd_kernel<false, 0><<<grid, kThreads, 0, s>>>(points, tb, res, corner, size, out, idx_out, B, L, (uint32_t)T, lpb);
    }
    SNRF_RETURN_LAUNCH("snrf_hash_fwd");
}

SNRF_API int snrf_hash_bwd(const float* points, const float* grad_in, const float* table, const int* res,
                           const float* corner, const float* size, float* grad_points, float* grad_table,
                           int B, int L, int T, int aggregate_levels, void* stream)
{
    SNRF_CHECK_ARG(B >= 0 && L > 0 && T > 0 && (T & (T - 1)) == 0, "snrf_hash_bwd: T must be a power of two (got B=%d L=%d T=%d)", B, L, T);
    SNRF_CHECK_ARG((corner == nullptr) == (size == nullptr), "snrf_hash_bwd: corner and size must be given together");
    SNRF_CHECK_ARG(grad_table != nullptr, "snrf_hash_bwd: grad_table is required");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int lpb = pick_lpb(L, T);
    const dim3 grid(grid_for(B), snrf_div_up(L, lpb));
    const float2* tb = (const float2*)table;
    const float2* gi = (const float2*)grad_in;
    float2* gt = (float2*)grad_table;
    if (aggregate_levels < 0) aggregate_levels = L / 2;
    if (corner) {
        if (grad_points) hash_bwd_kernel<true, true><<<grid, kThreads, 0, s>>>(points, gi, tb, res, corner, size, grad_points, gt, B, L, (uint32_t)T, aggregate_levels, lpb);
        else             hash_bwd_kernel<true, false><<<grid, kThreads, 0, s>>>(points, gi, tb, res, corner, size, grad_points, gt, B, L, (uint32_t)T, aggregate_levels, lpb);
    } else {
        if (grad_points) hash_bwd_kernel<false, true><<<grid, kThreads, 0, s>>>(points, gi, tb, res, corner, size, grad_points, gt, B, L, (uint32_t)T, aggregate_levels, lpb);
        else             hash_bwd_kernel<false, false><<<grid, kThreads, 0, s>>>(points, gi, tb, res, corner, size, grad_points, gt, B, L, (uint32_t)T, aggregate_levels, lpb);
    }
    SNRF_RETURN_LAUNCH("snrf_hash_bwd");
}
