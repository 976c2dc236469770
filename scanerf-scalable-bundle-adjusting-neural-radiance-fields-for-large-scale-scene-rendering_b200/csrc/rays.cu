// Ray generation (+ analytic bundle-adjustment backward), ray/box tests and the
// occupancy-grid bounded sample placement, sm_100a.
//
// Replaces (behaviour, not code) the reference operators
//   compute_ray_forward / compute_ray_backward      cuda/compute_ray_kernel.cu:95-136
//   ray_aabb_intersection / _v2                     cuda/helper_kernel.cu:130-197
//   sample_points_grid                              cuda/helper_kernel.cu:645-671
//   background_sampling_cuda, sample_insideout_block cuda/sample_kernel.cu:46-130
//
// All of these are one-thread-per-ray SIMT kernels bounded by HBM streaming of
// the per-ray rows; the interesting work is in keeping the output rows
// coalesced (z_vals / dists rows are staged through shared memory and written
// out by the whole warp) and in reducing the pose-gradient atomics.
#include "common.cuh"
#include "walk.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid1d(long long n, int threads = kThreads)
{
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)snrf_sm_count() * 32;
    if (g > cap) g = cap;
    return g > 0 ? (int)g : 1;
}

// ----------------------------- ray generation -------------------------------
// pinhole ray through the pixel centre (+0.5), cuda/include/cuda_utils.h:143-155
__global__ void __launch_bounds__(kThreads)
ray_fwd_kernel(float* __restrict__ rays_o, float* __restrict__ rays_d, const float* __restrict__ Ks,
               const float* __restrict__ C2Ws, const int* __restrict__ locs, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const int v = locs[3 * i], px = locs[3 * i + 1], py = locs[3 * i + 2];
        const float* K = Ks + 9 * (size_t)v;
        const float* M = C2Ws + 12 * (size_t)v;
        const float x = (1.0f * px + 0.5f - K[2]) / K[0];
        const float y = (1.0f * py + 0.5f - K[5]) / K[4];
        rays_d[3 * (size_t)i + 0] = M[0] * x + M[1] * y + M[2];
        rays_d[3 * (size_t)i + 1] = M[4] * x + M[5] * y + M[6];
        rays_d[3 * (size_t)i + 2] = M[8] * x + M[9] * y + M[10];
        rays_o[3 * (size_t)i + 0] = M[3];
        rays_o[3 * (size_t)i + 1] = M[7];
        rays_o[3 * (size_t)i + 2] = M[11];
    }
}

// dL/dC2W[v] += [ g_d (x) (x, y, 1) | g_o ]  (row-major 3x4).
// ref_index_bug = 1 reproduces the reference's read of the incoming gradients at
// index view_idx instead of the ray index (cuda/compute_ray_kernel.cu:71-72); the
// default (0) is the mathematically correct gradient.
// Rays of one camera are contiguous in the batch, so a warp usually holds a single
// view: reduce the 12 terms across the warp and issue 12 atomics per warp, not per ray.
__global__ void __launch_bounds__(kThreads)
ray_bwd_kernel(const float* __restrict__ g_o, const float* __restrict__ g_d, const float* __restrict__ Ks,
               float* __restrict__ grad_C2Ws, const int* __restrict__ locs, int B, int ref_index_bug)
{
    const int lane = threadIdx.x & 31;
    const int base0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
    for (int wb = base0; wb < B; wb += gridDim.x * blockDim.x) {
        const int i = wb + lane;
        const bool live = i < B;
        int v = -1;
        float t[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) t[k] = 0.f;
        if (live) {
            v = locs[3 * i];
            const int px = locs[3 * i + 1], py = locs[3 * i + 2];
            const float* K = Ks + 9 * (size_t)v;
            const float x = (1.0f * px + 0.5f - K[2]) / K[0];
            const float y = (1.0f * py + 0.5f - K[5]) / K[4];
            const size_t gi = ref_index_bug ? (size_t)v : (size_t)i;
            const f3 go = ld3(g_o + 3 * gi), gd = ld3(g_d + 3 * gi);
            t[0] = gd.x * x; t[1] = gd.x * y; t[2] = gd.x;  t[3] = go.x;
            t[4] = gd.y * x; t[5] = gd.y * y; t[6] = gd.y;  t[7] = go.y;
            t[8] = gd.z * x; t[9] = gd.z * y; t[10] = gd.z; t[11] = go.z;
        }
        const int v0 = __shfl_sync(0xffffffffu, v, 0);
        const bool uniform = __all_sync(0xffffffffu, v == v0 || !live);
        if (uniform) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                float s = t[k];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                t[k] = s;
            }
            if (lane < 12 && v0 >= 0) {
                float mine = 0.f;
#pragma unroll
                for (int k = 0; k < 12; ++k) if (lane == k) mine = t[k];
                atomicAdd(grad_C2Ws + 12 * (size_t)v0 + lane, mine);
            }
        } else if (live) {
#pragma unroll
            for (int k = 0; k < 12; ++k) atomicAdd(grad_C2Ws + 12 * (size_t)v + k, t[k]);
        }
    }
}

// ----------------------------- ray / box ------------------------------------
__global__ void __launch_bounds__(kThreads)
ray_aabb_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                const float* __restrict__ centers, const float* __restrict__ sizes,
                float2* __restrict__ bounds, int B, int K)
{
    const long long n = (long long)B * K;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / K), k = (int)(t % K);
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        const f3 c = ld3(centers + 3 * k);
        const f3 h = ld3(sizes + 3 * k) * 0.5f;   // size / 2.0f == size * (1/2) exactly
        bounds[t] = ray_aabb(o, d, c, h);
    }
}


// ----------------------------- background inverse-depth samples -----------------------------
// HashGrid.inverse_z_sampling + invalid_sampling_underground (hashgrid/__init__.py:287-293, 305-337) as one kernel;
// the reference evaluates it as ~25 small torch launches per step.  Every operation is rounded separately, in the order
// torch evaluates the expression  1 / (1 / (far + 1e-6) * (1 - t) + 1 / 1e6 * t), so the depths are bit-identical to it.
// t_lin = torch.linspace(0, 1, S) (an input: its construction is torch's).  One thread per sample.
__global__ void __launch_bounds__(kThreads)
bg_inverse_z_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ center_p,
                    const float* __restrict__ size_p, const float* __restrict__ t_lin, float* __restrict__ z_vals,
                    float* __restrict__ dists, unsigned char* __restrict__ valid, int B, int S, int invalid_underground)
{
    const f3 c = ld3(center_p), h = ld3(size_p) * 0.5f;
    const float floor_y = c.y - h.y;
    const long long total = (long long)B * S;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(n / S), k = (int)(n - (long long)i * S);
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        const float2 tb = ray_aabb(o, d, c, h);
        if (k == 0 && valid != nullptr) {
            bool ok = true;
            if (invalid_underground) ok = !(fabsf(__fadd_rn(o.y, __fmul_rn(tb.y, d.y)) - floor_y) < 0.0001f);
            valid[i] = ok ? 1 : 0;
        }
        const float far = (tb.x == -1.0f || tb.y == -1.0f) ? 0.1f : tb.y;
        const float inv_far = __fdiv_rn(1.0f, __fadd_rn(far, 1e-6f));
        auto depth = [&](int kk) {
            const float t = t_lin[kk];
            return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(inv_far, __fsub_rn(1.0f, t)), __fmul_rn(1e-6f, t)));
        };
        const float z = depth(k);
        z_vals[n] = z;
        dists[n] = (k == S - 1) ? 1e-6f : __fsub_rn(depth(k + 1), z);
    }
}

// ----------------------------- occupancy DDA --------------------------------
// One thread per ray, two passes of the walk (total occupied length, then
// proportional placement) exactly like cuda/helper_kernel.cu:539-615.  Output rows
// (S floats each, S up to 256) are written straight to global memory by the owning
// thread; rows that see nothing keep the caller's fill value.
__global__ void __launch_bounds__(128)
sample_grid_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int S,
                   float* __restrict__ z_vals, float* __restrict__ dists,
                   const float* __restrict__ corner_p, const float* __restrict__ size_p,
                   const unsigned char* __restrict__ occ, const int* __restrict__ log2dim,
                   int* __restrict__ counts, int B)
{
    const f3 corner = ld3(corner_p), size = ld3(size_p);
    const int lx = log2dim[0], ly = log2dim[1], lz = log2dim[2];
    const int rx = 1 << lx, ry = 1 << ly, rz = 1 << lz;
    const f3 cell = mk3(size.x / (float)rx, size.y / (float)ry, size.z / (float)rz);
    const f3 half = size * 0.5f;
    const f3 center = corner + half;

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        const float2 tb = ray_aabb(o, d, center, half);
        if (counts) counts[i] = 0;
        if (tb.x == -1.0f) continue;
        Walk w;
        w.init(o - corner, d, tb, rx, ry, rz, cell);
        float total = 0.0f;
        int count = 0;
        while (!w.done()) {
            w.next();
            const uint32_t n = ((uint32_t)w.cx << (ly + lz)) | ((uint32_t)w.cy << lz) | (uint32_t)w.cz;
            if (occ[n]) {
                const float len = w.t1 - w.t0;
                if (len > 0) { total += len; ++count; }
            }
            w.step();
        }
        if (counts) counts[i] = count;
        if (count == 0) continue;
        w.init(o - corner, d, tb, rx, ry, rz, cell);
        int left = S, seen = 0;
        float* zr = z_vals + (size_t)i * S;
        float* dr = dists + (size_t)i * S;
        while (!w.done()) {
            w.next();
            const uint32_t n = ((uint32_t)w.cx << (ly + lz)) | ((uint32_t)w.cy << lz) | (uint32_t)w.cz;
            if (occ[n]) {
                const float len = w.t1 - w.t0;
                if (len > 0) {
                    int num = min(max((int)(S * len / total), 1), left);
                    if (seen == count - 1) num = left;
                    const float interval = (w.t1 - w.t0) / num;   // uniform_sample_bound_v3
                    const int at = S - left;
                    for (int k = 0; k < num; ++k) {
                        zr[at + k] = w.t0 + k * interval;
                        dr[at + k] = interval;
                    }
                    left -= num;
                    ++seen;
                }
            }
            w.step();
        }
    }
}

// ----------------------------- simple samplers ------------------------------
__global__ void __launch_bounds__(kThreads)
bg_sampling_kernel(const float* __restrict__ starts, const float* __restrict__ bg_depth,
                   float* __restrict__ z_vals, int S, float range, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const float near = fmaxf(starts[i] + 0.00001f, bg_depth[i] - range * 0.5f);
        const float far = near + range;
        const float interval = (far - near) / (S - 1);     // uniform_sample_bound, cuda_utils.h:77-87
        float* zr = z_vals + (size_t)i * S;
        for (int k = 0; k < S; ++k) zr[k] = near + k * interval;
    }
}

__global__ void __launch_bounds__(kThreads)
insideout_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int S, int Sbg,
                 const float* __restrict__ center_p, const float* __restrict__ size_p, int B, float far,
                 float* __restrict__ z_vals, float* __restrict__ z_bg, int* __restrict__ miss_flag)
{
    const f3 center = ld3(center_p), half = ld3(size_p) * 0.5f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const float2 tb = ray_aabb(ld3(rays_o + 3 * (size_t)i), ld3(rays_d + 3 * (size_t)i), center, half);
        if (tb.x == -1.0f || tb.y == -1.0f) {   // the reference device-asserts here (sample_kernel.cu:91)
            if (miss_flag) atomicExch(miss_flag, 1);
            continue;
        }
        const float interval = (tb.y - tb.x) / (S - 1);
        float* zr = z_vals + (size_t)i * S;
        for (int k = 0; k < S; ++k) zr[k] = tb.x + k * interval;
        // inverse-z between the exit and `far`, cuda_utils.h:61-75
        const float inv_near = 1.0f / tb.y, inv_far = 1.0f / far;
        const float inv_bound = inv_far - inv_near;
        const float step = 1.0f / (Sbg - 1);
        float* zb = z_bg + (size_t)i * Sbg;
        for (int k = 0; k < Sbg; ++k) zb[k] = 1.0f / (step * k * inv_bound + inv_near);
    }
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
SNRF_API int snrf_compute_ray_fwd(float* rays_o, float* rays_d, const float* Ks, const float* C2Ws,
                                  const int* locs, int B, void* stream)
{
    if (B <= 0) return 0;
    ray_fwd_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, Ks, C2Ws, locs, B);
    SNRF_RETURN_LAUNCH("snrf_compute_ray_fwd");
}

SNRF_API int snrf_compute_ray_bwd(const float* grad_o, const float* grad_d, const float* Ks, float* grad_C2Ws,
                                  const int* locs, int B, int ref_index_bug, void* stream)
{
    if (B <= 0) return 0;
    ray_bwd_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(grad_o, grad_d, Ks, grad_C2Ws, locs, B, ref_index_bug);
    SNRF_RETURN_LAUNCH("snrf_compute_ray_bwd");
}

SNRF_API int snrf_ray_aabb(const float* rays_o, const float* rays_d, const float* centers, const float* sizes,
                           float* bounds, int B, int K, void* stream)
{
    if (B <= 0 || K <= 0) return 0;
    ray_aabb_kernel<<<grid1d((long long)B * K), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, centers, sizes, (float2*)bounds, B, K);
    SNRF_RETURN_LAUNCH("snrf_ray_aabb");
}

SNRF_API int snrf_sample_grid(const float* rays_o, const float* rays_d, float* z_vals, float* dists,
                              const float* corner, const float* size, const unsigned char* occupied,
                              const int* log2dim, int* counts, int B, int S, void* stream)
{
    SNRF_CHECK_ARG(S > 0, "snrf_sample_grid: num_sample must be positive");
    if (B <= 0) return 0;
    sample_grid_kernel<<<grid1d(B, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, S, z_vals, dists, corner, size, occupied, log2dim, counts, B);
    SNRF_RETURN_LAUNCH("snrf_sample_grid");
}

SNRF_API int snrf_bg_sampling(const float* starts, const float* bg_depth, float* z_vals, int B, int S,
                              float sample_range, void* stream)
{
    SNRF_CHECK_ARG(S > 1, "snrf_bg_sampling: num_sample must be > 1");
    if (B <= 0) return 0;
    bg_sampling_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(starts, bg_depth, z_vals, S, sample_range, B);
    SNRF_RETURN_LAUNCH("snrf_bg_sampling");
}

SNRF_API int snrf_sample_insideout(const float* rays_o, const float* rays_d, int S, int Sbg, const float* center,
                                   const float* size, float far, float* z_vals, float* z_vals_bg, int* miss_flag,
                                   int B, void* stream)
{
    SNRF_CHECK_ARG(S > 1 && Sbg > 1, "snrf_sample_insideout: sample counts must be > 1");
    if (B <= 0) return 0;
    insideout_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, S, Sbg, center, size, B, far, z_vals, z_vals_bg, miss_flag);
    SNRF_RETURN_LAUNCH("snrf_sample_insideout");
}

SNRF_API int snrf_bg_inverse_z(const float* rays_o, const float* rays_d, const float* center, const float* size, const float* t_lin,
                               float* z_vals, float* dists, unsigned char* valid, int B, int S, int invalid_underground, void* stream)
{
    SNRF_CHECK_ARG(S > 0 && t_lin != nullptr, "snrf_bg_inverse_z: S must be positive and t_lin given (S=%d)", S);
    if (B <= 0) return 0;
    bg_inverse_z_kernel<<<grid1d((long long)B * S), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, center, size, t_lin, z_vals, dists, valid,
                                                                                      B, S, invalid_underground);
    SNRF_RETURN_LAUNCH("snrf_bg_inverse_z");
}
