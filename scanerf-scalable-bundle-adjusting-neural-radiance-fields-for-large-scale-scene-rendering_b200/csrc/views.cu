// View selection, neighbour-view projection (+ analytic backward) and image sampling, sm_100a.
//
// Replaces (behaviour, not code) the reference operators
//   computeViewcost                         cuda/view_selection_kernel.cu:18-112
//   proj2neighbor_forward / _backward       cuda/view_selection_kernel.cu:115-352
//   grid_sample_forward/backward_cuda       cuda/grid_sample_kernel.cu:15-213
//   gaussian_grid_sample_forward/backward   cuda/grid_sample_kernel.cu:216-441
//   grid_sample_bool_cuda                   cuda/grid_sample_kernel.cu:445-493
//   proj2pixel_and_fetch_color              cuda/helper_kernel.cu:17-104
// Camera conventions are those of cuda/include/camera.h: K row-major 3x3, rt = world->camera
// row-major 3x4, camera centre = -R^T t.
//
// All of these are streaming kernels over (point, view) pairs: one thread per pair, per-view
// camera parameters staged once per CTA in shared memory (the reference re-reads the 21 floats
// per thread from global memory), pose-gradient atomics privatised per CTA in shared memory
// before they touch HBM (the reference: 12 global atomics per pair onto N_cam rows).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid1d(long long n)
{
    long long g = (n + kThreads - 1) / kThreads;
    const long long cap = (long long)snrf_sm_count() * 16;
    if (g > cap) g = cap;
    return g > 0 ? (int)g : 1;
}

struct Cam {
    float k[9], e[12];
    __device__ __forceinline__ void load(const float* __restrict__ ks, const float* __restrict__ rts, int i)
    {
#pragma unroll
        for (int j = 0; j < 9; ++j) k[j] = ks[9 * (size_t)i + j];
#pragma unroll
        for (int j = 0; j < 12; ++j) e[j] = rts[12 * (size_t)i + j];
    }
    __device__ __forceinline__ f3 rotate(f3 p) const
    {
        return mk3(e[0] * p.x + e[1] * p.y + e[2] * p.z, e[4] * p.x + e[5] * p.y + e[6] * p.z, e[8] * p.x + e[9] * p.y + e[10] * p.z);
    }
    __device__ __forceinline__ f3 world2cam(f3 p) const
    {
        const f3 r = rotate(p);
        return mk3(r.x + e[3], r.y + e[7], r.z + e[11]);
    }
    __device__ __forceinline__ f3 cam2pixel(f3 p) const
    {
        return mk3(k[0] * p.x + k[1] * p.y + k[2] * p.z, k[3] * p.x + k[4] * p.y + k[5] * p.z, k[6] * p.x + k[7] * p.y + k[8] * p.z);
    }
    __device__ __forceinline__ f3 center() const
    {   // translation of the inverted extrinsic (camera.h:83-95)
        return mk3(-(e[0] * e[3] + e[4] * e[7] + e[8] * e[11]), -(e[1] * e[3] + e[5] * e[7] + e[9] * e[11]),
                   -(e[2] * e[3] + e[6] * e[7] + e[10] * e[11]));
    }
    __device__ __forceinline__ f3 rotate_inv(f3 p) const
    {
        return mk3(e[0] * p.x + e[4] * p.y + e[8] * p.z, e[1] * p.x + e[5] * p.y + e[9] * p.z, e[2] * p.x + e[6] * p.y + e[10] * p.z);
    }
};

__device__ __forceinline__ f3 normalize3(f3 v) { return v * rsqrtf(dot3(v, v)); }   // cutil_math.h:471-475
__device__ __forceinline__ float norm3(f3 v) { return sqrtf(dot3(v, v)); }

// ------------------------------------------------------------------ view cost
// costs[n, b] = 0.9 (1 - cos angle(ray, neighbour ray)) + 0.1 max(0, 1 - |p - o| / |p - o_n|),
// 1 when p is behind camera n or projects outside its image.
__global__ void __launch_bounds__(kThreads)
view_cost_kernel(const float* __restrict__ pts, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                 const float* __restrict__ ks, const float* __restrict__ rts, float* __restrict__ costs, int height,
                 int width, int B)
{
    __shared__ Cam cam;
    if (threadIdx.x == 0) cam.load(ks, rts, blockIdx.y);
    __syncthreads();
    const f3 no = cam.center();
    float* row = costs + (size_t)blockIdx.y * B;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 p = ld3(pts + 3 * (size_t)i);
        const f3 uv = cam.cam2pixel(cam.world2cam(p));
        float cost = 1.0f;
        if (!(uv.z <= 0.001)) {
            const float x = uv.x / uv.z, y = uv.y / uv.z;
            if (!(x <= 0 || x >= width - 1 || y <= 0 || y >= height - 1)) {
                const f3 o = ld3(rays_o + 3 * (size_t)i);
                const f3 d = normalize3(ld3(rays_d + 3 * (size_t)i));
                const f3 nd = normalize3(p - no);
                const float angle_cost = 1.0f - dot3(d, nd);
                const float dis_cost = fmaxf(0.0f, 1.0f - norm3(p - o) / norm3(p - no));
                cost = (1.0f - 0.1f) * angle_cost + 0.1f * dis_cost;
            }
        }
        row[i] = cost;
    }
}

// ------------------------------------------------------------------ neighbour projection
__global__ void __launch_bounds__(kThreads)
proj_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ ks, const float* __restrict__ rts,
                const int* __restrict__ nei_views, const unsigned char* __restrict__ nei_valid, float* __restrict__ nei_origin,
                float* __restrict__ nei_direction, float* __restrict__ grid, long long total, int K)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        if (!nei_valid[t]) continue;
        Cam cam;
        cam.load(ks, rts, nei_views[t]);
        const f3 pc = cam.world2cam(ld3(pts + 3 * (size_t)(t / K)));
        st3(grid + 3 * (size_t)t, cam.cam2pixel(pc));
        st3(nei_origin + 3 * (size_t)t, cam.center());
        // the reference divides by (z + 1e-8) in double precision (1e-8 is a double literal)
        const double zz = (double)pc.z + 1e-8;
        const f3 dc = mk3((float)((double)pc.x / zz), (float)((double)pc.y / zz), 1.0f);
        st3(nei_direction + 3 * (size_t)t, cam.rotate_inv(dc));
    }
}

// d(grid)/d(pts), d(grid)/d(rt): grid = K (R p + t).  grad_pts[B,3], grad_rts[N,12] are ACCUMULATED.
// Pose gradients are summed per CTA in shared memory (n_cam * 12 floats) when they fit.
__global__ void __launch_bounds__(kThreads)
proj_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ ks, const float* __restrict__ rts,
                const int* __restrict__ nei_views, const unsigned char* __restrict__ nei_valid, const float* __restrict__ dgrid,
                float* __restrict__ grad_pts, float* __restrict__ grad_rts, long long total, int K, int n_cam, int use_smem)
{
    extern __shared__ float acc[];
    if (use_smem) {
        for (int i = threadIdx.x; i < n_cam * 12; i += blockDim.x) acc[i] = 0.0f;
        __syncthreads();
    }
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        if (!nei_valid[t]) continue;
        const int n = nei_views[t];
        const float* k = ks + 9 * (size_t)n;
        const float* e = rts + 12 * (size_t)n;
        const f3 g = ld3(dgrid + 3 * (size_t)t);
        const size_t b = (size_t)(t / K);
        const f3 p = ld3(pts + 3 * b);
        // dL/d(cam point) = K^T g
        const float c1 = k[0] * g.x + k[3] * g.y + k[6] * g.z;
        const float c2 = k[1] * g.x + k[4] * g.y + k[7] * g.z;
        const float c3 = k[2] * g.x + k[5] * g.y + k[8] * g.z;
        // dL/dp = R^T K^T g, written as the reference does: sum over image rows of (K R)_row * g_row
        const float dx = (k[0] * e[0] + k[1] * e[4] + k[2] * e[8]) * g.x + (k[3] * e[0] + k[4] * e[4] + k[5] * e[8]) * g.y +
                         (k[6] * e[0] + k[7] * e[4] + k[8] * e[8]) * g.z;
        const float dy = (k[0] * e[1] + k[1] * e[5] + k[2] * e[9]) * g.x + (k[3] * e[1] + k[4] * e[5] + k[5] * e[9]) * g.y +
                         (k[6] * e[1] + k[7] * e[5] + k[8] * e[9]) * g.z;
        const float dz = (k[0] * e[2] + k[1] * e[6] + k[2] * e[10]) * g.x + (k[3] * e[2] + k[4] * e[6] + k[5] * e[10]) * g.y +
                         (k[6] * e[2] + k[7] * e[6] + k[8] * e[10]) * g.z;
        float* dst = use_smem ? acc + 12 * n : grad_rts + 12 * (size_t)n;
        const float v[12] = {c1 * p.x, c1 * p.y, c1 * p.z, c1, c2 * p.x, c2 * p.y, c2 * p.z, c2, c3 * p.x, c3 * p.y, c3 * p.z, c3};
#pragma unroll
        for (int j = 0; j < 12; ++j) atomicAdd(dst + j, v[j]);
        atomicAdd(grad_pts + 3 * b + 0, dx);
        atomicAdd(grad_pts + 3 * b + 1, dy);
        atomicAdd(grad_pts + 3 * b + 2, dz);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_cam * 12; i += blockDim.x)
            if (acc[i] != 0.0f) atomicAdd(grad_rts + i, acc[i]);
    }
}

// ------------------------------------------------------------------ image sampling
struct Tap4 {
    bool inside;
    int i00;
    float x, y;
};
// normalised [-1,1] grid -> pixel space (align-corners), the four-tap footprint (grid_sample_kernel.cu:18-36)
__device__ __forceinline__ Tap4 locate_tap(float2& uv, int height, int width)
{
    uv.x = (uv.x + 1.0f) / 2.0f * (width - 1);
    uv.y = (uv.y + 1.0f) / 2.0f * (height - 1);
    Tap4 t;
    t.inside = !(uv.x < 0 || uv.x >= width - 1 || uv.y < 0 || uv.y >= height - 1);
    const int x0 = (int)uv.x, y0 = (int)uv.y;
    t.x = uv.x - x0; t.y = uv.y - y0;
    t.i00 = x0 + y0 * width;
    return t;
}
__device__ __forceinline__ f3 texel(const unsigned char* __restrict__ src, int idx)
{
    return mk3((float)src[idx * 3 + 0], (float)src[idx * 3 + 1], (float)src[idx * 3 + 2]);
}

__global__ void __launch_bounds__(kThreads)
grid_sample_fwd_kernel(const unsigned char* __restrict__ src, const float2* __restrict__ grid, float* __restrict__ out,
                       unsigned char* __restrict__ mask, int B, int height, int width)
{
    const unsigned char* img = src + (size_t)blockIdx.y * height * width * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const size_t loc = (size_t)blockIdx.y * B + i;
        float2 uv = grid[loc];
        const Tap4 t = locate_tap(uv, height, width);
        f3 c = mk3(0, 0, 0);
        bool ok = false;
        if (t.inside) {
            const f3 v00 = texel(img, t.i00), v01 = texel(img, t.i00 + width), v10 = texel(img, t.i00 + 1), v11 = texel(img, t.i00 + width + 1);
            c = v00 * (1.0f - t.x) * (1.0f - t.y) + v01 * (1.0f - t.x) * t.y + v10 * t.x * (1.0f - t.y) + v11 * t.x * t.y;
            ok = c.x != -1.0f;         // the reference signals "outside" through the colour value -1
            if (!ok) c = mk3(0, 0, 0);
        }
        mask[loc] = ok ? 1 : 0;
        st3(out + 3 * loc, c);
    }
}

__global__ void __launch_bounds__(kThreads)
grid_sample_bwd_kernel(const unsigned char* __restrict__ src, const float2* __restrict__ grid, const float* __restrict__ grad_in,
                       float2* __restrict__ grad_grid, int B, int height, int width)
{
    const unsigned char* img = src + (size_t)blockIdx.y * height * width * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const size_t loc = (size_t)blockIdx.y * B + i;
        float2 uv = grid[loc];
        const Tap4 t = locate_tap(uv, height, width);
        float2 g = make_float2(0.f, 0.f);
        if (t.inside) {
            const f3 v00 = texel(img, t.i00), v01 = texel(img, t.i00 + width), v10 = texel(img, t.i00 + 1), v11 = texel(img, t.i00 + width + 1);
            const f3 gx = (-1.0f * v00 * (1.0f - t.y) - v01 * t.y + v10 * (1.0f - t.y) + v11 * t.y) * (width - 1.0f) * (1.0f / 2.0f);
            const f3 gy = (-1.0f * v00 * (1.0f - t.x) + v01 * (1.0f - t.x) - v10 * t.x + v11 * t.x) * (height - 1.0f) * (1.0f / 2.0f);
            const f3 gi = ld3(grad_in + 3 * loc);
            g = make_float2(dot3(gi, gx), dot3(gi, gy));
        }
        grad_grid[loc] = g;
    }
}

// Gaussian-window resampling: weights exp(-d^2 / sigma^2) over an M x M window of pixel centres
template <bool BACKWARD>
__global__ void __launch_bounds__(kThreads)
gauss_sample_kernel(const unsigned char* __restrict__ src, const float2* __restrict__ grid, const float* __restrict__ grad_in,
                    float* __restrict__ out, unsigned char* __restrict__ mask, float2* __restrict__ grad_grid, int B, float sigma,
                    float max_dis, int height, int width)
{
    const unsigned char* img = src + (size_t)blockIdx.y * height * width * 3;
    const float item = -1.0f / (sigma * sigma);
    const int M = (int)(max_dis * 2) + 2, S = M / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const size_t loc = (size_t)blockIdx.y * B + i;
        float2 uv = grid[loc];
        const Tap4 t = locate_tap(uv, height, width);
        if (!t.inside) {
            if (BACKWARD) grad_grid[loc] = make_float2(0.f, 0.f);
            else { mask[loc] = 0; st3(out + 3 * loc, mk3(0, 0, 0)); }
            continue;
        }
        const int x0 = (int)uv.x, y0 = (int)uv.y;
        float total = 0.0f;
        f3 a = mk3(0, 0, 0), bsum = mk3(0, 0, 0);
        for (int ii = 0; ii < M; ++ii)
            for (int jj = 0; jj < M; ++jj) {
                const int lx = x0 + ii - S, ly = y0 + jj - S;
                if (lx < 0 || lx >= width || ly < 0 || ly >= height) continue;
                const float x = lx + 0.5f, y = ly + 0.5f;
                const float dis = (x - uv.x) * (x - uv.x) + (y - uv.y) * (y - uv.y);
                const float w = expf(item * dis);
                const f3 c = texel(img, ly * width + lx);
                if (BACKWARD) {
                    const f3 dc = c * w * item;
                    a = a + dc * (uv.x - x) * (width - 1.0f);
                    bsum = bsum + dc * (uv.y - y) * (height - 1.0f);
                } else {
                    a = a + w * c;
                }
                total += w;
            }
        if (total > 0) {
            const float inv = 1.0f / total;            // float3 / float = multiply by the reciprocal (cutil_math.h:411-415)
            a = a * inv;
            if (BACKWARD) bsum = bsum * inv;
        }
        if (BACKWARD) {
            const f3 gi = ld3(grad_in + 3 * loc);
            grad_grid[loc] = make_float2(dot3(gi, a), dot3(gi, bsum));
        } else {
            const bool ok = a.x != -1.0f;
            mask[loc] = ok ? 1 : 0;
            st3(out + 3 * loc, ok ? a : mk3(0, 0, 0));
        }
    }
}

// nearest-pixel fetch from boolean images; out-of-image locations keep the caller's value
__global__ void __launch_bounds__(kThreads)
grid_sample_bool_kernel(const unsigned char* __restrict__ src, const float2* __restrict__ grid, unsigned char* __restrict__ out,
                        int B, int height, int width)
{
    const unsigned char* img = src + (size_t)blockIdx.y * height * width;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const size_t loc = (size_t)blockIdx.y * B + i;
        float2 uv = grid[loc];
        uv.x = (uv.x + 1.0f) / 2.0f * (width - 1);
        uv.y = (uv.y + 1.0f) / 2.0f * (height - 1);
        const int x = (int)(uv.x + 0.5f), y = (int)(uv.y + 0.5f);
        if (x >= 0 && x < width && y >= 0 && y < height) out[loc] = img[y * width + x];
    }
}

// project every point into every view (camera-to-world poses) and fetch a bilinear colour from
// float images; (-1,-1,-1) / 0 when behind the camera or outside the image
__global__ void __launch_bounds__(kThreads)
proj_fetch_kernel(const float* __restrict__ pts, const float* __restrict__ Ks, const float* __restrict__ C2Ws,
                  const float* __restrict__ rgbs, float* __restrict__ fetched_pixels, float* __restrict__ fetched_colors, int B,
                  int n_cam, int height, int width)
{
    const int n = blockIdx.y;
    const float* K = Ks + 9 * (size_t)n;
    const float* M = C2Ws + 12 * (size_t)n;
    const float* img = rgbs + (size_t)n * height * width * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        f3 p = ld3(pts + 3 * (size_t)i);
        p.x -= M[3]; p.y -= M[7]; p.z -= M[11];
        const float xc = M[0] * p.x + M[4] * p.y + M[8] * p.z;
        const float yc = M[1] * p.x + M[5] * p.y + M[9] * p.z;
        const float zc = M[2] * p.x + M[6] * p.y + M[10] * p.z;
        float px = K[0] * xc + K[1] * yc + K[2] * zc;
        float py = K[3] * xc + K[4] * yc + K[5] * zc;
        const float pz = K[6] * xc + K[7] * yc + K[8] * zc;
        f3 col = mk3(0, 0, 0), pix = mk3(-1.0f, -1.0f, -1.0f);
        if (pz > 0) {
            px /= pz; py /= pz;
            if (px >= 0 && px <= width - 1 && py >= 0 && py <= height - 1) {
                pix = mk3(px, py, zc);
                if (!(px < 1 || px >= width - 1 || py < 1 || py >= height - 1)) {      // Bilinear<> border rule (interpolation.h:46-55)
                    const int x0 = (int)px, y0 = (int)py;
                    const float x = px - x0, y = py - y0;
                    const int i00 = x0 + y0 * width, i01 = i00 + width, i10 = i00 + 1, i11 = i01 + 1;
                    float c[3];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        c[ch] = img[i00 * 3 + ch] * (1.0f - x) * (1.0f - y) + img[i01 * 3 + ch] * (1.0f - x) * y +
                                img[i10 * 3 + ch] * x * (1.0f - y) + img[i11 * 3 + ch] * x * y;
                    col = mk3(c[0], c[1], c[2]);
                }
            }
        }
        const size_t o = ((size_t)i * n_cam + n) * 3;
        st3(fetched_colors + o, col);
        st3(fetched_pixels + o, pix);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Neighbour-view colour fetch of the warp loss (warp_loss.py:441-519 WarpLoss.sample_neighbor_color).  The reference
// keeps the training images on the HOST, builds four index tensors per batch on the GPU, copies them to the CPU,
// gathers there and copies the four colour tensors back (a D2H + H2D round trip and a synchronisation every step).
// Here the images stay resident in HBM as uint8 [N,H,W,3] and one kernel does, per (ray, neighbour):
//   lt = trunc(grid), offset = grid - lt, nearest = trunc(grid + 0.5)
//   valid_out = valid_in & occlusion[view][nearest]
//   colour    = bilinear blend of the pixels lt, lt + (1,0), lt + (0,1), lt + (1,1) of image `view`, scaled to [0,1]
// Pixel coordinates are clamped into the image (the reference indexes out of range at the right / bottom border and
// wraps around at negative indices); pairs with valid_in = 0 give zeros.
struct NeiTaps {
    int x0, y0, x1, y1;
    float fx, fy;
};
__device__ __forceinline__ NeiTaps nei_taps(float2 g, int height, int width)
{
    NeiTaps t;
    const long long lx = (long long)g.x, ly = (long long)g.y;          // .long(): truncation toward zero
    t.fx = g.x - (float)lx; t.fy = g.y - (float)ly;
    t.x0 = (int)min(max(lx, 0ll), (long long)width - 1); t.x1 = (int)min(max(lx + 1, 0ll), (long long)width - 1);
    t.y0 = (int)min(max(ly, 0ll), (long long)height - 1); t.y1 = (int)min(max(ly + 1, 0ll), (long long)height - 1);
    return t;
}
__device__ __forceinline__ f3 pixel_u8(const unsigned char* __restrict__ img, int y, int x, int width)
{
    const unsigned char* p = img + 3 * ((size_t)y * width + x);
    return mk3((float)p[0], (float)p[1], (float)p[2]);
}

__global__ void __launch_bounds__(kThreads)
nei_sample_fwd_kernel(const unsigned char* __restrict__ images, const unsigned char* __restrict__ occlusion,
                      const float2* __restrict__ grid, const int* __restrict__ nei_views, const unsigned char* __restrict__ nei_valid,
                      float* __restrict__ color, unsigned char* __restrict__ valid_out, int total, int height, int width)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        bool ok = nei_valid[i] != 0;
        f3 c = mk3(0, 0, 0);
        if (ok) {
            const float2 g = grid[i];
            const size_t view = (size_t)nei_views[i];
            if (occlusion) {
                const long long nx = (long long)(g.x + 0.5f), ny = (long long)(g.y + 0.5f);
                const int cx = (int)min(max(nx, 0ll), (long long)width - 1), cy = (int)min(max(ny, 0ll), (long long)height - 1);
                ok = occlusion[(view * height + cy) * width + cx] != 0;
            }
            const NeiTaps t = nei_taps(g, height, width);
            const unsigned char* img = images + view * (size_t)height * width * 3;
            const f3 lt = pixel_u8(img, t.y0, t.x0, width), rt = pixel_u8(img, t.y0, t.x1, width);
            const f3 lb = pixel_u8(img, t.y1, t.x0, width), rb = pixel_u8(img, t.y1, t.x1, width);
            const float s = 1.0f / 255.0f;
            c = ((1.0f - t.fx) * (1.0f - t.fy) * s) * lt + (t.fx * (1.0f - t.fy) * s) * rt + ((1.0f - t.fx) * t.fy * s) * lb + (t.fx * t.fy * s) * rb;
        }
        st3(color + 3 * (size_t)i, c);
        valid_out[i] = ok ? 1 : 0;
    }
}

// d colour / d grid: the blend weights are linear in the offsets, the taps are piecewise constant
__global__ void __launch_bounds__(kThreads)
nei_sample_bwd_kernel(const unsigned char* __restrict__ images, const float2* __restrict__ grid, const int* __restrict__ nei_views,
                      const unsigned char* __restrict__ nei_valid, const float* __restrict__ grad_color, float2* __restrict__ grad_grid,
                      int total, int height, int width)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float2 gg = make_float2(0.0f, 0.0f);
        if (nei_valid[i]) {
            const NeiTaps t = nei_taps(grid[i], height, width);
            const unsigned char* img = images + (size_t)nei_views[i] * (size_t)height * width * 3;
            const f3 lt = pixel_u8(img, t.y0, t.x0, width), rt = pixel_u8(img, t.y0, t.x1, width);
            const f3 lb = pixel_u8(img, t.y1, t.x0, width), rb = pixel_u8(img, t.y1, t.x1, width);
            const f3 g = ld3(grad_color + 3 * (size_t)i);
            const float s = 1.0f / 255.0f;
            const f3 dx = (1.0f - t.fy) * (rt - lt) + t.fy * (rb - lb);
            const f3 dy = (1.0f - t.fx) * (lb - lt) + t.fx * (rb - rt);
            gg = make_float2(s * dot3(g, dx), s * dot3(g, dy));
        }
        grad_grid[i] = gg;
    }
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
SNRF_API int snrf_view_cost(const float* rays_o, const float* rays_d, const float* pts, const float* ks, const float* rts,
                            float* costs, int n_cam, int B, int height, int width, void* stream)
{
    SNRF_CHECK_ARG(n_cam >= 0 && B >= 0, "snrf_view_cost: negative size");
    if (n_cam == 0 || B == 0) return 0;
    SNRF_CHECK_ARG(n_cam <= 65535, "snrf_view_cost: at most 65535 cameras per call (got %d)", n_cam);
    view_cost_kernel<<<dim3(grid1d(B), n_cam), kThreads, 0, (cudaStream_t)stream>>>(pts, rays_o, rays_d, ks, rts, costs, height, width, B);
    SNRF_RETURN_LAUNCH("snrf_view_cost");
}

SNRF_API int snrf_proj2nei_fwd(const float* pts, const float* ks, const float* rts, const int* nei_views,
                               const unsigned char* nei_valid, float* nei_origin, float* nei_direction, float* grid, int B, int K,
                               void* stream)
{
    SNRF_CHECK_ARG(B >= 0 && K >= 0, "snrf_proj2nei_fwd: negative size");
    const long long total = (long long)B * K;
    if (total == 0) return 0;
    proj_fwd_kernel<<<grid1d(total), kThreads, 0, (cudaStream_t)stream>>>(pts, ks, rts, nei_views, nei_valid, nei_origin, nei_direction, grid, total, K);
    SNRF_RETURN_LAUNCH("snrf_proj2nei_fwd");
}

SNRF_API int snrf_proj2nei_bwd(const float* pts, const float* ks, const float* rts, const int* nei_views,
                               const unsigned char* nei_valid, const float* dL_dgrid, float* grad_pts, float* grad_rts, int B, int K,
                               int n_cam, void* stream)
{
    SNRF_CHECK_ARG(B >= 0 && K >= 0 && n_cam > 0, "snrf_proj2nei_bwd: bad size");
    const long long total = (long long)B * K;
    if (total == 0) return 0;
    const size_t smem = (size_t)n_cam * 12 * sizeof(float);
    const int use_smem = smem <= 40 * 1024;
    int grid = grid1d(total);
    if (use_smem && grid > snrf_sm_count() * 4) grid = snrf_sm_count() * 4;       // fewer, longer CTAs: fewer flushes
    proj_bwd_kernel<<<grid, kThreads, use_smem ? smem : 0, (cudaStream_t)stream>>>(pts, ks, rts, nei_views, nei_valid, dL_dgrid, grad_pts,
                                                                                 grad_rts, total, K, n_cam, use_smem);
    SNRF_RETURN_LAUNCH("snrf_proj2nei_bwd");
}

SNRF_API int snrf_grid_sample_fwd(const unsigned char* src, const float* grid, float* out, unsigned char* mask, int n_img, int B,
                                  int height, int width, void* stream)
{
    if (n_img <= 0 || B <= 0) return 0;
    SNRF_CHECK_ARG(n_img <= 65535, "snrf_grid_sample_fwd: at most 65535 images per call");
    grid_sample_fwd_kernel<<<dim3(grid1d(B), n_img), kThreads, 0, (cudaStream_t)stream>>>(src, (const float2*)grid, out, mask, B, height, width);
    SNRF_RETURN_LAUNCH("snrf_grid_sample_fwd");
}

SNRF_API int snrf_grid_sample_bwd(const unsigned char* src, const float* grid, const float* grad_in, float* grad_grid, int n_img,
                                  int B, int height, int width, void* stream)
{
    if (n_img <= 0 || B <= 0) return 0;
    SNRF_CHECK_ARG(n_img <= 65535, "snrf_grid_sample_bwd: at most 65535 images per call");
    grid_sample_bwd_kernel<<<dim3(grid1d(B), n_img), kThreads, 0, (cudaStream_t)stream>>>(src, (const float2*)grid, grad_in, (float2*)grad_grid, B, height, width);
    SNRF_RETURN_LAUNCH("snrf_grid_sample_bwd");
}

SNRF_API int snrf_gauss_sample_fwd(const unsigned char* src, const float* grid, float* out, unsigned char* mask, int n_img, int B,
                                   int height, int width, float sigma, float max_dis, void* stream)
{
    if (n_img <= 0 || B <= 0) return 0;
    SNRF_CHECK_ARG(n_img <= 65535, "snrf_gauss_sample_fwd: at most 65535 images per call");
    gauss_sample_kernel<false><<<dim3(grid1d(B), n_img), kThreads, 0, (cudaStream_t)stream>>>(src, (const float2*)grid, nullptr, out, mask, nullptr, B, sigma, max_dis, height, width);
    SNRF_RETURN_LAUNCH("snrf_gauss_sample_fwd");
}

SNRF_API int snrf_gauss_sample_bwd(const unsigned char* src, const float* grid, const float* grad_in, float* grad_grid, int n_img,
                                   int B, int height, int width, float sigma, float max_dis, void* stream)
{
    if (n_img <= 0 || B <= 0) return 0;
    SNRF_CHECK_ARG(n_img <= 65535, "snrf_gauss_sample_bwd: at most 65535 images per call");
    gauss_sample_kernel<true><<<dim3(grid1d(B), n_img), kThreads, 0, (cudaStream_t)stream>>>(src, (const float2*)grid, grad_in, nullptr, nullptr, (float2*)grad_grid, B, sigma, max_dis, height, width);
    SNRF_RETURN_LAUNCH("snrf_gauss_sample_bwd");
}

SNRF_API int snrf_grid_sample_bool(const unsigned char* src, const float* grid, unsigned char* out, int n_img, int B, int height,
                                   int width, void* stream)
{
    if (n_img <= 0 || B <= 0) return 0;
    SNRF_CHECK_ARG(n_img <= 65535, "snrf_grid_sample_bool: at most 65535 images per call");
    grid_sample_bool_kernel<<<dim3(grid1d(B), n_img), kThreads, 0, (cudaStream_t)stream>>>(src, (const float2*)grid, out, B, height, width);
    SNRF_RETURN_LAUNCH("snrf_grid_sample_bool");
}

SNRF_API int snrf_proj2pixel_fetch(const float* pts, const float* Ks, const float* C2Ws, const float* rgbs, float* fetched_pixels,
                                   float* fetched_colors, int B, int n_cam, int height, int width, void* stream)
{
    if (n_cam <= 0 || B <= 0) return 0;
    SNRF_CHECK_ARG(n_cam <= 65535, "snrf_proj2pixel_fetch: at most 65535 cameras per call");
    proj_fetch_kernel<<<dim3(grid1d(B), n_cam), kThreads, 0, (cudaStream_t)stream>>>(pts, Ks, C2Ws, rgbs, fetched_pixels, fetched_colors, B, n_cam, height, width);
    SNRF_RETURN_LAUNCH("snrf_proj2pixel_fetch");
}

SNRF_API int snrf_nei_sample_fwd(const unsigned char* images, const unsigned char* occlusion, const float* grid, const int* nei_views,
                                 const unsigned char* nei_valid, float* color, unsigned char* valid_out, int B, int K, int height,
                                 int width, void* stream)
{
    if (B <= 0 || K <= 0) return 0;
    SNRF_CHECK_ARG(height > 0 && width > 0, "snrf_nei_sample_fwd: empty image (%d x %d)", height, width);
    SNRF_CHECK_ARG((long long)B * K <= 0x7fffffffll, "snrf_nei_sample_fwd: B * K exceeds 2^31");
    nei_sample_fwd_kernel<<<grid1d(B * K), kThreads, 0, (cudaStream_t)stream>>>(images, occlusion, (const float2*)grid, nei_views, nei_valid, color,
                                                                               valid_out, B * K, height, width);
    SNRF_RETURN_LAUNCH("snrf_nei_sample_fwd");
}

SNRF_API int snrf_nei_sample_bwd(const unsigned char* images, const float* grid, const int* nei_views, const unsigned char* nei_valid,
                                 const float* grad_color, float* grad_grid, int B, int K, int height, int width, void* stream)
{
    if (B <= 0 || K <= 0) return 0;
    SNRF_CHECK_ARG(height > 0 && width > 0, "snrf_nei_sample_bwd: empty image (%d x %d)", height, width);
    SNRF_CHECK_ARG((long long)B * K <= 0x7fffffffll, "snrf_nei_sample_bwd: B * K exceeds 2^31");
    nei_sample_bwd_kernel<<<grid1d(B * K), kThreads, 0, (cudaStream_t)stream>>>(images, (const float2*)grid, nei_views, nei_valid, grad_color,
                                                                               (float2*)grad_grid, B * K, height, width);
    SNRF_RETURN_LAUNCH("snrf_nei_sample_bwd");
}
