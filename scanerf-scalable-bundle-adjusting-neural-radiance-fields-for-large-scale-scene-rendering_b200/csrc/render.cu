// Multi-tile inference renderer, everything except the fused encode + decoder (infer.cu), sm_100a.
//
// Replaces (behaviour, not code) the reference operators of hashgrid/src/rendering_kernel.cu:
//   ray_block_intersection   :126-174      sample_points          :178-382
//   prepare_points           :391-449      accumulate_color       :623-702
//   ray_firsthit_block       :704-812      inverse_z_sampling     :815-868
//   get_last_block           :1211-1259    update_outgoing_bidx   :1262-1401 (+ _v2 :1405-1472)
//   process_occupied_grid    :1478-1565
// "block" is the reference's word for a tile in the renderer; tile boxes are (corner, size),
// occupancy grids of all tiles are concatenated bytes addressed through grid_starts (int64) and
// grid_log2dim (int3 per tile); `intersections[B, nb, 2]` holds (near, far) per ray and tile with
// 1e7 marking a miss.
//
// One thread per ray (or per sample) as in the reference -- these are streaming kernels whose
// cost is the per-ray rows they read and write; the 128-sample z / dists rows are written by the
// owning thread exactly as the reference does so that untouched entries keep the caller's fill.
#include "common.cuh"
#include "walk.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kMiss = 10000000.0f;      // INF_INTERSECTION
constexpr int kMaxPts = 4;                // MAX_PTS_BLOCKS

inline int grid1d(long long n, int threads = kThreads)
{
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)snrf_sm_count() * 32;
    if (g > cap) g = cap;
    return g > 0 ? (int)g : 1;
}

__global__ void __launch_bounds__(kThreads)
block_isect_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ corners,
                   const float* __restrict__ sizes, float2* __restrict__ out, int nb, int B)
{
    const int b = blockIdx.y;
    const f3 half = ld3(sizes + 3 * b) * 0.5f;            // "/ 2.0f" = multiply by the exact reciprocal
    const f3 center = ld3(corners + 3 * b) + half;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        float2 t = ray_aabb(ld3(rays_o + 3 * (size_t)i), ld3(rays_d + 3 * (size_t)i), center, half);
        if (t.x == -1.0f) t = make_float2(kMiss, kMiss);
        out[(size_t)i * nb + b] = t;
    }
}

struct TileGrid {
    f3 corner, cell;
    int lx, ly, lz;
    const unsigned char* occ;
    __device__ __forceinline__ void load(const float* corners, const float* sizes, const unsigned char* grid_occ,
                                         const long long* starts, const int* log2dim, int b)
    {
        corner = ld3(corners + 3 * b);
        const f3 size = ld3(sizes + 3 * b);
        lx = log2dim[3 * b]; ly = log2dim[3 * b + 1]; lz = log2dim[3 * b + 2];
        cell = mk3(size.x / (float)(1 << lx), size.y / (float)(1 << ly), size.z / (float)(1 << lz));
        occ = grid_occ + starts[b];
    }
    __device__ __forceinline__ void start(Walk& w, f3 o, f3 d, float2 t) const { w.init(o - corner, d, t, 1 << lx, 1 << ly, 1 << lz, cell); }
    __device__ __forceinline__ bool on(const Walk& w) const
    {
        return occ[((uint32_t)w.cx << (ly + lz)) | ((uint32_t)w.cy << lz) | (uint32_t)w.cz] != 0;
    }
};

// Per ray: advance through the tiles in near-to-far order (tracing_blocks) until one has occupied
// segments beyond z_start, then place S samples there proportionally to segment length.
// Resumable through tracing_idx / z_start.
__global__ void __launch_bounds__(128)
render_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ corners,
                     const float* __restrict__ sizes, const unsigned char* __restrict__ grid_occ,
                     const long long* __restrict__ grid_starts, const int* __restrict__ grid_log2dim, int S, int nb,
                     const int* __restrict__ tracing_blocks, const float2* __restrict__ intersections, int* __restrict__ tracing_idx,
                     float* __restrict__ z_start, float* __restrict__ z_vals, float* __restrict__ dists, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        const int* order = tracing_blocks + (size_t)i * nb;
        const float2* isect = intersections + (size_t)i * nb;
        float* zr = z_vals + (size_t)i * S;
        float* dr = dists + (size_t)i * S;
        int step = tracing_idx[i];
        float2 t = make_float2(z_start[i], 0.0f);
        while (step < nb) {
            const int b = order[step];
            const float2 bound = isect[b];
            if (bound.x == kMiss) break;
            if (t.x >= bound.y) { ++step; continue; }
            if (step == 0) t.x = bound.x;
            TileGrid g;
            g.load(corners, sizes, grid_occ, grid_starts, grid_log2dim, b);
            Walk w;
            g.start(w, o, d, t);
            int num_seg = 0;
            float total = 0.0f;
            while (!w.done()) {
                w.next();
                if (g.on(w)) { const float len = w.t1 - w.t0; if (len > 0) { total += len; ++num_seg; } }
                w.step();
            }
            if (num_seg == 0) { t.x = bound.y; ++step; continue; }
            g.start(w, o, d, t);
            int num = 0, count = 0;
            while (!w.done()) {
                w.next();
                if (g.on(w)) {
                    const float len = w.t1 - w.t0;
                    if (len > 0) {
                        int n = min(max((int)(len / total * S), 1), S - num);
                        if (count == num_seg - 1) n = S - num;
                        if (n > 0) {
                            const float interval = (w.t1 - w.t0) / n;           // uniform_sample_bound_v3
                            for (int k = 0; k < n; ++k) { zr[num + k] = w.t0 + k * interval; dr[num + k] = interval; }
                        }
                        num += n;
                        ++count;
                    }
                }
                w.step();
            }
            t.x = bound.y;
            ++step;
            break;
        }
        tracing_idx[i] = step;
        z_start[i] = t.x;
    }
}

// Per sample of a running ray: the (up to 4) tiles whose [near, far] contains the sample depth.
__global__ void __launch_bounds__(kThreads)
prepare_points_kernel(const float* __restrict__ z_vals, const unsigned char* __restrict__ running, short* __restrict__ block_idxs,
                      const float2* __restrict__ intersections, int S, int nb, long long total)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long ray = t / S;
        if (!running[ray]) continue;
        const float z = z_vals[t];
        if (z == -1.0f) continue;
        const float2* isect = intersections + ray * nb;
        short* dst = block_idxs + t * kMaxPts;
        int index = 0;
        for (int b = 0; b < nb; ++b) {
            const float2 bound = isect[b];
            // the reference writes a fifth match out of bounds; a sample lies in at most 4 tiles by construction
            if (z >= bound.x && z <= bound.y && index < kMaxPts) dst[index++] = (short)b;
        }
    }
}

// Front-to-back accumulation of pre-multiplied (alpha * colour) samples; rays below T = 1e-5 are skipped.
__global__ void __launch_bounds__(256)
accumulate_kernel(const float* __restrict__ pts_diffuse, const float* __restrict__ pts_specular, const float* __restrict__ pts_alpha,
                  float* __restrict__ transparency, const float* __restrict__ z_vals, int S, int B, float* __restrict__ out_diffuse,
                  float* __restrict__ out_specular, float* __restrict__ out_depth)
{
    // One warp per ray.  The transmittance chain is sequential in the sample index (and is kept in that order: the
    // result is defined by it), so the warp stages 32 samples at a time with coalesced loads and lanes 0..6 each carry
    // one of the seven accumulators (3 diffuse, 3 specular, depth) through the chunk, every lane with its own copy of T.
    __shared__ float stage[8][256];                       // per warp: 96 diffuse | 96 specular | 32 alpha | 32 depth
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* s = stage[wib];
    for (int i = blockIdx.x * 8 + wib; i < B; i += gridDim.x * 8) {
        float T = transparency[i];
        if (T < 0.00001f) continue;
        const float* cd = pts_diffuse + 3 * (size_t)i * S;
        const float* cs = pts_specular + 3 * (size_t)i * S;
        const float* al = pts_alpha + (size_t)i * S;
        const float* zz = z_vals + (size_t)i * S;
        float acc = 0.0f;
        if (lane < 3) acc = out_diffuse[3 * (size_t)i + lane];
        else if (lane < 6) acc = out_specular[3 * (size_t)i + lane - 3];
        else if (lane == 6) acc = out_depth[i];
        for (int k0 = 0; k0 < S; k0 += 32) {
            const int n = min(32, S - k0);
            __syncwarp();
            for (int j = lane; j < 3 * n; j += 32) { s[j] = cd[3 * k0 + j]; s[96 + j] = cs[3 * k0 + j]; }
            if (lane < n) { s[192 + lane] = al[k0 + lane]; s[224 + lane] = zz[k0 + lane]; }
            __syncwarp();
            if (lane < 6) {
                const float* v = s + (lane < 3 ? lane : 93 + lane);
                for (int k = 0; k < n; ++k) {
                    acc = acc + T * v[3 * k];
                    T = T * (1 - s[192 + k]);
                }
            } else if (lane == 6) {
                for (int k = 0; k < n; ++k) {
                    const float a = s[192 + k];
                    acc += T * a * s[224 + k];
                    T = T * (1 - a);
                }
            }
        }
        if (lane == 0) transparency[i] = T;
        if (lane < 3) out_diffuse[3 * (size_t)i + lane] = acc;
        else if (lane < 6) out_specular[3 * (size_t)i + lane - 3] = acc;
        else if (lane == 6) out_depth[i] = acc;
    }
}

__global__ void __launch_bounds__(128)
firsthit_block_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ corners,
                      const float* __restrict__ sizes, const unsigned char* __restrict__ grid_occ,
                      const long long* __restrict__ grid_starts, const int* __restrict__ grid_log2dim,
                      const int* __restrict__ tracing_blocks, const float2* __restrict__ intersections, short* __restrict__ hit, int nb,
                      int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        const int* order = tracing_blocks + (size_t)i * nb;
        const float2* isect = intersections + (size_t)i * nb;
        float dis = kMiss;
        int last = -1;
        for (int s = 0; s < nb; ++s) {
            const int b = order[s];
            const float2 bound = isect[b];
            if (bound.x == kMiss) break;
            TileGrid g;
            g.load(corners, sizes, grid_occ, grid_starts, grid_log2dim, b);
            Walk w;
            g.start(w, o, d, bound);
            bool any = false;
            while (!w.done()) {
                w.next();
                if (g.on(w)) { any = true; break; }
                w.step();
            }
            if (any && dis > bound.y) { hit[i] = (short)b; dis = bound.y; }
            last = b;
        }
        if (last != -1 && hit[i] == -1) hit[i] = (short)last;
    }
}

__global__ void __launch_bounds__(kThreads)
inverse_z_kernel(const float2* __restrict__ intersections, const short* __restrict__ related, int S, int nb, float range,
                 float* __restrict__ z_vals, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const short b = related[i];
        if (b == -1) continue;
        const float2 bound = intersections[(size_t)i * nb + b];
        if (bound.x == kMiss) continue;
        const float near = bound.y, far = near + range;
        const float inv_near = 1.0f / near, inv_far = 1.0f / far;       // inverse_z_sample_bound, cuda_utils.h:61-75
        const float inv_bound = inv_far - inv_near;
        const float step = 1.0f / (S - 1);
        float* zr = z_vals + (size_t)i * S;
        for (int k = 0; k < S; ++k) zr[k] = 1.0f / (step * k * inv_bound + inv_near);
    }
}

__global__ void __launch_bounds__(kThreads)
last_block_kernel(const int* __restrict__ tracing_blocks, int* __restrict__ bidxs, const float2* __restrict__ intersections, int nb,
                  int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        int idx = -1;
        for (int s = 0; s < nb; ++s) {
            const int b = tracing_blocks[(size_t)i * nb + s];
            if (intersections[(size_t)i * nb + b].x == kMiss) break;
            idx = b;
        }
        bidxs[i] = idx;
    }
}

// distance-to-face blend weight of a point inside a tile box (x and z faces only, as the reference)
__device__ __forceinline__ float face_weight(f3 dis)
{
    if (dis.x != 0 && dis.z != 0) return dis.x * dis.z;
    if (dis.x != 0) return dis.x;
    if (dis.z != 0) return dis.z;
    return 0.0f;
}
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// Which tile(s) a ray leaves the scene through (the tiles sharing the largest exit depth) and
// their blend weights at the exit point.
__global__ void __launch_bounds__(kThreads)
outgoing_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ corners,
                const float* __restrict__ sizes, const int* __restrict__ tracing_blocks, const float2* __restrict__ intersections,
                short* __restrict__ out_bidx, float* __restrict__ out_w, int skip, int nb, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        float far = -1.0f;
        int index = 0;
        short ids[kMaxPts] = {-1, -1, -1, -1};
        for (int s = 0; s < nb; ++s) {
            const int b = tracing_blocks[(size_t)i * nb + s];
            const float2 bound = intersections[(size_t)i * nb + b];
            if (bound.x == kMiss) break;
            if (!skip && (bound.x > far && far != -1.0f)) break;
            if (bound.y > far) {
                far = bound.y;
                ids[0] = (short)b; ids[1] = ids[2] = ids[3] = -1;
                index = 1;
            } else if (bound.y == far && index < kMaxPts) {
                ids[index++] = (short)b;
            }
        }
        if (far == -1.0f) continue;
        short* ob = out_bidx + (size_t)i * kMaxPts;
        float* ow = out_w + (size_t)i * kMaxPts;
        if (index == 1) { ow[0] = 1.0f; ob[0] = ids[0]; continue; }
        const f3 p = ld3(rays_o + 3 * (size_t)i) + far * ld3(rays_d + 3 * (size_t)i);
#pragma unroll
        for (int k = 0; k < kMaxPts; ++k) {
            const int b = ids[k];
            if (b == -1) break;
            const f3 c = ld3(corners + 3 * b), sz = ld3(sizes + 3 * b);
            const f3 q = mk3(clamp01((p.x - c.x) / sz.x), clamp01((p.y - c.y) / sz.y), clamp01((p.z - c.z) / sz.z));
            const f3 dis = mk3((0.5f - fabsf(q.x - 0.5f)) * sz.x, (0.5f - fabsf(q.y - 0.5f)) * sz.y, (0.5f - fabsf(q.z - 0.5f)) * sz.z);
            ow[k] = face_weight(dis);
            ob[k] = (short)b;
        }
    }
}

// tiles that contain the ray origin, with volume-style blend weights
__global__ void __launch_bounds__(kThreads)
inside_kernel(const float* __restrict__ rays_o, const float* __restrict__ corners, const float* __restrict__ sizes,
              short* __restrict__ out_bidx, float* __restrict__ out_w, int nb, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 o = ld3(rays_o + 3 * (size_t)i);
        int index = 0;
        for (int b = 0; b < nb && index < kMaxPts; ++b) {
            const f3 c = ld3(corners + 3 * b), sz = ld3(sizes + 3 * b);
            const f3 loc = mk3((o.x - c.x) / sz.x, (o.y - c.y) / sz.y, (o.z - c.z) / sz.z);
            if (loc.x >= 0 && loc.x <= 1 && loc.y >= 0 && loc.y <= 1 && loc.z >= 0 && loc.z <= 1) {
                out_bidx[(size_t)i * kMaxPts + index] = (short)b;
                out_w[(size_t)i * kMaxPts + index] = ((0.5f - fabsf(loc.x - 0.5f)) * sz.x) * ((0.5f - fabsf(loc.y - 0.5f)) * sz.y) *
                                                      ((0.5f - fabsf(loc.z - 0.5f)) * sz.z);
                ++index;
            }
        }
    }
}

// Setup: every occupied cell of tile `bidx` marks the cells of the OTHER tiles that contain one of
// its 8 corners (benign races: all writers store `true`).
__global__ void __launch_bounds__(kThreads)
process_occupied_kernel(int bidx, int nb, const float* __restrict__ corners, const float* __restrict__ sizes,
                        const unsigned char* __restrict__ grid_occ, const long long* __restrict__ grid_starts,
                        const int* __restrict__ grid_log2dim, unsigned char* __restrict__ tgt, int total)
{
    const f3 bc = ld3(corners + 3 * bidx), bs = ld3(sizes + 3 * bidx);
    const unsigned char* occ = grid_occ + grid_starts[bidx];
    const int rx = 1 << grid_log2dim[3 * bidx], ry = 1 << grid_log2dim[3 * bidx + 1], rz = 1 << grid_log2dim[3 * bidx + 2];
    (void)rx;
    const f3 cell = mk3(bs.x / (float)rx, bs.y / (float)ry, bs.z / (float)rz);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        if (!occ[t]) continue;
        const int x = t / (ry * rz), rem = t - x * (ry * rz), y = rem / rz, z = rem % rz;
        const f3 p0 = mk3((float)x * cell.x + bc.x, (float)y * cell.y + bc.y, (float)z * cell.z + bc.z);
        for (int b = 0; b < nb; ++b) {
            if (b == bidx) continue;
            const f3 c = ld3(corners + 3 * b), sz = ld3(sizes + 3 * b);
            const int lx = grid_log2dim[3 * b], ly = grid_log2dim[3 * b + 1], lz = grid_log2dim[3 * b + 2];
            unsigned char* g = tgt + grid_starts[b];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // vertex order of the reference is irrelevant: all eight corners are visited
                const f3 v = mk3((float)((j >> 2) & 1), (float)((j >> 1) & 1), (float)(j & 1));
                const f3 q = mk3((p0.x + v.x * cell.x - c.x) / sz.x, (p0.y + v.y * cell.y - c.y) / sz.y, (p0.z + v.z * cell.z - c.z) / sz.z);
                if (q.x >= 0 && q.x < 1 && q.y >= 0 && q.y < 1 && q.z >= 0 && q.z < 1) {
                    const int ix = (int)(q.x * (float)(1 << lx)), iy = (int)(q.y * (float)(1 << ly)), iz = (int)(q.z * (float)(1 << lz));
                    g[(ix << (ly + lz)) | (iy << lz) | iz] = 1;
                }
            }
        }
    }
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
SNRF_API int snrf_ray_block_isect(const float* rays_o, const float* rays_d, const float* corners, const float* sizes,
                                  float* intersections, int B, int nb, void* stream)
{
    if (B <= 0 || nb <= 0) return 0;
    SNRF_CHECK_ARG(nb <= 65535, "snrf_ray_block_isect: at most 65535 tiles");
    block_isect_kernel<<<dim3(grid1d(B), nb), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, corners, sizes, (float2*)intersections, nb, B);
    SNRF_RETURN_LAUNCH("snrf_ray_block_isect");
}

SNRF_API int snrf_render_sample(const float* rays_o, const float* rays_d, const float* corners, const float* sizes,
                                const unsigned char* grid_occupied, const long long* grid_starts, const int* grid_log2dim,
                                const int* tracing_blocks, const float* intersections, int* tracing_idx, float* z_start,
                                float* z_vals, float* dists, int B, int nb, int S, void* stream)
{
    SNRF_CHECK_ARG(S > 0, "snrf_render_sample: num_sample must be positive");
    if (B <= 0 || nb <= 0) return 0;
    render_sample_kernel<<<grid1d(B, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, corners, sizes, grid_occupied, grid_starts,
                                                                         grid_log2dim, S, nb, tracing_blocks, (const float2*)intersections,
                                                                         tracing_idx, z_start, z_vals, dists, B);
    SNRF_RETURN_LAUNCH("snrf_render_sample");
}

SNRF_API int snrf_prepare_points(const float* z_vals, const unsigned char* running_mask, const float* intersections,
                                 short* block_idxs, int B, int S, int nb, void* stream)
{
    const long long total = (long long)B * S;
    if (total <= 0 || nb <= 0) return 0;
    prepare_points_kernel<<<grid1d(total), kThreads, 0, (cudaStream_t)stream>>>(z_vals, running_mask, block_idxs, (const float2*)intersections, S, nb, total);
    SNRF_RETURN_LAUNCH("snrf_prepare_points");
}

SNRF_API int snrf_accumulate(const float* pts_diffuse, const float* pts_specular, const float* pts_alpha, float* transparency,
                             const float* z_vals, float* diffuse, float* specular, float* depth, int B, int S, void* stream)
{
    if (B <= 0) return 0;
    accumulate_kernel<<<min(snrf_div_up(B, 8), snrf_sm_count() * 64), 256, 0, (cudaStream_t)stream>>>(pts_diffuse, pts_specular, pts_alpha, transparency, z_vals, S, B, diffuse, specular, depth);
    SNRF_RETURN_LAUNCH("snrf_accumulate");
}

SNRF_API int snrf_ray_firsthit_block(const float* rays_o, const float* rays_d, const float* corners, const float* sizes,
                                     const unsigned char* grid_occupied, const long long* grid_starts, const int* grid_log2dim,
                                     const int* tracing_blocks, const float* intersections, short* hit_block_idxs, int B, int nb,
                                     void* stream)
{
    if (B <= 0 || nb <= 0) return 0;
    firsthit_block_kernel<<<grid1d(B, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, corners, sizes, grid_occupied, grid_starts, grid_log2dim,
                                                                          tracing_blocks, (const float2*)intersections, hit_block_idxs, nb, B);
    SNRF_RETURN_LAUNCH("snrf_ray_firsthit_block");
}

SNRF_API int snrf_inverse_z(const float* intersections, const short* related_bidx, float* z_vals, float sample_range, int B,
                            int nb, int S, void* stream)
{
    SNRF_CHECK_ARG(S > 1, "snrf_inverse_z: num_sample must be > 1");
    if (B <= 0) return 0;
    inverse_z_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>((const float2*)intersections, related_bidx, S, nb, sample_range, z_vals, B);
    SNRF_RETURN_LAUNCH("snrf_inverse_z");
}

SNRF_API int snrf_get_last_block(const int* tracing_blocks, int* bidxs, const float* intersections, int B, int nb, void* stream)
{
    if (B <= 0) return 0;
    last_block_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(tracing_blocks, bidxs, (const float2*)intersections, nb, B);
    SNRF_RETURN_LAUNCH("snrf_get_last_block");
}

SNRF_API int snrf_outgoing_bidx(const float* rays_o, const float* rays_d, const float* corners, const float* sizes,
                                const int* tracing_blocks, const float* intersections, short* outgoing_bidxs, float* blend_weights,
                                int skip, int B, int nb, void* stream)
{
    if (B <= 0) return 0;
    outgoing_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, corners, sizes, tracing_blocks, (const float2*)intersections,
                                                                    outgoing_bidxs, blend_weights, skip, nb, B);
    SNRF_RETURN_LAUNCH("snrf_outgoing_bidx");
}

SNRF_API int snrf_inside_bidx(const float* rays_o, const float* corners, const float* sizes, short* inside_bidxs,
                              float* blend_weights, int B, int nb, void* stream)
{
    if (B <= 0) return 0;
    inside_kernel<<<grid1d(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, corners, sizes, inside_bidxs, blend_weights, nb, B);
    SNRF_RETURN_LAUNCH("snrf_inside_bidx");
}

SNRF_API int snrf_process_occupied(int bidx, int total_grid, const float* corners, const float* sizes,
                                   const unsigned char* grid_occupied, const long long* grid_starts, const int* grid_log2dim,
                                   unsigned char* tgt_grid_occupied, int nb, void* stream)
{
    SNRF_CHECK_ARG(bidx >= 0 && bidx < nb, "snrf_process_occupied: tile index %d out of range [0,%d)", bidx, nb);
    if (total_grid <= 0) return 0;
    process_occupied_kernel<<<grid1d(total_grid), kThreads, 0, (cudaStream_t)stream>>>(bidx, nb, corners, sizes, grid_occupied, grid_starts,
                                                                                     grid_log2dim, tgt_grid_occupied, total_grid);
    SNRF_RETURN_LAUNCH("snrf_process_occupied");
}
