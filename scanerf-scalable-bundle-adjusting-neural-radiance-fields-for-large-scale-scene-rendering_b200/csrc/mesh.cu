// Ray <-> proxy-mesh queries over a 64^3 uniform grid of per-cell triangle lists, sm_100a.
//
// Replaces (behaviour, not code) the reference's fastMesh extension:
//   build        fastMesh/include/tile_structure.h:13-106 (break_into_tiles), buffer.h, tile.h
//   fisrtHit     fastMesh/src/fastMesh_kernel.cu:230-329
//   firstEnter   fastMesh/src/fastMesh_kernel.cu:125-227
//   sample_points fastMesh/src/fastMesh_kernel.cu:23-122
//
// The answers are defined by the reference's cell walk, so the acceleration structure keeps its
// geometry (cubic grid anchored at the vertex AABB minimum, cell = max_extent / 64, a face is
// listed in every cell its AABB overlaps, cells are visited in DDA order from the ray origin and
// the nearest hit of the FIRST cell that has any hit is returned).  What changes is the data
// layout the walk touches on the GPU:
//   * cell rank = popcount prefix table + one popcount (the reference scans up to 4096 words
//     per occupied cell: fastMesh_kernel.cu:9-14);
//   * per-cell triangle records (9 floats, contiguous) instead of face index -> int3 -> 3 float3
//     (three dependent loads per triangle become one streaming read);
//   * per-cell face AABBs are precomputed for firstEnter;
//   * the grid header travels as a kernel argument (the reference re-uploads a __constant__ per call).
#include "common.cuh"
#include "mesh_io.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int kLog2 = 6, kSide = 64, kCells = kSide * kSide * kSide, kWords = kCells / 64;
constexpr float kInf = 100000000.0f;      // the reference's INF (macros.h)
constexpr int kThreads = 128;

struct MeshDev {
    float ox, oy, oz, cell;               // scene_info: grid origin and cell size
    const unsigned long long* bitmask;    // [kWords]
    const unsigned* word_rank;            // [kWords] number of set bits before each word
    const uint2* cell_range;              // [occupied] (start, num) into tris / face_ids
    const float* tris;                    // [entries][9]
    const float* cell_aabb;               // [occupied][6] min, max over the cell's faces
    const unsigned* face_ids;             // [entries] original face index (extension: hit id output)
};

struct Mesh {
    MeshDev dev;
    void* blob = nullptr;
    float bound[6];
    size_t n_faces = 0, n_entries = 0, n_occupied = 0;
};

// ------------------------------------------------------------------ host build
bool build_mesh(const HostMesh& hm, Mesh& m, std::string& err)
{
    const size_t nv = hm.verts.size() / 3, nf = hm.faces.size() / 3;
    if (nv == 0 || nf == 0) { err = "mesh has no vertices / faces"; return false; }
    float mn[3] = {hm.verts[0], hm.verts[1], hm.verts[2]}, mx[3] = {mn[0], mn[1], mn[2]};
    for (size_t i = 0; i < nv; ++i)
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], hm.verts[3 * i + a]); mx[a] = fmaxf(mx[a], hm.verts[3 * i + a]); }
    for (int a = 0; a < 3; ++a) { m.bound[a] = mn[a]; m.bound[3 + a] = mx[a]; }
    const float max_size = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
    const float cell = max_size / (float)kSide;
    float smax[3];
    for (int a = 0; a < 3; ++a) smax[a] = mn[a] + (float)kSide * cell;
    const float inv_cell = 1.0f / cell;            // float3 / float = multiply by the reciprocal (cutil_math.h:203-207)

    std::vector<std::vector<unsigned>> lists(kCells);
    for (size_t i = 0; i < nf; ++i) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            const float A = hm.verts[3 * hm.faces[3 * i] + a], B = hm.verts[3 * hm.faces[3 * i + 1] + a], C = hm.verts[3 * hm.faces[3 * i + 2] + a];
            const float l = fminf(fminf(A, B), C), h = fmaxf(fmaxf(A, B), C);
            const float ctr = (l + h) * 0.5f, half = (h - l) * 0.5f;      // "/ 2.0f" = * (1/2), exact
            lo[a] = ctr - half; hi[a] = ctr + half;
        }
        if (hi[0] <= mn[0] || hi[1] <= mn[1] || hi[2] <= mn[2] || lo[0] >= smax[0] || lo[1] >= smax[1] || lo[2] >= smax[2]) continue;
        int i0[3], i1[3];
        for (int a = 0; a < 3; ++a) {
            i0[a] = std::min(std::max((int)((lo[a] - mn[a]) * inv_cell), 0), kSide - 1);
            i1[a] = std::min(std::max((int)((hi[a] - mn[a]) * inv_cell), 0), kSide - 1);
        }
        for (int x = i0[0]; x <= i1[0]; ++x)
            for (int y = i0[1]; y <= i1[1]; ++y)
                for (int z = i0[2]; z <= i1[2]; ++z) lists[(x << (2 * kLog2)) | (y << kLog2) | z].push_back((unsigned)i);
    }
    std::vector<unsigned long long> bits(kWords, 0ull);
    std::vector<unsigned> rank(kWords, 0u);
    std::vector<uint2> range;
    std::vector<float> tris, aabb;
    std::vector<unsigned> ids;
    for (int n = 0; n < kCells; ++n) {
        if (lists[n].empty()) continue;
        bits[n >> 6] |= 1ull << (n & 63);
        uint2 r; r.x = (unsigned)ids.size(); r.y = (unsigned)lists[n].size();
        range.push_back(r);
        float bmin[3] = {kInf, kInf, kInf}, bmax[3] = {-kInf, -kInf, -kInf};
        for (unsigned f : lists[n]) {
            ids.push_back(f);
            for (int v = 0; v < 3; ++v)
                for (int a = 0; a < 3; ++a) {
                    const float c = hm.verts[3 * hm.faces[3 * f + v] + a];
                    tris.push_back(c);
                    bmin[a] = fminf(bmin[a], c); bmax[a] = fmaxf(bmax[a], c);
                }
        }
        for (int a = 0; a < 3; ++a) aabb.push_back(bmin[a]);
        for (int a = 0; a < 3; ++a) aabb.push_back(bmax[a]);
    }
    unsigned acc = 0;
    for (int w = 0; w < kWords; ++w) { rank[w] = acc; acc += (unsigned)__builtin_popcountll(bits[w]); }

    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_bits = 0, o_rank = align(o_bits + kWords * 8), o_range = align(o_rank + kWords * 4),
                 o_tris = align(o_range + range.size() * 8), o_aabb = align(o_tris + tris.size() * 4),
                 o_ids = align(o_aabb + aabb.size() * 4), total = align(o_ids + ids.size() * 4);
    std::vector<unsigned char> host(total, 0);
    memcpy(host.data() + o_bits, bits.data(), kWords * 8);
    memcpy(host.data() + o_rank, rank.data(), kWords * 4);
    if (!range.empty()) memcpy(host.data() + o_range, range.data(), range.size() * 8);
    if (!tris.empty()) memcpy(host.data() + o_tris, tris.data(), tris.size() * 4);
    if (!aabb.empty()) memcpy(host.data() + o_aabb, aabb.data(), aabb.size() * 4);
    if (!ids.empty()) memcpy(host.data() + o_ids, ids.data(), ids.size() * 4);
    cudaError_t e = cudaMalloc(&m.blob, total);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return false; }
    e = cudaMemcpy(m.blob, host.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); cudaFree(m.blob); m.blob = nullptr; return false; }
    unsigned char* b = (unsigned char*)m.blob;
    m.dev = MeshDev{mn[0], mn[1], mn[2], cell, (const unsigned long long*)(b + o_bits), (const unsigned*)(b + o_rank),
                    (const uint2*)(b + o_range), (const float*)(b + o_tris), (const float*)(b + o_aabb), (const unsigned*)(b + o_ids)};
    m.n_faces = nf; m.n_entries = ids.size(); m.n_occupied = range.size();
    return true;
}

// ------------------------------------------------------------------ device walk
// Cell walk of fastMesh/include/dda.h:7-131 with cell size 1 in index space (o is in cell
// units, d stays in world units so t is world t / cell).  Tie rule as written there; its
// all-zero mask on an exact three-way tie would never advance, here that case steps in z.
struct CellWalk {
    int cx, cy, cz, sx, sy, sz, mx, my, mz;
    float tmx, tmy, tmz, tdx, tdy, tdz, t0, t1;
    __device__ __forceinline__ void init(f3 o, f3 d, float t_start)
    {
        o = o + t_start * d;
        cx = min(max((int)o.x, 0), kSide - 1);
        cy = min(max((int)o.y, 0), kSide - 1);
        cz = min(max((int)o.z, 0), kSide - 1);
        sx = sign_pos0(d.x); sy = sign_pos0(d.y); sz = sign_pos0(d.z);
        float bx = (float)(cx + sx), by = (float)(cy + sy), bz = (float)(cz + sz);
        if (sx < 0) bx += 1.0f;
        if (sy < 0) by += 1.0f;
        if (sz < 0) bz += 1.0f;
        t0 = t_start; t1 = 0.0f;
        tmx = fmaxf(safe_div(bx - o.x, d.x), 0.0f) + t0;
        tmy = fmaxf(safe_div(by - o.y, d.y), 0.0f) + t0;
        tmz = fmaxf(safe_div(bz - o.z, d.z), 0.0f) + t0;
        tdx = fabsf(safe_div(1.0f, d.x)); tdy = fabsf(safe_div(1.0f, d.y)); tdz = fabsf(safe_div(1.0f, d.z));
    }
    __device__ __forceinline__ void next()
    {
        mx = (tmx < tmy) & (tmx <= tmz);
        my = (tmy < tmz) & (tmy <= tmx);
        mz = (tmz < tmx) & (tmz <= tmy);
        t1 = mx ? tmx : (my ? tmy : tmz);
        if (!(mx | my | mz)) mz = 1;
    }
    __device__ __forceinline__ void step()
    {
        t0 = t1;
        tmx += (float)mx * tdx; tmy += (float)my * tdy; tmz += (float)mz * tdz;
        cx += mx * sx; cy += my * sy; cz += mz * sz;
    }
    __device__ __forceinline__ bool done() const
    {
        return cx < 0 || cy < 0 || cz < 0 || cx >= kSide || cy >= kSide || cz >= kSide || (tmx <= 0 && tmy <= 0 && tmz <= 0);
    }
    __device__ __forceinline__ unsigned cell() const { return ((unsigned)cx << (2 * kLog2)) | ((unsigned)cy << kLog2) | (unsigned)cz; }
};

__device__ __forceinline__ bool cell_on(const MeshDev& m, unsigned n) { return (m.bitmask[n >> 6] >> (n & 63)) & 1ull; }
__device__ __forceinline__ unsigned cell_rank(const MeshDev& m, unsigned n)
{
    return m.word_rank[n >> 6] + (unsigned)__popcll(m.bitmask[n >> 6] & ((1ull << (n & 63)) - 1ull));
}

// Moeller-Trumbore with back-face rejection, fastMesh/include/cuda_utils.h:403-436 (blur = 0)
__device__ __forceinline__ float ray_triangle(f3 o, f3 d, f3 v0, f3 v1, f3 v2)
{
    const f3 e1 = v1 - v0, e2 = v2 - v0, s = o - v0;
    const f3 p = cross3(d, e2);
    float det = dot3(e1, p);
    if (det < 0.00000001f) return -kInf;
    det = 1.0f / det;
    const float u = dot3(s, p) * det;
    if (u < 0.0f || u > 1.0f) return -kInf;
    const f3 q = cross3(s, e1);
    const float v = dot3(d, q) * det;
    if (v < 0.0f || v > 1.0f) return -kInf;
    if ((u + v) < 0.0f || (u + v) > 1.0f) return -kInf;
    return dot3(e2, q) * det;
}

__device__ __forceinline__ f3 to_index_space(const MeshDev& m, f3 o)
{
    const float inv = 1.0f / m.cell;                  // float3 / float in the reference = * (1/s)
    return mk3((o.x - m.ox) * inv, (o.y - m.oy) * inv, (o.z - m.oz) * inv);
}

__global__ void __launch_bounds__(kThreads)
first_hit_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ z_depth,
                 int* __restrict__ hit_face, MeshDev m, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        CellWalk w;
        w.init(to_index_space(m, o), d, 0.0f);
        float t_near = kInf;
        int face = -1;
        while (!w.done()) {
            w.next();
            const unsigned n = w.cell();
            if (cell_on(m, n)) {
                const uint2 r = m.cell_range[cell_rank(m, n)];
                const float* tri = m.tris + 9 * (size_t)r.x;
                for (unsigned k = 0; k < r.y; ++k, tri += 9) {
                    const f3 A = ld3(tri), Bv = ld3(tri + 3), C = ld3(tri + 6);
                    // two-sided: flip the winding when the ray looks along the face normal
                    const f3 nrm = cross3(Bv - A, C - Bv);
                    const float t = dot3(d, nrm) > 0.0f ? ray_triangle(o, d, A, C, Bv) : ray_triangle(o, d, A, Bv, C);
                    if (t > 0.0f && t < t_near) { t_near = t; face = (int)m.face_ids[r.x + k]; }
                }
                if (t_near != kInf) break;
            }
            w.step();
        }
        z_depth[i] = t_near != kInf ? t_near : 0.0f;
        if (hit_face) hit_face[i] = t_near != kInf ? face : -1;
    }
}

__global__ void __launch_bounds__(kThreads)
first_enter_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ z_depth,
                   MeshDev m, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const f3 o = ld3(rays_o + 3 * (size_t)i), d = ld3(rays_d + 3 * (size_t)i);
        CellWalk w;
        w.init(to_index_space(m, o), d, 0.0f);
        float out = 0.0f;
        while (!w.done()) {
            w.next();
            const unsigned n = w.cell();
            if (cell_on(m, n)) {
                const float* bb = m.cell_aabb + 6 * (size_t)cell_rank(m, n);
                const f3 lo = ld3(bb), hi = ld3(bb + 3);
                const f3 center = (hi + lo) * 0.5f, half = (hi - lo) * 0.5f;
                const float2 tb = ray_aabb(o, d, center, half);
                if (tb.x >= 0.0f) { out = tb.x; break; }
            }
            w.step();
        }
        z_depth[i] = out;
    }
}

__global__ void __launch_bounds__(kThreads)
mesh_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ t_start,
                   float* __restrict__ z_vals, int S, MeshDev m, int B)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        float ts = t_start[i];
        if (ts == -1.0f) continue;
        const f3 o = to_index_space(m, ld3(rays_o + 3 * (size_t)i)), d = ld3(rays_d + 3 * (size_t)i);
        ts = ts / m.cell;                                  // scalar / scalar: a true divide in the reference
        CellWalk w;
        w.init(o, d, ts);
        int count = 0;
        float total = 0.0f;
        while (!w.done()) {
            w.next();
            if (cell_on(m, w.cell())) { const float len = w.t1 - w.t0; if (len > 0) { total += len; ++count; } }
            w.step();
        }
        if (count == 0) continue;
        w.init(o, d, ts);
        int left = S, seen = 0;
        float* zr = z_vals + (size_t)i * S;
        while (!w.done()) {
            w.next();
            if (cell_on(m, w.cell())) {
                const float len = w.t1 - w.t0;
                if (len > 0) {
                    int num = min(max((int)(S * len / total), 1), left);
                    if (seen == count - 1) num = left;
                    const float interval = (w.t1 - w.t0) / num;          // uniform_sample_bound_v2
                    for (int k = 0; k < num; ++k) zr[S - left + k] = w.t0 + k * interval;
                    left -= num;
                    ++seen;
                }
            }
            w.step();
        }
        for (int k = 0; k < S; ++k) zr[k] = zr[k] * m.cell;
    }
}

inline int grid_for(int B)
{
    int g = snrf_div_up(B, kThreads);
    const int cap = snrf_sm_count() * 16;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
// fastMesh::build (fastMesh/include/fastMesh.h:22-26): PLY -> grid on the current device.
SNRF_API int snrf_mesh_create(const char* ply_path, void** handle)
{
    SNRF_CHECK_ARG(ply_path != nullptr && handle != nullptr, "snrf_mesh_create: null argument");
    HostMesh hm;
    std::string err;
    if (!snrf_read_ply(ply_path, hm, err)) { snrf_set_error("snrf_mesh_create: %s", err.c_str()); return (int)cudaErrorInvalidValue; }
    Mesh* m = new Mesh();
    if (!build_mesh(hm, *m, err)) { snrf_set_error("snrf_mesh_create: %s", err.c_str()); delete m; return (int)cudaErrorInvalidValue; }
    *handle = m;
    return 0;
}
// Same from host arrays (verts[nv,3] f32, faces[nf,3] i32) -- used by tests.
SNRF_API int snrf_mesh_create_from_arrays(const float* verts_host, int nv, const int* faces_host, int nf, void** handle)
{
    SNRF_CHECK_ARG(verts_host && faces_host && handle && nv > 0 && nf > 0, "snrf_mesh_create_from_arrays: bad argument");
    HostMesh hm;
    hm.verts.assign(verts_host, verts_host + 3 * (size_t)nv);
    hm.faces.assign(faces_host, faces_host + 3 * (size_t)nf);
    for (int v : hm.faces)
        if (v < 0 || v >= nv) { snrf_set_error("snrf_mesh_create_from_arrays: face index out of range"); return (int)cudaErrorInvalidValue; }
    std::string err;
    Mesh* m = new Mesh();
    if (!build_mesh(hm, *m, err)) { snrf_set_error("snrf_mesh_create_from_arrays: %s", err.c_str()); delete m; return (int)cudaErrorInvalidValue; }
    *handle = m;
    return 0;
}
// fastMesh::destroy (fastMesh.h:52-55)
SNRF_API int snrf_mesh_destroy(void* handle)
{
    Mesh* m = (Mesh*)handle;
    if (!m) return 0;
    if (m->blob) cudaFree(m->blob);
    delete m;
    return 0;
}
// fastMesh::getSceneBound (fastMesh.h:28-38): min xyz, max xyz of the vertices (host floats)
SNRF_API int snrf_mesh_bounds(void* handle, float* bound6_host)
{
    SNRF_CHECK_ARG(handle && bound6_host, "snrf_mesh_bounds: null argument");
    memcpy(bound6_host, ((Mesh*)handle)->bound, 6 * sizeof(float));
    return 0;
}
// occupied cells, face-list entries, faces (host ints) -- introspection for tests
SNRF_API int snrf_mesh_stats(void* handle, long long* stats3_host)
{
    SNRF_CHECK_ARG(handle && stats3_host, "snrf_mesh_stats: null argument");
    Mesh* m = (Mesh*)handle;
    stats3_host[0] = (long long)m->n_occupied; stats3_host[1] = (long long)m->n_entries; stats3_host[2] = (long long)m->n_faces;
    return 0;
}
// fastMesh::fisrtHit (sic): z_depth[B] = nearest t > 0 in the first cell with a hit, 0 = miss;
// hit_face[B] (optional, extension) = index of that face or -1.
SNRF_API int snrf_mesh_first_hit(void* handle, const float* rays_o, const float* rays_d, float* z_depth, int* hit_face,
                                 int B, void* stream)
{
    SNRF_CHECK_ARG(handle, "snrf_mesh_first_hit: null mesh");
    if (B <= 0) return 0;
    first_hit_kernel<<<grid_for(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z_depth, hit_face, ((Mesh*)handle)->dev, B);
    SNRF_RETURN_LAUNCH("snrf_mesh_first_hit");
}
// fastMesh::firstEnter: entry t of the face AABB of the first occupied cell whose AABB is hit, 0 = none
SNRF_API int snrf_mesh_first_enter(void* handle, const float* rays_o, const float* rays_d, float* z_depth, int B, void* stream)
{
    SNRF_CHECK_ARG(handle, "snrf_mesh_first_enter: null mesh");
    if (B <= 0) return 0;
    first_enter_kernel<<<grid_for(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z_depth, ((Mesh*)handle)->dev, B);
    SNRF_RETURN_LAUNCH("snrf_mesh_first_enter");
}
// fastMesh::sample_points: occupancy-proportional samples from t_start[B] (-1 = skip ray); z_vals[B,S]
SNRF_API int snrf_mesh_sample(void* handle, const float* rays_o, const float* rays_d, const float* t_start, float* z_vals,
                              int B, int S, void* stream)
{
    SNRF_CHECK_ARG(handle && S > 0, "snrf_mesh_sample: null mesh or S <= 0");
    if (B <= 0) return 0;
    mesh_sample_kernel<<<grid_for(B), kThreads, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_start, z_vals, S, ((Mesh*)handle)->dev, B);
    SNRF_RETURN_LAUNCH("snrf_mesh_sample");
}
