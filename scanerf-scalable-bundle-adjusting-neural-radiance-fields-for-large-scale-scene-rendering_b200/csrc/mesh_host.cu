// Host-side mesh ingest: PLY reader and the CPU voxeliser that seeds a tile's
// occupancy grid.  Replaces (behaviour, not code) cuda/include/voxelize.h:12-119
// (voxelize_mesh) and the tinyply front end fastMesh/include/plyIO.h.
#include "common.cuh"
#include "mesh_io.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace {

struct Prop { std::string name; std::string type; bool is_list = false; std::string count_type; };
struct Elem { std::string name; size_t count = 0; std::vector<Prop> props; };

int type_size(const std::string& t)
{
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64" || t == "int64" || t == "uint64") return 8;
    return 0;
}

double read_scalar_bin(const unsigned char* p, const std::string& t)
{
    if (t == "float" || t == "float32") { float v; memcpy(&v, p, 4); return v; }
    if (t == "double" || t == "float64") { double v; memcpy(&v, p, 8); return v; }
    if (t == "uchar" || t == "uint8") return *p;
    if (t == "char" || t == "int8") return *(const signed char*)p;
    if (t == "short" || t == "int16") { int16_t v; memcpy(&v, p, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, p, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; memcpy(&v, p, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, p, 4); return v; }
    if (t == "int64") { int64_t v; memcpy(&v, p, 8); return (double)v; }
    if (t == "uint64") { uint64_t v; memcpy(&v, p, 8); return (double)v; }
    return 0.0;
}

}  // namespace

static bool read_ply_impl(const std::string& path, HostMesh& mesh, std::string& err);

// Never throws: a malformed / truncated file (or an allocation failure on a lying header) comes back as an error string,
// so nothing propagates out of the extern "C" entry points that call this.
bool snrf_read_ply(const std::string& path, HostMesh& mesh, std::string& err)
{
    try {
        return read_ply_impl(path, mesh, err);
    } catch (const std::exception& e) {
        err = std::string("malformed PLY file (") + e.what() + "): " + path;
    } catch (...) {
        err = "malformed PLY file: " + path;
    }
    return false;
}

static bool read_ply_impl(const std::string& path, HostMesh& mesh, std::string& err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    f.seekg(0, std::ios::end);
    const long long file_size = (long long)f.tellg();
    f.seekg(0, std::ios::beg);
    std::string line;
    std::getline(f, line);
    if (line.substr(0, 3) != "ply") { err = "not a PLY file: " + path; return false; }
    bool ascii = false, big = false;
    std::vector<Elem> elems;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "format") {
            std::string fmt; ss >> fmt;
            ascii = (fmt == "ascii");
            big = (fmt == "binary_big_endian");
        } else if (tok == "element") {
            Elem e; ss >> e.name >> e.count; elems.push_back(e);
        } else if (tok == "property") {
            if (elems.empty()) { err = "property before element"; return false; }
            Prop p; std::string t; ss >> t;
            if (t == "list") { p.is_list = true; ss >> p.count_type >> p.type >> p.name; }
            else { p.type = t; ss >> p.name; }
            elems.back().props.push_back(p);
        } else if (tok == "end_header") {
            break;
        }
    }
    if (big) { err = "binary_big_endian PLY is not supported"; return false; }
    // the header is untrusted: every element needs at least one byte per property and row, and every type must be known
    {
        const long long body = file_size - (long long)f.tellg();
        long long need = 0;
        for (const Elem& e : elems) {
            if (e.props.empty() && e.count > 0) { err = "element " + e.name + " has no properties"; return false; }
            for (const Prop& p : e.props) {
                if (!ascii && type_size(p.type) == 0) { err = "unknown PLY type " + p.type; return false; }
                if (!ascii && p.is_list && type_size(p.count_type) == 0) { err = "unknown PLY list count type " + p.count_type; return false; }
            }
            if ((long long)e.count < 0 || (long long)e.count > body) { err = "element count of " + e.name + " exceeds the file size"; return false; }
            need += (long long)e.count * (long long)e.props.size();
        }
        if (need > body) { err = "PLY header declares more data than the file holds"; return false; }
    }

    for (const Elem& e : elems) {
        const bool is_v = e.name == "vertex", is_f = e.name == "face";
        int xi = -1, yi = -1, zi = -1;
        for (size_t k = 0; k < e.props.size(); ++k) {
            if (e.props[k].name == "x") xi = (int)k;
            if (e.props[k].name == "y") yi = (int)k;
            if (e.props[k].name == "z") zi = (int)k;
        }
        if (is_v) {
            if (xi < 0 || yi < 0 || zi < 0) { err = "vertex element lacks x/y/z"; return false; }
            mesh.verts.resize(e.count * 3);
        }
        if (is_f) mesh.faces.reserve(e.count * 3);
        for (size_t i = 0; i < e.count; ++i) {
            for (size_t k = 0; k < e.props.size(); ++k) {
                const Prop& p = e.props[k];
                if (!p.is_list) {
                    double v;
                    if (ascii) { f >> v; }
                    else {
                        unsigned char buf[8];
                        const int n = type_size(p.type);
                        if (n == 0) { err = "unknown PLY type " + p.type; return false; }
                        f.read((char*)buf, n);
                        v = read_scalar_bin(buf, p.type);
                    }
                    if (is_v) {
                        if ((int)k == xi) mesh.verts[3 * i + 0] = (float)v;
                        if ((int)k == yi) mesh.verts[3 * i + 1] = (float)v;
                        if ((int)k == zi) mesh.verts[3 * i + 2] = (float)v;
                    }
                } else {
                    long long n;
                    if (ascii) { f >> n; }
                    else {
                        unsigned char buf[8];
                        const int cs = type_size(p.count_type);
                        f.read((char*)buf, cs);
                        n = (long long)read_scalar_bin(buf, p.count_type);
                    }
                    const bool idx = is_f && (p.name == "vertex_indices" || p.name == "vertex_index");
                    if (!f || n < 0 || n > 255 + (file_size - (long long)f.tellg())) { err = "bad list length in element " + e.name; return false; }
                    std::vector<long long> vals((size_t)n);
                    for (long long j = 0; j < n; ++j) {
                        if (ascii) { double v; f >> v; vals[j] = (long long)v; }
                        else {
                            unsigned char buf[8];
                            const int ts = type_size(p.type);
                            f.read((char*)buf, ts);
                            vals[j] = (long long)read_scalar_bin(buf, p.type);
                        }
                    }
                    if (idx) {
                        if (n != 3) { err = "only triangle faces are supported"; return false; }
                        mesh.faces.push_back((int)vals[0]); mesh.faces.push_back((int)vals[1]); mesh.faces.push_back((int)vals[2]);
                    }
                }
            }
            if (!f) { err = "unexpected end of file in element " + e.name; return false; }
        }
    }
    const int nv = (int)(mesh.verts.size() / 3);
    for (int v : mesh.faces)
        if (v < 0 || v >= nv) { err = "face index out of range"; return false; }
    return true;
}

// Occupancy seeding on the host (all pointers are HOST pointers).
//   log2dim[3], corner[3], size[3]; vis / outside: bool[2^lx * 2^ly * 2^lz], cell index
//   n = (x << (ly+lz)) | (y << lz) | z.  Every face's AABB, inflated x1.5 about its centre, is
//   rasterised into `vis`; with init_out every cell whose centre lies outside the AABB of
//   the (inflated) in-box geometry is marked in both `vis` and `outside`.  An empty
//   model_path marks everything visible.  cuda/include/voxelize.h:12-119.
SNRF_API int snrf_voxelize_mesh_host(const int* log2dim, const float* corner, const float* size,
                                     const char* model_path, unsigned char* vis, int init_out,
                                     unsigned char* outside)
{
    const int lx = log2dim[0], ly = log2dim[1], lz = log2dim[2];
    const int rx = 1 << lx, ry = 1 << ly, rz = 1 << lz;
    const int res[3] = {rx, ry, rz};
    float cell[3], bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) {
        cell[a] = size[a] / (float)res[a];
        bmin[a] = corner[a];
        bmax[a] = corner[a] + size[a];
    }
    if (model_path == nullptr || model_path[0] == 0) {
        memset(vis, 1, (size_t)rx * ry * rz);
        return 0;
    }
    HostMesh mesh;
    std::string err;
    if (!snrf_read_ply(model_path, mesh, err)) {
        snrf_set_error("snrf_voxelize_mesh_host: %s", err.c_str());
        return (int)cudaErrorInvalidValue;
    }
    const float INF_ = 100000000.0f;
    float gmin[3] = {INF_, INF_, INF_}, gmax[3] = {-INF_, -INF_, -INF_};
    const size_t nf = mesh.faces.size() / 3;
    for (size_t i = 0; i < nf; ++i) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            const float A = mesh.verts[3 * mesh.faces[3 * i] + a], B = mesh.verts[3 * mesh.faces[3 * i + 1] + a],
                        C = mesh.verts[3 * mesh.faces[3 * i + 2] + a];
            const float mn = fminf(fminf(A, B), C), mx = fmaxf(fmaxf(A, B), C);
            const float ctr = (mn + mx) * 0.5f;          // "/ 2.0f" = multiply by the exact reciprocal
            const float half = ((mx - mn) * 1.5f) * 0.5f;
            lo[a] = ctr - half; hi[a] = ctr + half;
        }
        if (hi[0] <= bmin[0] || hi[1] <= bmin[1] || hi[2] <= bmin[2] ||
            lo[0] >= bmax[0] || lo[1] >= bmax[1] || lo[2] >= bmax[2]) continue;
        int i0[3], i1[3];
        for (int a = 0; a < 3; ++a) {
            gmin[a] = fminf(lo[a], gmin[a]); gmax[a] = fmaxf(hi[a], gmax[a]);
            i0[a] = std::min(std::max((int)((lo[a] - bmin[a]) / cell[a]), 0), res[a] - 1);
            i1[a] = std::min(std::max((int)((hi[a] - bmin[a]) / cell[a]), 0), res[a] - 1);
        }
        for (int x = i0[0]; x <= i1[0]; ++x)
            for (int y = i0[1]; y <= i1[1]; ++y)
                for (int z = i0[2]; z <= i1[2]; ++z)
                    vis[((size_t)x << (ly + lz)) | ((size_t)y << lz) | (size_t)z] = 1;
    }
    if (init_out) {
        for (int x = 0; x < rx; ++x)
            for (int y = 0; y < ry; ++y)
                for (int z = 0; z < rz; ++z) {
                    const float px = bmin[0] + (float)x * cell[0] + cell[0] * 0.5f;
                    const float py = bmin[1] + (float)y * cell[1] + cell[1] * 0.5f;
                    const float pz = bmin[2] + (float)z * cell[2] + cell[2] * 0.5f;
                    if (px < gmin[0] || py < gmin[1] || pz < gmin[2] || px > gmax[0] || py > gmax[1] || pz > gmax[2]) {
                        const size_t n = ((size_t)x << (ly + lz)) | ((size_t)y << lz) | (size_t)z;
                        vis[n] = 1;
                        if (outside) outside[n] = 1;
                    }
                }
    }
    return 0;
}
