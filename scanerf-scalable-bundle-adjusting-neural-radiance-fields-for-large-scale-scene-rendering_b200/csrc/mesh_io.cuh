// Minimal PLY triangle-mesh reader (host).  Accepts what the reference's tinyply-based
// read_plyFile accepts for this project (fastMesh/include/plyIO.h:215-232,262-278):
// element `vertex` with float32 (or float64) x,y,z among arbitrary scalar properties and
// element `face` with a list property `vertex_indices` (or `vertex_index`) of 3 ints;
// ascii or binary_little_endian.
#pragma once
#include <string>
#include <vector>

struct HostMesh {
    std::vector<float> verts;   // 3 floats per vertex
    std::vector<int> faces;     // 3 ints per triangle
};

// returns false and fills `err` on failure
bool snrf_read_ply(const std::string& path, HostMesh& mesh, std::string& err);
