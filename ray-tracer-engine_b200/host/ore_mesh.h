// ore_mesh.h - host-side OBJ loading and flat-BVH construction for the shim build.
//
// In the reference both live inside kernel.cu (mesh::mesh, kernel.cu:577-780; createBvhMesh, :782-936), which the
// drop-in replaces, so the shim carries its own restatement.  It produces exactly the arrays ore_set_mesh takes
// (include/ore_render.h) and is pinned against the reference's own loader/builder: identical triangles, leaf
// boxes and leaf index lists on every test file (tests/test_host_mesh.py).
#pragma once
#include <string>
#include <vector>

struct OreMesh {
    std::vector<float> tris;        // 27 floats per triangle: points[3], normal, vecNormal[3], vt[3] (kernel.cu:206-212)
    bool has_normals = false;       // mesh::has_normals
    std::vector<float> box_bounds;  // 6 floats per leaf: bounds[0], bounds[1]
    std::vector<int> box_offsets;   // leaves + 1
    std::vector<int> box_indices;   // triangle indices per leaf, in leaf order
    int n_tris() const { return (int)(tris.size() / 27); }
    int n_boxes() const { return (int)(box_bounds.size() / 6); }
};

// Parses the OBJ dialect the reference understands: v / vt / vn records and f records with 3 or 4 corners written
// as a/b/c, a//c or a.  Returns false if the file cannot be opened.
bool ore_load_obj(const std::string& path, OreMesh& out);
// Splits the triangle set `layers` times (reference: bvhLayer_count = 10) into the flat list of leaf boxes.
void ore_build_flat_bvh(OreMesh& mesh, int layers = 10);
