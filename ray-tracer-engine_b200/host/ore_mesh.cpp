// ore_mesh.cpp - see ore_mesh.h.  Behaviour follows the reference's loader and BVH builder, quirks included,
// because the leaf boxes decide which triangles a ray is tested against and the leaf order decides ties.
#include "ore_mesh.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace {
struct V3 {
    float x, y, z;
};
struct V2 {
    float u, v;
};
struct Tri {
    V3 p[3];
    V3 n;
    V3 vn[3];
    V2 vt[3];
};
static_assert(sizeof(Tri) == 27 * sizeof(float), "triangle layout");

V3 sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 cross(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// kernel.cu:102-108: divide by the DOUBLE length; a zero vector stays zero
V3 normalise(V3 v) {
    double l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (l != 0) return {(float)(v.x / l), (float)(v.y / l), (float)(v.z / l)};
    return {0, 0, 0};
}
int slash_count(const std::string& s) {
    int n = 0;
    for (char c : s) n += c == '/';
    return n;
}
// "a/b/c" -> "a b c" (kernel.cu:1073-1109)
std::istringstream split(const std::string& s) {
    std::string t = s;
    for (char& c : t)
        if (c == '/') c = ' ';
    return std::istringstream(t);
}
V3 face_normal(const std::vector<V3>& p, const int* f) {
    return normalise(cross(sub(p[f[1] - 1], p[f[0] - 1]), sub(p[f[2] - 1], p[f[0] - 1])));
}
void grow(const Tri& t, V3& lo, V3& hi) {  // getMinMaxP, kernel.cu:973-997
    for (int j = 0; j < 3; j++) {
        if (t.p[j].x > hi.x) hi.x = t.p[j].x;
        if (t.p[j].y > hi.y) hi.y = t.p[j].y;
        if (t.p[j].z > hi.z) hi.z = t.p[j].z;
        if (t.p[j].x < lo.x) lo.x = t.p[j].x;
        if (t.p[j].y < lo.y) lo.y = t.p[j].y;
        if (t.p[j].z < lo.z) lo.z = t.p[j].z;
    }
}
struct Box {
    V3 lo, hi;
    std::vector<int> idx;
};
}  // namespace

bool ore_load_obj(const std::string& path, OreMesh& out) {
    std::ifstream file(path);
    if (!file.is_open()) return false;
    std::vector<Tri> tris;
    std::vector<V3> p, vn;
    std::vector<V2> vt;
    bool has_normals = false;
    std::string line;
    char c = 0;  // like the reference's, `c` keeps its value when a line yields no token
    while (std::getline(file, line)) {
        const char type = line.size() > 1 ? line[1] : '\0';
        std::istringstream ss(line);
        char junk;
        ss >> c;
        if (c == 'v' || c == 'V') {
            if (type == 'n' || type == 'N') {
                V3 n{};
                ss >> junk >> n.x >> n.y >> n.z;
                vn.push_back(n);
            } else if (type == 't' || type == 'T') {
                V2 v{};
                ss >> junk >> v.u >> v.v;
                vt.push_back(v);
            } else {
                V3 v{};
                ss >> v.x >> v.y >> v.z;
                p.push_back(v);
            }
        }
        if (c == 'f' || c == 'F') {
            std::string in[4];
            ss >> in[0] >> in[1] >> in[2] >> in[3];
            const int sc = slash_count(line);
            if (!vn.empty() && !vt.empty()) {  // a/b/c corners
                if (sc <= 6) {
                    int f[3], pt[3], pn[3];
                    for (int k = 0; k < 3; k++) split(in[k]) >> f[k] >> pt[k] >> pn[k];
                    const V3 n = face_normal(p, f);
                    tris.push_back({{p[f[0] - 1], p[f[1] - 1], p[f[2] - 1]}, n, {vn[pn[0] - 1], vn[pn[1] - 1], vn[pn[2] - 1]},
                                    {vt[pt[0] - 1], vt[pt[1] - 1], vt[pt[2] - 1]}});
                }
                if (sc >= 8) {
                    int f[4], pt[4], pn[4];
                    for (int k = 0; k < 4; k++) split(in[k]) >> f[k] >> pt[k] >> pn[k];
                    const V3 n = face_normal(p, f);
                    tris.push_back({{p[f[0] - 1], p[f[1] - 1], p[f[2] - 1]}, n, {vn[pn[0] - 1], vn[pn[1] - 1], vn[pn[2] - 1]},
                                    {vt[pt[0] - 1], vt[pt[1] - 1], vt[pt[2] - 1]}});
                    // the second half of a quad repeats vt[2] for its last corner (kernel.cu:675)
                    tris.push_back({{p[f[0] - 1], p[f[2] - 1], p[f[3] - 1]}, n, {vn[pn[0] - 1], vn[pn[2] - 1], vn[pn[3] - 1]},
                                    {vt[pt[0] - 1], vt[pt[2] - 1], vt[pt[2] - 1]}});
                }
                has_normals = true;
            } else if (!vn.empty()) {  // a//c corners
                const V2 t0{0, 0}, t1{0, 1}, t2{1, 0};
                if (sc <= 6) {
                    int f[3], pn[3];
                    for (int k = 0; k < 3; k++) split(in[k]) >> f[k] >> pn[k];
                    const V3 n = face_normal(p, f);
                    tris.push_back({{p[f[0] - 1], p[f[1] - 1], p[f[2] - 1]}, n, {vn[pn[0] - 1], vn[pn[1] - 1], vn[pn[2] - 1]}, {t0, t1, t2}});
                }
                if (sc >= 8) {
                    int f[4], pn[4];
                    for (int k = 0; k < 4; k++) split(in[k]) >> f[k] >> pn[k];
                    const V3 n = face_normal(p, f);
                    tris.push_back({{p[f[0] - 1], p[f[1] - 1], p[f[2] - 1]}, n, {vn[pn[0] - 1], vn[pn[1] - 1], vn[pn[2] - 1]}, {t0, t1, t2}});
                    // the second half keeps the FIRST triangle's vertex normals (kernel.cu:728)
                    tris.push_back({{p[f[0] - 1], p[f[2] - 1], p[f[3] - 1]}, n, {vn[pn[0] - 1], vn[pn[1] - 1], vn[pn[2] - 1]}, {t0, t1, t2}});
                }
                has_normals = true;
            } else {  // bare indices
                const V3 z{0, 0, 0};
                if (sc <= 6) {
                    int f[3] = {std::stoi(in[0]), std::stoi(in[1]), std::stoi(in[2])};
                    const V3 n = face_normal(p, f);
                    tris.push_back({{p[f[0] - 1], p[f[1] - 1], p[f[2] - 1]}, n, {z, z, z},
                                    {V2{0.666413, 0.250594}, V2{0.333587, 0.250594}, V2{0.333587, 0.000975}}});
                }
                if (sc == 8) {
                    int f[4] = {std::stoi(in[0]), std::stoi(in[1]), std::stoi(in[2]), std::stoi(in[3])};
                    const V3 n = face_normal(p, f);
                    const V2 t0{0, 0}, t1{0, 1}, t2{1, 0};
                    tris.push_back({{p[f[0] - 1], p[f[1] - 1], p[f[2] - 1]}, n, {z, z, z}, {t0, t1, t2}});
                    tris.push_back({{p[f[0] - 1], p[f[2] - 1], p[f[3] - 1]}, n, {z, z, z}, {t0, t1, t2}});
                }
                has_normals = false;
            }
        }
    }
    out = OreMesh();
    out.has_normals = has_normals;
    out.tris.resize(tris.size() * 27);
    if (!tris.empty()) memcpy(out.tris.data(), tris.data(), tris.size() * sizeof(Tri));
    ore_build_flat_bvh(out);
    return true;
}

void ore_build_flat_bvh(OreMesh& mesh, int layers) {
    const int n = mesh.n_tris();
    mesh.box_bounds.clear();
    mesh.box_offsets.assign(1, 0);
    mesh.box_indices.clear();
    if (n == 0) return;
    const Tri* t = reinterpret_cast<const Tri*>(mesh.tris.data());
    auto bounds_of = [&](Box& b) {  // first / last triangle seed the bounds, then every vertex grows them
        b.lo = t[b.idx.front()].p[0];
        b.hi = t[b.idx.back()].p[0];
        for (int i : b.idx) grow(t[i], b.lo, b.hi);
    };
    std::vector<Box> prev(1), next;
    prev[0].idx.resize(n);
    for (int i = 0; i < n; i++) prev[0].idx[i] = i;
    bounds_of(prev[0]);
    int split_dir = 0;  // advances after every split box, not after every layer
    for (int layer = 0; layer < layers; layer++) {
        for (Box& b : prev) {
            if (b.idx.size() > 5) {
                Box lo_side, hi_side;
                float split_p;
                if (split_dir == 0)
                    split_p = (b.hi.y + b.lo.y) / 2;
                else if (split_dir == 1)
                    split_p = (b.hi.x + b.lo.x) / 2;
                else
                    split_p = (b.hi.z + b.lo.z) / 2;
                for (int i : b.idx) {
                    const V3& q = t[i].p[0];
                    const float key = split_dir == 0 ? q.y : (split_dir == 1 ? q.x : q.z);
                    (key <= split_p ? lo_side : hi_side).idx.push_back(i);
                }
                if (!lo_side.idx.empty()) {
                    bounds_of(lo_side);
                    next.push_back(std::move(lo_side));
                }
                if (!hi_side.idx.empty()) {
                    bounds_of(hi_side);
                    next.push_back(std::move(hi_side));
                }
                split_dir = (split_dir == 0 ? 1 : split_dir == 1 ? 2 : 0);
            } else {
                next.push_back(b);
            }
        }
        prev.swap(next);
        next.clear();
    }
    for (const Box& b : prev) {
        const float bb[6] = {b.lo.x, b.lo.y, b.lo.z, b.hi.x, b.hi.y, b.hi.z};
        mesh.box_bounds.insert(mesh.box_bounds.end(), bb, bb + 6);
        mesh.box_indices.insert(mesh.box_indices.end(), b.idx.begin(), b.idx.end());
        mesh.box_offsets.push_back((int)mesh.box_indices.size());
    }
}

// C hook with the signature of oracle_ref_build_mesh (tests compare the two)
extern "C" int ore_host_build_mesh(const char* obj_path, float* tris, int cap_tris, int* n_tris, int* has_normals,
                                   float* box_bounds, int* box_offsets, int cap_boxes, int* n_boxes, int* box_indices,
                                   int cap_indices) {
    OreMesh m;
    if (!ore_load_obj(obj_path, m)) return 1;
    if (m.n_tris() > cap_tris || m.n_boxes() > cap_boxes || (int)m.box_indices.size() > cap_indices) return 1;
    *n_tris = m.n_tris();
    *has_normals = m.has_normals ? 1 : 0;
    *n_boxes = m.n_boxes();
    memcpy(tris, m.tris.data(), m.tris.size() * sizeof(float));
    memcpy(box_bounds, m.box_bounds.data(), m.box_bounds.size() * sizeof(float));
    memcpy(box_offsets, m.box_offsets.data(), m.box_offsets.size() * sizeof(int));
    memcpy(box_indices, m.box_indices.data(), m.box_indices.size() * sizeof(int));
    return 0;
}
