// ore_clusters.h - host-only: shadow-sweep clusters of the sphere set (no CUDA types; also compiled by the CPU tests).
//
// The spheres in Morton order of their centres with two levels of bounding balls over them: LEAVES of 8 consecutive
// spheres and SUPER-clusters of 32 consecutive leaves (build_hierarchy below; build_clusters also returns round 1's
// 32-sphere cluster balls, which only the containment test still looks at).  Neither the any-hit result of a shadow ray
// nor - with candidates adjudicated by (t, original index) - the nearest hit depends on the order spheres are visited in,
// so the shadow sweep and the primary kernel test supers first, open the leaves of the survivors and only then single
// spheres.  Soundness rests on one property, which tests/test_clusters_cpu.py checks: every member ball
// (centre, R') lies inside the ball of its leaf AND inside the ball of its super-cluster, or that radius is +inf
// ("always open": a member with non-finite or >= 1e15 coordinates / radius).
#ifndef ORE_CLUSTERS_H
#define ORE_CLUSTERS_H
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace ore_host {

struct Rec4 {  // layout of float4
    float x, y, z, w;
};

inline bool tame_value(float v) { return std::isfinite(v) && std::fabs(v) < 1e15f; }

// ex / sh: n exact records (cx,cy,cz,radius member) and n shadow records (cx,cy,cz,R').  Outputs: sorted_shadow and
// sorted_exact with ceil(n/32)*32 entries (the tail repeats the last sphere; the kernels never read index >= n),
// bounds with ceil(n/32) rounded up to a multiple of 4 entries (centre, radius; padding = zeros) - the 32-sphere
// clusters of round 1, kept for the containment test - and, optionally, the permutation itself.
inline void build_clusters(const Rec4* ex, const Rec4* sh, int n, std::vector<Rec4>& sorted_shadow,
                           std::vector<Rec4>& sorted_exact, std::vector<Rec4>& bounds, std::vector<int>* sort_index = nullptr) {
    const int n_clu = (n + 31) / 32;
    const size_t n_sort = (size_t)n_clu * 32, n_clu_pad = ((size_t)n_clu + 3) & ~(size_t)3;
    sorted_shadow.assign(n_sort, Rec4{0.f, 0.f, 0.f, 0.f});
    sorted_exact.assign(n_sort, Rec4{0.f, 0.f, 0.f, 0.f});
    bounds.assign(n_clu_pad, Rec4{0.f, 0.f, 0.f, 0.f});
    if (sort_index) sort_index->assign(n_sort, 0);
    if (n <= 0) return;

    // Morton keys (10 bits per axis) over the bounding box of the tame centres
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++) {
        const float c[3] = {sh[i].x, sh[i].y, sh[i].z};
        for (int k = 0; k < 3; k++)
            if (tame_value(c[k])) {
                lo[k] = std::min(lo[k], (double)c[k]);
                hi[k] = std::max(hi[k], (double)c[k]);
            }
    }
    auto spread = [](uint32_t v) {
        v &= 1023u;
        v = (v | (v << 16)) & 0x030000FFu;
        v = (v | (v << 8)) & 0x0300F00Fu;
        v = (v | (v << 4)) & 0x030C30C3u;
        v = (v | (v << 2)) & 0x09249249u;
        return v;
    };
    std::vector<std::pair<uint32_t, int>> order((size_t)n);
    for (int i = 0; i < n; i++) {
        const float c[3] = {sh[i].x, sh[i].y, sh[i].z};
        uint32_t q[3];
        for (int k = 0; k < 3; k++) {
            double t = 0.0;
            if (tame_value(c[k]) && hi[k] > lo[k]) t = ((double)c[k] - lo[k]) / (hi[k] - lo[k]);
            q[k] = (uint32_t)std::min(1023.0, std::max(0.0, t * 1023.0));
        }
        order[(size_t)i] = {spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2), i};
    }
    std::stable_sort(order.begin(), order.end(), [](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) {
        return a.first < b.first;
    });

    for (size_t p = 0; p < n_sort; p++) {
        const int i = order[p < (size_t)n ? p : (size_t)n - 1].second;
        sorted_shadow[p] = sh[i];
        sorted_exact[p] = ex[i];
        if (sort_index) (*sort_index)[p] = i;   // original index of the sphere at sorted position p
    }
    for (size_t j = 0; j < (size_t)n_clu; j++) {
        const size_t p0 = j * 32, p1 = std::min(p0 + 32, (size_t)n);
        double cx = 0, cy = 0, cz = 0;
        bool tame = true;
        for (size_t p = p0; p < p1; p++) {
            const Rec4& s = sorted_shadow[p];
            tame = tame && tame_value(s.x) && tame_value(s.y) && tame_value(s.z) && tame_value(s.w);
            cx += s.x;
            cy += s.y;
            cz += s.z;
        }
        const double m = (double)(p1 - p0);
        const float fx = (float)(cx / m), fy = (float)(cy / m), fz = (float)(cz / m);
        double rad = 0;
        for (size_t p = p0; p < p1 && tame; p++) {
            const Rec4& s = sorted_shadow[p];
            const double dx = (double)s.x - fx, dy = (double)s.y - fy, dz = (double)s.z - fz;
            rad = std::max(rad, std::sqrt(dx * dx + dy * dy + dz * dz) + (double)s.w);
        }
        float fr = INFINITY;  // "always open": a member the float tests cannot bound
        if (tame && std::isfinite(rad) && rad < 1e15) fr = std::nextafter((float)(rad * (1.0 + 1e-6) + 1e-30), INFINITY);
        bounds[j] = Rec4{fx, fy, fz, fr};
    }
}

// Bounding ball of the member balls sorted[p0, p1): centre = the better of (mean of the centres, centre of the box
// around the member balls), radius = max |c_i - centre| + R'_i, rounded up.  +inf ("always open") when a member
// cannot be bounded by the float tests.
inline Rec4 bounding_ball(const std::vector<Rec4>& sorted, size_t p0, size_t p1) {
    if (p1 <= p0) return Rec4{0.f, 0.f, 0.f, 0.f};
    double mean[3] = {0, 0, 0}, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    bool tame = true;
    for (size_t p = p0; p < p1; p++) {
        const Rec4& s = sorted[p];
        tame = tame && tame_value(s.x) && tame_value(s.y) && tame_value(s.z) && tame_value(s.w);
        const double c[3] = {s.x, s.y, s.z};
        for (int k = 0; k < 3; k++) {
            mean[k] += c[k];
            lo[k] = std::min(lo[k], c[k] - std::fabs((double)s.w));
            hi[k] = std::max(hi[k], c[k] + std::fabs((double)s.w));
        }
    }
    if (!tame) return Rec4{(float)0, (float)0, (float)0, INFINITY};
    const double m = (double)(p1 - p0);
    Rec4 best{0.f, 0.f, 0.f, INFINITY};
    for (int cand = 0; cand < 2; cand++) {
        const float f[3] = {(float)(cand ? 0.5 * (lo[0] + hi[0]) : mean[0] / m), (float)(cand ? 0.5 * (lo[1] + hi[1]) : mean[1] / m),
                            (float)(cand ? 0.5 * (lo[2] + hi[2]) : mean[2] / m)};
        double rad = 0;
        for (size_t p = p0; p < p1; p++) {
            const Rec4& s = sorted[p];
            const double dx = (double)s.x - f[0], dy = (double)s.y - f[1], dz = (double)s.z - f[2];
            rad = std::max(rad, std::sqrt(dx * dx + dy * dy + dz * dz) + std::fabs((double)s.w));
        }
        if (std::isfinite(rad) && rad < 1e15) {
            const float fr = std::nextafter((float)(rad * (1.0 + 1e-6) + 1e-30), INFINITY);
            if (fr < best.w) best = Rec4{f[0], f[1], f[2], fr};
        }
    }
    return best;
}

// Two more levels over the same Morton order (round 2, shadow_sweep_kernel): LEAVES of 8 consecutive spheres and
// SUPER-clusters of 32 consecutive leaves (256 spheres), one bounding ball each.  leaves: ceil(n/8) entries rounded up
// to a multiple of 32 (padding = radius -1: never touched); supers: ceil(n/256) entries rounded up to a multiple of 4.
constexpr int LEAF_SPHERES = 8;
constexpr int SUPER_LEAVES = 32;
inline void build_hierarchy(const std::vector<Rec4>& sorted_shadow, int n, std::vector<Rec4>& leaves, std::vector<Rec4>& supers) {
    const size_t n_leaf = ((size_t)std::max(n, 0) + LEAF_SPHERES - 1) / LEAF_SPHERES;
    const size_t n_sup = (n_leaf + SUPER_LEAVES - 1) / SUPER_LEAVES;
    leaves.assign((n_leaf + 31) & ~(size_t)31, Rec4{0.f, 0.f, 0.f, -1.f});
    supers.assign((n_sup + 3) & ~(size_t)3, Rec4{0.f, 0.f, 0.f, -1.f});
    for (size_t j = 0; j < n_leaf; j++)
        leaves[j] = bounding_ball(sorted_shadow, j * LEAF_SPHERES, std::min((j + 1) * LEAF_SPHERES, (size_t)n));
    const size_t per = (size_t)LEAF_SPHERES * SUPER_LEAVES;
    for (size_t k = 0; k < n_sup; k++) supers[k] = bounding_ball(sorted_shadow, k * per, std::min((k + 1) * per, (size_t)n));
}

}  // namespace ore_host
#endif
