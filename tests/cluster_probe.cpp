// Test-only C wrapper around the host-side cluster builder of the shadow sweep (csrc/ore_clusters.h).
#include "ore_clusters.h"

#include <cstring>

extern "C" int ore_probe_build_clusters(const float* ex, const float* sh, int n, float* sorted_shadow, float* sorted_exact,
                                        float* bounds, int cap_sorted, int cap_bounds) {
    std::vector<ore_host::Rec4> ss, xs, cl;
    ore_host::build_clusters(reinterpret_cast<const ore_host::Rec4*>(ex), reinterpret_cast<const ore_host::Rec4*>(sh), n, ss, xs, cl);
    if ((int)ss.size() > cap_sorted || (int)cl.size() > cap_bounds) return -1;
    if (!ss.empty()) {
        std::memcpy(sorted_shadow, ss.data(), ss.size() * sizeof(ore_host::Rec4));
        std::memcpy(sorted_exact, xs.data(), xs.size() * sizeof(ore_host::Rec4));
    }
    if (!cl.empty()) std::memcpy(bounds, cl.data(), cl.size() * sizeof(ore_host::Rec4));
    return (int)ss.size() / 32;
}

// leaves of 8 and super-clusters of 256 over the same order (round 2): returns the leaf count; bounds as float4 arrays
extern "C" int ore_probe_build_hierarchy(const float* ex, const float* sh, int n, float* sorted_shadow, int* sort_index,
                                         float* leaves, float* supers, int cap_sorted, int cap_leaves, int cap_supers) {
    std::vector<ore_host::Rec4> ss, xs, cl, lv, sp;
    std::vector<int> order;
    ore_host::build_clusters(reinterpret_cast<const ore_host::Rec4*>(ex), reinterpret_cast<const ore_host::Rec4*>(sh), n, ss, xs, cl, &order);
    ore_host::build_hierarchy(ss, n, lv, sp);
    if ((int)ss.size() > cap_sorted || (int)lv.size() > cap_leaves || (int)sp.size() > cap_supers) return -1;
    if (!ss.empty()) {
        std::memcpy(sorted_shadow, ss.data(), ss.size() * sizeof(ore_host::Rec4));
        std::memcpy(sort_index, order.data(), order.size() * sizeof(int));
    }
    std::memcpy(leaves, lv.data(), lv.size() * sizeof(ore_host::Rec4));
    std::memcpy(supers, sp.data(), sp.size() * sizeof(ore_host::Rec4));
    return (n + ore_host::LEAF_SPHERES - 1) / ore_host::LEAF_SPHERES;
}
