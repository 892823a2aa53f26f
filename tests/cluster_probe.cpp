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
