// Host-only: the permutation std::sort (libstdc++ introsort) applies to n elements whose keys
// all compare equal — see pg2_render.cuh `g_sort_perm` and SURVEY Q5. Uses the same element
// type and comparator as the reference (std::pair<float, Entity>, compare .first).
#pragma once
#include <stdint.h>
#include <algorithm>
#include <utility>
#include <vector>

namespace pg2 {
inline std::vector<uint8_t> build_sort_perm(int maxn) {
    std::vector<uint8_t> table((size_t)(maxn + 1) * maxn, 0);
    for (int n = 0; n <= maxn; n++) {
        std::vector<std::pair<float, int>> v(n);
        for (int i = 0; i < n; i++) v[i] = std::make_pair(1.0f, i);
        std::sort(v.begin(), v.end(), [](const std::pair<float, int>& l, const std::pair<float, int>& r) { return l.first < r.first; });
        for (int i = 0; i < n; i++) table[(size_t)n * maxn + i] = (uint8_t)v[i].second;
    }
    return table;
}
}  // namespace pg2
