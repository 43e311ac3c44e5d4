/*
 * oracle_sort.cpp — TEST INFRASTRUCTURE ONLY.  The reference orders the calls of a merged read
 * with std::sort and a position-only comparator (src/shared/Util.cpp:3-5, Util.h:100-106); for
 * equal positions the result is whatever libstdc++'s introsort produces, so the oracle calls the
 * very same std::sort on (position, payload) records.
 */
#include <algorithm>
#include <cstdint>
#include <vector>
#include "oracle.h"

namespace {
struct rec { int32_t position; int32_t payload; };
struct by_position { bool operator()(const rec &a, const rec &b) const { return a.position < b.position; } };
}

extern "C" void orc_std_sort_by_pos(int32_t *pos, int32_t *perm, int32_t n) {
    std::vector<rec> v((size_t)n);
    for (int32_t i = 0; i < n; i++) { v[i].position = pos[i]; v[i].payload = perm[i]; }
    std::sort(v.begin(), v.end(), by_position());
    for (int32_t i = 0; i < n; i++) { pos[i] = v[i].position; perm[i] = v[i].payload; }
}
