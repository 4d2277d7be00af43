/*
 * sharding.h -- slice ranges of the multi-GPU path.  The R rows, the Clustering object and the pixels of a slice depend only
 * on that slice (the reference itself deals contiguous slice ranges to its Rbuilder / ClusterRefiner threads,
 * vrlIntegrator.cpp:1049-1051, Preprocessor.cpp:739-741), so rank r owns a contiguous range of slice ids.  The reference cuts
 * by slice COUNT (id*S/w); slice sizes differ several-fold, so the ranges here are cut by PIXELS: boundary r is the slice
 * whose cumulative pixel count is closest to r/world of the total.  Host-only, shared with the CPU test shim.
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace alvrl {

inline void balanced_slice_range(const uint32_t *sizes, uint32_t S, int world, int rank, uint32_t &begin, uint32_t &end) {
    std::vector<uint64_t> cum(S + 1, 0);
    for (uint32_t i = 0; i < S; i++) cum[i + 1] = cum[i] + sizes[i];
    std::vector<uint32_t> bounds(1, 0);
    for (int r = 1; r < world; r++) {
        const double target = (double) cum[S] * r / world;
        uint32_t b = 0;
        while (b <= S && (double) cum[b] < target) b++;                  /* first index with cum >= target */
        if (b > S) b = S;
        if (b > 0 && std::abs((double) cum[b - 1] - target) <= std::abs((double) cum[b] - target)) b--;
        if (b < bounds.back()) b = bounds.back();
        bounds.push_back(b);
    }
    bounds.push_back(S);
    begin = bounds[rank]; end = bounds[rank + 1];
}

} // namespace alvrl
