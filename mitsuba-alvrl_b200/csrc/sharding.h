/*
 * sharding.h -- slice ranges of the multi-GPU path.  The R rows, the Clustering object and the pixels of a slice depend only
 * on that slice (the reference itself deals contiguous slice ranges to its Rbuilder / ClusterRefiner threads,
 * vrlIntegrator.cpp:1049-1051, Preprocessor.cpp:739-741), so rank r owns a contiguous range of slice ids.  The reference cuts
 * by slice COUNT (id*S/w); slice sizes differ several-fold, so the ranges here are cut by PIXELS: boundary r is the slice
 * whose cumulative pixel count is closest to r/world of the total.  Host-only, shared with the CPU test shim.
 */
#pragma once
#include <cstdint>
#include <cmath>
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

/*
 * Measured-time balancing of the ranges (group.cu).  Every rank keeps the same per-slice cost estimates; a frame is cut on
 * them (cut_weights: integer weights for balanced_slice_range), the ranks exchange what their range took, and
 * correct_slice_costs scales the estimates of each rank's slices by (measured share) / (predicted share), damped: the model
 * error is taken out in the first two corrections (0.5 each), later ones follow slowly (0.2), and deviations inside a 5 % band
 * are left alone -- a rank's time varies by +-15 % from frame to frame with the order in which k_refine_mt draws its tickets.
 * Pure functions of (estimates, measured times): all ranks compute identical results from the all-reduced times.
 */
inline void initial_slice_costs(const uint32_t *sliceSize, uint32_t S, std::vector<double> &cost) {
    /* cost of a slice: its pixels (rows of R, refinement sweeps and render work grow with them) plus a constant per Clustering
     * object (picks, sorts and queue work do not depend on the rows): measured on C2, one object weighs about 1/500 of all pixels */
    uint64_t totalPix = 0;
    for (uint32_t i = 0; i < S; i++) totalPix += sliceSize[i];
    cost.resize(S);
    for (uint32_t i = 0; i < S; i++) cost[i] = (double) sliceSize[i] + (double) (totalPix / 500u);
}
inline void cut_weights(const std::vector<double> &cost, std::vector<uint32_t> &weights) {
    const uint32_t S = (uint32_t) cost.size();
    double mx = 0;
    for (uint32_t i = 0; i < S; i++) mx = cost[i] > mx ? cost[i] : mx;
    weights.resize(S);
    for (uint32_t i = 0; i < S; i++) {
        const double w = mx > 0 ? cost[i] / mx * 1048576.0 : 1.0;
        weights[i] = w < 1.0 ? 1u : (uint32_t) w;
    }
}
/* weights: what the frame was cut on; t[r]: what rank r's range took.  Returns true when a correction was counted. */
inline bool correct_slice_costs(std::vector<double> &cost, const uint32_t *weights, int world, const float *t, uint32_t corrections) {
    const uint32_t S = (uint32_t) cost.size();
    double tot = 0, totCost = 0;
    for (int r = 0; r < world; r++) tot += t[r];
    for (uint32_t i = 0; i < S; i++) totCost += cost[i];
    if (!(tot > 0) || !(totCost > 0)) return false;
    /* the scales are computed from the estimates as they were when the frame was cut, then applied */
    std::vector<double> scale((size_t) world, 1.0);
    std::vector<uint32_t> b0((size_t) world), e0((size_t) world);
    for (int r = 0; r < world; r++) {
        balanced_slice_range(weights, S, world, r, b0[r], e0[r]);
        double sum = 0;
        for (uint32_t i = b0[r]; i < e0[r]; i++) sum += cost[i];
        if (!(sum > 0) || !(t[r] > 0)) continue;
        scale[r] = (t[r] / tot) / (sum / totCost);                              /* measured share / predicted share */
    }
    const double alpha = corrections < 2u ? 0.5 : 0.2;
    for (int r = 0; r < world; r++) {
        if (std::abs(scale[r] - 1.0) < 0.05) continue;
        for (uint32_t i = b0[r]; i < e0[r]; i++) cost[i] *= (1.0 - alpha) + alpha * scale[r];
    }
    return true;
}

} // namespace alvrl
