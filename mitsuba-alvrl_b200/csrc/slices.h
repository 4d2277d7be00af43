/*
 * slices.h -- host side of Preprocessor::buildSlices / getSlices / getSlicesPQ and Slice::sampleRepresentativePixels
 * (src/integrators/vrl/Preprocessor.cpp:66-121,1200-1227,1295-1341,1349-1418,1420-1487).
 *
 * Slice numbering is defined by the *array order* of boost::heap::priority_queue (a std::vector kept as a binary
 * max-heap with std::push_heap/std::pop_heap on operator<; Boost is a system dependency of the reference, version
 * unpinned), and by a Hoare partition whose outcome depends on the visiting order -- so this stays a sequential host
 * algorithm over the gather points the GPU produced (P x 24 bytes), exactly in the reference's order.
 */
#pragma once
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <future>
#include <memory>
#include "host_sampler.h"
#include "../../include/alvrl.h"

namespace alvrl {

struct P3 { float x, y, z; };

struct SliceInfo {
    std::vector<uint32_t> pixels;      /* gather-point (= pixel) indices in partition order */
};

class SliceTree {
    struct Node {
        uint32_t lo, hi; float diag; unsigned char dim; float split;
        float centroid[6];                                         /* positionCentroid, directionCentroid: box midpoints, 1337-1338 */
        bool operator<(const Node &o) const { return diag < o.diag; }
    };
    /* a gather point travels with its pixel index: the partition permutes these records themselves (the reference permutes
     * an index array and gathers through it), so that every pass over a node streams contiguous memory */
    struct Rec { float v[6]; uint32_t idx; };
    const std::vector<P3> &pos, &dir;
    std::unique_ptr<Rec[]> rec;                                    /* not value-initialised: every record is written before it is read */
    static bool finite3(const P3 &p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

    /* findSplitPoint, 1451-1487: midpoint of the largest extent; x wins only if strictly larger than y and z */
    static void splitPoint(const float mn[3], const float mx[3], unsigned char &dim, float &split, float &extent) {
        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        if (dx == 0 && dy == 0 && dz == 0) { extent = 0; dim = 0; split = std::numeric_limits<float>::quiet_NaN(); return; }
        int c;
        if (dx > dy) c = (dx > dz) ? 0 : 2; else c = (dy > dz) ? 1 : 2;
        const float d = c == 0 ? dx : (c == 1 ? dy : dz);
        dim = (unsigned char) c;
        split = (float) ((double) mn[c] + 0.5 * (double) d);     /* `min + 0.5*diff` is evaluated in double, 1469 */
        extent = d;
    }
    Node makeNode(uint32_t lo, uint32_t hi) const {                /* SliceNode ctor, 1301-1339 */
        Node n; n.lo = lo; n.hi = hi;
        for (int c = 0; c < 6; c++) n.centroid[c] = std::numeric_limits<float>::quiet_NaN();
        if (lo + 1 == hi) { n.diag = 0; n.dim = 0; n.split = std::numeric_limits<float>::quiet_NaN(); return n; }
        const float inf = std::numeric_limits<float>::infinity();
        float mn[6] = {inf, inf, inf, inf, inf, inf}, mx[6] = {-inf, -inf, -inf, -inf, -inf, -inf};
        const Rec *r = rec.get();
        for (uint32_t i = lo; i < hi; i++)
            for (int c = 0; c < 6; c++) {
                const float x = r[i].v[c];
                mn[c] = x < mn[c] ? x : mn[c];                     /* the reference's `if (x < min) min = x` */
                mx[c] = x > mx[c] ? x : mx[c];
            }
        const float *pmin = mn, *pmax = mx, *dmin = mn + 3, *dmax = mx + 3;
        /* sliceDistance(minPos, minDir, maxPos, maxDir), 1230-1234 */
        float dp = 0, dd = 0;
        { const float a = pmin[0] - pmax[0], b = pmin[1] - pmax[1], c = pmin[2] - pmax[2]; dp = a * a + b * b + c * c; }
        { const float a = dmin[0] - dmax[0], b = dmin[1] - dmax[1], c = dmin[2] - dmax[2]; dd = a * a + b * b + c * c; }
        n.diag = std::sqrt(dp + dd);
        unsigned char dimP, dimD; float splitP, splitD, extP, extD;
        splitPoint(pmin, pmax, dimP, splitP, extP);
        splitPoint(dmin, dmax, dimD, splitD, extD);
        if (extP > extD) { n.dim = dimP; n.split = splitP; } else { n.dim = 3 + dimD; n.split = splitD; }   /* 1442-1448 */
        for (int c = 0; c < 6; c++) n.centroid[c] = mn[c] + 0.5f * (mx[c] - mn[c]);
        return n;
    }
public:
    std::vector<float> centroids;                                  /* 6 per slice, filled by build() */
    SliceTree(const std::vector<P3> &p, const std::vector<P3> &d) : pos(p), dir(d) {}

    /* returns pixel -> slice; fills `slices` in slice-id order */
    std::vector<uint32_t> build(uint32_t targetNumSlices, std::vector<SliceInfo> &slices) {
        const uint32_t n = (uint32_t) pos.size();
        std::vector<uint32_t> toSlice(n, ALVRL_NO_SLICE);
        rec.reset(new Rec[n]);
        for (uint32_t i = 0; i < n; i++) { Rec &r = rec[i]; r.v[0] = pos[i].x; r.v[1] = pos[i].y; r.v[2] = pos[i].z; r.v[3] = dir[i].x; r.v[4] = dir[i].y; r.v[5] = dir[i].z; r.idx = i; }
        /* move the misses (non-finite gather points) to the front, 1206-1221 */
        uint32_t firstGood = 0;
        while (firstGood < n && !finite3(pos[firstGood])) firstGood++;
        for (uint32_t i = firstGood + 1; i < n; i++)
            if (!finite3(pos[i])) { std::swap(rec[i], rec[firstGood]); firstGood++; }
        slices.clear();
        if (firstGood >= n) return toSlice;
        std::vector<Node> heap;
        heap.push_back(makeNode(firstGood, n));
        while (heap.size() < targetNumSlices && heap.front().diag > 0) {      /* 1364 */
            std::pop_heap(heap.begin(), heap.end());
            const Node top = heap.back();
            heap.pop_back();
            /* Hoare partition, 1368-1393; isLarger, 1420-1430: component `dim` of (position, scaled normal) against the split */
            const int dim = top.dim; const float split = top.split;
            Rec *r = rec.get();
            size_t lo = top.lo, hi = top.hi - 1, i = lo - 1, j = hi + 1;
            for (;;) {
                do { i++; } while (!(r[i].v[dim] > split || i == hi));
                do { j--; } while (!(!(r[j].v[dim] > split) || j == lo));
                if (i >= j) break;
                std::swap(r[i], r[j]);
            }
            /* the two children's extrema are independent scans: a large node's second child is scanned by another thread */
            const uint32_t mid = (uint32_t) j + 1;
            if (top.hi - top.lo >= (1u << 16)) {
                std::future<Node> right = std::async(std::launch::async, [this, mid, &top]() { return makeNode(mid, top.hi); });
                const Node left = makeNode(top.lo, mid);
                heap.push_back(left); std::push_heap(heap.begin(), heap.end());
                heap.push_back(right.get()); std::push_heap(heap.begin(), heap.end());
            } else {
                heap.push_back(makeNode(top.lo, mid)); std::push_heap(heap.begin(), heap.end());
                heap.push_back(makeNode(mid, top.hi)); std::push_heap(heap.begin(), heap.end());
            }
        }
        centroids.clear();
        for (const Node &nd : heap) {                                         /* slice id = heap array position, 1400-1417 */
            centroids.insert(centroids.end(), nd.centroid, nd.centroid + 6);
            SliceInfo si;
            si.pixels.resize(nd.hi - nd.lo);
            for (uint32_t k = nd.lo; k < nd.hi; k++) { const uint32_t g = rec[k].idx; si.pixels[k - nd.lo] = g; toSlice[g] = (uint32_t) slices.size(); }
            slices.push_back(std::move(si));
        }
        rec.reset();
        return toSlice;
    }
};

/* Slice::sampleRepresentativePixels, Preprocessor.cpp:66-121, in terms of positions inside the slice's pixel list (the device
 * path gathers the pixel ids itself): all of them, a shuffled prefix, or rejection-sampled distinct positions */
inline std::vector<uint32_t> sampleRepresentativeIndices(size_t numPixels, float targetUndersampling, HostSampler *smp) {
    size_t targetNum = (size_t) (0.5 + numPixels / targetUndersampling);
    if (targetNum < 2) targetNum = std::min((size_t) 2, numPixels);
    std::vector<uint32_t> idx;
    if (numPixels <= targetNum) {
        idx.resize(numPixels);
        for (size_t i = 0; i < numPixels; i++) idx[i] = (uint32_t) i;
        return idx;
    }
    if (numPixels <= 2 * targetNum) {
        idx.resize(numPixels);
        for (size_t i = 0; i < numPixels; i++) idx[i] = (uint32_t) i;
        for (size_t i = numPixels - 1; i > 0; i--) std::swap(idx[i], idx[(size_t) ((i + 1) * smp->next1D())]);   /* quirk B8 */
        idx.resize(targetNum);
    } else {
        idx.resize(targetNum);
        for (size_t n = 0; n < targetNum; n++) {
            bool unique;
            do {
                idx[n] = (uint32_t) (smp->next1D() * numPixels);
                unique = true;
                for (size_t i = 0; i < n; i++) if (idx[i] == idx[n]) { unique = false; break; }
            } while (!unique);
        }
    }
    return idx;
}
inline std::vector<uint32_t> sampleRepresentativePixels(const SliceInfo &s, float targetUndersampling, HostSampler *smp) {
    const std::vector<uint32_t> idx = sampleRepresentativeIndices(s.pixels.size(), targetUndersampling, smp);
    std::vector<uint32_t> out(idx.size());
    for (size_t i = 0; i < idx.size(); i++) out[i] = s.pixels[idx[i]];
    return out;
}

} // namespace alvrl
