/*
 * oracle_prep.hpp -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of src/integrators/vrl/Preprocessor.cpp: slice construction, representative
 * pixels, and VRL column clustering.  Third-party arithmetic the reference pulls from Boost
 * (system package, version unpinned by the reference: build/config-linux-gcc.py:19) is restated
 * from its published behaviour:
 *   - boost::heap::priority_queue<T>  = std::vector<T> + std::push_heap / std::pop_heap on
 *     operator<, iterated in underlying-array order;
 *   - boost::numeric::ublas vector ops (inner_prod, norm_1, norm_2, element_prod) = plain
 *     left-to-right loops in the vector's value type.
 * No reference test pins any of this ("parity unpinned"); this file is the definition.
 * The two by-value copies of R in the reference (Preprocessor.cpp:146-155,780; quirk B6) are
 * replaced by views -- results are identical.
 */
#pragma once
#include "oracle_core.hpp"
#include <list>
#include <utility>
#include <numeric>

namespace orc {

struct VrlContribution { Float mean, var; };                               // VRL.h:8-12

/* boost::heap::priority_queue restated */
template <typename T> struct BoostPQ {
    std::vector<T> q;
    void push(const T &v) { q.push_back(v); std::push_heap(q.begin(), q.end()); }
    const T &top() const { return q.front(); }
    void pop() { std::pop_heap(q.begin(), q.end()); q.pop_back(); }
    size_t size() const { return q.size(); }
};

/* A matrix view: rows are pointers to N contiguous VrlContributions */
struct MatView {
    std::vector<const VrlContribution *> rows;
    uint32_t nVrls = 0;
    size_t numRows() const { return rows.size(); }
};

struct ClusterStats { uint64_t splits = 0, varianceSteps = 0; };

/* Preprocessor::weightedSample, Preprocessor.cpp:1534-1580 */
inline size_t weightedSample(const std::vector<Float> &weights, Sampler *sampler, Float *prob,
                             size_t begin, size_t end, const std::vector<uint32_t> *ind) {
    if (begin >= end) fail("Trying to take weighted sample of empty set!");
    if (end == begin + 1) { if (prob) *prob = 1; return begin; }
    Float weightSum = 0.0f;
    for (size_t i = begin; i < end; i++) weightSum += weights[ind ? ind->at(i) : i];
    Float probability; size_t idx;
    if (weightSum <= 0) {
        do { idx = begin + sampler->next1D() * (end - begin); } while (idx >= end);
        probability = 1.0 / (end - begin);
    } else {
        Float alpha = sampler->next1D() * weightSum;
        Float accum = 0.0f;
        idx = begin;
        for (size_t i = begin; i < end; i++) {
            accum += weights[ind ? ind->at(i) : i];
            if (accum >= alpha) { idx = i; break; }
        }
        probability = weights[ind ? ind->at(idx) : idx] / weightSum;
    }
    if (prob) *prob = probability;
    return idx;
}

/* Preprocessor.cpp:985-1008 */
inline void calculateColumnWeights(const MatView &M, const std::vector<double> &lw, std::vector<Float> &cw, Float safetyFraction = 1e-2) {
    cw.resize(M.nVrls);
    for (size_t vrl = 0; vrl < M.nVrls; vrl++) {
        double acc = 0;
        for (size_t r = 0; r < M.numRows(); r++) {
            double mean = M.rows[r][vrl].mean, var = M.rows[r][vrl].var;
            double x = mean * mean + var;
            acc += lw[r] * x;
        }
        cw[vrl] = (Float) safe_sqrt(acc);
        if (!std::isfinite(cw[vrl])) fail("Invalid calculated average column weight");
    }
    Float averageWeight = std::accumulate(cw.begin(), cw.end(), 0.0f) / M.nVrls;
    if (averageWeight == 0) averageWeight = 1.0;
    for (size_t vrl = 0; vrl < M.nVrls; vrl++) cw[vrl] += averageWeight * safetyFraction;
}

/* Preprocessor.cpp:1022-1048 */
template <typename It>
inline void calculateUnclusteredVariance(const MatView &R, const std::vector<double> &lw, It begin, It end,
                                         Float &tracerVariance, Float &vrlIntegrationVariance) {
    size_t n = 0, nr = R.numRows();
    std::vector<double> mean(nr, 0), M2(nr, 0), summedVars(nr, 0);
    for (It it = begin; it != end; ++it) {
        n++;
        uint32_t v = *it;
        for (size_t r = 0; r < nr; r++) {
            summedVars[r] += (double) R.rows[r][v].var;
            double x = R.rows[r][v].mean;
            double delta = x - mean[r];
            mean[r] += delta / n;
            M2[r] += delta * (x - mean[r]);
        }
    }
    if (n <= 1) fail("Need at least 2 VRLs to estimate variance");
    double a = 0, b = 0;
    for (size_t r = 0; r < nr; r++) a += lw[r] * summedVars[r];
    vrlIntegrationVariance = (Float) a;
    for (size_t r = 0; r < nr; r++) b += lw[r] * M2[r];
    tracerVariance = (Float) (b - vrlIntegrationVariance);
}

/* Preprocessor.cpp:1058-1120 */
template <typename It>
inline std::pair<Float, Float> calculateClusterVariance(const MatView &R, const std::vector<Float> &cw,
        const std::vector<double> &lw, It begin, It end, std::vector<std::pair<Float, Float>> *incrementalVar = nullptr,
        ClusterStats *stats = nullptr) {
    if (begin == end) fail("taking variance of empty cluster!");
    size_t nr = R.numRows();
    std::vector<double> sum(nr, 0), M(nr, 0), sumVars(nr, 0);
    double weightSum = 0;
    int n = 0;
    auto innerM = [&]() { double t = 0; for (size_t r = 0; r < nr; r++) t += lw[r] * (M[r] / weightSum); return t; };
    auto innerV = [&]() { double t = 0; for (size_t r = 0; r < nr; r++) t += lw[r] * (sumVars[r] * weightSum); return t; };
    for (It it = begin; it != end; ++it) {
        uint32_t vrl = *it;
        double weight = cw[vrl];
        if (!std::isfinite(weight) || weight <= 0) fail("Invalid weight in calculateClusterVariance()");
        double newWeightSum = weightSum + weight;
        double c1 = 0, c2 = 0;
        if (n > 0) { c1 = (newWeightSum * newWeightSum) / (weightSum * weightSum); c2 = (1.0 / weight + 1.0 / weightSum); }
        for (size_t r = 0; r < nr; r++) {
            double x = R.rows[r][vrl].mean;
            double newSum = sum[r] + x;
            double tmp = weight * sum[r] - weightSum * x;
            if (n > 0) M[r] = c1 * M[r] + c2 * (tmp * tmp);
            sumVars[r] += (double) R.rows[r][vrl].var / weight;
            sum[r] = newSum;
        }
        weightSum = newWeightSum;
        if (incrementalVar) {
            incrementalVar->at(n).first = (n == 0) ? 0 : (Float) innerM();
            incrementalVar->at(n).second = (Float) innerV();
        }
        if (stats) stats->varianceSteps++;
        n++;
    }
    std::pair<Float, Float> result;
    result.first = (Float) innerM();
    result.second = (Float) innerV();
    if (!std::isfinite(result.first) || result.first < 0) fail("invalid undersampled VRL cluster variance");
    if (!std::isfinite(result.second) || result.second < 0) fail("invalid undersampled VRL integration cluster variance");
    return result;
}

/* Preprocessor::Clustering, Preprocessor.cpp:287-720 */
class Clustering {
    struct ClusterNode {
        Float undersamplingVar, integrationVar; uint32_t begin, end;
        bool operator<(const ClusterNode &o) const { return undersamplingVar + integrationVar < o.undersamplingVar + o.integrationVar; }
    };
public:
    Clustering(const std::vector<std::vector<uint32_t>> &vrlsPerCluster, const MatView &M, const std::vector<double> &lw,
               Float pixelUndersampling, Float depthCorrection = 1, ClusterStats *stats = nullptr)
        : m_M(M), m_lw(lw), m_pixelUndersampling(pixelUndersampling), m_depthCorrection(depthCorrection), m_stats(stats) {
        double n1 = 0; for (double w : lw) n1 += std::abs(w);
        Float norm = (Float) n1;
        if (std::fabs(norm - 1) > 1e-3) fail("Incorrect normalization in localityWeights");
        if (m_pixelUndersampling <= 0 || m_pixelUndersampling > 1) fail("Invalid pixel undersampling");
        calculateColumnWeights(m_M, m_lw, m_columnWeights);
        m_clusterUndersamplingVariance = 0;
        m_clusterVrlIntegrationVariance = 0;
        uint32_t numVrlsTotal = 0;
        for (auto &c : vrlsPerCluster) numVrlsTotal += c.size();
        m_vrls.resize(numVrlsTotal);
        uint32_t begin = 0;
        for (auto &c : vrlsPerCluster) {
            for (uint32_t j = 0; j < c.size(); j++) m_vrls[begin + j] = c[j];
            addCluster(begin, begin + c.size());
            begin += c.size();
        }
        calculateUnclusteredVariance(m_M, m_lw, m_vrls.begin(), m_vrls.end(), m_vrlTracingVariance, m_unclusteredVrlIntegrationVariance);
    }
    uint32_t numSingletonClusters() const { return m_singletons.size(); }
    uint32_t numMultiClusters() const { return m_pq.size(); }
    uint32_t numClusters() const { return numSingletonClusters() + numMultiClusters(); }

    void sampleRepresentatives(std::vector<uint32_t> &reprVrls, std::vector<Float> &weights, Sampler *sampler) const { // 354-378
        reprVrls.resize(numClusters()); weights.resize(numClusters());
        int i = 0;
        for (uint32_t v : m_singletons) { reprVrls[i] = v; weights[i] = 1; i++; }
        for (const ClusterNode &cn : m_pq.q) {
            Float prob;
            uint32_t j = weightedSample(m_columnWeights, sampler, &prob, cn.begin, cn.end, &m_vrls);
            reprVrls[i] = m_vrls[j];
            weights[i] = 1.0f / prob;
            i++;
        }
    }
    bool refine(Float undersampling, Sampler *sampler) {                   // 380-385
        if (undersampling <= 0) return refineAdaptively(sampler, m_depthCorrection);
        return refineFixedDepth(undersampling, sampler);
    }
    bool refineFixedDepth(Float undersampling, Sampler *sampler) {         // 387-399
        uint32_t targetClusters = 0.5 + numVrls() / undersampling;
        if (numClusters() >= targetClusters || numMultiClusters() <= 0) return true;
        while (numClusters() < targetClusters && numMultiClusters() > 0) {
            ClusterNode cn = popMultiCluster();
            if (!split(cn.begin, cn.end, sampler)) fail("couldn't split cluster!");
        }
        return true;
    }
    bool refineAdaptively(Sampler *sampler, Float depthCorrection) {       // 402-489
        /* depthCorrection != 1 (403-408, 455-470): the reference draws the split decisions from a fresh ReplayableSampler
         * (src/libbidir/rsampler.cpp:23-80; seeded by Random() = /dev/urandom, so not reproducible run to run), keeps only the
         * INITIAL snapshot, and afterwards replays the uniforms for 0.5 + depthCorrection * bestNumberOfSplits splits from the
         * restored queue.  The VRL list is NOT part of the snapshot (686-699): the second pass partitions the ranges as the
         * first pass left them, so -- the centres being picked by running sums in list order -- it does not retrace the first
         * pass's splits even though it sees the same uniforms.  That quirk is kept.  Restated for the counter stream, where a
         * split draws from the sub-stream of the cluster [begin, end) it splits (Sampler::enterNode): "the same uniforms
         * again" then holds by construction.  A sequential SFMT stream has no defined second pass and refuses. */
        if (depthCorrection != 1 && !sampler->replayable()) fail("depthCorrection != 1 needs the counter sample stream (rngMode = COUNTER)");
        if (numMultiClusters() <= 0) return true;
        if (unclusteredVariance() == 0) return false;
        Float bestConstant = convergenceConstant();
        int numberOfSplits = 0, bestNumberOfSplits = 0;
        makeRefinementSnapshot();
        while (numMultiClusters() > 0) {
            ClusterNode cn = popMultiCluster();
            if (!split(cn.begin, cn.end, sampler)) fail("couldn't split cluster!");
            numberOfSplits++;
            Float currConstant = convergenceConstant();
            if (currConstant < bestConstant) {
                if (depthCorrection == 1) makeRefinementSnapshot();
                bestConstant = currConstant;
                bestNumberOfSplits = numberOfSplits;
            }
            if (lowerBoundOfFutureConvergenceConstants() >= bestConstant) break;
        }
        restoreSnapshot();
        if (depthCorrection != 1) {                                        // 455-470
            int correctedNumberOfSplits = 0.5 + depthCorrection * bestNumberOfSplits;
            for (int i = 0; i < correctedNumberOfSplits; i++) {
                if (numMultiClusters() == 0) break;                        // EWarn in the reference
                ClusterNode cn = popMultiCluster();
                if (!split(cn.begin, cn.end, sampler)) fail("couldn't split cluster in second pass!");
            }
        }
        return true;
    }
    Float unclusteredVariance() const { return m_vrlTracingVariance + m_unclusteredVrlIntegrationVariance; }
    Float clusteredVariance() const { return m_vrlTracingVariance + m_clusterUndersamplingVariance + m_clusterVrlIntegrationVariance; }
    Float convergenceConstant() const {                                     // 503-509
        Float c = (numVrls() * m_pixelUndersampling + numClusters()) * clusteredVariance();
        if (!std::isfinite(c) || c <= 0) fail("invalid convergence constant");
        return c;
    }
    Float lowerBoundOfFutureConvergenceConstants() const {                  // 511-517
        Float c = (numVrls() * m_pixelUndersampling + numClusters()) * unclusteredVariance();
        if (!std::isfinite(c) || c <= 0) fail("invalid lower bound on convergence constant");
        return c;
    }
    std::vector<std::vector<uint32_t>> getVrlsPerCluster() const {          // 526-543
        std::vector<std::vector<uint32_t>> out;
        for (uint32_t v : m_singletons) out.push_back(std::vector<uint32_t>(1, v));
        for (const ClusterNode &cn : m_pq.q) out.push_back(std::vector<uint32_t>(m_vrls.begin() + cn.begin, m_vrls.begin() + cn.end));
        return out;
    }
    const std::vector<Float> &columnWeights() const { return m_columnWeights; }
    Float tracingVariance() const { return m_vrlTracingVariance; }
    Float unclusteredIntegrationVariance() const { return m_unclusteredVrlIntegrationVariance; }
    uint32_t nearTieSplits = 0;   // splits whose best and second-best variance agree to 1e-6 rel (G5 flag)
private:
    uint32_t numVrls() const { return m_M.nVrls; }
    void addCluster(uint32_t begin, uint32_t end, Float undersampVar, Float integrationVar) {   // 549-572
        if (end == begin) fail("Trying to add empty cluster!");
        if (end == begin + 1) {
            m_singletons.push_front(m_vrls[begin]);
            if (undersampVar != 0) fail("Trying to add singleton cluster with non-zero undersampling variance");
            m_clusterVrlIntegrationVariance += integrationVar;
        } else {
            m_pq.push(ClusterNode{undersampVar, integrationVar, begin, end});
            m_clusterUndersamplingVariance += undersampVar;
            m_clusterVrlIntegrationVariance += integrationVar;
        }
    }
    void addCluster(uint32_t begin, uint32_t end) {
        std::pair<Float, Float> v = calculateClusterVariance(m_M, m_columnWeights, m_lw, m_vrls.begin() + begin, m_vrls.begin() + end, nullptr, m_stats);
        addCluster(begin, end, v.first, v.second);
    }
    ClusterNode popMultiCluster() {                                         // 581-587
        ClusterNode cn = m_pq.top();
        m_pq.pop();
        m_clusterUndersamplingVariance -= cn.undersamplingVar;
        m_clusterVrlIntegrationVariance -= cn.integrationVar;
        return cn;
    }
    static Float norm2f(const std::vector<Float> &v) { Float t = 0; for (Float u : v) { Float a = std::abs(u); t += a * a; } return std::sqrt(t); }
    void extractMeanF(uint32_t vrl, std::vector<Float> &col) const { col.resize(m_M.numRows()); for (size_t j = 0; j < col.size(); j++) col[j] = m_M.rows[j][vrl].mean; }

    bool split(uint32_t begin, uint32_t end, Sampler *sampler) {            // 590-684
        uint32_t clusterSize = end - begin;
        if (clusterSize < 2) return false;
        if (m_stats) m_stats->splits++;
        sampler->enterNode(begin, end);                                     /* counter stream: this cluster's own draws */
        struct Leave { Sampler *s; ~Leave() { s->leaveNode(); } } leave{sampler};
        uint32_t vrl1 = m_vrls[weightedSample(m_columnWeights, sampler, nullptr, begin, end, &m_vrls)];
        Float weight1 = m_columnWeights[vrl1];
        m_columnWeights[vrl1] = 0.0f;
        uint32_t vrl2 = m_vrls[weightedSample(m_columnWeights, sampler, nullptr, begin, end, &m_vrls)];
        m_columnWeights[vrl1] = weight1;

        size_t nr = m_M.numRows();
        std::vector<Float> direction(nr), vrl1col, vrl2col, diff(nr);
        extractMeanF(vrl1, vrl1col); Float vrl1len = norm2f(vrl1col);
        extractMeanF(vrl2, vrl2col); Float vrl2len = norm2f(vrl2col);
        for (size_t i = 0; i < nr; i++) diff[i] = vrl2col[i] - vrl1col[i];
        Float diffLen = norm2f(diff);
        if (vrl1len != 0 && vrl2len != 0 && diffLen != 0) {
            for (size_t i = 0; i < nr; i++) direction[i] = diff[i] / diffLen;
        } else {
            do {
                for (size_t i = 0; i < nr; i++) {
                    /* warp::squareToStdNormal(sampler->next2D()).x, src/libcore/warp.cpp:131-137 */
                    Float s1 = sampler->next1D(), s2 = sampler->next1D();
                    Float r = std::sqrt(-2 * (Float) ::log((double) (1 - s1))), phi = (Float) (2 * M_PI * s2);
                    direction[i] = (Float) ::cos((double) phi) * r;     /* sincosf pinned: evaluated in double, then rounded (libm-independent; the product does the same) */
                }
            } while (norm2f(direction) == 0);
            Float n = norm2f(direction);
            for (size_t i = 0; i < nr; i++) direction[i] = direction[i] / n;
        }
        std::vector<std::pair<Float, uint32_t>> proj(end - begin);
        std::vector<Float> col;
        for (uint32_t j = begin; j < end; j++) {
            uint32_t vrl = m_vrls[j];
            extractMeanF(vrl, col);
            Float projection, len = norm2f(col);
            if (len == 0) projection = 0;
            else {
                Float t = 0;
                for (size_t i = 0; i < nr; i++) { col[i] = col[i] / len; }
                for (size_t i = 0; i < nr; i++) t += direction[i] * col[i];
                projection = t;
            }
            proj[j - begin] = std::pair<Float, uint32_t>(projection, vrl);
        }
        std::sort(proj.begin(), proj.end());
        for (uint32_t j = begin; j < end; j++) m_vrls[j] = proj[j - begin].second;

        std::vector<std::pair<Float, Float>> fromStart(clusterSize), fromEnd(clusterSize);
        calculateClusterVariance(m_M, m_columnWeights, m_lw, m_vrls.begin() + begin, m_vrls.begin() + end, &fromStart, m_stats);
        calculateClusterVariance(m_M, m_columnWeights, m_lw, m_vrls.rend() - end, m_vrls.rend() - begin, &fromEnd, m_stats);
        Float bestVariance = std::numeric_limits<Float>::infinity(), second = std::numeric_limits<Float>::infinity();
        uint32_t bestIndex = 0xffffffffu;
        for (uint32_t i = 1; i < clusterSize; ++i) {
            std::pair<Float, Float> varHead = fromStart[i - 1], varTail = fromEnd[clusterSize - 1 - i];
            Float thisVar = varHead.first + varHead.second + varTail.first + varTail.second;
            if (thisVar < bestVariance) { second = bestVariance; bestVariance = thisVar; bestIndex = i; }
            else if (thisVar < second) second = thisVar;
        }
        if (bestIndex == 0xffffffffu) fail("Couldn't find best splitting index!");
        if (std::isfinite(second) && std::fabs(second - bestVariance) <= 1e-6f * std::fabs(bestVariance)) nearTieSplits++;
        uint32_t splitIndex = begin + bestIndex;
        addCluster(begin, splitIndex, fromStart[bestIndex - 1].first, fromStart[bestIndex - 1].second);
        addCluster(splitIndex, end, fromEnd[clusterSize - 1 - bestIndex].first, fromEnd[clusterSize - 1 - bestIndex].second);
        return true;
    }
    void makeRefinementSnapshot() { s_uv = m_clusterUndersamplingVariance; s_iv = m_clusterVrlIntegrationVariance; s_pq = m_pq; s_single = m_singletons; }
    void restoreSnapshot() { m_clusterUndersamplingVariance = s_uv; m_clusterVrlIntegrationVariance = s_iv; m_pq = s_pq; m_singletons = s_single; }

    std::vector<uint32_t> m_vrls;
    std::vector<Float> m_columnWeights;
    const MatView &m_M;
    const std::vector<double> &m_lw;
    Float m_vrlTracingVariance, m_unclusteredVrlIntegrationVariance, m_clusterUndersamplingVariance, m_clusterVrlIntegrationVariance;
    Float m_pixelUndersampling, m_depthCorrection;
    BoostPQ<ClusterNode> m_pq;
    std::list<uint32_t> m_singletons;
    Float s_uv, s_iv; BoostPQ<ClusterNode> s_pq; std::list<uint32_t> s_single;
    ClusterStats *m_stats;
};

/* ---- slices: Preprocessor.cpp:1130-1499 ------------------------------------------------ */
struct SliceData {
    std::vector<uint32_t> gatherIdx;        // indices into the pixel arrays (pixel index = y + H*x)
    V3 positionCentroid, directionCentroid;
};

struct SliceBuilder {
    static void updateMin(const V3 &p, V3 &m) { if (p.x < m.x) m.x = p.x; if (p.y < m.y) m.y = p.y; if (p.z < m.z) m.z = p.z; }
    static void updateMax(const V3 &p, V3 &m) { if (p.x > m.x) m.x = p.x; if (p.y > m.y) m.y = p.y; if (p.z > m.z) m.z = p.z; }
    static Float sliceDistance(const V3 &p1, const V3 &d1, const V3 &p2, const V3 &d2) { return std::sqrt(distanceSquared(p1, p2) + distanceSquared(d1, d2)); }
    static void findSplitPoint(const V3 &mx, const V3 &mn, unsigned char &dim, Float &split, Float &extent) {   // 1451-1487
        Float diffx = mx.x - mn.x, diffy = mx.y - mn.y, diffz = mx.z - mn.z;
        if (diffx < 0 || diffy < 0 || diffz < 0) fail("findSplitPoint: min not smaller than max!");
        if (diffx == 0 && diffy == 0 && diffz == 0) { extent = 0; dim = 0; split = std::numeric_limits<Float>::quiet_NaN(); return; }
        if (diffx > diffy) {
            if (diffx > diffz) { dim = 0; split = mn.x + 0.5 * diffx; extent = diffx; }
            else { dim = 2; split = mn.z + 0.5 * diffz; extent = diffz; }
        } else {
            if (diffy > diffz) { dim = 1; split = mn.y + 0.5 * diffy; extent = diffy; }
            else { dim = 2; split = mn.z + 0.5 * diffz; extent = diffz; }
        }
    }
    static void findSplit(const V3 &maxPos, const V3 &minPos, const V3 &maxDir, const V3 &minDir, unsigned char &dim, Float &split) { // 1432-1449
        unsigned char dimPos, dimDir; Float splitPos, splitDir, extPos, extDir;
        findSplitPoint(maxPos, minPos, dimPos, splitPos, extPos);
        findSplitPoint(maxDir, minDir, dimDir, splitDir, extDir);
        if (extPos == 0 && extDir == 0) fail("findSplit: min equal to max!");
        if (extPos > extDir) { dim = dimPos; split = splitPos; } else { dim = 3 + dimDir; split = splitDir; }
    }
    static bool isLarger(const V3 &p, const V3 &d, int dim, Float split) {  // 1420-1430
        switch (dim) { case 0: return p.x > split; case 1: return p.y > split; case 2: return p.z > split;
                       case 3: return d.x > split; case 4: return d.y > split; case 5: return d.z > split; }
        fail("isLarger: invalid split dimension");
    }
    struct SliceNode {                                                       // 1295-1341
        uint32_t minInd, maxInd; Float distance; unsigned char dim; Float split; V3 positionCentroid, directionCentroid;
        SliceNode(uint32_t minI, uint32_t maxI, const std::vector<V3> &positions, const std::vector<V3> &directions, const std::vector<uint32_t> &indices)
            : minInd(minI), maxInd(maxI) {
            if (minInd >= maxInd) fail("trying to create empty SliceNode");
            const Float nan = std::numeric_limits<Float>::quiet_NaN();
            if (minInd + 1 == maxInd) { distance = 0; dim = 0; split = nan; positionCentroid = V3(nan); directionCentroid = V3(nan); return; }
            Float posInf = std::numeric_limits<Float>::infinity(), negInf = -1 * posInf;
            V3 maxPos(negInf), minPos(posInf), maxDir(negInf), minDir(posInf);
            for (size_t i = minInd; i < maxInd; i++) {
                const V3 position = positions[indices[i]];
                updateMin(position, minPos); updateMax(position, maxPos);
                const V3 direction = directions[indices[i]];
                updateMin(direction, minDir); updateMax(direction, maxDir);
            }
            distance = sliceDistance(minPos, minDir, maxPos, maxDir);
            findSplit(maxPos, minPos, maxDir, minDir, dim, split);
            positionCentroid = minPos + 0.5f * (maxPos - minPos);
            directionCentroid = minDir + 0.5f * (maxDir - minDir);
        }
        bool operator<(const SliceNode &o) const { return distance < o.distance; }
    };
    /* getSlices + getSlicesPQ, Preprocessor.cpp:1200-1227,1349-1418 */
    static std::vector<uint32_t> getSlices(const std::vector<V3> &gatherPoints, const std::vector<V3> &directions,
                                           uint32_t targetNumSlices, std::vector<SliceData> &slices) {
        std::vector<uint32_t> gatherPointToSlice(gatherPoints.size(), ALVRL_NO_SLICE);
        std::vector<uint32_t> indices(gatherPoints.size());
        for (uint32_t i = 0; i < indices.size(); i++) indices[i] = i;
        uint32_t firstGood = 0;
        while (!gatherPoints[firstGood].isFinite()) { firstGood++; if (firstGood >= gatherPoints.size()) break; }
        for (uint32_t i = firstGood + 1; i < indices.size(); i++) {
            if (!gatherPoints[i].isFinite()) { indices[i] = indices[firstGood]; indices[firstGood] = i; firstGood++; }
        }
        slices.clear();
        size_t minInd = firstGood, maxInd = indices.size();
        if (maxInd <= minInd) return gatherPointToSlice;
        BoostPQ<SliceNode> pq;
        pq.push(SliceNode(minInd, maxInd, gatherPoints, directions, indices));
        while (pq.size() < targetNumSlices && pq.top().distance > 0) {
            SliceNode sn = pq.top();
            pq.pop();
            size_t lo = sn.minInd, hi = sn.maxInd - 1, i = lo - 1, j = hi + 1;
            while (true) {
                while (true) { i++; if (isLarger(gatherPoints[indices[i]], directions[indices[i]], sn.dim, sn.split) || i == hi) break; }
                while (true) { j--; if (!isLarger(gatherPoints[indices[j]], directions[indices[j]], sn.dim, sn.split) || j == lo) break; }
                if (i >= j) break;
                std::swap(indices[i], indices[j]);
            }
            pq.push(SliceNode(sn.minInd, j + 1, gatherPoints, directions, indices));
            pq.push(SliceNode(j + 1, sn.maxInd, gatherPoints, directions, indices));
        }
        for (const SliceNode &sn : pq.q) {
            SliceData sd;
            sd.gatherIdx.assign(indices.begin() + sn.minInd, indices.begin() + sn.maxInd);
            sd.positionCentroid = sn.positionCentroid; sd.directionCentroid = sn.directionCentroid;
            slices.push_back(sd);
            uint32_t slice = slices.size() - 1;
            for (size_t i = sn.minInd; i < sn.maxInd; i++) gatherPointToSlice[indices[i]] = slice;
        }
        return gatherPointToSlice;
    }
    /* Slice::sampleRepresentativePixels, Preprocessor.cpp:66-121: returns positions in gatherIdx */
    static std::vector<uint32_t> sampleRepresentativePixels(const SliceData &sd, Float targetUndersampling, Sampler *sampler) {
        size_t numPixels = sd.gatherIdx.size();
        std::vector<uint32_t> pixels;
        size_t targetNum = 0.5 + numPixels / targetUndersampling;
        if (targetNum < 2) targetNum = std::min((size_t) 2, numPixels);
        if (numPixels <= targetNum) { pixels = sd.gatherIdx; return pixels; }
        std::vector<uint32_t> indices;
        if (numPixels <= 2 * targetNum) {
            indices.resize(numPixels);
            for (size_t i = 0; i < numPixels; i++) indices[i] = i;
            for (size_t i = numPixels - 1; i > 0; i--) std::swap(indices[i], indices[(size_t) ((i + 1) * sampler->next1D())]);
        } else {
            indices.resize(targetNum);
            size_t n = 0;
            while (n < targetNum) {
                bool unique;
                do {
                    indices[n] = sampler->next1D() * numPixels;
                    unique = true;
                    for (size_t i = 0; i < n; i++) if (indices[i] == indices[n]) { unique = false; break; }
                } while (!unique);
                n++;
            }
        }
        pixels.resize(targetNum);
        for (size_t i = 0; i < targetNum; i++) pixels[i] = sd.gatherIdx[indices[i]];
        return pixels;
    }
};

} // namespace orc
