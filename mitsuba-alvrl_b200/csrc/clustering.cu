/*
 * clustering.cu -- Preprocessor::buildClusters on the device (src/integrators/vrl/Preprocessor.cpp:133-283,
 * 287-720, 838-912, 985-1120).
 *
 * Split of labour.  Everything that is O(#VRLs x #rows) runs in kernels over the column-major R:
 *     k_total_contribution   totalVrlContribution (936-945)            zero / non-zero columns
 *     k_column_weights       calculateColumnWeigths (985-1008)         sqrt(sum_r w_r (mean^2 + var))
 *     k_unclustered          calculateUnclusteredVariance (1022-1048)  Welford across VRLs, per row
 *     k_direction/k_project  Clustering::split (604-640)               split direction + column projections
 *     k_cluster_variance     calculateClusterVariance (1058-1120)      forward / reverse prefix variances
 *     k_combine              reduction of the per-row-block partials into the (float, float) prefix pairs
 * The decision logic whose *order* defines the result -- the binary max-heap of multi-clusters (boost::heap::
 * priority_queue = std::vector + push_heap/pop_heap), the front-inserted singleton list, the sequential fp32 prefix
 * sums of weightedSample (1534-1580), std::sort of (projection, vrl) pairs and the first-minimum argmin (664-675) --
 * stays on the host in reference order.  All Clustering objects (one per slice) advance in lock step, one split per
 * round, so that every round is a handful of batched launches and three batched copies, not a launch per slice.
 *
 * Arithmetic.  Per-column quantities that decide the sort order (norms, projections, column weights) are
 * accumulated sequentially in the reference's order and type (this file is compiled with -fmad=false), so they are
 * bit-identical to the reference.  The sums over rows of the prefix variances are block-reduced in double and only
 * then rounded to float: they can differ from the sequential double sum in the last double bit, which survives the
 * rounding to float with probability ~1e-9 per value (documented near-tie flips, gate G5).
 */
#include <list>
#include <thread>
#include <algorithm>
#include <numeric>
#include <cmath>
#include "context.h"

namespace alvrl {

#define CL_THREADS 256
#define CL_CHUNK 8

struct ClTask {                 /* one Clustering object's piece of work in a batched launch */
    uint32_t r0, nr, rowBlocks; /* rows of the local matrix L_i (getLocalMatrix with neighbourWeight <= 0, 779-794) */
    uint32_t begin, end;        /* range in the instance's vrl list */
    uint32_t reverse, finalOnly;
    uint32_t vrl1, vrl2;        /* split centres */
    uint64_t listOff;           /* instance's vrl list in dLists */
    uint64_t cwOff;             /* instance's column weights in dCw */
    uint64_t outOff;            /* first step / projection slot of this task in the round scratch */
    uint64_t partOff;           /* first partial slot (steps x rowBlocks) */
    uint64_t dirOff;            /* direction vector slot */
    double lw;                  /* uniform locality weight 1 / nr */
};

__global__ void k_total_contribution(const float2 *__restrict__ R, uint32_t ldR, uint32_t G, uint32_t N, uint8_t *__restrict__ nonZero) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    float sum = 0;
    const float2 *col = R + (size_t) v * ldR;
    for (uint32_t r = 0; r < G; r++) sum += col[r].x;
    nonZero[v] = sum != 0;
}

__global__ void k_column_weights(const float2 *__restrict__ R, uint32_t ldR, uint32_t N, const ClTask *__restrict__ tasks, float *__restrict__ cw) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const float2 *col = R + (size_t) v * ldR + t.r0;
    double acc = 0;
    for (uint32_t r = 0; r < t.nr; r++) {
        const double mean = col[r].x, var = col[r].y;
        acc += t.lw * (mean * mean + var);
    }
    cw[t.cwOff + v] = (float) sqrt(fmax(0.0, acc));
}

/* block reduction of two doubles; result valid in thread 0 */
__device__ __forceinline__ void block_reduce2(double &a, double &b, double *sh) {
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[2 * w] = a; sh[2 * w + 1] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0; b = 0;
        for (int i = 0; i < CL_THREADS / 32; i++) { a += sh[2 * i]; b += sh[2 * i + 1]; }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(CL_THREADS) k_unclustered(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                            const uint32_t *__restrict__ lists, double2 *__restrict__ out, uint32_t maxRowBlocks) {
    __shared__ double sh[2 * CL_THREADS / 32];
    const ClTask t = tasks[blockIdx.y];
    if (blockIdx.x >= t.rowBlocks) return;
    const uint32_t lr = blockIdx.x * CL_THREADS + threadIdx.x;
    const bool active = lr < t.nr;
    const uint32_t row = t.r0 + (active ? lr : 0);
    const uint32_t *list = lists + t.listOff;
    double mean = 0, M2 = 0, summedVars = 0;
    size_t n = 0;
    for (uint32_t k = t.begin; k < t.end; k++) {
        n++;
        const float2 e = R[(size_t) list[k] * ldR + row];
        summedVars += (double) e.y;
        const double x = e.x, delta = x - mean;
        mean += delta / (double) n;
        M2 += delta * (x - mean);
    }
    double a = active ? summedVars : 0.0, b = active ? M2 : 0.0;
    block_reduce2(a, b, sh);
    if (threadIdx.x == 0) out[(size_t) blockIdx.y * maxRowBlocks + blockIdx.x] = make_double2(a, b);
}

/* split direction (604-623): direction = (col2 - col1) / |col2 - col1|; flag = 1 when a norm is zero */
__global__ void k_direction(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks, float *__restrict__ dir, uint32_t *__restrict__ flags) {
    const ClTask t = tasks[blockIdx.x];
    const float2 *c1 = R + (size_t) t.vrl1 * ldR + t.r0, *c2 = R + (size_t) t.vrl2 * ldR + t.r0;
    __shared__ float sDiffLen; __shared__ uint32_t sFlag;
    if (threadIdx.x == 0) {
        float t1 = 0, t2 = 0, td = 0;
        for (uint32_t r = 0; r < t.nr; r++) {
            const float a = c1[r].x, b = c2[r].x, d = b - a;
            t1 += fabsf(a) * fabsf(a); t2 += fabsf(b) * fabsf(b); td += fabsf(d) * fabsf(d);
        }
        const float l1 = sqrtf(t1), l2 = sqrtf(t2), ld = sqrtf(td);
        sDiffLen = ld;
        sFlag = !(l1 != 0 && l2 != 0 && ld != 0);
        flags[blockIdx.x] = sFlag;
    }
    __syncthreads();
    if (sFlag) return;
    for (uint32_t r = threadIdx.x; r < t.nr; r += blockDim.x) dir[t.dirOff + r] = (c2[r].x - c1[r].x) / sDiffLen;
}

/* projections of the normalised columns on the direction (625-640), sequential fp32 in row order */
__global__ void k_project(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks, const uint32_t *__restrict__ lists,
                          const float *__restrict__ dir, float *__restrict__ proj) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.end - t.begin) return;
    const uint32_t vid = lists[t.listOff + t.begin + j];
    const float2 *col = R + (size_t) vid * ldR + t.r0;
    const float *d = dir + t.dirOff;
    float s = 0;
    for (uint32_t r = 0; r < t.nr; r++) { const float u = fabsf(col[r].x); s += u * u; }
    const float len = sqrtf(s);
    float p = 0;
    if (len != 0) for (uint32_t r = 0; r < t.nr; r++) p += d[r] * (col[r].x / len);
    proj[t.outOff + j] = p;
}

__global__ void k_scatter_lists(const ClTask *__restrict__ tasks, const uint32_t *__restrict__ staged, uint32_t *__restrict__ lists) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < t.end - t.begin) lists[t.listOff + t.begin + j] = staged[t.outOff + j];
}

/*
 * calculateClusterVariance (1058-1120).  thread = row, sequential over the VRLs of the range (reverse: from the end),
 * CL_CHUNK steps at a time: the loads of a chunk are issued together, the weight-only factors of the recurrence are
 * computed once per step by one thread, and the per-step sums over rows are reduced from shared memory by one warp
 * per step instead of a block-wide reduction per step.
 */
__global__ void __launch_bounds__(CL_THREADS) k_cluster_variance(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                                 const uint32_t *__restrict__ lists, const float *__restrict__ cw,
                                                                 double2 *__restrict__ partial, double *__restrict__ Wk) {
    __shared__ double sM[CL_CHUNK][CL_THREADS];
    __shared__ double sV[CL_CHUNK][CL_THREADS];
    __shared__ double sC1[CL_CHUNK], sC2[CL_CHUNK], sRw[CL_CHUNK], sWgt[CL_CHUNK], sW[CL_CHUNK + 1];
    __shared__ uint32_t sVid[CL_CHUNK];
    const ClTask t = tasks[blockIdx.y];
    if (blockIdx.x >= t.rowBlocks) return;
    const uint32_t tid = threadIdx.x;
    const uint32_t lr = blockIdx.x * CL_THREADS + tid;
    const bool active = lr < t.nr;
    const uint32_t row = t.r0 + (active ? lr : 0);
    const uint32_t n = t.end - t.begin;
    const uint32_t *list = lists + t.listOff;
    const float *w = cw + t.cwOff;
    double sum = 0, M = 0, sumVars = 0;
    if (tid == 0) sW[0] = 0;
    for (uint32_t k0 = 0; k0 < n; k0 += CL_CHUNK) {
        const uint32_t cnt = min((uint32_t) CL_CHUNK, n - k0);
        /* weight-only part: prefix weight sums (sequential, double) and the factors of the M recurrence */
        if (tid == 0) {
            double W = sW[0];
            for (uint32_t j = 0; j < cnt; j++) {
                const uint32_t idx = t.reverse ? (t.end - 1 - (k0 + j)) : (t.begin + k0 + j);
                const uint32_t vid = list[idx];
                const double weight = w[vid];
                const double newW = W + weight;
                sVid[j] = vid;
                sWgt[j] = weight;
                sRw[j] = 1.0 / weight;
                if (k0 + j > 0) { sC1[j] = (newW * newW) / (W * W); sC2[j] = 1.0 / weight + 1.0 / W; }
                else { sC1[j] = 0; sC2[j] = 0; }
                W = newW;
                sW[j + 1] = W;
            }
        }
        __syncthreads();
        float2 x[CL_CHUNK];
#pragma unroll
        for (uint32_t j = 0; j < CL_CHUNK; j++) x[j] = (j < cnt) ? R[(size_t) sVid[j] * ldR + row] : make_float2(0, 0);
#pragma unroll
        for (uint32_t j = 0; j < CL_CHUNK; j++) {
            if (j < cnt) {
                const double xm = x[j].x;
                if (k0 + j > 0) {
                    const double tmp = sWgt[j] * sum - sW[j] * xm;
                    M = sC1[j] * M + sC2[j] * (tmp * tmp);
                }
                sumVars += (double) x[j].y * sRw[j];
                sum = sum + xm;
                sM[j][tid] = active ? M : 0.0;
                sV[j][tid] = active ? sumVars : 0.0;
            }
        }
        __syncthreads();
        /* one warp per step: sum over the rows of this block */
        const uint32_t warp = tid >> 5, lane = tid & 31;
        for (uint32_t j = warp; j < cnt; j += CL_THREADS / 32) {
            double a = 0, b = 0;
            for (uint32_t i = lane; i < CL_THREADS; i += 32) { a += sM[j][i]; b += sV[j][i]; }
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
            if (lane == 0) {
                const uint32_t step = k0 + j;
                if (!t.finalOnly || step == n - 1) {
                    const uint64_t slot = t.finalOnly ? 0 : step;
                    partial[t.partOff + slot * t.rowBlocks + blockIdx.x] = make_double2(a, b);
                    if (blockIdx.x == 0) Wk[t.outOff + slot] = sW[j + 1];
                }
            }
        }
        __syncthreads();
        if (tid == 0) sW[0] = sW[cnt];
    }
}

/* (undersampling variance, integration variance) prefix pairs: inner_prod(localityWeights, M / weightSum),
 * inner_prod(localityWeights, sumVars * weightSum) (1098-1106), uniform locality weights */
__global__ void k_combine(const ClTask *__restrict__ tasks, const double2 *__restrict__ partial, const double *__restrict__ Wk, float2 *__restrict__ out) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t n = t.finalOnly ? 1 : (t.end - t.begin);
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double a = 0, b = 0;
    for (uint32_t rb = 0; rb < t.rowBlocks; rb++) { const double2 p = partial[t.partOff + (uint64_t) s * t.rowBlocks + rb]; a += p.x; b += p.y; }
    const double W = Wk[t.outOff + s];
    const bool first = !t.finalOnly && s == 0;
    const bool single = t.finalOnly && (t.end - t.begin) == 1;
    float2 r;
    r.x = (first || single) ? 0.0f : (float) (t.lw * (a / W));
    r.y = (float) (t.lw * (b * W));
    out[t.outOff + s] = r;
}

/* ---- host side ------------------------------------------------------------------------------------ */
namespace {

struct ClusterNode {                                                           /* Preprocessor.cpp:289-298 */
    float undersamplingVar, integrationVar; uint32_t begin, end;
    bool operator<(const ClusterNode &o) const { return undersamplingVar + integrationVar < o.undersamplingVar + o.integrationVar; }
};

/* weightedSample, Preprocessor.cpp:1534-1580 (sequential fp32 prefix sums, reference order) */
size_t weighted_sample(const std::vector<float> &weights, HostSampler *smp, float *prob, size_t begin, size_t end, const std::vector<uint32_t> &ind) {
    if (begin >= end) throw Error(ALVRL_ERR_ARG, "Trying to take weighted sample of empty set!");
    if (end == begin + 1) { if (prob) *prob = 1; return begin; }
    float weightSum = 0.0f;
    for (size_t i = begin; i < end; i++) weightSum += weights[ind[i]];
    float probability; size_t idx;
    if (weightSum <= 0) {
        do { idx = (size_t) (begin + smp->next1D() * (end - begin)); } while (idx >= end);
        probability = (float) (1.0 / (end - begin));
    } else {
        const float alpha = smp->next1D() * weightSum;
        float accum = 0.0f;
        idx = begin;
        for (size_t i = begin; i < end; i++) { accum += weights[ind[i]]; if (accum >= alpha) { idx = i; break; } }
        probability = weights[ind[idx]] / weightSum;
    }
    if (prob) *prob = probability;
    return idx;
}

struct Inst {                                                                  /* one Clustering object, 287-720 */
    uint32_t id = 0, r0 = 0, nr = 0, rowBlocks = 1;
    double lw = 0; float pixelUndersampling = 1;
    HostSampler *smp = nullptr; std::unique_ptr<HostSampler> ownSmp; int group = 0;
    std::vector<uint32_t> vrls; std::vector<float> cw;
    uint64_t listOff = 0, cwOff = 0;
    std::vector<ClusterNode> pq; std::list<uint32_t> singletons;
    float tracingVar = 0, unclIntVar = 0, underVar = 0, intVar = 0;
    std::vector<ClusterNode> s_pq; std::list<uint32_t> s_single; float s_under = 0, s_int = 0;
    uint32_t numVrlsTotal = 0;
    /* refinement state */
    bool refining = false, done = true, failed = false, adaptive = false;
    uint32_t targetClusters = 0; float bestConstant = 0;
    ClusterNode cur{0, 0, 0, 0}; uint32_t vrl1 = 0, vrl2 = 0;
    uint32_t nearTies = 0;

    uint32_t numMulti() const { return (uint32_t) pq.size(); }
    uint32_t numClusters() const { return (uint32_t) (pq.size() + singletons.size()); }
    float unclusteredVariance() const { return tracingVar + unclIntVar; }
    float clusteredVariance() const { return tracingVar + underVar + intVar; }
    float convergenceConstant() const {                                        /* 503-509 */
        const float c = (numVrlsTotal * pixelUndersampling + numClusters()) * clusteredVariance();
        if (!std::isfinite(c) || c <= 0) throw Error(ALVRL_ERR_ARG, "invalid convergence constant");
        return c;
    }
    float lowerBound() const {                                                 /* 511-517 */
        const float c = (numVrlsTotal * pixelUndersampling + numClusters()) * unclusteredVariance();
        if (!std::isfinite(c) || c <= 0) throw Error(ALVRL_ERR_ARG, "invalid lower bound on convergence constant");
        return c;
    }
    void addCluster(uint32_t begin, uint32_t end, float uvar, float ivar) {    /* 549-572 */
        if (end == begin) throw Error(ALVRL_ERR_ARG, "Trying to add empty cluster!");
        if (end == begin + 1) {
            singletons.push_front(vrls[begin]);
            if (uvar != 0) throw Error(ALVRL_ERR_ARG, "Trying to add singleton cluster with non-zero undersampling variance");
            intVar += ivar;
        } else {
            pq.push_back(ClusterNode{uvar, ivar, begin, end}); std::push_heap(pq.begin(), pq.end());
            underVar += uvar; intVar += ivar;
        }
    }
    ClusterNode popMulti() {                                                   /* 581-587 */
        std::pop_heap(pq.begin(), pq.end());
        ClusterNode cn = pq.back(); pq.pop_back();
        underVar -= cn.undersamplingVar; intVar -= cn.integrationVar;
        return cn;
    }
    void snapshot() { s_under = underVar; s_int = intVar; s_pq = pq; s_single = singletons; }
    void restore() { underVar = s_under; intVar = s_int; pq = s_pq; singletons = s_single; }
    void sampleRepresentatives(std::vector<uint32_t> &repr, std::vector<float> &weights) {   /* 354-378 */
        repr.resize(numClusters()); weights.resize(numClusters());
        size_t i = 0;
        for (uint32_t v : singletons) { repr[i] = v; weights[i] = 1; i++; }
        for (const ClusterNode &cn : pq) {
            float prob;
            const size_t j = weighted_sample(cw, smp, &prob, cn.begin, cn.end, vrls);
            repr[i] = vrls[j]; weights[i] = 1.0f / prob; i++;
        }
    }
    std::vector<std::vector<uint32_t>> vrlsPerCluster() const {                /* 526-543 */
        std::vector<std::vector<uint32_t>> out;
        for (uint32_t v : singletons) out.push_back(std::vector<uint32_t>(1, v));
        for (const ClusterNode &cn : pq) out.push_back(std::vector<uint32_t>(vrls.begin() + cn.begin, vrls.begin() + cn.end));
        return out;
    }
};

/* device-side workspace shared by all Clustering objects of one buildClusters call */
struct Workspace {
    alvrl_ctx *c; cudaStream_t st; uint32_t N, ldR; const float2 *R;
    DevBuf<uint32_t> dLists; DevBuf<float> dCw;
    DevBuf<ClTask> dTasks; DevBuf<float> dDir, dProj; DevBuf<uint32_t> dFlags, dStaged;
    DevBuf<double2> dPartial, dUncl; DevBuf<double> dWk; DevBuf<float2> dPairs;
    std::vector<Inst *> insts;

    template <typename T> static void ensure(DevBuf<T> &b, size_t n) { if (b.n < n) b.alloc(n + n / 4 + 16); }
    void launches(uint32_t k) { c->stats.kernelLaunches += k; }

    void allocInstances() {
        ensure(dLists, insts.size() * (size_t) N); ensure(dCw, insts.size() * (size_t) N);
        for (size_t i = 0; i < insts.size(); i++) { insts[i]->listOff = i * (uint64_t) N; insts[i]->cwOff = i * (uint64_t) N; }
    }
    ClTask baseTask(const Inst &in) const {
        ClTask t; memset(&t, 0, sizeof(t));
        t.r0 = in.r0; t.nr = in.nr; t.rowBlocks = in.rowBlocks; t.listOff = in.listOff; t.cwOff = in.cwOff; t.lw = in.lw;
        return t;
    }
    /* calculateColumnWeigths for every instance (985-1008); the float average is summed on the host in index order */
    void columnWeights() {
        std::vector<ClTask> tasks;
        for (Inst *in : insts) tasks.push_back(baseTask(*in));
        dTasks.upload(tasks, st);
        k_column_weights<<<dim3((N + 127) / 128, (uint32_t) tasks.size()), 128, 0, st>>>(R, ldR, N, dTasks.p, dCw.p);
        launches(1);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<float> all(insts.size() * (size_t) N);
        dCw.download(all.data(), all.size(), st);
        for (size_t i = 0; i < insts.size(); i++) {
            Inst &in = *insts[i];
            in.cw.assign(all.begin() + i * (size_t) N, all.begin() + (i + 1) * (size_t) N);
            for (float w : in.cw) if (!std::isfinite(w)) throw Error(ALVRL_ERR_ARG, "Invalid calculated average column weight");
            float averageWeight = std::accumulate(in.cw.begin(), in.cw.end(), 0.0f) / N;
            if (averageWeight == 0) averageWeight = 1.0;
            const float safetyFraction = 1e-2;
            for (float &w : in.cw) w += averageWeight * safetyFraction;
            std::copy(in.cw.begin(), in.cw.end(), all.begin() + i * (size_t) N);
        }
        ALVRL_CUDA(cudaMemcpyAsync(dCw.p, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        ALVRL_CUDA(cudaStreamSynchronize(st));
    }
    void uploadLists() {
        std::vector<uint32_t> all(insts.size() * (size_t) N, 0);
        for (size_t i = 0; i < insts.size(); i++) std::copy(insts[i]->vrls.begin(), insts[i]->vrls.end(), all.begin() + i * (size_t) N);
        ALVRL_CUDA(cudaMemcpyAsync(dLists.p, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ALVRL_CUDA(cudaStreamSynchronize(st));
    }
    /* variance of whole ranges (addCluster(begin, end), 576-579): one (uvar, ivar) pair per task */
    std::vector<float2> rangeVariances(std::vector<ClTask> &tasks) {
        uint64_t out = 0, part = 0; uint32_t maxRb = 1;
        for (ClTask &t : tasks) { t.finalOnly = 1; t.reverse = 0; t.outOff = out; t.partOff = part; out += 1; part += t.rowBlocks; maxRb = std::max(maxRb, t.rowBlocks); }
        ensure(dPartial, part); ensure(dWk, out); ensure(dPairs, out);
        dTasks.upload(tasks, st);
        k_cluster_variance<<<dim3(maxRb, (uint32_t) tasks.size()), CL_THREADS, 0, st>>>(R, ldR, dTasks.p, dLists.p, dCw.p, dPartial.p, dWk.p);
        k_combine<<<dim3(1, (uint32_t) tasks.size()), 32, 0, st>>>(dTasks.p, dPartial.p, dWk.p, dPairs.p);
        launches(2);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<float2> res(out);
        dPairs.download(res.data(), out, st);
        return res;
    }
    /* Clustering constructor (301-341) for all instances: initial clusters + unclustered variance */
    void construct(const std::vector<std::vector<uint32_t>> &vrlsPerCluster) {
        uint32_t total = 0;
        for (auto &cl : vrlsPerCluster) total += (uint32_t) cl.size();
        for (Inst *in : insts) {
            in->vrls.clear();
            for (auto &cl : vrlsPerCluster) in->vrls.insert(in->vrls.end(), cl.begin(), cl.end());
            in->numVrlsTotal = N; in->underVar = 0; in->intVar = 0; in->pq.clear(); in->singletons.clear();
            if (std::fabs((float) (in->lw * in->nr) - 1) > 1e-3) throw Error(ALVRL_ERR_ARG, "Incorrect normalization in localityWeights");
            if (in->pixelUndersampling <= 0 || in->pixelUndersampling > 1) throw Error(ALVRL_ERR_ARG, "Invalid pixel undersampling");
        }
        uploadLists();
        std::vector<ClTask> tasks;
        for (Inst *in : insts) {
            uint32_t begin = 0;
            for (auto &cl : vrlsPerCluster) { ClTask t = baseTask(*in); t.begin = begin; t.end = begin + (uint32_t) cl.size(); tasks.push_back(t); begin = t.end; }
        }
        std::vector<float2> res = rangeVariances(tasks);
        size_t k = 0;
        for (Inst *in : insts) {
            uint32_t begin = 0;
            for (auto &cl : vrlsPerCluster) {
                const float2 v = res[k++];
                if (!std::isfinite(v.x) || v.x < 0) throw Error(ALVRL_ERR_ARG, "invalid undersampled VRL cluster variance");
                if (!std::isfinite(v.y) || v.y < 0) throw Error(ALVRL_ERR_ARG, "invalid undersampled VRL integration cluster variance");
                in->addCluster(begin, begin + (uint32_t) cl.size(), v.x, v.y);
                begin += (uint32_t) cl.size();
            }
        }
        /* calculateUnclusteredVariance over all VRLs of the list (334-335, 1022-1048) */
        if (total <= 1) throw Error(ALVRL_ERR_ARG, "Need at least 2 VRLs to estimate variance");
        std::vector<ClTask> ut; uint32_t maxRb = 1;
        for (Inst *in : insts) { ClTask t = baseTask(*in); t.begin = 0; t.end = total; ut.push_back(t); maxRb = std::max(maxRb, t.rowBlocks); }
        ensure(dUncl, ut.size() * (size_t) maxRb);
        dTasks.upload(ut, st);
        k_unclustered<<<dim3(maxRb, (uint32_t) ut.size()), CL_THREADS, 0, st>>>(R, ldR, dTasks.p, dLists.p, dUncl.p, maxRb);
        launches(1);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<double2> u(ut.size() * (size_t) maxRb);
        dUncl.download(u.data(), u.size(), st);
        for (size_t i = 0; i < insts.size(); i++) {
            double a = 0, b = 0;
            for (uint32_t rb = 0; rb < insts[i]->rowBlocks; rb++) { a += u[i * maxRb + rb].x; b += u[i * maxRb + rb].y; }
            insts[i]->unclIntVar = (float) (insts[i]->lw * a);
            insts[i]->tracingVar = (float) (insts[i]->lw * b - (double) insts[i]->unclIntVar);
        }
    }
    /* Clustering::refine for the given instances, all advancing one split per round (380-489, 590-684) */
    void refine(const std::vector<Inst *> &which, float undersampling) {
        for (Inst *in : which) {
            in->failed = false; in->done = false; in->refining = true;
            in->adaptive = undersampling <= 0;
            if (in->adaptive) {                                                 /* refineAdaptively, 402-423 */
                if (in->numMulti() <= 0) { in->done = true; continue; }
                if (in->unclusteredVariance() == 0) { in->done = true; in->failed = true; continue; }
                in->bestConstant = in->convergenceConstant();
                in->snapshot();
            } else {                                                            /* refineFixedDepth, 387-399 */
                in->targetClusters = (uint32_t) (0.5 + in->numVrlsTotal / undersampling);
                if (in->numClusters() >= in->targetClusters || in->numMulti() <= 0) in->done = true;
            }
        }
        for (;;) {
            /* one runnable instance per sampler group: a shared sequential stream (SFMT) serialises its instances */
            std::vector<Inst *> round; std::vector<int> groupsBusy;
            for (Inst *in : which) {
                if (in->done) continue;
                if (std::find(groupsBusy.begin(), groupsBusy.end(), in->group) != groupsBusy.end()) continue;
                groupsBusy.push_back(in->group);
                round.push_back(in);
            }
            if (round.empty()) break;
            splitRound(round);
            for (Inst *in : round) {
                if (in->adaptive) {
                    const float curr = in->convergenceConstant();               /* 436-452 */
                    if (curr < in->bestConstant) { in->snapshot(); in->bestConstant = curr; }
                    if (in->lowerBound() >= in->bestConstant || in->numMulti() == 0) { in->restore(); in->done = true; }
                } else if (!(in->numClusters() < in->targetClusters && in->numMulti() > 0)) in->done = true;
            }
        }
        for (Inst *in : which) in->refining = false;
    }
    /* Clustering::split (590-684) for one cluster of every instance in `round` */
    void splitRound(const std::vector<Inst *> &round) {
        const size_t T = round.size();
        std::vector<ClTask> tasks(T);
        uint64_t out = 0, dirOff = 0;
        for (size_t i = 0; i < T; i++) {
            Inst &in = *round[i];
            in.cur = in.popMulti();
            const uint32_t begin = in.cur.begin, end = in.cur.end;
            if (end - begin < 2) throw Error(ALVRL_ERR_ARG, "couldn't split cluster!");
            in.vrl1 = in.vrls[weighted_sample(in.cw, in.smp, nullptr, begin, end, in.vrls)];   /* 597-602 */
            const float weight1 = in.cw[in.vrl1];
            in.cw[in.vrl1] = 0.0f;
            in.vrl2 = in.vrls[weighted_sample(in.cw, in.smp, nullptr, begin, end, in.vrls)];
            in.cw[in.vrl1] = weight1;
            ClTask t = baseTask(in);
            t.begin = begin; t.end = end; t.vrl1 = in.vrl1; t.vrl2 = in.vrl2; t.outOff = out; t.dirOff = dirOff;
            out += end - begin; dirOff += in.nr;
            tasks[i] = t;
        }
        ensure(dDir, dirOff); ensure(dProj, out); ensure(dFlags, T); ensure(dStaged, out);
        dTasks.upload(tasks, st);
        uint32_t maxN = 0; for (const ClTask &t : tasks) maxN = std::max(maxN, t.end - t.begin);
        k_direction<<<(uint32_t) T, 128, 0, st>>>(R, ldR, dTasks.p, dDir.p, dFlags.p);
        k_project<<<dim3((maxN + 127) / 128, (uint32_t) T), 128, 0, st>>>(R, ldR, dTasks.p, dLists.p, dDir.p, dProj.p);
        launches(2);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<uint32_t> flags(T);
        dFlags.download(flags.data(), T, st);
        bool redo = false;
        for (size_t i = 0; i < T; i++) if (flags[i]) {
            /* degenerate centres: direction uniform on the n-sphere, squareToStdNormal(next2D()).x per row (616-622) */
            Inst &in = *round[i];
            std::vector<float> d(in.nr);
            float nrm;
            do {
                for (uint32_t r = 0; r < in.nr; r++) {
                    const float s1 = in.smp->next1D(), s2 = in.smp->next1D();
                    const float rr = std::sqrt(-2 * (float) std::log((double) (1 - s1))), phi = (float) (2 * M_PI * s2);
                    d[r] = cosf(phi) * rr;
                }
                float t = 0; for (float u : d) t += std::fabs(u) * std::fabs(u);
                nrm = std::sqrt(t);
            } while (nrm == 0);
            for (float &u : d) u = u / nrm;
            ALVRL_CUDA(cudaMemcpyAsync(dDir.p + tasks[i].dirOff, d.data(), in.nr * sizeof(float), cudaMemcpyHostToDevice, st));
            ALVRL_CUDA(cudaStreamSynchronize(st));
            redo = true;
        }
        if (redo) {
            k_project<<<dim3((maxN + 127) / 128, (uint32_t) T), 128, 0, st>>>(R, ldR, dTasks.p, dLists.p, dDir.p, dProj.p);
            launches(1);
            ALVRL_CUDA(cudaGetLastError());
        }
        std::vector<float> proj(out);
        dProj.download(proj.data(), out, st);
        /* std::sort of (projection, vrl) pairs (641-646), instances in parallel on host threads */
        std::vector<uint32_t> staged(out);
        auto sortOne = [&](size_t i) {
            Inst &in = *round[i];
            const ClTask &t = tasks[i];
            const uint32_t n = t.end - t.begin;
            std::vector<std::pair<float, uint32_t>> pr(n);
            for (uint32_t j = 0; j < n; j++) pr[j] = std::make_pair(proj[t.outOff + j], in.vrls[t.begin + j]);
            std::sort(pr.begin(), pr.end());
            for (uint32_t j = 0; j < n; j++) { in.vrls[t.begin + j] = pr[j].second; staged[t.outOff + j] = pr[j].second; }
        };
        {
            const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            if (T == 1 || hw == 1) for (size_t i = 0; i < T; i++) sortOne(i);
            else {
                std::vector<std::thread> th;
                for (unsigned w = 0; w < hw; w++) th.emplace_back([&, w]() { for (size_t i = w; i < T; i += hw) sortOne(i); });
                for (auto &x : th) x.join();
            }
        }
        ALVRL_CUDA(cudaMemcpyAsync(dStaged.p, staged.data(), out * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        k_scatter_lists<<<dim3((maxN + 127) / 128, (uint32_t) T), 128, 0, st>>>(dTasks.p, dStaged.p, dLists.p);
        /* forward and reverse prefix variances (648-657) */
        std::vector<ClTask> vt(2 * T);
        uint64_t part = 0, steps = 0; uint32_t maxRb = 1;
        for (size_t i = 0; i < T; i++)
            for (int rev = 0; rev < 2; rev++) {
                ClTask t = tasks[i];
                t.reverse = rev; t.finalOnly = 0; t.outOff = steps; t.partOff = part;
                steps += t.end - t.begin; part += (uint64_t) (t.end - t.begin) * t.rowBlocks; maxRb = std::max(maxRb, t.rowBlocks);
                vt[2 * i + rev] = t;
            }
        ensure(dPartial, part); ensure(dWk, steps); ensure(dPairs, steps);
        DevBuf<ClTask> &dT2 = dTasks2;
        dT2.upload(vt, st);
        k_cluster_variance<<<dim3(maxRb, (uint32_t) vt.size()), CL_THREADS, 0, st>>>(R, ldR, dT2.p, dLists.p, dCw.p, dPartial.p, dWk.p);
        k_combine<<<dim3((maxN + 127) / 128, (uint32_t) vt.size()), 128, 0, st>>>(dT2.p, dPartial.p, dWk.p, dPairs.p);
        launches(3);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<float2> pairs(steps);
        dPairs.download(pairs.data(), steps, st);
        for (size_t i = 0; i < T; i++) {
            Inst &in = *round[i];
            const uint32_t begin = tasks[i].begin, end = tasks[i].end, n = end - begin;
            const float2 *fromStart = &pairs[vt[2 * i].outOff], *fromEnd = &pairs[vt[2 * i + 1].outOff];
            float bestVariance = INFINITY, second = INFINITY; uint32_t bestIndex = 0xffffffffu;
            for (uint32_t k = 1; k < n; ++k) {                                   /* 664-675: first minimum wins */
                const float2 h = fromStart[k - 1], tl = fromEnd[n - 1 - k];
                const float thisVar = h.x + h.y + tl.x + tl.y;
                if (thisVar < bestVariance) { second = bestVariance; bestVariance = thisVar; bestIndex = k; }
                else if (thisVar < second) second = thisVar;
            }
            if (bestIndex == 0xffffffffu) throw Error(ALVRL_ERR_ARG, "Couldn't find best splitting index!");
            if (std::isfinite(second) && std::fabs(second - bestVariance) <= 1e-6f * std::fabs(bestVariance)) in.nearTies++;
            const uint32_t splitIndex = begin + bestIndex;
            in.addCluster(begin, splitIndex, fromStart[bestIndex - 1].x, fromStart[bestIndex - 1].y);
            in.addCluster(splitIndex, end, fromEnd[n - 1 - bestIndex].x, fromEnd[n - 1 - bestIndex].y);
        }
    }
    DevBuf<ClTask> dTasks2;
};

} // namespace

void column_nonzero_device(alvrl_ctx *c, std::vector<uint8_t> &flags) {
    const uint32_t N = (uint32_t) c->vrlHost.size(), G = (uint32_t) c->rowPixel.size();
    DevBuf<uint8_t> dNz; dNz.alloc(N);
    k_total_contribution<<<(N + 127) / 128, 128, 0, c->stream>>>(c->dR.p, c->ldR, G, N, dNz.p);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    flags.resize(N);
    dNz.download(flags.data(), N, c->stream);
}

/* register-resident FFMA chains: 8 independent accumulators per thread, all SMs full */
__global__ void k_ffma_peak(float *out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 0.999f, b = 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
            a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
float measure_fp32_peak_tflops() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float *d = nullptr;
    if (cudaMalloc(&d, (size_t) blocks * threads * sizeof(float)) != cudaSuccess) return 0;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_ffma_peak<<<blocks, threads>>>(d, 64);
    float best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        k_ffma_peak<<<blocks, threads>>>(d, iters);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 64.0 * iters * (double) blocks * threads;
        best = std::max(best, (float) (flops / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return best;
}

void build_clusters_device(alvrl_ctx *c, bool needFallback) {
    const uint32_t N = (uint32_t) c->vrlHost.size(), G = (uint32_t) c->rowPixel.size(), S = (uint32_t) c->slices.size();
    if (c->globalPixelUndersampling < 0) throw Error(ALVRL_ERR_STATE, "Invalid pixel undersampling. Did you forget to call buildSlices first?");
    cudaStream_t st = c->stream;
    const bool sfmt = c->P.rngMode == ALVRL_RNG_MODE_SFMT;
    const bool lazyFallbackCall = c->haveClusters && !c->haveFallback;       /* called again only for the global / fallback lists */

    Workspace ws; ws.c = c; ws.st = st; ws.N = N; ws.ldR = c->ldR; ws.R = c->dR.p;

    /* the global stream: cluster() -> global representatives -> fallback refinement, in this order (159-179) */
    HostSampler *globalSmp = c->mainSampler.get();
    if (!sfmt) {
        if (!lazyFallbackCall || !c->globalStream) { c->globalStream.reset(new CounterStream(c->P.seed)); c->globalStream->setContext(ALVRL_RNG_CLUSTER, ALVRL_RNG_GLOBAL_ID, 0); }
        globalSmp = c->globalStream.get();
    }

    auto makeGlobalInst = [&](Inst &g) {
        g.id = ALVRL_RNG_GLOBAL_ID; g.r0 = 0; g.nr = G; g.rowBlocks = (G + CL_THREADS - 1) / CL_THREADS; g.lw = 1.0 / G;
        g.pixelUndersampling = c->globalPixelUndersampling; g.smp = globalSmp; g.group = -1;
    };

    if (!lazyFallbackCall) {
        /* cluster(), 838-898: zero / non-zero columns over all rows */
        std::vector<uint8_t> nz;
        if (c->columnFlagsOverride.size() == N) nz = c->columnFlagsOverride;
        else column_nonzero_device(c, nz);
        std::vector<uint32_t> nonZero, zero;
        for (uint32_t i = 0; i < N; i++) (nz[i] ? nonZero : zero).push_back(i);
        c->globalVrlsPerCluster.clear();
        if (!nonZero.empty()) {
            if (c->P.globalCluster) {                                          /* clusterRefinement, 899-912 */
                Inst g; makeGlobalInst(g);
                ws.insts = {&g}; ws.allocInstances(); ws.columnWeights();
                ws.construct(std::vector<std::vector<uint32_t>>(1, nonZero));
                ws.refine({&g}, c->P.globalUndersampling);
                if (g.failed) throw Error(ALVRL_ERR_ARG, "Couldn't refine global clustering!");
                c->globalVrlsPerCluster = g.vrlsPerCluster();
                c->nearTieSplits += g.nearTies;
            } else c->globalVrlsPerCluster.assign(1, nonZero);
        }
        if (!zero.empty()) c->globalVrlsPerCluster.push_back(zero);
    }

    auto computeFallback = [&]() {
        Inst g; makeGlobalInst(g);
        Workspace w2; w2.c = c; w2.st = st; w2.N = N; w2.ldR = c->ldR; w2.R = c->dR.p;
        w2.insts = {&g}; w2.allocInstances(); w2.columnWeights();
        w2.construct(c->globalVrlsPerCluster);
        g.sampleRepresentatives(c->gcVrls, c->gcWeight);                         /* 169 */
        w2.refine({&g}, c->P.fallBackUndersampling);                             /* 177 */
        if (g.failed) throw Error(ALVRL_ERR_ARG, "couldn't refine global clustering! (but all VRLs should be non-zero!)");
        g.sampleRepresentatives(c->fallBackVrls, c->fallBackWeight);             /* 179 */
        c->nearTieSplits += g.nearTies;
        c->haveFallback = true;
    };
    if (lazyFallbackCall) { computeFallback(); return; }
    if (needFallback) computeFallback();

    /* refinePerSlice, 199-283: one Clustering per slice of this handle's range */
    const uint32_t sb = std::min(c->sliceBegin, S), se = std::min(c->sliceEnd, S);
    c->selectedVrls.assign(S, {}); c->clusterWeight.assign(S, {});
    std::vector<std::unique_ptr<Inst>> store;
    std::vector<std::unique_ptr<HostSampler>> clones;
    const int w = std::max(1, c->P.workerCount);
    if (sfmt && w > 1) for (int i = 0; i < w; i++) clones.emplace_back(c->mainSampler->clone());   /* ClusterRefiner ctor, 738 */
    for (uint32_t i = sb; i < se; i++) {
        std::unique_ptr<Inst> in(new Inst());
        in->id = i; in->r0 = c->rowOffset[i]; in->nr = c->rowOffset[i + 1] - c->rowOffset[i];
        in->rowBlocks = std::max(1u, (in->nr + CL_THREADS - 1) / CL_THREADS);
        in->lw = 1.0 / in->nr; in->pixelUndersampling = c->sliceUndersampling[i];
        if (sfmt) {
            int id = 0;
            if (w > 1) { for (id = 0; id < w; id++) if (i >= ((uint64_t) id * S) / w && i < ((uint64_t) (id + 1) * S) / w) break; }
            in->smp = w > 1 ? clones[id].get() : c->mainSampler.get();
            in->group = w > 1 ? id : 0;
        } else {
            in->ownSmp.reset(new CounterStream(c->P.seed));
            in->ownSmp->setContext(ALVRL_RNG_CLUSTER, i, 0);
            in->smp = in->ownSmp.get(); in->group = (int) i;
        }
        store.push_back(std::move(in));
    }
    ws.insts.clear();
    for (auto &p : store) ws.insts.push_back(p.get());
    if (!ws.insts.empty()) {
        ws.allocInstances();
        ws.columnWeights();
        ws.construct(c->globalVrlsPerCluster);
        if (sfmt) {
            /* a shared sequential stream: refine and sample slice after slice, in slice order (230-233, 746-752) */
            for (Inst *in : ws.insts) {
                if (c->P.localRefinement) ws.refine({in}, c->P.localUndersampling);
                if (!in->failed) in->sampleRepresentatives(c->selectedVrls[in->id], c->clusterWeight[in->id]);
            }
        } else {
            if (c->P.localRefinement) ws.refine(ws.insts, c->P.localUndersampling);
            for (Inst *in : ws.insts) if (!in->failed) in->sampleRepresentatives(c->selectedVrls[in->id], c->clusterWeight[in->id]);
        }
        bool anyFailed = false;
        for (Inst *in : ws.insts) { anyFailed |= in->failed; c->nearTieSplits += in->nearTies; }
        if (anyFailed) {                                                        /* 279-281: fall-back clustering */
            if (!c->haveFallback) computeFallback();
            for (Inst *in : ws.insts) if (in->failed) { c->selectedVrls[in->id] = c->fallBackVrls; c->clusterWeight[in->id] = c->fallBackWeight; }
        }
    }
}

} // namespace alvrl
