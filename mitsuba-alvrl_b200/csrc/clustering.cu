/*
 * clustering.cu -- Preprocessor::buildClusters on the device (src/integrators/vrl/Preprocessor.cpp:133-283,
 * 287-720, 838-912, 985-1120).
 *
 * Split of labour.  Everything that is O(#VRLs x #rows) runs in kernels over the column-major R:
 *     k_total_contribution   totalVrlContribution (936-945)            zero / non-zero columns
 *     k_column_weights       calculateColumnWeigths (985-1008)         sqrt(sum_r w_r (mean^2 + var))
 *     k_unclustered          calculateUnclusteredVariance (1022-1048)  Welford across VRLs, per row
 *     k_direction/k_project  Clustering::split (604-640)               split direction + column projections
 *     k_weights/k_seg_sums/k_carry/k_seg_main/k_final
 *                            calculateClusterVariance (1058-1120)      forward / reverse prefix variances (segmented)
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
#include <set>
#include <thread>
#include <atomic>
#include <algorithm>
#include <numeric>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_segmented_sort.cuh>
#include "context.h"

namespace alvrl {

#define CL_THREADS 256
#define CL_CHUNK 8

struct ClTask {                 /* one Clustering object's piece of work in a batched launch */
    uint32_t r0, nr, rowBlocks; /* rows of the local matrix L_i (getLocalMatrix with neighbourWeight <= 0, 779-794) */
    uint32_t begin, end;        /* range in the instance's vrl list */
    uint32_t reverse, finalOnly;
    uint32_t vrl1, vrl2;        /* split centres (host rounds) */
    float u1, u2;               /* uniforms of the two weightedSample calls (device rounds) */
    uint64_t listOff;           /* instance's vrl list in dLists */
    uint64_t cwOff;             /* instance's column weights in dCw */
    uint64_t outOff;            /* first step / projection slot of this task in the round scratch */
    uint64_t stepOff;           /* first step slot of this task in the w / W / partial arrays of the launch */
    uint64_t partOff;           /* first partial slot (steps x rowBlocks double2) */
    uint64_t carryOff;          /* first carry slot (nseg x rowBlocks x CL_THREADS doubles) */
    uint32_t nseg, segOff;      /* number of segLen-step segments, first global segment index */
    uint32_t segLen;            /* steps per segment: short for small ranges (latency), CL_SEG for large ones */
    uint64_t dirOff;            /* direction vector slot */
    double lw;                  /* uniform locality weight 1 / nr */
};

/* totalVrlContribution != 0 (Preprocessor.cpp:846-855, 936-945).  The means are sums of valid -- finite, non-negative --
 * contributions, so the row sum of a column is non-zero iff some entry is: an OR over the rows this handle owns (the other
 * rows of a sharded R are zero), one warp per column, rows coalesced. */
__global__ void k_total_contribution(const float2 *__restrict__ R, uint32_t ldR, uint32_t r0, uint32_t r1, uint32_t N, uint8_t *__restrict__ nonZero) {
    const uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= N) return;
    const float2 *col = R + (size_t) v * ldR;
    bool any = false;
    for (uint32_t r = r0 + lane; r < r1; r += 32) any |= col[r].x != 0.0f;
    const uint32_t b = __ballot_sync(0xffffffffu, any);
    if (lane == 0) nonZero[v] = b != 0u;
}

__global__ void k_column_weights(const float2 *__restrict__ R, uint32_t ldR, uint32_t N, const ClTask *__restrict__ tasks, float *__restrict__ cw,
                                 const double *__restrict__ rowW) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const float2 *col = R + (size_t) v * ldR + t.r0;
    double acc = 0;
    for (uint32_t r = 0; r < t.nr; r++) {
        const double mean = col[r].x, var = col[r].y;
        acc += (rowW ? rowW[t.r0 + r] : t.lw) * (mean * mean + var);
    }
    cw[t.cwOff + v] = (float) sqrt(fmax(0.0, acc));
}

/* The tail of calculateColumnWeigths (1000-1008) for one Clustering object per block: the fp32 average of the weights, summed
 * sequentially in index order like std::accumulate (one thread walks chunks the block stages in shared memory), then
 * w += average * safetyFraction for every column.  flags[object] = 1 when a weight is not finite. */
__global__ void __launch_bounds__(CL_THREADS) k_cw_finish(uint32_t N, const ClTask *__restrict__ tasks, float *__restrict__ cw, uint32_t *__restrict__ flags) {
    __shared__ __align__(16) float stage[2][2048];
    __shared__ float shAdd;
    __shared__ uint32_t shBad;
    float *w = cw + tasks[blockIdx.x].cwOff;
    if (threadIdx.x == 0) shBad = 0;
    uint32_t bad = 0;
    float acc = 0.0f;
    const uint32_t nChunks = (N + 2047u) / 2048u;
    for (uint32_t i = threadIdx.x; i < min(N, 2048u); i += CL_THREADS) { const float x = w[i]; stage[0][i] = x; bad |= !isfinite(x); }
    for (uint32_t c = 0; c < nChunks; c++) {
        __syncthreads();
        if (c + 1 < nChunks)
            for (uint32_t i = threadIdx.x; i < min(N - (c + 1) * 2048u, 2048u); i += CL_THREADS) { const float x = w[(c + 1) * 2048u + i]; stage[(c + 1) & 1][i] = x; bad |= !isfinite(x); }
        if (threadIdx.x == 0) {
            const float *sp = stage[c & 1];
            const uint32_t cnt = min(N - c * 2048u, 2048u);
            uint32_t i = 0;
            for (; i + 4 <= cnt; i += 4) { const float4 v = *reinterpret_cast<const float4 *>(sp + i); acc += v.x; acc += v.y; acc += v.z; acc += v.w; }
            for (; i < cnt; i++) acc += sp[i];
        }
    }
    if (bad) atomicOr(&shBad, 1u);
    if (threadIdx.x == 0) {
        float averageWeight = acc / N;
        if (averageWeight == 0) averageWeight = 1.0f;
        const float safetyFraction = 1e-2;
        shAdd = averageWeight * safetyFraction;
    }
    __syncthreads();
    const float add = shAdd;
    for (uint32_t i = threadIdx.x; i < N; i += CL_THREADS) w[i] += add;
    if (threadIdx.x == 0) flags[blockIdx.x] = shBad;
}

/* Clustering::sampleRepresentatives (354-378) for the multi-clusters of many objects, thread = cluster: weightedSample
 * (1534-1580) with its sequential fp32 sums in list order; the uniform of the i-th multi-cluster of an object is draw
 * rngPos + i of the object's counter stream (one draw per cluster, as on the host).  flags[object] = 1 when a cluster has a
 * non-positive weight sum (the uniform-pick branch draws a data-dependent number of uniforms: the host handles that object). */
struct RepObj { uint64_t listOff, cwOff; uint32_t firstCluster, numClusters, rngKey, rngPos; };
__global__ void __launch_bounds__(128) k_sample_representatives(const RepObj *__restrict__ objs, const uint2 *__restrict__ clusters,
                                                                const uint32_t *__restrict__ lists, const float *__restrict__ cw,
                                                                uint32_t *__restrict__ reprOut, float *__restrict__ weightOut, uint32_t *__restrict__ flags) {
    const RepObj o = objs[blockIdx.y];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= o.numClusters) return;
    const uint2 cl = clusters[o.firstCluster + i];
    const uint32_t *list = lists + o.listOff;
    const float *w = cw + o.cwOff;
    float weightSum = 0.0f;
    for (uint32_t k0 = cl.x; k0 < cl.y; k0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (k0 + u < cl.y) ? w[list[k0 + u]] : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; u++) if (k0 + u < cl.y) weightSum += v[u];
    }
    if (weightSum <= 0) { atomicOr(flags + blockIdx.y, 1u); return; }
    const float alpha = alvrl_rng_uniform(o.rngKey, o.rngPos + i) * weightSum;
    float accum = 0.0f;
    uint32_t idx = cl.x;
    bool found = false;
    for (uint32_t k0 = cl.x; k0 < cl.y && !found; k0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (k0 + u < cl.y) ? w[list[k0 + u]] : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (!found && k0 + u < cl.y) { accum += v[u]; if (accum >= alpha) { idx = k0 + u; found = true; } }
    }
    const uint32_t vrl = list[idx];
    const float probability = w[vrl] / weightSum;
    reprOut[o.firstCluster + i] = vrl;
    weightOut[o.firstCluster + i] = 1.0f / probability;
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
                                                            const uint32_t *__restrict__ lists, double2 *__restrict__ out, uint32_t maxRowBlocks,
                                                            const double *__restrict__ rowW) {
    __shared__ double sh[2 * CL_THREADS / 32];
    const ClTask t = tasks[blockIdx.y];
    if (blockIdx.x >= t.rowBlocks) return;
    const uint32_t lr = blockIdx.x * CL_THREADS + threadIdx.x;
    const bool active = lr < t.nr;
    const uint32_t row = t.r0 + (active ? lr : 0);
    const uint32_t *list = lists + t.listOff;
    double mean = 0, M2 = 0, summedVars = 0;
    /* The Welford chain is sequential (reference order); the loads are not: eight columns in flight per thread.  The chain's
     * division delta / n is the long pole (an IEEE double division is a ~40-instruction routine), and n is the same for every
     * row: the block computes the correctly rounded reciprocals y = 1 / n of a chunk of steps once, and the chain uses
     * q0 = delta * y, q = fma(fma(-q0, n, delta), y, q0), which IS the correctly rounded quotient (Markstein: one fma
     * correction of a faithful quotient with the correctly rounded reciprocal; n is a small integer, far from the theorem's
     * exceptional all-ones significands) -- bit-identical to delta / n, three dependent operations instead of forty. */
    __shared__ double rcpN[CL_THREADS];
    for (uint32_t c0 = t.begin; c0 < t.end; c0 += CL_THREADS) {
        __syncthreads();
        rcpN[threadIdx.x] = 1.0 / (double) (c0 - t.begin + threadIdx.x + 1u);
        __syncthreads();
        const uint32_t cEnd = min(t.end, c0 + CL_THREADS);
        for (uint32_t k0 = c0; k0 < cEnd; k0 += 8) {
            float2 e[8];
#pragma unroll
            for (int u = 0; u < 8; u++) e[u] = (k0 + u < cEnd) ? R[(size_t) list[k0 + u] * ldR + row] : make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (k0 + u < cEnd) {
                    const double nD = (double) (k0 + u - t.begin + 1u), y = rcpN[k0 + u - c0];
                    summedVars += (double) e[u].y;
                    const double x = e[u].x, delta = x - mean;
                    const double q0 = delta * y;
                    mean += fma(fma(-q0, nD, delta), y, q0);          /* == delta / n */
                    M2 += delta * (x - mean);
                }
            }
        }
    }
    const double rw = rowW ? rowW[row] : 1.0;              /* inner_prod(localityWeights, .), 1044-1046 (uniform: the host applies lw) */
    double a = active ? rw * summedVars : 0.0, b = active ? rw * M2 : 0.0;
    block_reduce2(a, b, sh);
    if (threadIdx.x == 0) out[(size_t) blockIdx.y * maxRowBlocks + blockIdx.x] = make_double2(a, b);
}

/* The same Welford statistics with the VRL list cut into `segs` segments (grid.z) that run concurrently and are merged in list
 * order afterwards (Chan et al.'s pairwise update: n = na + nb, d = mean_b - mean_a, M2 = M2a + M2b + d^2 na nb / n): a rank of
 * a multi-GPU job holds a few hundred rows, and one thread per row walking 10^5 VRLs leaves the GPU empty for ~18 ms whatever
 * the number of rows.  The double sums are associated differently from the sequential chain (last bits of a double that is
 * rounded to float afterwards: the same class as the segment carries of the variance sweeps). */
__global__ void __launch_bounds__(CL_THREADS) k_unclustered_seg(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                                const uint32_t *__restrict__ lists, uint32_t segs, double *__restrict__ part,
                                                                uint32_t maxRowBlocks) {
    const ClTask t = tasks[blockIdx.y];
    if (blockIdx.x >= t.rowBlocks) return;
    const uint32_t lr = blockIdx.x * CL_THREADS + threadIdx.x;
    const bool active = lr < t.nr;
    const uint32_t row = t.r0 + (active ? lr : 0);
    const uint32_t *list = lists + t.listOff;
    const uint32_t total = t.end - t.begin, seg = blockIdx.z;
    const uint32_t sb = t.begin + (uint32_t) (((uint64_t) total * seg) / segs), se = t.begin + (uint32_t) (((uint64_t) total * (seg + 1)) / segs);
    double mean = 0, M2 = 0, summedVars = 0;
    __shared__ double rcpN[CL_THREADS];
    for (uint32_t c0 = sb; c0 < se; c0 += CL_THREADS) {
        __syncthreads();
        rcpN[threadIdx.x] = 1.0 / (double) (c0 - sb + threadIdx.x + 1u);
        __syncthreads();
        const uint32_t cEnd = min(se, c0 + CL_THREADS);
        for (uint32_t k0 = c0; k0 < cEnd; k0 += 8) {
            float2 e[8];
#pragma unroll
            for (int u = 0; u < 8; u++) e[u] = (k0 + u < cEnd) ? R[(size_t) list[k0 + u] * ldR + row] : make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (k0 + u < cEnd) {
                    const double nD = (double) (k0 + u - sb + 1u), y = rcpN[k0 + u - c0];
                    summedVars += (double) e[u].y;
                    const double x = e[u].x, delta = x - mean;
                    const double q0 = delta * y;
                    mean += fma(fma(-q0, nD, delta), y, q0);          /* == delta / n */
                    M2 += delta * (x - mean);
                }
            }
        }
    }
    /* part[((task * segs + seg) * maxRowBlocks + rowBlock) * CL_THREADS + thread] x 3 */
    double *o = part + (((size_t) blockIdx.y * segs + seg) * maxRowBlocks + blockIdx.x) * CL_THREADS * 3 + threadIdx.x;
    o[0] = mean; o[CL_THREADS] = M2; o[2 * CL_THREADS] = summedVars;
}
__global__ void __launch_bounds__(CL_THREADS) k_unclustered_merge(const ClTask *__restrict__ tasks, uint32_t segs, const double *__restrict__ part,
                                                                  double2 *__restrict__ out, uint32_t maxRowBlocks, const double *__restrict__ rowW) {
    __shared__ double sh[2 * CL_THREADS / 32];
    const ClTask t = tasks[blockIdx.y];
    if (blockIdx.x >= t.rowBlocks) return;
    const uint32_t lr = blockIdx.x * CL_THREADS + threadIdx.x;
    const bool active = lr < t.nr;
    const uint32_t total = t.end - t.begin;
    double mean = 0, M2 = 0, sv = 0, n = 0;
    for (uint32_t seg = 0; seg < segs; seg++) {
        const uint32_t sb = (uint32_t) (((uint64_t) total * seg) / segs), se = (uint32_t) (((uint64_t) total * (seg + 1)) / segs);
        const double nb = (double) (se - sb);
        if (nb == 0) continue;
        const double *o = part + (((size_t) blockIdx.y * segs + seg) * maxRowBlocks + blockIdx.x) * CL_THREADS * 3 + threadIdx.x;
        const double mb = o[0], Mb = o[CL_THREADS];
        sv += o[2 * CL_THREADS];
        const double nn = n + nb, d = mb - mean;
        M2 = M2 + Mb + d * d * (n * nb / nn);
        mean = mean + d * (nb / nn);
        n = nn;
    }
    const double rw = rowW ? rowW[t.r0 + (active ? lr : 0)] : 1.0;
    double a = active ? rw * sv : 0.0, b = active ? rw * M2 : 0.0;
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

/* projections of the normalised columns on the direction (625-640), sequential fp32 in row order.
 * warp = VRL: the lanes fetch 32 consecutive rows of the column at a time (coalesced, all chunks in flight), and the strictly
 * ordered sums are carried by a shuffle-broadcast chain, so a column costs a few microseconds instead of 2 x nr dependent
 * DRAM round trips.  keys != NULL: also emit the 64-bit sort key that orders like std::sort on pair<float, uint32_t> (641). */
#define CL_PROJ_WARPS 4
#define CL_PROJ_MAXCH 8          /* columns of up to 256 rows stay in registers; longer ones are re-read per pass */
__global__ void __launch_bounds__(CL_PROJ_WARPS * 32) k_project(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                               const uint32_t *__restrict__ lists, const float *__restrict__ dir,
                                                               float *__restrict__ proj, uint64_t *__restrict__ keys) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t j = blockIdx.x * CL_PROJ_WARPS + (threadIdx.x >> 5);
    if (j >= t.end - t.begin) return;
    const uint32_t vid = lists[t.listOff + t.begin + j];
    const float2 *col = R + (size_t) vid * ldR + t.r0;
    const float *d = dir + t.dirOff;
    const uint32_t nch = (t.nr + 31) / 32;
    float x[CL_PROJ_MAXCH], dv[CL_PROJ_MAXCH];
    const bool inReg = nch <= CL_PROJ_MAXCH;
    if (inReg) {
#pragma unroll
        for (uint32_t c = 0; c < CL_PROJ_MAXCH; c++) {
            const uint32_t r = c * 32 + lane;
            x[c] = (c < nch && r < t.nr) ? col[r].x : 0.0f;
            dv[c] = (c < nch && r < t.nr) ? d[r] : 0.0f;
        }
    }
    float s = 0;
    for (uint32_t c = 0; c < nch; c++) {
        float xv;
        if (inReg) {
            xv = 0;
#pragma unroll
            for (uint32_t k = 0; k < CL_PROJ_MAXCH; k++) if (k == c) xv = x[k];
        } else { const uint32_t r = c * 32 + lane; xv = r < t.nr ? col[r].x : 0.0f; }
        const float u = fabsf(xv), sq = u * u;
        const uint32_t cnt = min(32u, t.nr - c * 32);
        for (uint32_t i = 0; i < cnt; i++) s += __shfl_sync(0xffffffffu, sq, i);
    }
    const float len = sqrtf(s);
    float p = 0;
    if (len != 0) {
        for (uint32_t c = 0; c < nch; c++) {
            float xv, dd;
            if (inReg) {
                xv = 0; dd = 0;
#pragma unroll
                for (uint32_t k = 0; k < CL_PROJ_MAXCH; k++) if (k == c) { xv = x[k]; dd = dv[k]; }
            } else { const uint32_t r = c * 32 + lane; xv = r < t.nr ? col[r].x : 0.0f; dd = r < t.nr ? d[r] : 0.0f; }
            const float term = dd * (xv / len);
            const uint32_t cnt = min(32u, t.nr - c * 32);
            for (uint32_t i = 0; i < cnt; i++) p += __shfl_sync(0xffffffffu, term, i);
        }
    }
    if (lane == 0) {
        if (proj) proj[t.outOff + j] = p;
        if (keys) {
            const float q = p + 0.0f;                                   /* -0.0 and +0.0 compare equal in the pair order */
            uint32_t b = __float_as_uint(q);
            b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
            keys[t.outOff + j] = ((uint64_t) b << 32) | vid;
        }
    }
}

__global__ void k_scatter_lists(const ClTask *__restrict__ tasks, const uint32_t *__restrict__ staged, uint32_t *__restrict__ lists) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < t.end - t.begin) lists[t.listOff + t.begin + j] = staged[t.outOff + j];
}

/*
 * calculateClusterVariance (1058-1120) as a segmented, fully parallel pipeline.
 *
 * Reference recurrence per row r over the ordered VRLs k (x = mean, w = column weight, W = prefix weight):
 *     M_r(k) = (W_k / W_{k-1})^2 M_r(k-1) + (1/w_k + 1/W_{k-1}) (w_k S_r(k-1) - W_{k-1} x_r(k))^2,   S_r = prefix sum of x_r
 *     first(k)  = sum_r lw M_r(k) / W_k          second(k) = sum_r lw W_k sumVars_r(k),  sumVars_r = prefix sum of var_r / w
 * The factors (W_k / W_{k-1})^2 telescope: M_r(k) = W_k^2 sum_{j<=k} b_r(j) / W_j^2, and the sum over rows commutes with the
 * prefix over k.  So a range is cut into segments of CL_SEG steps that run concurrently:
 *     k_weights    w_k (gather) and W_k (block scan)                                  per task
 *     k_seg_sums   column sums of x over each segment                                  per segment x row block
 *     k_carry      exclusive prefix of those sums over the segments -> S carry-in      per task x row block
 *     k_seg_main   thread = row walks its segment: B(k) = sum_r tmp_r(k)^2, V(k) = sum_r var_r(k) / w_k   (row-reduced per step)
 *     k_final      Q = scan(c2_k B_k / W_k^2), SV = scan(V_k); first = lw W_k Q_k, second = lw W_k SV_k -> (float, float)
 * All sums are double; they differ from the reference's sequential order in the last double bits only.
 */
#define CL_SEG 256

struct SegDesc { uint32_t task, seg; };

/* inclusive block scan of one value per thread with a running carry (all threads get the same carry back) */
__device__ __forceinline__ double block_scan_incl(double v, double &carry, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const double n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    if (lane == 31) sh[warp] = v;
    __syncthreads();
    double off = 0, total = 0;
    for (int i = 0; i < CL_THREADS / 32; i++) { const double t = sh[i]; if (i < warp) off += t; total += t; }
    const double r = carry + off + v;
    __syncthreads();
    carry += total;
    return r;
}

__global__ void __launch_bounds__(CL_THREADS) k_weights(const ClTask *__restrict__ tasks, const uint32_t *__restrict__ lists, const float *__restrict__ cw,
                                                        double *__restrict__ wArr, double *__restrict__ WArr) {
    __shared__ double sh[CL_THREADS / 32];
    const ClTask t = tasks[blockIdx.x];
    const uint32_t n = t.end - t.begin;
    double carry = 0;
    for (uint32_t k0 = 0; k0 < n; k0 += CL_THREADS) {
        const uint32_t k = k0 + threadIdx.x;
        double w = 0;
        if (k < n) {
            const uint32_t idx = t.reverse ? (t.end - 1 - k) : (t.begin + k);
            w = (double) cw[t.cwOff + lists[t.listOff + idx]];
        }
        const double W = block_scan_incl(w, carry, sh);
        if (k < n) { wArr[t.stepOff + k] = w; WArr[t.stepOff + k] = W; }
    }
}

__global__ void __launch_bounds__(CL_THREADS) k_seg_sums(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                         const SegDesc *__restrict__ segs, const uint32_t *__restrict__ lists, double *__restrict__ carry) {
    const SegDesc sd = segs[blockIdx.x];
    const ClTask t = tasks[sd.task];
    if (blockIdx.y >= t.rowBlocks || t.nseg <= 1) return;
    const uint32_t lr = blockIdx.y * CL_THREADS + threadIdx.x;
    const uint32_t row = t.r0 + (lr < t.nr ? lr : 0);
    const uint32_t n = t.end - t.begin, k0 = sd.seg * t.segLen, k1 = min(n, k0 + t.segLen);
    const uint32_t *list = lists + t.listOff;
    double s = 0;
    for (uint32_t k = k0; k < k1; k++) {
        const uint32_t idx = t.reverse ? (t.end - 1 - k) : (t.begin + k);
        s += (double) R[(size_t) list[idx] * ldR + row].x;
    }
    carry[t.carryOff + (uint64_t) sd.seg * (t.rowBlocks * CL_THREADS) + lr] = s;
}

__global__ void __launch_bounds__(CL_THREADS) k_carry(const ClTask *__restrict__ tasks, double *__restrict__ carry) {
    const ClTask t = tasks[blockIdx.x];
    if (blockIdx.y >= t.rowBlocks || t.nseg <= 1) return;
    const uint32_t lr = blockIdx.y * CL_THREADS + threadIdx.x;
    const uint64_t stride = (uint64_t) t.rowBlocks * CL_THREADS;
    double run = 0;
    for (uint32_t sgi = 0; sgi < t.nseg; sgi++) {
        double *p = &carry[t.carryOff + sgi * stride + lr];
        const double v = *p;
        *p = run;
        run += v;
    }
}

__global__ void __launch_bounds__(CL_THREADS) k_seg_main(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                         const SegDesc *__restrict__ segs, const uint32_t *__restrict__ lists,
                                                         const double *__restrict__ wArr, const double *__restrict__ WArr,
                                                         const double *__restrict__ carry, double2 *__restrict__ partial, const double *__restrict__ rowW) {
    __shared__ double sB[CL_CHUNK][CL_THREADS];
    __shared__ double sV[CL_CHUNK][CL_THREADS];
    const SegDesc sd = segs[blockIdx.x];
    const ClTask t = tasks[sd.task];
    if (blockIdx.y >= t.rowBlocks) return;
    const uint32_t tid = threadIdx.x;
    const uint32_t lr = blockIdx.y * CL_THREADS + tid;
    const bool active = lr < t.nr;
    const uint32_t row = t.r0 + (active ? lr : 0);
    const uint32_t n = t.end - t.begin, kBeg = sd.seg * t.segLen, kEnd = min(n, kBeg + t.segLen);
    const uint32_t *list = lists + t.listOff;
    const double *w = wArr + t.stepOff, *W = WArr + t.stepOff;
    double S = (t.nseg > 1) ? carry[t.carryOff + (uint64_t) sd.seg * (t.rowBlocks * CL_THREADS) + lr] : 0.0;
    const double rw = rowW ? rowW[row] : 1.0;              /* inner_prod(localityWeights, .), 1103-1104 (uniform: k_final applies lw) */
    for (uint32_t k0 = kBeg; k0 < kEnd; k0 += CL_CHUNK) {
        const uint32_t cnt = min((uint32_t) CL_CHUNK, kEnd - k0);
        float2 x[CL_CHUNK];
#pragma unroll
        for (uint32_t j = 0; j < CL_CHUNK; j++) {
            if (j < cnt) {
                const uint32_t k = k0 + j;
                const uint32_t idx = t.reverse ? (t.end - 1 - k) : (t.begin + k);
                x[j] = R[(size_t) list[idx] * ldR + row];
            } else x[j] = make_float2(0, 0);
        }
#pragma unroll
        for (uint32_t j = 0; j < CL_CHUNK; j++) {
            if (j < cnt) {
                const uint32_t k = k0 + j;
                const double wk = w[k], Wprev = k ? W[k - 1] : 0.0;
                const double xm = x[j].x;
                const double tmp = wk * S - Wprev * xm;
                S += xm;
                sB[j][tid] = active ? rw * (tmp * tmp) : 0.0;
                sV[j][tid] = active ? rw * ((double) x[j].y / wk) : 0.0;
            }
        }
        __syncthreads();
        const uint32_t warp = tid >> 5, lane = tid & 31;
        for (uint32_t j = warp; j < cnt; j += CL_THREADS / 32) {
            double a = 0, b = 0;
            for (uint32_t i = lane; i < CL_THREADS; i += 32) { a += sB[j][i]; b += sV[j][i]; }
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
            if (lane == 0) partial[t.partOff + (uint64_t) (k0 + j) * t.rowBlocks + blockIdx.y] = make_double2(a, b);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CL_THREADS) k_final(const ClTask *__restrict__ tasks, const double *__restrict__ wArr, const double *__restrict__ WArr,
                                                      const double2 *__restrict__ partial, float2 *__restrict__ out, int weightedRows) {
    __shared__ double sh[CL_THREADS / 32];
    const ClTask t = tasks[blockIdx.x];
    const uint32_t n = t.end - t.begin;
    const double *w = wArr + t.stepOff, *W = WArr + t.stepOff;
    double carryQ = 0, carryV = 0;
    for (uint32_t k0 = 0; k0 < n; k0 += CL_THREADS) {
        const uint32_t k = k0 + threadIdx.x;
        double tq = 0, tv = 0, Wk = 1;
        if (k < n) {
            double B = 0, V = 0;
            for (uint32_t rb = 0; rb < t.rowBlocks; rb++) { const double2 p = partial[t.partOff + (uint64_t) k * t.rowBlocks + rb]; B += p.x; V += p.y; }
            Wk = W[k];
            if (k > 0) tq = (1.0 / w[k] + 1.0 / W[k - 1]) * B / (Wk * Wk);
            tv = V;
        }
        const double Q = block_scan_incl(tq, carryQ, sh);
        const double SV = block_scan_incl(tv, carryV, sh);
        if (k < n) {
            float2 r;
            const double lwF = weightedRows ? 1.0 : t.lw;
            r.x = (k == 0) ? 0.0f : (float) (lwF * (Wk * Q));
            r.y = (float) (lwF * (SV * Wk));
            if (!t.finalOnly) out[t.outOff + k] = r;
            else if (k == n - 1) out[t.outOff] = r;
        }
    }
}

/* ---- device rounds (counter stream): centre picking, sort and argmin without a host round trip ----------------- */
struct RoundResult {            /* what the host needs back from one split */
    uint32_t bestIndex, vrl1, vrl2, flags;      /* flags: 1 = degenerate centres (host path), 2 = non-positive weight sum */
    float2 head, tail;                          /* variance pairs of the two halves */
    float best, second;
    uint32_t firstVrl, lastVrl;                 /* VRL ids at the ends of the sorted range (singleton halves) */
};

/*
 * One block per split: weightedSample x 2 (Preprocessor.cpp:597-602,1534-1580) -- sequential fp32 sums in list order by
 * one thread over shared-memory chunks that the block stages (the second draw sees the first centre's weight as 0) --
 * followed by the split direction (604-623): three threads accumulate |col1|, |col2|, |col2 - col1| in row order.
 */
#define CL_PICK_CHUNK 2048
__global__ void __launch_bounds__(128) k_pick_direction(const float2 *__restrict__ R, uint32_t ldR, const ClTask *__restrict__ tasks,
                                                        const uint32_t *__restrict__ lists, const float *__restrict__ cw,
                                                        RoundResult *__restrict__ res, float *__restrict__ dir) {
    __shared__ float sw[CL_PICK_CHUNK];
    __shared__ float sAcc; __shared__ uint32_t sIdx, sFound;
    __shared__ float sNorm[3];
    const ClTask t = tasks[blockIdx.x];
    const uint32_t n = t.end - t.begin;
    const uint32_t *list = lists + t.listOff + t.begin;
    const float *w = cw + t.cwOff;
    uint32_t idx1 = 0xffffffffu, picks[2] = {0, 0}, flags = 0;
    for (int draw = 0; draw < 2; draw++) {
        if (threadIdx.x == 0) sAcc = 0.0f;
        for (uint32_t c0 = 0; c0 < n; c0 += CL_PICK_CHUNK) {
            const uint32_t cnt = min((uint32_t) CL_PICK_CHUNK, n - c0);
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) sw[i] = (c0 + i == idx1) ? 0.0f : w[list[c0 + i]];
            __syncthreads();
            if (threadIdx.x == 0) { float a = sAcc; for (uint32_t i = 0; i < cnt; i++) a += sw[i]; sAcc = a; }
        }
        __syncthreads();
        const float weightSum = sAcc;
        if (!(weightSum > 0)) flags |= 2u;
        const float alpha = (draw == 0 ? t.u1 : t.u2) * weightSum;
        __syncthreads();
        if (threadIdx.x == 0) { sAcc = 0.0f; sIdx = 0; sFound = 0; }
        for (uint32_t c0 = 0; c0 < n; c0 += CL_PICK_CHUNK) {
            const uint32_t cnt = min((uint32_t) CL_PICK_CHUNK, n - c0);
            __syncthreads();
            if (sFound) break;
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) sw[i] = (c0 + i == idx1) ? 0.0f : w[list[c0 + i]];
            __syncthreads();
            if (threadIdx.x == 0) {
                float a = sAcc;
                for (uint32_t i = 0; i < cnt; i++) { a += sw[i]; if (a >= alpha) { sIdx = c0 + i; sFound = 1; break; } }
                sAcc = a;
            }
        }
        __syncthreads();
        picks[draw] = sIdx;
        if (draw == 0) idx1 = sIdx;
        __syncthreads();
    }
    const uint32_t vrl1 = list[picks[0]], vrl2 = list[picks[1]];
    /* direction: norms in row order, one accumulator per thread (three independent sequential chains) */
    const float2 *c1 = R + (size_t) vrl1 * ldR + t.r0, *c2 = R + (size_t) vrl2 * ldR + t.r0;
    float *cA = sw, *cB = sw + CL_PICK_CHUNK / 2;
    float acc = 0;
    for (uint32_t r0 = 0; r0 < t.nr; r0 += CL_PICK_CHUNK / 2) {
        const uint32_t cnt = min((uint32_t) CL_PICK_CHUNK / 2, t.nr - r0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) { cA[i] = c1[r0 + i].x; cB[i] = c2[r0 + i].x; }
        __syncthreads();
        if (threadIdx.x == 0) for (uint32_t i = 0; i < cnt; i++) acc += fabsf(cA[i]) * fabsf(cA[i]);
        else if (threadIdx.x == 32) for (uint32_t i = 0; i < cnt; i++) acc += fabsf(cB[i]) * fabsf(cB[i]);
        else if (threadIdx.x == 64) for (uint32_t i = 0; i < cnt; i++) { const float d = cB[i] - cA[i]; acc += fabsf(d) * fabsf(d); }
    }
    if (threadIdx.x == 0) sNorm[0] = sqrtf(acc);
    if (threadIdx.x == 32) sNorm[1] = sqrtf(acc);
    if (threadIdx.x == 64) sNorm[2] = sqrtf(acc);
    __syncthreads();
    const bool degenerate = !(sNorm[0] != 0 && sNorm[1] != 0 && sNorm[2] != 0);
    if (degenerate) flags |= 1u;
    const float diffLen = sNorm[2];
    for (uint32_t r = threadIdx.x; r < t.nr; r += blockDim.x) dir[t.dirOff + r] = degenerate ? 0.0f : (c2[r].x - c1[r].x) / diffLen;
    if (threadIdx.x == 0) {
        RoundResult r;
        r.bestIndex = 0xffffffffu; r.flags = flags; r.vrl1 = vrl1; r.vrl2 = vrl2;
        r.head = make_float2(0, 0); r.tail = make_float2(0, 0); r.best = 0; r.second = 0; r.firstVrl = 0; r.lastVrl = 0;
        res[blockIdx.x] = r;
    }
}

__global__ void k_scatter_sorted(const ClTask *__restrict__ tasks, const uint64_t *__restrict__ keys, uint32_t *__restrict__ lists) {
    const ClTask t = tasks[blockIdx.y];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < t.end - t.begin) lists[t.listOff + t.begin + j] = (uint32_t) (keys[t.outOff + j] & 0xffffffffull);
}

/* first minimum of head + tail variance over the split index (664-675), plus the runner-up for the near-tie flag */
__global__ void __launch_bounds__(CL_THREADS) k_argmin(const ClTask *__restrict__ vt, const float2 *__restrict__ pairs, const uint32_t *__restrict__ lists,
                                                       RoundResult *__restrict__ res) {
    __shared__ float sb[CL_THREADS], ss[CL_THREADS]; __shared__ uint32_t si[CL_THREADS];
    const ClTask f = vt[2 * blockIdx.x], r = vt[2 * blockIdx.x + 1];
    const uint32_t n = f.end - f.begin;
    const float2 *fromStart = pairs + f.outOff, *fromEnd = pairs + r.outOff;
    float best = INFINITY, second = INFINITY; uint32_t bi = 0xffffffffu;
    for (uint32_t k = 1 + threadIdx.x; k < n; k += CL_THREADS) {
        const float2 h = fromStart[k - 1], tl = fromEnd[n - 1 - k];
        const float v = h.x + h.y + tl.x + tl.y;
        if (v < best) { second = best; best = v; bi = k; }
        else if (v < second) second = v;
    }
    sb[threadIdx.x] = best; ss[threadIdx.x] = second; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = CL_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const float b1 = sb[threadIdx.x], b2 = sb[threadIdx.x + o], s1 = ss[threadIdx.x], s2 = ss[threadIdx.x + o];
            const uint32_t i1 = si[threadIdx.x], i2 = si[threadIdx.x + o];
            if (b2 < b1 || (b2 == b1 && i2 < i1)) { sb[threadIdx.x] = b2; si[threadIdx.x] = i2; ss[threadIdx.x] = fminf(s2, b1); }
            else ss[threadIdx.x] = fminf(s1, b2);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        RoundResult &o = res[blockIdx.x];
        o.bestIndex = si[0]; o.best = sb[0]; o.second = ss[0];
        if (si[0] != 0xffffffffu) { o.head = fromStart[si[0] - 1]; o.tail = fromEnd[n - 1 - si[0]]; }
        o.firstVrl = lists[f.listOff + f.begin]; o.lastVrl = lists[f.listOff + f.end - 1];
    }
}

/* ---- host side ------------------------------------------------------------------------------------ */
namespace {

/* optional phase timers of the clustering rounds (ALVRL_PROFILE=1) */
struct Prof {
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; uint64_t rounds = 0, tasks = 0, steps = 0;
    bool on = getenv("ALVRL_PROFILE") != nullptr;
    static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    double *evMs = nullptr;
    ~Prof() {
        if (on && rounds && evMs) fprintf(stderr, "[alvrl clustering] device ms: pick+dir %.1f project %.1f sort %.1f scatter %.1f variance(incl. uploads) %.1f argmin %.1f\n", evMs[0], evMs[1], evMs[2], evMs[3], evMs[4], evMs[5]);
        if (on && rounds) fprintf(stderr, "[alvrl clustering] rounds %llu tasks %llu steps %llu | ms: sample+tasks %.1f dir+proj %.1f d2h-proj %.1f sort %.1f "
                                  "h2d+variance %.1f d2h-pairs %.1f argmin %.1f\n", (unsigned long long) rounds, (unsigned long long) tasks,
                                  (unsigned long long) steps, t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
    }
};

struct ClusterNode {                                                           /* Preprocessor.cpp:289-298 */
    float undersamplingVar, integrationVar; uint32_t begin, end;
    bool operator<(const ClusterNode &o) const { return undersamplingVar + integrationVar < o.undersamplingVar + o.integrationVar; }
};

} // namespace (anonymous)
} // namespace alvrl
#include "refine.cuh"
namespace alvrl {
namespace {

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
    bool listsStale = false;      /* the device holds a newer permutation than `vrls` */
    bool currentIsBest = true;    /* adaptive refinement: the current state is the best-so-far snapshot */
    /* depthCorrection != 1 (403-408, 455-470): splits so far, splits at the best convergence constant, split count of pass 2 */
    uint32_t splitsDone = 0, bestSplits = 0, fixedSplits = 0;

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
    alvrl_ctx *c = nullptr; cudaStream_t st = nullptr; uint32_t N = 0, ldR = 0; const float2 *R = nullptr;
    /* per-row locality weights (neighbourWeight > 0: getLocalMatrix 796-820), indexed like the rows of R; nullptr: uniform lw */
    const double *rowW = nullptr;
    DevBuf<uint32_t> dLists; DevBuf<float> dCw;
    DevBuf<ClTask> dTasks; DevBuf<float> dDir, dProj; DevBuf<uint32_t> dFlags, dStaged;
    DevBuf<double2> dPartial, dUncl; DevBuf<double> dUnclPart; DevBuf<double> dW1, dW2, dCarry; DevBuf<float2> dPairs; DevBuf<SegDesc> dSegs;
    std::vector<Inst *> insts;
    Prof prof;

    template <typename T> static void ensure(DevBuf<T> &b, size_t n) { if (b.n < n) b.alloc(n + n / 4 + 16); }
    uint32_t launchCount = 0; bool serialSort = false;
    bool deviceRounds = false;        /* counter stream: centres / sort / argmin on the device (no host round trip per phase) */
    void launches(uint32_t k) { launchCount += k; }

    void allocInstances() {
        prof.evMs = evMs;
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
        k_column_weights<<<dim3((N + 127) / 128, (uint32_t) tasks.size()), 128, 0, st>>>(R, ldR, N, dTasks.p, dCw.p, rowW);
        launches(1);
        ALVRL_CUDA(cudaGetLastError());
        ensure(dFlags, tasks.size());
        k_cw_finish<<<(uint32_t) tasks.size(), CL_THREADS, 0, st>>>(N, dTasks.p, dCw.p, dFlags.p);
        launches(1);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<uint32_t> bad(tasks.size());
        dFlags.download(bad.data(), bad.size(), st);
        for (uint32_t b : bad) if (b) throw Error(ALVRL_ERR_ARG, "Invalid calculated average column weight");
        cwOnHost = false;
        if (!lazyMirrors) ensureHostCw();
    }
    /* Host mirrors of the column weights and of the VRL permutations are only needed by the host-side paths (weightedSample of
     * the representatives on the host, host-driven rounds, getVrlsPerCluster): with lazyMirrors they are fetched on demand
     * (2 x 40 MB of pageable copies per frame at C2 otherwise). */
    bool lazyMirrors = false, cwOnHost = false;
    void ensureHostCw() {
        if (cwOnHost) return;
        for (Inst *in : insts) {
            in->cw.resize(N);
            ALVRL_CUDA(cudaMemcpyAsync(in->cw.data(), dCw.p + in->cwOff, (size_t) N * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        ALVRL_CUDA(cudaStreamSynchronize(st));
        cwOnHost = true;
    }
    /* sameLists: every instance starts from the same list (construct): one upload, replicated on the device */
    void uploadLists(bool sameLists = false) {
        if (sameLists && !insts.empty()) {
            const size_t n = insts[0]->vrls.size();
            ALVRL_CUDA(cudaMemsetAsync(dLists.p, 0, insts.size() * (size_t) N * sizeof(uint32_t), st));
            ALVRL_CUDA(cudaMemcpyAsync(dLists.p, insts[0]->vrls.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            for (size_t i = 1; i < insts.size(); i++)
                ALVRL_CUDA(cudaMemcpyAsync(dLists.p + i * (size_t) N, dLists.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            ALVRL_CUDA(cudaStreamSynchronize(st));
            return;
        }
        std::vector<uint32_t> all(insts.size() * (size_t) N, 0);
        for (size_t i = 0; i < insts.size(); i++) std::copy(insts[i]->vrls.begin(), insts[i]->vrls.end(), all.begin() + i * (size_t) N);
        ALVRL_CUDA(cudaMemcpyAsync(dLists.p, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ALVRL_CUDA(cudaStreamSynchronize(st));
    }
    /* calculateClusterVariance for a batch of ranges; results in dPairs at task.outOff (n pairs, or 1 when finalOnly).
     * prepareVariance assigns offsets and uploads the task / segment tables (host-known: can run before the kernels that
     * produce the sorted lists are even launched); launchVariance enqueues the pipeline. */
    struct VarPlan { uint64_t out = 0, carry = 0; uint32_t segTotal = 0, maxRb = 1, T = 0; };
    VarPlan prepareVariance(std::vector<ClTask> &tasks, DevBuf<ClTask> &dT) {
        VarPlan pl;
        uint64_t steps = 0, part = 0;
        std::vector<SegDesc> segs;
        for (size_t i = 0; i < tasks.size(); i++) {
            ClTask &t = tasks[i];
            const uint32_t n = t.end - t.begin;
            t.outOff = pl.out; t.stepOff = steps; t.carryOff = pl.carry; t.partOff = part;
            part += (uint64_t) n * t.rowBlocks;
            t.segLen = n <= 8192 ? 32 : CL_SEG;
            t.nseg = (n + t.segLen - 1) / t.segLen; t.segOff = pl.segTotal;
            pl.out += t.finalOnly ? 1 : n; steps += n;
            if (t.nseg > 1) pl.carry += (uint64_t) t.nseg * t.rowBlocks * CL_THREADS;
            for (uint32_t sgi = 0; sgi < t.nseg; sgi++) segs.push_back(SegDesc{(uint32_t) i, sgi});
            pl.segTotal += t.nseg; pl.maxRb = std::max(pl.maxRb, t.rowBlocks);
        }
        pl.T = (uint32_t) tasks.size();
        ensure(dW1, steps); ensure(dW2, steps); ensure(dCarry, std::max<uint64_t>(pl.carry, 1)); ensure(dPartial, part); ensure(dPairs, pl.out);
        dT.upload(tasks, st);
        dSegs.upload(segs, st);
        return pl;
    }
    void launchVariance(const VarPlan &pl, DevBuf<ClTask> &dT) {
        k_weights<<<pl.T, CL_THREADS, 0, st>>>(dT.p, dLists.p, dCw.p, dW1.p, dW2.p);
        if (pl.carry) {
            k_seg_sums<<<dim3(pl.segTotal, pl.maxRb), CL_THREADS, 0, st>>>(R, ldR, dT.p, dSegs.p, dLists.p, dCarry.p);
            k_carry<<<dim3(pl.T, pl.maxRb), CL_THREADS, 0, st>>>(dT.p, dCarry.p);
            launches(2);
        }
        k_seg_main<<<dim3(pl.segTotal, pl.maxRb), CL_THREADS, 0, st>>>(R, ldR, dT.p, dSegs.p, dLists.p, dW1.p, dW2.p, dCarry.p, dPartial.p, rowW);
        k_final<<<pl.T, CL_THREADS, 0, st>>>(dT.p, dW1.p, dW2.p, dPartial.p, dPairs.p, rowW ? 1 : 0);
        launches(3);
        ALVRL_CUDA(cudaGetLastError());
    }
    uint64_t runVariance(std::vector<ClTask> &tasks, DevBuf<ClTask> &dT) {
        const VarPlan pl = prepareVariance(tasks, dT);
        launchVariance(pl, dT);
        return pl.out;
    }
    /* variance of whole ranges (addCluster(begin, end), 576-579): one (uvar, ivar) pair per task */
    std::vector<float2> rangeVariances(std::vector<ClTask> &tasks) {
        for (ClTask &t : tasks) { t.finalOnly = 1; t.reverse = 0; }
        const uint64_t out = runVariance(tasks, dTasks);
        std::vector<float2> res(out);
        dPairs.download(res.data(), out, st);
        return res;
    }
    /* Clustering constructor (301-341) for all instances: initial clusters + unclustered variance */
    void construct(const std::vector<std::vector<uint32_t>> &vrlsPerCluster) {
        double c0_ = Prof::now();
        auto clap = [&](const char *what) { if (prof.on) { cudaStreamSynchronize(st); const double n_ = Prof::now(); fprintf(stderr, "[alvrl clustering]   construct %s %.1f ms\n", what, n_ - c0_); c0_ = n_; } };
        uint32_t total = 0;
        for (auto &cl : vrlsPerCluster) total += (uint32_t) cl.size();
        {   /* every object starts from the same list: the host mirrors are filled by a few threads */
            std::vector<uint32_t> flat; flat.reserve(total);
            for (auto &cl : vrlsPerCluster) flat.insert(flat.end(), cl.begin(), cl.end());
            const unsigned nt = insts.size() >= 8 ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1u;
            std::atomic<size_t> next(0);
            auto fill = [&]() { for (size_t k = next++; k < insts.size(); k = next++) insts[k]->vrls = flat; };
            std::vector<std::thread> pool;
            for (unsigned ti = 1; ti < nt; ti++) pool.emplace_back(fill);
            fill();
            for (auto &t : pool) t.join();
        }
        for (Inst *in : insts) {
            in->numVrlsTotal = N; in->underVar = 0; in->intVar = 0; in->pq.clear(); in->singletons.clear();
            if (!rowW && std::fabs((float) (in->lw * in->nr) - 1) > 1e-3) throw Error(ALVRL_ERR_ARG, "Incorrect normalization in localityWeights");
            if (in->pixelUndersampling <= 0 || in->pixelUndersampling > 1) throw Error(ALVRL_ERR_ARG, "Invalid pixel undersampling");
        }
        clap("host lists");
        uploadLists(true);
        clap("uploadLists");
        std::vector<ClTask> tasks;
        for (Inst *in : insts) {
            uint32_t begin = 0;
            for (auto &cl : vrlsPerCluster) { ClTask t = baseTask(*in); t.begin = begin; t.end = begin + (uint32_t) cl.size(); tasks.push_back(t); begin = t.end; }
        }
        std::vector<float2> res = rangeVariances(tasks);
        clap("rangeVariances");
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
        /* few row blocks (a rank of a multi-GPU job): the VRL list is cut into segments that run concurrently */
        uint32_t totalRb = 0; for (const ClTask &t : ut) totalRb += t.rowBlocks;
        uint32_t segs = 1;
        if (!getenv("ALVRL_UNCL_SEQ")) { while (segs < 64u && totalRb * segs < 592u && (total / (2u * segs)) >= 512u) segs <<= 1; }
        if (segs > 1) {
            ensure(dUnclPart, ut.size() * (size_t) segs * maxRb * CL_THREADS * 3);
            k_unclustered_seg<<<dim3(maxRb, (uint32_t) ut.size(), segs), CL_THREADS, 0, st>>>(R, ldR, dTasks.p, dLists.p, segs, dUnclPart.p, maxRb);
            k_unclustered_merge<<<dim3(maxRb, (uint32_t) ut.size()), CL_THREADS, 0, st>>>(dTasks.p, segs, dUnclPart.p, dUncl.p, maxRb, rowW);
            launches(2);
        } else {
            k_unclustered<<<dim3(maxRb, (uint32_t) ut.size()), CL_THREADS, 0, st>>>(R, ldR, dTasks.p, dLists.p, dUncl.p, maxRb, rowW);
            launches(1);
        }
        ALVRL_CUDA(cudaGetLastError());
        std::vector<double2> u(ut.size() * (size_t) maxRb);
        dUncl.download(u.data(), u.size(), st);
        clap("unclustered");
        for (size_t i = 0; i < insts.size(); i++) {
            double a = 0, b = 0;
            for (uint32_t rb = 0; rb < insts[i]->rowBlocks; rb++) { a += u[i * maxRb + rb].x; b += u[i * maxRb + rb].y; }
            const double lwU = rowW ? 1.0 : insts[i]->lw;                     /* per-row weights are applied by the kernel */
            insts[i]->unclIntVar = (float) (lwU * a);
            insts[i]->tracingVar = (float) (lwU * b - (double) insts[i]->unclIntVar);
        }
    }
    /* Clustering::refine for the given instances, all advancing one split per round (380-489, 590-684) */
    float depthCorrection = 1.0f;     /* of the objects of this workspace (the per-slice ones: Preprocessor.cpp:268-269) */
    void refine(const std::vector<Inst *> &which, float undersampling) {
        for (Inst *in : which) {
            in->failed = false; in->done = false; in->refining = true;
            in->adaptive = undersampling <= 0;
            if (in->adaptive) {                                                 /* refineAdaptively, 402-423 */
                if (in->numMulti() <= 0) { in->done = true; continue; }
                if (in->unclusteredVariance() == 0) { in->done = true; in->failed = true; continue; }
                in->bestConstant = in->convergenceConstant();
                in->snapshot(); in->currentIsBest = true;
            } else {                                                            /* refineFixedDepth, 387-399 */
                in->targetClusters = (uint32_t) (0.5 + in->numVrlsTotal / undersampling);
                if (in->numClusters() >= in->targetClusters || in->numMulti() <= 0) in->done = true;
            }
        }
        auto afterSplit = [](Inst *in) {
            in->splitsDone++;
            if (in->adaptive) {
                const float curr = in->convergenceConstant();                   /* 436-452 */
                in->currentIsBest = curr < in->bestConstant;
                if (in->currentIsBest) { in->snapshot(); in->bestConstant = curr; in->bestSplits = in->splitsDone; }
                if (in->lowerBound() >= in->bestConstant || in->numMulti() == 0) { in->restore(); in->done = true; }
            } else if (in->fixedSplits) { if (in->splitsDone >= in->fixedSplits || in->numMulti() == 0) in->done = true; }   /* 455-470 */
            else if (!(in->numClusters() < in->targetClusters && in->numMulti() > 0)) in->done = true;
        };
        /* depthCorrection != 1 (per-slice objects only, 268-269): the queue before the first split is what the second pass starts from */
        const bool secondPass = depthCorrection != 1.0f && undersampling <= 0;
        struct Initial { std::vector<ClusterNode> pq; std::list<uint32_t> singletons; float underVar, intVar; };
        std::vector<Initial> initial;
        if (secondPass) for (Inst *in : which) { initial.push_back(Initial{in->pq, in->singletons, in->underVar, in->intVar}); in->splitsDone = in->bestSplits = in->fixedSplits = 0; }
        auto drive = [&]() {
            if (deviceRounds && !getenv("ALVRL_HOST_ROUNDS")) {
                refineDevice(which);
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
                if (deviceRounds) splitRoundDevice(round); else splitRound(round);
                for (Inst *in : round) afterSplit(in);
            }
        };
        drive();
        if (secondPass) {
            /* the VRL lists stay as the first pass left them (makeRefinementSnapshot does not cover m_vrls, 686-699) */
            for (size_t i = 0; i < which.size(); i++) {
                Inst *in = which[i];
                if (in->failed || initial[i].pq.empty()) continue;
                const int corrected = (int) (0.5 + depthCorrection * (float) in->bestSplits);        /* 460 */
                in->pq = initial[i].pq; in->singletons = initial[i].singletons; in->underVar = initial[i].underVar; in->intVar = initial[i].intVar;
                in->adaptive = false; in->fixedSplits = (uint32_t) std::max(0, corrected); in->splitsDone = 0;
                in->done = corrected <= 0 || in->numMulti() == 0;
            }
            drive();
            for (Inst *in : which) in->fixedSplits = 0;
        }
        if (!lazyMirrors) syncLists(which);
        for (Inst *in : which) in->refining = false;
    }
    /* sampleRepresentatives for every object of this workspace on the device (counter stream); an object whose weights the
     * kernel declines is sampled on the host from the mirrors */
    void sampleRepresentativesDevice(std::vector<std::vector<uint32_t>> &selected, std::vector<std::vector<float>> &weights) {
        std::vector<RepObj> objs; std::vector<uint2> clusters; std::vector<Inst *> who;
        for (Inst *in : insts) {
            if (in->failed) continue;
            RepObj o; memset(&o, 0, sizeof(o));
            if (!in->smp->counterState(o.rngKey, o.rngPos)) { hostSample(in, selected, weights); continue; }
            o.listOff = in->listOff; o.cwOff = in->cwOff; o.firstCluster = (uint32_t) clusters.size(); o.numClusters = (uint32_t) in->pq.size();
            for (const ClusterNode &cn : in->pq) clusters.push_back(make_uint2(cn.begin, cn.end));
            objs.push_back(o); who.push_back(in);
        }
        if (objs.empty()) return;
        uint32_t maxC = 1; for (const RepObj &o : objs) maxC = std::max(maxC, o.numClusters);
        clusters.push_back(make_uint2(0, 0));                                   /* never empty */
        DevBuf<RepObj> dObjs; DevBuf<uint2> dCl; DevBuf<uint32_t> dRepr, dBad; DevBuf<float> dWt;
        dObjs.upload(objs, st); dCl.upload(clusters, st);
        dRepr.alloc(clusters.size()); dWt.alloc(clusters.size()); dBad.alloc(objs.size());
        ALVRL_CUDA(cudaMemsetAsync(dBad.p, 0, objs.size() * sizeof(uint32_t), st));
        k_sample_representatives<<<dim3((maxC + 127) / 128, (uint32_t) objs.size()), 128, 0, st>>>(dObjs.p, dCl.p, dLists.p, dCw.p, dRepr.p, dWt.p, dBad.p);
        launches(1);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<uint32_t> repr(clusters.size()), bad(objs.size()); std::vector<float> wt(clusters.size());
        dRepr.download(repr.data(), repr.size(), st); dWt.download(wt.data(), wt.size(), st); dBad.download(bad.data(), bad.size(), st);
        for (size_t k = 0; k < who.size(); k++) {
            Inst *in = who[k];
            if (bad[k]) { hostSample(in, selected, weights); continue; }
            std::vector<uint32_t> &r = selected[in->id]; std::vector<float> &w = weights[in->id];
            r.resize(in->numClusters()); w.resize(in->numClusters());
            size_t i = 0;
            for (uint32_t v : in->singletons) { r[i] = v; w[i] = 1; i++; }
            std::copy(repr.begin() + objs[k].firstCluster, repr.begin() + objs[k].firstCluster + objs[k].numClusters, r.begin() + i);
            std::copy(wt.begin() + objs[k].firstCluster, wt.begin() + objs[k].firstCluster + objs[k].numClusters, w.begin() + i);
            in->smp->setCounterPos(objs[k].rngPos + objs[k].numClusters);
        }
    }
    void hostSample(Inst *in, std::vector<std::vector<uint32_t>> &selected, std::vector<std::vector<float>> &weights) {
        ensureHostCw();
        in->listsStale = true;
        syncLists({in});
        in->sampleRepresentatives(selected[in->id], weights[in->id]);
    }
    /* Device-resident refinement (refine.cuh): every object of `which` that is not done and that the kernel supports runs to
     * completion in one launch; objects it hands back (RF_RESUME_HOST: queue larger than the shared-memory heap) and the ones
     * it does not take (more than RF_MAXROWS rows, sequential sample streams) continue in the host-driven rounds below. */
    void refineDevice(const std::vector<Inst *> &which) {
        std::vector<Inst *> dev;
        std::vector<RfInst> hi;
        std::vector<ClusterNode> initNodes; std::vector<uint32_t> initSingles;
        uint64_t xFloats = 0;
        size_t freeB = 0, totalB = 0;
        ALVRL_CUDA(cudaMemGetInfo(&freeB, &totalB));
        for (Inst *in : which) {
            uint32_t key = 0, pos = 0;
            if (in->done || in->nr > RF_MAXROWS_BIG || in->nr == 0 || !in->smp->counterState(key, pos)) continue;
            if (2 * (xFloats + (uint64_t) N * (in->nr + 4)) * sizeof(float) + (dev.size() + 1) * (uint64_t) N * 160 > freeB / 2) continue;   /* the two compacted copies would not fit */
            RfInst r; memset(&r, 0, sizeof(r));
            r.r0 = in->r0; r.nr = in->nr; r.lw = in->lw; r.listOff = in->listOff; r.cwOff = in->cwOff;
            r.nrP = (in->nr + 3u) & ~3u; r.xOff = xFloats; r.vOff = (uint64_t) dev.size() * N;
            xFloats += (uint64_t) N * r.nrP;
            r.numVrlsTotal = in->numVrlsTotal; r.pixelUndersampling = in->pixelUndersampling; r.tracingVar = in->tracingVar; r.unclIntVar = in->unclIntVar;
            r.adaptive = in->adaptive ? 1u : 0u; r.targetClusters = in->targetClusters; r.rngKey = key; r.rngPos = pos;
            r.splitsBase = in->splitsDone; r.bestSplits = in->bestSplits; r.fixedSplits = in->fixedSplits;
            r.underVar = in->underVar; r.intVar = in->intVar; r.bestConstant = in->bestConstant;
            r.heapCount = r.nodeCount = (uint32_t) in->pq.size(); r.singleCount = (uint32_t) in->singletons.size();
            r.initNodeOff = (uint32_t) initNodes.size(); r.initSingleOff = (uint32_t) initSingles.size();
            initNodes.insert(initNodes.end(), in->pq.begin(), in->pq.end());
            for (auto it = in->singletons.rbegin(); it != in->singletons.rend(); ++it) initSingles.push_back(*it);   /* insertion order */
            dev.push_back(in); hi.push_back(r);
        }
        if (dev.empty()) return;
        double p0 = Prof::now(), pr0 = p0;
        auto rlap = [&](const char *what) { if (prof.on) { cudaStreamSynchronize(st); const double n_ = Prof::now(); fprintf(stderr, "[alvrl clustering]   refineDevice %s %.1f ms\n", what, n_ - pr0); pr0 = n_; } };
        int devId = 0, sms = 0;
        ALVRL_CUDA(cudaGetDevice(&devId));
        ALVRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devId));
        uint64_t keyStride = 2; while (keyStride < N) keyStride <<= 1;
        /* a queue holds at most N / 2 multi-clusters, a split tree at most N - 1 nodes on top of the initial ones, N singletons */
        size_t initMax = 0; for (Inst *in : dev) initMax = std::max(initMax, in->pq.size());
        const uint32_t heapCap = std::max<uint32_t>(RF_HEAP_CAP + 8, N / 2 + (uint32_t) initMax + 8), nodeCap = N + (uint32_t) initMax + 8;
        /* many CTAs per object (k_refine_mt) unless asked for the one-CTA-per-object kernel or the task encoding does not fit */
        const bool mt = !getenv("ALVRL_REFINE_ST") && dev.size() < 0x8000u && nodeCap < 0x1000000u;
        const uint32_t grid = mt ? (uint32_t) std::max(1, sms) : (uint32_t) std::min<size_t>(dev.size(), (size_t) std::max(1, sms));
        const size_t pool = mt ? dev.size() : (size_t) grid;                   /* node tables and queues: per object / per CTA */
        DevBuf<HeapEntry> dHeapOv;
        DevBuf<RfInst> dInst; DevBuf<ClusterNode> dInitNodes, dOutNodes, dNodes; DevBuf<uint32_t> dInitSingles, dOutSingles, dSingles, dCursors;
        DevBuf<HeapEntry> dSnap; DevBuf<unsigned long long> dKeysG; DevBuf<double> dWG, dVcol; DevBuf<float2> dPairsG; DevBuf<float> dX, dX2; DevBuf<uint32_t> dSrcPos;
        initNodes.push_back(ClusterNode{0, 0, 0, 0}); initSingles.push_back(0);                  /* never empty */
        dInst.upload(hi, st); dInitNodes.upload(initNodes, st); dInitSingles.upload(initSingles, st);
        dOutNodes.alloc(dev.size() * (size_t) (2 * heapCap)); dOutSingles.alloc(dev.size() * (size_t) nodeCap);
        DevBuf<MtNode> dMtNodes; DevBuf<HeapEntry> dMtHeap; DevBuf<uint32_t> dOutstanding, dLists1, dCtr; DevBuf<unsigned long long> dSlots, dMtClk;
        DevBuf<uint32_t> dCtl, dWaitNode;
        if (mt) { dMtNodes.alloc(pool * nodeCap); dMtHeap.alloc(pool * heapCap); dOutstanding.alloc(pool); dLists1.alloc(dLists.n); dCtr.alloc(4);
                  dCtl.alloc(pool); dWaitNode.alloc(pool); }
        else { dNodes.alloc((size_t) grid * nodeCap); dHeapOv.alloc((size_t) grid * (heapCap - RF_HEAP_CAP)); }
        dSingles.alloc(pool * nodeCap); dSnap.alloc(pool * heapCap);
        dKeysG.alloc((size_t) grid * keyStride); dWG.alloc((size_t) grid * 3 * N); dPairsG.alloc((size_t) grid * 2 * N);
        dCursors.alloc(4);
        dX.alloc(xFloats); dX2.alloc(xFloats); dVcol.alloc(dev.size() * (size_t) N); dSrcPos.alloc((size_t) grid * 2 * N);
        rlap("alloc+upload");
        k_rf_compact<<<dim3((N + 7) / 8, (uint32_t) dev.size()), 256, 0, st>>>(R, ldR, N, dInst.p, dLists.p, dCw.p, dX.p, dVcol.p, rowW);
        rlap("compact");
        ALVRL_CUDA(cudaMemsetAsync(dCursors.p, 0, 4 * sizeof(uint32_t), st));
        RfScratch scr;
        scr.keys = dKeysG.p; scr.w = dWG.p; scr.Wf = dWG.p + (size_t) grid * N; scr.Wr = dWG.p + (size_t) grid * 2 * N; scr.pairs = dPairsG.p;
        scr.keyStride = keyStride; scr.stepStride = N; scr.unfoldRows = getenv("ALVRL_RF_UNFOLD") ? 1u : 0u;
        /* gangs: clusters of >= gangMin columns are split by several CTAs together (refine_split.inl); ALVRL_GANG_MIN=0 turns them off */
        scr.gangMin = 8192u; scr.gangMax = RF_GANG_MAX;
        if (const char *e = getenv("ALVRL_GANG_MIN")) scr.gangMin = (uint32_t) std::max(0, atoi(e));
        if (const char *e = getenv("ALVRL_GANG_MAX")) scr.gangMax = (uint32_t) std::min(RF_GANG_MAX, std::max(1, atoi(e)));
        /* a gang trades CTA time (members wait for each other) for latency: worth it while there are fewer large clusters than
         * CTAs, i.e. when a GPU holds few objects (one rank of a multi-GPU job).  With `load` objects per CTA a root cluster of N
         * columns gets about 1 / load CTAs: gangMin = load x N (ALVRL_GANG_LOAD scales it) */
        if (scr.gangMin && !getenv("ALVRL_GANG_MIN")) {
            double scale = 1.0;
            if (const char *e = getenv("ALVRL_GANG_LOAD")) scale = atof(e);
            const double perCta = (double) dev.size() / (double) std::max(1, sms);
            scr.gangMin = std::max<uint32_t>(scr.gangMin, (uint32_t) std::min(4e9, scale * perCta * (double) N));
            if (perCta * scale >= 0.5) scr.gangMin = 0;        /* the GPU is full of objects: throughput, not latency, bounds the kernel */
        }
        if (scr.gangMin && scr.gangMin < 2u * RF_SMALL) scr.gangMin = 2u * RF_SMALL;          /* gangs use the large-cluster code path */
        DevBuf<double> dCarry; dCarry.alloc((size_t) grid * 2 * RF_GANG_MAX * RF_MAXROWS);
        scr.carry = dCarry.p;
        scr.rowW = rowW;
        scr.snapHeap = dSnap.p; scr.nodes = dNodes.p; scr.singles = dSingles.p; scr.heapOv = dHeapOv.p; scr.heapCap = heapCap; scr.nodeCap = nodeCap;
        scr.srcPos = dSrcPos.p; scr.posTmp = dSrcPos.p + (size_t) grid * N;
        scr.initNodes = dInitNodes.p; scr.initSingles = dInitSingles.p; scr.outNodes = dOutNodes.p; scr.outSingles = dOutSingles.p; scr.cursors = dCursors.p;
        if (mt) {
            /* the ticket ring: at most MT_K split tasks or one control task per object are in flight */
            const size_t gangTickets = scr.gangMin ? (size_t) (N / scr.gangMin + 1) * RF_GANG_MAX : 0;   /* per object, on top of MT_K */
            /* splits of one object kept in flight: 32 when the GPU is full of objects; a rank of a multi-GPU job holds few, and
             * more speculation keeps its CTAs busy (the clusters with the largest keys are the ones the queue pops next) */
            uint32_t inflight = 32;
            { const double load = (double) dev.size() / std::max(1, sms); if (load < 0.25) inflight = 128; else if (load < 0.5) inflight = 64; }
            if (const char *e = getenv("ALVRL_MT_K")) inflight = (uint32_t) std::min(MT_K_MAX, std::max(1, atoi(e)));
            uint32_t qcap = 65536; while (qcap < 4 * (inflight + 1 + gangTickets) * dev.size()) qcap <<= 1;
            dSlots.alloc(qcap); dMtClk.alloc((size_t) grid * 32);
            ALVRL_CUDA(cudaMemsetAsync(dSlots.p, 0, qcap * sizeof(unsigned long long), st));
            ALVRL_CUDA(cudaMemsetAsync(dOutstanding.p, 0, pool * sizeof(uint32_t), st));
            ALVRL_CUDA(cudaMemsetAsync(dWaitNode.p, 0xff, pool * sizeof(uint32_t), st));            /* MT_NONE */
            dCtl.upload(std::vector<uint32_t>(pool, 1u), st);                                        /* the first control passes are in the ring */
            ALVRL_CUDA(cudaMemsetAsync(dMtClk.p, 0, (size_t) grid * 32 * sizeof(unsigned long long), st));
            std::vector<unsigned long long> first(dev.size());                  /* one control task per object to start with */
            for (size_t i = 0; i < dev.size(); i++) first[i] = (1ull << 40) | (1ull << 39) | ((unsigned long long) i << 24);
            dSlots.upload(first, st);
            const std::vector<uint32_t> ctr = {0u, (uint32_t) dev.size(), (uint32_t) dev.size(), 0u};
            dCtr.upload(ctr, st);
            MtPools mp;
            mp.nodes = dMtNodes.p; mp.heap = dMtHeap.p; mp.snap = dSnap.p; mp.singles = dSingles.p; mp.outstanding = dOutstanding.p;
            mp.ctl = dCtl.p; mp.waitNode = dWaitNode.p;
            mp.inflight = inflight;
            mp.slots = dSlots.p; mp.qmask = qcap - 1; mp.ctr = dCtr.p; mp.nodeCap = nodeCap; mp.heapCap = heapCap; mp.clk = dMtClk.p;
            ALVRL_CUDA(cudaFuncSetAttribute(k_refine_mt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(RfShared)));
            k_refine_mt<<<grid, RF_THREADS, sizeof(RfShared), st>>>(dX.p, dX2.p, dVcol.p, dInst.p, (uint32_t) dev.size(), dLists.p, dLists1.p, dCw.p, scr, mp);
            ALVRL_CUDA(cudaGetLastError());
            if (prof.on) {
                std::vector<unsigned long long> ck((size_t) grid * 32);
                dMtClk.download(ck.data(), ck.size(), st);
                unsigned long long t[32] = {0};
                for (size_t i = 0; i < ck.size(); i++) t[i % 32] += ck[i];
                fprintf(stderr, "[alvrl clustering]   k_refine_mt on %u CTAs: control passes %llu (%.1f Mcycles), split tasks %llu, waiting for tickets %.1f Mcycles\n",
                        grid, t[26], t[24] * 1e-6, t[27], t[25] * 1e-6);
                fprintf(stderr, "[alvrl clustering]   gangs (>= %u columns, up to %u CTAs): %llu syncs, %.1f Mcycles waiting in them; sweeps: carries %.1f, main pass %.1f Mcycles\n", scr.gangMin, scr.gangMax, t[29], t[28] * 1e-6, t[30] * 1e-6, t[31] * 1e-6);
                for (int k = 0; k < 2; k++)
                    fprintf(stderr, "[alvrl clustering]   %s split tasks %llu, Mcycles summed over CTAs: pick %.1f direction %.1f stage %.1f project %.1f sort %.1f weights %.1f sweep %.1f pairs %.1f result %.1f\n",
                            k ? "large" : "small", t[12 * k + 9], t[12 * k + 0] * 1e-6, t[12 * k + 1] * 1e-6, t[12 * k + 2] * 1e-6, t[12 * k + 3] * 1e-6, t[12 * k + 4] * 1e-6,
                            t[12 * k + 5] * 1e-6, t[12 * k + 6] * 1e-6, t[12 * k + 7] * 1e-6, t[12 * k + 8] * 1e-6);
            }
        } else {
            ALVRL_CUDA(cudaFuncSetAttribute(k_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(RfShared)));
            k_refine<<<grid, RF_THREADS, sizeof(RfShared), st>>>(dX.p, dX2.p, dVcol.p, dInst.p, (uint32_t) dev.size(), dLists.p, dCw.p, scr);
        }
        launches(2);
        ALVRL_CUDA(cudaGetLastError());
        rlap("k_refine");
        dInst.download(hi.data(), hi.size(), st);
        uint32_t cursors[4];
        dCursors.download(cursors, 4, st);
        std::vector<ClusterNode> outNodes(std::max<uint32_t>(cursors[1], 1)); std::vector<uint32_t> outSingles(std::max<uint32_t>(cursors[2], 1));
        if (cursors[1]) dOutNodes.download(outNodes.data(), cursors[1], st);
        if (cursors[2]) dOutSingles.download(outSingles.data(), cursors[2], st);
        uint64_t splits = 0, degenerate = 0, resumed = 0;
        unsigned long long clk[2][12] = {{0}};
        std::vector<std::pair<unsigned long long, size_t>> perObj;
        for (size_t i = 0; i < dev.size(); i++) {
            Inst &in = *dev[i];
            const RfInst &r = hi[i];
            switch (r.status) {
                case RF_DONE: case RF_RESUME_HOST: break;
                case RF_ERR_CONSTANT: throw Error(ALVRL_ERR_ARG, "invalid convergence constant");
                case RF_ERR_LOWER: throw Error(ALVRL_ERR_ARG, "invalid lower bound on convergence constant");
                case RF_ERR_SPLIT: throw Error(ALVRL_ERR_ARG, "couldn't split cluster!");
                case RF_ERR_NOBEST: throw Error(ALVRL_ERR_ARG, "Couldn't find best splitting index!");
                case RF_ERR_SINGLETON_VAR: throw Error(ALVRL_ERR_ARG, "Trying to add singleton cluster with non-zero undersampling variance");
                default: throw Error(ALVRL_ERR_ARG, "non-positive column weight sum in a cluster split");
            }
            const ClusterNode *cur = outNodes.data() + r.outNodeOff, *snap = cur + r.heapCount;
            const uint32_t *sg = outSingles.data() + r.outSingleOff;
            in.pq.assign(cur, cur + r.heapCount); in.s_pq.assign(snap, snap + r.sHeapCount);
            in.singletons.clear(); in.s_single.clear();
            for (uint32_t k = 0; k < r.singleCount; k++) in.singletons.push_front(sg[k]);
            for (uint32_t k = 0; k < r.sSingleCount; k++) in.s_single.push_front(sg[k]);
            in.underVar = r.underVar; in.intVar = r.intVar; in.s_under = r.sUnder; in.s_int = r.sInt; in.bestConstant = r.bestConstant;
            in.smp->setCounterPos(r.rngPos);
            in.nearTies += r.nearTies; in.listsStale = true;
            in.splitsDone += r.splits; in.bestSplits = r.bestSplits;
            splits += r.splits; degenerate += r.degenerate;
            { unsigned long long tot = 0; for (int a = 0; a < 24; a++) { clk[a / 12][a % 12] += r.clk[a / 12][a % 12]; if (a % 12 != 9) tot += r.clk[a / 12][a % 12]; } perObj.push_back(std::make_pair(tot, i)); }
            if (r.status == RF_DONE) { if (in.adaptive) in.restore(); in.done = true; }
            else resumed++;
        }
        rlap("readback");
        prof.tasks += splits;
        if (prof.on) fprintf(stderr, "[alvrl clustering] device refinement: %zu objects on %u CTAs, %llu splits (%llu random directions, %llu handed back) in %.1f ms\n",
                             dev.size(), grid, (unsigned long long) splits, (unsigned long long) degenerate, (unsigned long long) resumed, Prof::now() - p0);
        if (prof.on) {
            uint32_t nrMax = 0; size_t skippedRows = 0;
            for (Inst *in : which) { nrMax = std::max(nrMax, in->nr); if (in->nr > RF_MAXROWS_BIG) skippedRows++; }
            fprintf(stderr, "[alvrl clustering]   objects %zu, %zu with more than %d rows (max %u)\n", which.size(), skippedRows, RF_MAXROWS_BIG, nrMax);
            for (int k = 0; k < 2; k++)
                fprintf(stderr, "[alvrl clustering]   %s splits %llu, Mcycles summed over CTAs: pick %.1f direction %.1f stage %.1f project %.1f sort %.1f weights %.1f sweep %.1f pairs %.1f argmin+queue %.1f\n",
                        k ? "large" : "small", clk[k][9], clk[k][0] * 1e-6, clk[k][1] * 1e-6, clk[k][2] * 1e-6, clk[k][3] * 1e-6, clk[k][4] * 1e-6, clk[k][5] * 1e-6, clk[k][6] * 1e-6,
                        clk[k][7] * 1e-6, clk[k][8] * 1e-6);
            fprintf(stderr, "[alvrl clustering]   columns visited (sum of n over splits, all objects): small %.3e (%.3e not tile-resident) | large %.3e\n", (double) clk[0][10], (double) clk[0][11], (double) clk[1][10]);
            std::sort(perObj.rbegin(), perObj.rend());
            for (size_t k = 0; k < std::min<size_t>(4, perObj.size()); k++) {
                const RfInst &r = hi[perObj[k].second];
                fprintf(stderr, "[alvrl clustering]   slowest #%zu: %.1f Mcycles, nr %u, splits %u (large %llu): stage %.1f project %.1f sort %.1f sweep %.1f | large: stage %.1f project %.1f sort %.1f sweep %.1f pick %.1f\n",
                        k, perObj[k].first * 1e-6, r.nr, r.splits, r.clk[1][9], r.clk[0][2] * 1e-6, r.clk[0][3] * 1e-6, r.clk[0][4] * 1e-6, r.clk[0][6] * 1e-6,
                        r.clk[1][2] * 1e-6, r.clk[1][3] * 1e-6, r.clk[1][4] * 1e-6, r.clk[1][6] * 1e-6, r.clk[1][0] * 1e-6);
            }
            fprintf(stderr, "[alvrl clustering]   mean per object %.1f Mcycles\n", (double) std::accumulate(perObj.begin(), perObj.end(), 0ull, [](unsigned long long a, const std::pair<unsigned long long, size_t> &b) { return a + b.first; }) * 1e-6 / perObj.size());
        }
    }
    /* Clustering::split (590-684) for one cluster of every instance in `round` */
    void splitRound(const std::vector<Inst *> &round, const std::vector<std::pair<uint32_t, uint32_t>> *centres = nullptr) {
        ensureHostCw();
        const size_t T = round.size();
        std::vector<ClTask> tasks(T);
        uint64_t out = 0, dirOff = 0;
        double p0 = Prof::now(), p1;
        prof.rounds++; prof.tasks += T;
        for (size_t i = 0; i < T; i++) {
            Inst &in = *round[i];
            if (!centres) in.cur = in.popMulti();
            const uint32_t begin = in.cur.begin, end = in.cur.end;
            if (end - begin < 2) throw Error(ALVRL_ERR_ARG, "couldn't split cluster!");
            if (centres) { in.vrl1 = (*centres)[i].first; in.vrl2 = (*centres)[i].second; }
            else {
                in.smp->enterNode(begin, end);                                      /* counter stream: the cluster's own draws (alvrl_rng.h) */
                in.vrl1 = in.vrls[weighted_sample(in.cw, in.smp, nullptr, begin, end, in.vrls)];   /* 597-602 */
                const float weight1 = in.cw[in.vrl1];
                in.cw[in.vrl1] = 0.0f;
                in.vrl2 = in.vrls[weighted_sample(in.cw, in.smp, nullptr, begin, end, in.vrls)];
                in.cw[in.vrl1] = weight1;
            }
            ClTask t = baseTask(in);
            t.begin = begin; t.end = end; t.vrl1 = in.vrl1; t.vrl2 = in.vrl2; t.outOff = out; t.dirOff = dirOff;
            out += end - begin; dirOff += in.nr;
            tasks[i] = t;
        }
        ensure(dDir, dirOff); ensure(dProj, out); ensure(dFlags, T); ensure(dStaged, out);
        prof.steps += out;
        dTasks.upload(tasks, st);
        p1 = Prof::now(); prof.t[0] += p1 - p0; p0 = p1;
        uint32_t maxN = 0; for (const ClTask &t : tasks) maxN = std::max(maxN, t.end - t.begin);
        k_direction<<<(uint32_t) T, 128, 0, st>>>(R, ldR, dTasks.p, dDir.p, dFlags.p);
        k_project<<<dim3((maxN + CL_PROJ_WARPS - 1) / CL_PROJ_WARPS, (uint32_t) T), CL_PROJ_WARPS * 32, 0, st>>>(R, ldR, dTasks.p, dLists.p, dDir.p, dProj.p, nullptr);
        launches(2);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<uint32_t> flags(T);
        dFlags.download(flags.data(), T, st);
        p1 = Prof::now(); prof.t[1] += p1 - p0; p0 = p1;
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
                    d[r] = (float) std::cos((double) phi) * rr;          /* pinned transcendental: double, then rounded (as refine.cuh and the oracle) */
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
            k_project<<<dim3((maxN + CL_PROJ_WARPS - 1) / CL_PROJ_WARPS, (uint32_t) T), CL_PROJ_WARPS * 32, 0, st>>>(R, ldR, dTasks.p, dLists.p, dDir.p, dProj.p, nullptr);
            launches(1);
            ALVRL_CUDA(cudaGetLastError());
        }
        std::vector<float> proj(out);
        dProj.download(proj.data(), out, st);
        p1 = Prof::now(); prof.t[2] += p1 - p0; p0 = p1;
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
            if (T == 1 || hw == 1 || serialSort || out < 4096) for (size_t i = 0; i < T; i++) sortOne(i);
            else {
                std::vector<std::thread> th;
                for (unsigned w = 0; w < hw; w++) th.emplace_back([&, w]() { for (size_t i = w; i < T; i += hw) sortOne(i); });
                for (auto &x : th) x.join();
            }
        }
        p1 = Prof::now(); prof.t[3] += p1 - p0; p0 = p1;
        ALVRL_CUDA(cudaMemcpyAsync(dStaged.p, staged.data(), out * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        k_scatter_lists<<<dim3((maxN + 127) / 128, (uint32_t) T), 128, 0, st>>>(dTasks.p, dStaged.p, dLists.p);
        /* forward and reverse prefix variances (648-657) */
        std::vector<ClTask> vt(2 * T);
        for (size_t i = 0; i < T; i++)
            for (int rev = 0; rev < 2; rev++) { ClTask t = tasks[i]; t.reverse = rev; t.finalOnly = 0; vt[2 * i + rev] = t; }
        const uint64_t steps = runVariance(vt, dTasks2);
        ALVRL_CUDA(cudaStreamSynchronize(st));
        p1 = Prof::now(); prof.t[4] += p1 - p0; p0 = p1;
        std::vector<float2> pairs(steps);
        dPairs.download(pairs.data(), steps, st);
        p1 = Prof::now(); prof.t[5] += p1 - p0; p0 = p1;
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
        for (Inst *in : round) in->smp->leaveNode();
        p1 = Prof::now(); prof.t[6] += p1 - p0;
    }
    DevBuf<ClTask> dTasks2;
    cudaEvent_t ev[8] = {nullptr}; double evMs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void evRec(int i) { if (prof.on) { if (!ev[i]) cudaEventCreate(&ev[i]); cudaEventRecord(ev[i], st); } }
    DevBuf<float> dWtmp; DevBuf<uint64_t> dKeys, dKeysOut; DevBuf<RoundResult> dRes; DevBuf<int> dSegOff; DevBuf<unsigned char> dSortTemp;

    /* Clustering::split for one cluster of every instance in `round`, device rounds: the host only pops the heap nodes, draws
     * the two uniforms per split and pushes the two halves; centres, sort and argmin never leave the device. */
    void splitRoundDevice(const std::vector<Inst *> &round) {
        const size_t T = round.size();
        std::vector<ClTask> tasks(T);
        std::vector<int> segOff(2 * T);
        uint64_t out = 0, dirOff = 0;
        double p0 = Prof::now(), p1;
        prof.rounds++; prof.tasks += T;
        uint32_t maxN = 0;
        for (size_t i = 0; i < T; i++) {
            Inst &in = *round[i];
            in.cur = in.popMulti();
            const uint32_t begin = in.cur.begin, end = in.cur.end;
            if (end - begin < 2) throw Error(ALVRL_ERR_ARG, "couldn't split cluster!");
            ClTask t = baseTask(in);
            t.begin = begin; t.end = end; t.outOff = out; t.dirOff = dirOff;
            in.smp->enterNode(begin, end);                                      /* counter stream: the cluster's own draws (alvrl_rng.h) */
            t.u1 = in.smp->next1D(); t.u2 = in.smp->next1D();                   /* the two weightedSample draws, 597-602 */
            segOff[i] = (int) out; segOff[T + i] = (int) (out + (end - begin));
            out += end - begin; dirOff += in.nr;
            maxN = std::max(maxN, end - begin);
            tasks[i] = t;
        }
        if (out > 0x7fffffffull) throw Error(ALVRL_ERR_UNSUPPORTED, "clustering round larger than 2^31 entries");
        ensure(dDir, dirOff); ensure(dKeys, out); ensure(dKeysOut, out); ensure(dRes, T);
        prof.steps += out;
        dTasks.upload(tasks, st);
        dSegOff.upload(segOff, st);
        std::vector<ClTask> vt(2 * T);
        for (size_t i = 0; i < T; i++)
            for (int rev = 0; rev < 2; rev++) { ClTask t = tasks[i]; t.reverse = rev; t.finalOnly = 0; vt[2 * i + rev] = t; }
        const VarPlan plan = prepareVariance(vt, dTasks2);                  /* everything the host knows goes up before the first launch */
        size_t tempBytes = 0;
        cub::DeviceSegmentedSort::SortKeys(nullptr, tempBytes, dKeys.p, dKeysOut.p, (int) out, (int) T, dSegOff.p, dSegOff.p + T, st);
        ensure(dSortTemp, tempBytes + 16);
        p1 = Prof::now(); prof.t[0] += p1 - p0; p0 = p1;
        const dim3 gN((maxN + 127) / 128, (uint32_t) T);
        evRec(0);
        k_pick_direction<<<(uint32_t) T, 128, 0, st>>>(R, ldR, dTasks.p, dLists.p, dCw.p, dRes.p, dDir.p);
        evRec(1);
        k_project<<<dim3((maxN + CL_PROJ_WARPS - 1) / CL_PROJ_WARPS, (uint32_t) T), CL_PROJ_WARPS * 32, 0, st>>>(R, ldR, dTasks.p, dLists.p, dDir.p, nullptr, dKeys.p);
        evRec(2);
        cub::DeviceSegmentedSort::SortKeys(dSortTemp.p, tempBytes, dKeys.p, dKeysOut.p, (int) out, (int) T, dSegOff.p, dSegOff.p + T, st);
        evRec(3);
        k_scatter_sorted<<<gN, 128, 0, st>>>(dTasks.p, dKeysOut.p, dLists.p);
        launches(6);
        ALVRL_CUDA(cudaGetLastError());
        p1 = Prof::now(); prof.t[1] += p1 - p0; p0 = p1;
        evRec(4);
        launchVariance(plan, dTasks2);
        evRec(5);
        k_argmin<<<(uint32_t) T, CL_THREADS, 0, st>>>(dTasks2.p, dPairs.p, dLists.p, dRes.p);
        evRec(6);
        launches(1);
        ALVRL_CUDA(cudaGetLastError());
        std::vector<RoundResult> res(T);
        dRes.download(res.data(), T, st);
        if (prof.on) for (int i = 0; i < 6; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); evMs[i] += ms; }
        p1 = Prof::now(); prof.t[4] += p1 - p0; p0 = p1;
        std::vector<Inst *> hostRound; std::vector<std::pair<uint32_t, uint32_t>> hostCentres;
        for (size_t i = 0; i < T; i++) {
            Inst &in = *round[i];
            const RoundResult &r = res[i];
            const uint32_t begin = tasks[i].begin, end = tasks[i].end;
            if (r.flags) {                                                      /* degenerate centres: random direction on the host (616-622) */
                std::vector<uint32_t> seg(end - begin);
                dLists.download(seg.data(), end - begin, st, in.listOff + begin);
                std::copy(seg.begin(), seg.end(), in.vrls.begin() + begin);
                hostRound.push_back(&in); hostCentres.push_back(std::make_pair(r.vrl1, r.vrl2));
                continue;
            }
            if (r.bestIndex == 0xffffffffu) throw Error(ALVRL_ERR_ARG, "Couldn't find best splitting index!");
            if (std::isfinite(r.second) && std::fabs(r.second - r.best) <= 1e-6f * std::fabs(r.best)) in.nearTies++;
            const uint32_t splitIndex = begin + r.bestIndex;
            in.listsStale = true;
            in.vrls[begin] = r.firstVrl; in.vrls[end - 1] = r.lastVrl;          /* singleton halves record their VRL id (561) */
            in.addCluster(begin, splitIndex, r.head.x, r.head.y);
            in.addCluster(splitIndex, end, r.tail.x, r.tail.y);
        }
        p1 = Prof::now(); prof.t[6] += p1 - p0;
        if (!hostRound.empty()) splitRound(hostRound, &hostCentres);
        for (Inst *in : round) in->smp->leaveNode();
    }
    /* bring the host mirrors of the VRL permutations up to date (device rounds reorder them on the device only) */
    void syncLists(const std::vector<Inst *> &which) {
        for (Inst *in : which) {
            if (!in->listsStale) continue;
            dLists.download(in->vrls.data(), in->vrls.size(), st, in->listOff);
            in->listsStale = false;
        }
    }
};

} // namespace

/* the same flags left in a caller-owned device buffer (N bytes), for the all-reduce of the multi-GPU path */
void column_nonzero_into(alvrl_ctx *c, uint8_t *dFlags) {
    const uint32_t N = (uint32_t) c->vrlHost.size(), S = c->numSlices();
    const uint32_t r0 = c->rowOffset[std::min(c->sliceBegin, S)], r1 = c->rowOffset[std::min(c->sliceEnd, S)];
    k_total_contribution<<<(N + 7) / 8, 256, 0, c->stream>>>(c->dR.p, c->ldR, r0, r1, N, dFlags);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
}
void column_nonzero_device(alvrl_ctx *c, std::vector<uint8_t> &flags) {
    const uint32_t N = (uint32_t) c->vrlHost.size(), S = c->numSlices();
    const uint32_t r0 = c->rowOffset[std::min(c->sliceBegin, S)], r1 = c->rowOffset[std::min(c->sliceEnd, S)];
    DevBuf<uint8_t> dNz; dNz.alloc(N);
    k_total_contribution<<<(N + 7) / 8, 256, 0, c->stream>>>(c->dR.p, c->ldR, r0, r1, N, dNz.p);
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

/* getLocalMatrix with neighbourWeight > 0 (779-827): the rows of slice i followed by the rows of its neighbour slices are
 * gathered into one contiguous block of a second matrix (the reference copies them too), so that every clustering kernel
 * keeps addressing "rows [r0, r0 + nr) of every column"; warp = column, lanes = rows */
__global__ void k_gather_local_rows(const float2 *__restrict__ R, uint32_t ldR, const uint32_t *__restrict__ rowMap, uint32_t rows, uint32_t N,
                                    float2 *__restrict__ out, uint32_t ldOut) {
    const uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= N) return;
    const float2 *col = R + (size_t) v * ldR;
    float2 *dst = out + (size_t) v * ldOut;
    for (uint32_t j = lane; j < rows; j += 32) dst[j] = col[rowMap[j]];
}

/* Preprocessor::buildLocalities (1241-1293): the neighbourCount slices closest in the 6-D (position centroid, direction
 * centroid) distance; quirk B7 kept (maxInd is not reset per slice; entries never filled keep index 0).  std::set order:
 * by slice index, then distance. */
static void build_localities(const alvrl_ctx *c, std::vector<std::set<std::pair<uint32_t, float>>> &loc) {
    const uint32_t S = c->numSlices(), nc = (uint32_t) std::max(0, c->P.neighbourCount);
    loc.assign(S, {});
    auto dist = [&](uint32_t i, uint32_t j) {                                   /* sliceDistance, 1230-1239 */
        const float *a = &c->sliceCentroid[6 * (size_t) i], *b = &c->sliceCentroid[6 * (size_t) j];
        float dp, dd;
        { const float x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2]; dp = x * x + y * y + z * z; }
        { const float x = a[3] - b[3], y = a[4] - b[4], z = a[5] - b[5]; dd = x * x + y * y + z * z; }
        return std::sqrt(dp + dd);
    };
    if (S <= nc) {
        for (uint32_t i = 0; i < S; i++) for (uint32_t j = 0; j < S; j++) if (i != j) loc[i].insert(std::make_pair(j, dist(i, j)));
        return;
    }
    if (nc == 0) return;
    std::vector<float> distances(nc, std::numeric_limits<float>::infinity()); std::vector<uint32_t> indices(nc, 0);
    uint32_t maxInd = 0;
    for (uint32_t i = 0; i < S; i++) {
        for (uint32_t j = 0; j < nc; j++) distances[j] = std::numeric_limits<float>::infinity();
        for (uint32_t j = 0; j < S; j++) {
            if (i == j) continue;
            const float d = dist(i, j);
            if (d < distances[maxInd]) {
                distances[maxInd] = d; indices[maxInd] = j;
                for (uint32_t k = 0; k < nc; k++) if (distances[k] > distances[maxInd]) maxInd = k;
            }
        }
        for (uint32_t x = 0; x < nc; x++) loc[i].insert(std::make_pair(indices[x], distances[x]));
    }
}

void build_clusters_device(alvrl_ctx *c, bool needFallback) {
    const uint32_t N = (uint32_t) c->vrlHost.size(), G = (uint32_t) c->rowPixel.size(), S = c->numSlices();
    if (c->globalPixelUndersampling < 0) throw Error(ALVRL_ERR_STATE, "Invalid pixel undersampling. Did you forget to call buildSlices first?");
    cudaStream_t st = c->stream;
    const bool sfmt = c->P.rngMode == ALVRL_RNG_MODE_SFMT;
    const bool lazyFallbackCall = c->haveClusters && !c->haveFallback;       /* called again only for the global / fallback lists */

    Workspace ws; ws.c = c; ws.st = st; ws.N = N; ws.ldR = c->ldR; ws.R = c->dR.p; ws.deviceRounds = !sfmt;

    /* the global stream: cluster() -> global representatives -> fallback refinement, in this order (159-179) */
    HostSampler *globalSmp = c->mainSampler.get();
    if (!sfmt) {
        if (!lazyFallbackCall || !c->globalStream) { c->globalStream.reset(new CounterStream(c->P.seed)); c->globalStream->setContext(ALVRL_RNG_CLUSTER, ALVRL_RNG_GLOBAL_ID, 0); }
        globalSmp = c->globalStream.get();
    }

    /* the global Clustering object spans ALL rows of R (1/G weights, Preprocessor.cpp:159-165): a handle that owns a slice
     * range holds zeros in the other ranks' rows, so it must not build it (the group entry points gather what they need) */
    const bool ranged = !(std::min(c->sliceBegin, S) == 0 && std::min(c->sliceEnd, S) == S);
    auto makeGlobalInst = [&](Inst &g) {
        if (ranged) throw Error(ALVRL_ERR_UNSUPPORTED, "global / fallback clustering needs all rows of R, but this handle owns only a slice range "
                                                        "(globalCluster, a slice with zero unclustered variance, or alvrl_get_clusters on a sharded handle)");
        g.id = ALVRL_RNG_GLOBAL_ID; g.r0 = 0; g.nr = G; g.rowBlocks = (G + CL_THREADS - 1) / CL_THREADS; g.lw = 1.0 / G;
        g.pixelUndersampling = c->globalPixelUndersampling; g.smp = globalSmp; g.group = -1;
    };

    const double tb0 = Prof::now();
    if (!lazyFallbackCall) {
        /* cluster(), 838-898: zero / non-zero columns over all rows */
        std::vector<uint8_t> nz;
        if (c->columnFlagsOverride.size() == N) nz = c->columnFlagsOverride;
        else column_nonzero_device(c, nz);
        if (ws.prof.on) fprintf(stderr, "[alvrl clustering] column_nonzero %.1f ms\n", Prof::now() - tb0);
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
                c->stats.kernelLaunches += ws.launchCount; ws.launchCount = 0;
            } else c->globalVrlsPerCluster.assign(1, nonZero);
        }
        if (!zero.empty()) c->globalVrlsPerCluster.push_back(zero);
    }

    auto computeFallback = [&]() {
        Inst g; makeGlobalInst(g);
        Workspace w2; w2.c = c; w2.st = st; w2.N = N; w2.ldR = c->ldR; w2.R = c->dR.p; w2.deviceRounds = !sfmt;
        w2.insts = {&g}; w2.allocInstances(); w2.columnWeights();
        w2.construct(c->globalVrlsPerCluster);
        g.sampleRepresentatives(c->gcVrls, c->gcWeight);                         /* 169 */
        w2.refine({&g}, c->P.fallBackUndersampling);                             /* 177 */
        if (g.failed) throw Error(ALVRL_ERR_ARG, "couldn't refine global clustering! (but all VRLs should be non-zero!)");
        g.sampleRepresentatives(c->fallBackVrls, c->fallBackWeight);             /* 179 */
        c->nearTieSplits += g.nearTies;
        c->stats.kernelLaunches += w2.launchCount;
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
    /* neighbour slices in the local matrices (neighbourWeight > 0; getLocalMatrix 796-820, buildLocalities 1241-1293) */
    const bool neighbours = c->P.neighbourWeight > 0;
    DevBuf<float2> dRnb; DevBuf<double> dRowW; DevBuf<uint32_t> dRowMap;
    std::vector<uint32_t> nbR0(S + 1, 0);                     /* first row of slice i's block in the gathered matrix */
    uint32_t ldNb = 0;
    if (neighbours) {
        if (sfmt) throw Error(ALVRL_ERR_UNSUPPORTED, "neighbourWeight > 0 needs the counter sample stream on the device path");
        if (ranged) throw Error(ALVRL_ERR_UNSUPPORTED, "neighbourWeight > 0: the neighbour slices' rows of R live on other ranks (no halo exchange); use one handle");
        std::vector<std::set<std::pair<uint32_t, float>>> loc;
        build_localities(c, loc);
        std::vector<uint32_t> rowMap; std::vector<double> rowW;
        for (uint32_t i = 0; i < S; i++) {
            nbR0[i] = (uint32_t) rowMap.size();
            const uint32_t nr = c->rowOffset[i + 1] - c->rowOffset[i];
            for (uint32_t r = c->rowOffset[i]; r < c->rowOffset[i + 1]; r++) rowMap.push_back(r);
            /* Float arithmetic as written (796-820): 1.0 / dist is a double quotient stored to Float, the rest is Float, the
             * final weights are widened to double */
            std::vector<float> neighbourWeights(loc[i].size());
            float summedNeighbourWeight = 0;
            size_t j = 0;
            for (auto it = loc[i].begin(); it != loc[i].end(); ++it, ++j) {
                for (uint32_t r = c->rowOffset[it->first]; r < c->rowOffset[it->first + 1]; r++) rowMap.push_back(r);
                neighbourWeights[j] = (float) (1.0 / it->second);
                summedNeighbourWeight += neighbourWeights[j];
            }
            const float sliceWeight = summedNeighbourWeight * (1 - c->P.neighbourWeight) / c->P.neighbourWeight;
            const float normalization = 1 / (sliceWeight + summedNeighbourWeight);
            for (uint32_t k = 0; k < nr; k++) rowW.push_back(sliceWeight * normalization / nr);
            j = 0;
            for (auto it = loc[i].begin(); it != loc[i].end(); ++it, ++j) {
                const uint32_t nrj = c->rowOffset[it->first + 1] - c->rowOffset[it->first];
                for (uint32_t k = 0; k < nrj; k++) rowW.push_back(neighbourWeights[j] * normalization / nrj);
            }
        }
        nbR0[S] = (uint32_t) rowMap.size();
        const uint32_t rows = (uint32_t) rowMap.size();
        ldNb = (rows + 31u) & ~31u;
        dRowMap.upload(rowMap, st); dRowW.upload(rowW, st);
        dRnb.alloc((size_t) N * ldNb);
        k_gather_local_rows<<<(N + 7) / 8, 256, 0, st>>>(c->dR.p, c->ldR, dRowMap.p, rows, N, dRnb.p, ldNb);
        c->stats.kernelLaunches++;
        ALVRL_CUDA(cudaGetLastError());
        ALVRL_CUDA(cudaStreamSynchronize(st));
    }
    for (uint32_t i = sb; i < se; i++) {
        std::unique_ptr<Inst> in(new Inst());
        in->id = i; in->r0 = c->rowOffset[i]; in->nr = c->rowOffset[i + 1] - c->rowOffset[i];
        if (neighbours) { in->r0 = nbR0[i]; in->nr = nbR0[i + 1] - nbR0[i]; }
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
        if (sfmt) {
            /* a shared sequential stream: refine and sample slice after slice, in slice order (230-233, 746-752) */
            ws.allocInstances();
            ws.columnWeights();
            ws.construct(c->globalVrlsPerCluster);
            for (Inst *in : ws.insts) {
                if (c->P.localRefinement) ws.refine({in}, c->P.localUndersampling);
                if (!in->failed) in->sampleRepresentatives(c->selectedVrls[in->id], c->clusterWeight[in->id]);
            }
            c->stats.kernelLaunches += ws.launchCount;
        } else {
            /* independent per-slice streams: the Clustering objects are dealt to a few host threads, each driving its own CUDA
             * stream, so that one group's host work (sorting, heap updates) overlaps the other groups' kernels and copies */
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            /* device rounds leave little host work per round, and every round costs ~20 launches whatever its size: few groups
             * (more Clustering objects per launch) beat many */
            size_t wantGroups = getenv("ALVRL_HOST_ROUNDS") ? 3 : 1;      /* the device-resident refinement needs no host overlap */
            if (const char *e = getenv("ALVRL_CLUSTER_GROUPS")) wantGroups = (size_t) std::max(1, atoi(e));
            const size_t nGroups = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(wantGroups, hw), ws.insts.size() / 4 + 1));
            std::vector<std::unique_ptr<Workspace>> groups(nGroups);
            std::vector<std::string> errors(nGroups);
            std::vector<std::thread> threads;
            for (size_t gI = 0; gI < nGroups; gI++) {
                groups[gI].reset(new Workspace());
                Workspace &w2 = *groups[gI];
                w2.c = c; w2.N = N; w2.ldR = c->ldR; w2.R = c->dR.p; w2.serialSort = nGroups > 1; w2.deviceRounds = true;
                if (neighbours) { w2.R = dRnb.p; w2.ldR = ldNb; w2.rowW = dRowW.p; }
                for (size_t i = gI; i < ws.insts.size(); i += nGroups) w2.insts.push_back(ws.insts[i]);
            }
            for (size_t gI = 0; gI < nGroups; gI++) threads.emplace_back([&, gI]() {
                Workspace &w2 = *groups[gI];
                try {
                    ALVRL_CUDA(cudaSetDevice(c->device));
                    ALVRL_CUDA(cudaStreamCreateWithFlags(&w2.st, cudaStreamNonBlocking));
                    double q0 = Prof::now(), q1;
                    auto lap = [&](const char *what) { if (w2.prof.on) { cudaStreamSynchronize(w2.st); q1 = Prof::now(); fprintf(stderr, "[alvrl clustering] group %zu %s %.1f ms\n", gI, what, q1 - q0); q0 = q1; } };
                    w2.allocInstances(); lap("allocInstances");
                    w2.lazyMirrors = true;
                    w2.columnWeights(); lap("columnWeights");
                    w2.construct(c->globalVrlsPerCluster); lap("construct");
                    w2.depthCorrection = c->P.depthCorrection;
                    if (c->P.localRefinement) w2.refine(w2.insts, c->P.localUndersampling);
                    lap("refine");
                    w2.sampleRepresentativesDevice(c->selectedVrls, c->clusterWeight);
                    cudaStreamSynchronize(w2.st); lap("sampleRepresentatives");
                } catch (const std::exception &e) { errors[gI] = e.what(); }
                if (w2.st) { cudaStreamDestroy(w2.st); w2.st = nullptr; }
            });
            for (auto &t : threads) t.join();
            for (size_t gI = 0; gI < nGroups; gI++) {
                c->stats.kernelLaunches += groups[gI]->launchCount;
                if (!errors[gI].empty()) throw Error(ALVRL_ERR_ARG, errors[gI]);
            }
            const double tf0 = Prof::now();
            groups.clear();          /* frees the group workspaces before the fallback, if any */
            if (ws.prof.on) fprintf(stderr, "[alvrl clustering] free workspaces %.1f ms, build_clusters so far %.1f ms\n", Prof::now() - tf0, Prof::now() - tb0);
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
