/*
 * Device-resident cluster refinement (counter sample stream): ONE persistent CTA runs the whole refinement loop of one
 * Clustering object -- Clustering::refineAdaptively / refineFixedDepth / split / addCluster / calculateClusterVariance,
 * Preprocessor.cpp:387-452, 549-587, 590-684, 1058-1120 -- with no host round trip between splits.
 *
 * Why: the Clustering objects (one per slice) are independent latency chains of ~10^3 dependent splits each.  Driving
 * them from the host in lock-step rounds costs a launch pipeline + a device->host sync per round (the round is as slow as
 * its largest split); here every object advances at its own pace on its own SM and the host sees only the final queues.
 *
 * What stays identical to the other paths (and to the oracle), because it decides the result:
 *   - the multi-cluster queue order: heap_order.h, the libstdc++ sift sequences on (key, node) entries kept in shared memory;
 *   - the singleton list order (front insertion: here an append-only array that the host reads back to front);
 *   - the fp32 running sums underVar / intVar in the order popMulti / addCluster apply them, and the convergence test;
 *   - weightedSample: sequential fp32 running sums in list order by one thread (chunked through shared memory; the
 *     index search over the stored sums is parallel), the second draw with the first centre's weight counted as 0;
 *   - direction norms and projections: sequential fp32 in row order, -fmad=false;
 *   - the sort: (projection, vrl) keys are unique, so the bitonic network yields std::sort's order;
 *   - the first-minimum argmin over head + tail variance.
 * The prefix variances are double sums associated differently from the reference's (block scans across steps instead of a
 * sequential sweep), the same class of last-bit-of-double differences as the batched pipeline (file header of clustering.cu).
 *
 * Layout: thread = column for the projections, thread = step for the variances; for a step chunk every row needs one
 * block-wide exclusive scan of the means (S_r(k-1)), done RF_ROWS rows at a time with one barrier per scan; each thread
 * accumulates its own B_k = sum_r (w_k S_r(k-1) - W_{k-1} x_r(k))^2.  Clusters whose local matrix fits the shared-memory
 * tile (n * nr <= 24576, i.e. almost all of them) read R exactly once per split; larger ones stream columns from L2.
 */
#pragma once
#include <type_traits>
#include "heap_order.h"

namespace alvrl {

#define RF_THREADS 512
#define RF_WARPS (RF_THREADS / 32)
#define RF_MAXROWS 512              /* rows whose sweeps fit one pass (thread = 2 rows); gangs and resident tiles need this */
#define RF_MAXROWS_BIG 4096         /* more rows than RF_MAXROWS: row-block passes in the sweeps, direction in the tile */
#define RF_TILE_FLOATS 24576
#define RF_HEAP_CAP 4096           /* queue entries kept in shared memory; deeper levels spill to global memory (SplitHeap) */
#define RF_SMALL 1024
#define RF_CHUNK 1024
#define RF_MAXCHUNKS 1024
#define RF_ROWS 2
#define RF_SORT_BLOCK 8192          /* keys sorted in one piece out of the tile (64 KB) */

enum { RF_DONE = 0, RF_RESUME_HOST = 1, RF_ERR_CONSTANT = 2, RF_ERR_LOWER = 3, RF_ERR_SPLIT = 4, RF_ERR_NOBEST = 5,
       RF_ERR_SINGLETON_VAR = 6, RF_ERR_WEIGHTS = 7 };

struct RfInst {
    /* constants of the Clustering object */
    uint32_t r0, nr; double lw; uint64_t listOff, cwOff;
    uint32_t nrP; uint64_t xOff, vOff;  /* the compacted local matrix: X[xOff + vrl * nrP + row], nrP = roundup(nr, 4); Vcol[vOff + vrl] */
    uint32_t numVrlsTotal; float pixelUndersampling, tracingVar, unclIntVar;
    uint32_t adaptive, targetClusters, rngKey;
    /* depthCorrection != 1 (Preprocessor.cpp:403-408, 455-470): pass 1 records the split count of the best convergence constant
     * (bestSplits, counted from splitsBase), pass 2 -- not adaptive -- runs exactly fixedSplits splits from the initial queue */
    uint32_t splitsBase, bestSplits, fixedSplits;
    /* state, in and out */
    uint32_t rngPos; float underVar, intVar, bestConstant;
    uint32_t heapCount, nodeCount, singleCount;
    uint32_t sHeapCount, sSingleCount; float sUnder, sInt;
    uint32_t nearTies, status, splits, degenerate;
    /* in: the initial queue (nodes in array order) and singletons (insertion order) in the compact init arrays;
     * out: the final queue followed by the best-so-far snapshot (nodes in array order) and the singletons, compacted */
    uint32_t initNodeOff, initSingleOff, outNodeOff, outSingleOff;
    unsigned long long clk[2][12];      /* cycles per phase, [small | large cluster]: pick, direction, stage, project, sort, weights, sweep, pairs, argmin+queue, count */
    uint32_t mtInit, mtWaves, mtFinishing;   /* k_refine_mt: the object's queue has been set up / control passes so far / converged, waiting for splits in flight */
};

struct RfScratch {                      /* per CTA */
    unsigned long long *keys; double *w, *Wf, *Wr; float2 *pairs; uint64_t keyStride, stepStride;   /* clusters too large for shared memory */
    uint32_t *srcPos, *posTmp;          /* [stepStride] source position of every sorted step / position of every VRL id (large clusters) */
    uint32_t heapCap, nodeCap;          /* capacities per CTA: queue entries (>= RF_HEAP_CAP), nodes / singletons */
    HeapEntry *heapOv;                  /* [heapCap - RF_HEAP_CAP] queue entries beyond the shared-memory part */
    HeapEntry *snapHeap;                /* [heapCap] best-so-far snapshot of the queue */
    ClusterNode *nodes;                 /* [nodeCap] append-only node table the queue entries index */
    uint32_t *singles;                  /* [nodeCap] singleton VRL ids in insertion order */
    const ClusterNode *initNodes; const uint32_t *initSingles; ClusterNode *outNodes; uint32_t *outSingles;
    uint32_t *cursors;                  /* [0] next object, [1] output node cursor, [2] output singleton cursor */
    uint32_t unfoldRows;                /* experiment (ALVRL_RF_UNFOLD): one row per thread in the variance sweeps when nr <= 256 */
    /* gangs (k_refine_mt): clusters of >= gangMin columns are split by 2 CTAs, >= 2 gangMin by 4, ... up to gangMax (0: off) */
    uint32_t gangMin, gangMax;
    double *carry;                      /* [grid][2][RF_GANG_MAX][RF_MAXROWS] row sums of the members' step ranges */
    const double *rowW;                 /* per-row locality weights indexed like the rows of R (neighbour slices in L_i), or nullptr: uniform lw */
};
#define RF_GANG_MAX 16
/* CTAs that split a cluster of n columns together: a pure function of n (and of the launch), so that every member and the
 * control pass that hands the cluster out agree */
__host__ __device__ __forceinline__ uint32_t rf_gang_size(uint32_t n, uint32_t gangMin, uint32_t gangMax) {
    if (!gangMin || n < gangMin) return 1u;
    uint32_t g = 2u;
    while (g < gangMax && (uint64_t) n >= (uint64_t) gangMin * g) g <<= 1;
    return g;
}

struct RfShared {
    float tile[RF_TILE_FLOATS];                         /* staged columns of the local matrix (swizzled granules, see k_refine) */
    HeapEntry heap[RF_HEAP_CAP];
    unsigned long long keys[RF_SMALL];
    double w[RF_SMALL], Wf[RF_SMALL], Wr[RF_SMALL];
    float2 pairs[2][RF_SMALL];
    float sw[2][RF_CHUNK], acc[RF_CHUNK], chunkEnd[RF_MAXCHUNKS];
    uint16_t pos[RF_SMALL];
    float c1[RF_MAXROWS], c2[RF_MAXROWS], sdir[RF_MAXROWS];
    double scan[2][RF_WARPS][RF_ROWS];
    double stepW[2][2][32], stepWp[2][2][32];
    uint64_t mbar[2][2];                                /* [half][stage]: completion of the bulk copies of the variance ring */           /* [half][stage][step]: w_k and W_{k-1} of the staged steps (large clusters) */
    float rb[RF_WARPS], rs[RF_WARPS]; uint32_t ri[RF_WARPS];
    /* control block (written by thread 0 between barriers) */
    uint32_t inst, begin, end, srcBuf, found, pick[2], flags, done, snap, err, nodeKey, nodePos;
    uint32_t task[2], stop, selCount, sel[128];             /* k_refine_mt: the task in hand, the control pass */
    uint32_t gang[3];                                      /* k_refine_mt: gang size, this CTA's member index, the leader's block */
    unsigned long long mtClk[8];                          /* control cycles, ticket wait, control passes, split tasks, gang sync wait, gang syncs */
    float u1, u2, norm[3];
    /* refinement state (thread 0) */
    uint32_t rngPos, heapCount, nodeCount, singleCount, sHeapCount, sSingleCount, nearTies, splits, degenerate, bestSplits;
    float underVar, intVar, bestConstant, sUnder, sInt;
    unsigned long long clk[2][12];
};

/* block-wide scan of NV doubles per thread with ONE barrier (callers alternate the scratch buffer); warps >= activeWarps
 * hold zeros and skip the shuffles.  total[] = sum over the block (valid when wantTotal). */
template <int NV, bool EXCL>
__device__ __forceinline__ void rf_scan(double (&v)[NV], double (*scratch)[RF_ROWS], uint32_t activeWarps, bool wantTotal, double (&total)[NV]) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < activeWarps) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < NV; q++) { const double nb = __shfl_up_sync(0xffffffffu, v[q], o); if (lane >= (uint32_t) o) v[q] += nb; }
        }
        if (lane == 31) {
#pragma unroll
            for (int q = 0; q < NV; q++) scratch[warp][q] = v[q];
        }
        if (EXCL) {
#pragma unroll
            for (int q = 0; q < NV; q++) { const double e = __shfl_up_sync(0xffffffffu, v[q], 1); v[q] = lane ? e : 0.0; }
        }
    }
    __syncthreads();
    double off[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) { off[q] = 0; total[q] = 0; }
    const uint32_t upto = wantTotal ? activeWarps : min(warp, activeWarps);
    for (uint32_t i = 0; i < upto; i++) {
#pragma unroll
        for (int q = 0; q < NV; q++) { const double s = scratch[i][q]; if (i < warp) off[q] += s; total[q] += s; }
    }
#pragma unroll
    for (int q = 0; q < NV; q++) v[q] = off[q] + v[q];
}

__device__ __forceinline__ void rf_cp_async16(void *smemDst, const void *gmemSrc) {
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void rf_cp_async8(void *smemDst, const void *gmemSrc) {
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void rf_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void rf_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void rf_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void rf_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rf_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
                 ::"r"((uint32_t) __cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
/* one bulk asynchronous copy (TMA, 1-D) of a contiguous block of global memory into shared memory */
__device__ __forceinline__ void rf_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t) __cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t) __cvta_generic_to_shared(bar)) : "memory");
}

/* The four quotients e / len of the projection loop (Preprocessor.cpp:632), correctly rounded without four IEEE division
 * routines: with y = RN(1 / len), q0 = RN(e y) and one fma correction, q = RN(q0 + RN(e - len q0) y) IS RN(e / len) (Markstein)
 * as long as nothing on the way leaves the normal range and len's significand is not all ones.  lenOk states the conditions on
 * len (1e-30 <= len < 1e6, significand != 0x7fffff; then |e| <= len keeps q0 and y normal), a non-zero |e| below 1e-30 (whose
 * residual could underflow) takes the division.  Checked against the division on 5e9 random pairs (tools/micro/div_exact.cpp). */
__device__ __forceinline__ float4 rf_div4(const float4 e, const float len, const float y, const bool lenOk) {
    float4 q = make_float4(e.x * y, e.y * y, e.z * y, e.w * y);
    q.x = fmaf(fmaf(-q.x, len, e.x), y, q.x); q.y = fmaf(fmaf(-q.y, len, e.y), y, q.y);
    q.z = fmaf(fmaf(-q.z, len, e.z), y, q.z); q.w = fmaf(fmaf(-q.w, len, e.w), y, q.w);
    const float tiny = 1e-30f;
    const bool slow = !lenOk || (fabsf(e.x) < tiny && e.x != 0.0f) || (fabsf(e.y) < tiny && e.y != 0.0f) ||
                      (fabsf(e.z) < tiny && e.z != 0.0f) || (fabsf(e.w) < tiny && e.w != 0.0f);
    if (slow) q = make_float4(e.x / len, e.y / len, e.z / len, e.w / len);
    return q;
}
__device__ __forceinline__ bool rf_div_len_ok(float len) {
    return len >= 1e-30f && len < 1e6f && (__float_as_uint(len) & 0x7fffffu) != 0x7fffffu;
}

/* sums over the 32 lanes of 16 values per lane with 16 shuffles: afterwards b[0] of lane l holds the total of value (l & 15) */
__device__ __forceinline__ void rf_reduce16(double (&b)[16], uint32_t lane) {
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (lane & (uint32_t) s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const double send = up ? b[i] : b[i + s], keep = up ? b[i + s] : b[i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    b[0] += __shfl_xor_sync(0xffffffffu, b[0], 16);
}

/* the compare-exchange steps j = jStart, jStart / 2, .., 1 of bitonic stage k on `len` keys in shared memory whose first key has
 * global index `base` (the sort direction of a pair depends on its global index); `pos` is an optional 16-bit payload */
__device__ __forceinline__ void rf_sort_steps(unsigned long long *sk, uint32_t len, uint32_t base, uint32_t k, uint32_t jStart, uint16_t *pos) {
    for (uint32_t j = jStart; j > 0; j >>= 1) {
        for (uint32_t t = threadIdx.x; t < (len >> 1); t += RF_THREADS) {
            const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
            const unsigned long long a = sk[lo], b = sk[hi];
            if ((a > b) == (((base + lo) & k) == 0)) {
                sk[lo] = b; sk[hi] = a;
                if (pos) { const uint16_t pa = pos[lo]; pos[lo] = pos[hi]; pos[hi] = pa; }
            }
        }
        __syncthreads();
    }
}

/* The local matrices of the objects (rows [r0, r0 + nr) of every column of R) gathered once into contiguous, zero-padded column
 * blocks IN LIST ORDER -- column at list position i of the object at XA[xOff + i * nrP + row] -- and the per-column sums
 * sum_r var_r / w (the "second" term of calculateClusterVariance, order independent), indexed by VRL id.  From here on the
 * refinement keeps the columns of every cluster physically contiguous (each split writes its sorted order into the other
 * copy), so that a split streams n * nrP * 4 contiguous bytes instead of gathering columns from all over R.  Warp = column. */
__global__ void __launch_bounds__(256) k_rf_compact(const float2 *__restrict__ R, uint32_t ldR, uint32_t N, const RfInst *__restrict__ insts,
                                                    const uint32_t *__restrict__ lists, const float *__restrict__ cw, float *__restrict__ X,
                                                    double *__restrict__ Vcol, const double *__restrict__ rowW) {
    const RfInst &I = insts[blockIdx.y];
    const uint32_t lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= N) return;
    const uint32_t v = lists[I.listOff + i];
    const float2 *col = R + (size_t) v * ldR + I.r0;
    float *out = X + I.xOff + (size_t) i * I.nrP;
    const double wc = (double) cw[I.cwOff + v];
    double vy = 0;
    for (uint32_t r = lane; r < I.nrP; r += 32) {
        float2 e = make_float2(0, 0);
        if (r < I.nr) { e = col[r]; vy += (rowW ? rowW[I.r0 + r] : 1.0) * ((double) e.y / wc); }
        out[r] = e.x;
    }
    for (int o = 16; o > 0; o >>= 1) vy += __shfl_down_sync(0xffffffffu, vy, o);
    if (lane == 0) Vcol[I.vOff + v] = vy;
}

__global__ void __launch_bounds__(RF_THREADS, 1) k_refine(float *XA, float *XB, const double *__restrict__ Vcol, RfInst *insts, uint32_t numInst,
                                                          uint32_t *lists, const float *__restrict__ cw, RfScratch scr) {
    extern __shared__ __align__(16) unsigned char rfRaw[];
    RfShared &sm = *reinterpret_cast<RfShared *>(rfRaw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long *keysG = scr.keys + (uint64_t) blockIdx.x * scr.keyStride;
    double *wG = scr.w + (uint64_t) blockIdx.x * scr.stepStride, *WfG = scr.Wf + (uint64_t) blockIdx.x * scr.stepStride,
           *WrG = scr.Wr + (uint64_t) blockIdx.x * scr.stepStride;
    float2 *pairsG = scr.pairs + (uint64_t) blockIdx.x * 2 * scr.stepStride;
    HeapEntry *snapHeap = scr.snapHeap + (uint64_t) blockIdx.x * scr.heapCap;
    ClusterNode *nodes = scr.nodes + (uint64_t) blockIdx.x * scr.nodeCap;
    uint32_t *singles = scr.singles + (uint64_t) blockIdx.x * scr.nodeCap;
    SplitHeap heap; heap.lo = sm.heap; heap.hi = scr.heapOv + (uint64_t) blockIdx.x * (scr.heapCap - RF_HEAP_CAP); heap.cap = RF_HEAP_CAP;
    uint32_t *srcG = scr.srcPos + (uint64_t) blockIdx.x * scr.stepStride, *posTmp = scr.posTmp + (uint64_t) blockIdx.x * scr.stepStride;
    uint32_t scanIt = 0;
    uint32_t ringPhase = 0;                             /* parity of the two ring barriers of this thread's half */
    if (tid == 0) {
        for (int a = 0; a < 4; a++) rf_mbar_init(&sm.mbar[a >> 1][a & 1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    long long tPhase = 0;
#define RF_TICK(ph) do { if (tid == 0) { const long long now_ = clock64(); sm.clk[small ? 0 : 1][ph] += (unsigned long long) (now_ - tPhase); tPhase = now_; } } while (0)

    for (;;) {                                                      /* persistent: objects are handed out dynamically */
        __syncthreads();
        if (tid == 0) sm.inst = atomicAdd(scr.cursors, 1u);
        __syncthreads();
        if (sm.inst >= numInst) break;
        RfInst *I = insts + sm.inst;
        const uint32_t nr = I->nr, key = I->rngKey;
        const double lw = scr.rowW ? 1.0 : I->lw;                     /* per-row weights enter B_k and Vcol instead */
        uint32_t *ilist = lists + I->listOff;
        const float *icw = cw + I->cwOff;
        /* padded column length, 16-byte granules per column, tile column stride (a multiple of 8 granules: element r of tile
         * column c sits in granule (r / 4) ^ (c & 7), which makes both the thread = column LDS.128 of the projections and the
         * thread = row LDS.32 of the variance sweep bank-conflict free) */
        const uint32_t nrP = I->nrP, nq = nrP >> 2, tS = (nrP + 31u) & ~31u;
        const double *Vi = Vcol + I->vOff;
        const uint32_t TV = min((uint32_t) RF_SMALL, (uint32_t) RF_TILE_FLOATS / tS);
        for (uint32_t i = tid; i < I->heapCount; i += RF_THREADS) {
            const ClusterNode cn = scr.initNodes[I->initNodeOff + i];
            HeapEntry e; e.key = cn.undersamplingVar + cn.integrationVar; e.id = i;
            nodes[i] = cn; heap[i] = e; snapHeap[i] = e;
        }
        for (uint32_t i = tid; i < I->singleCount; i += RF_THREADS) singles[i] = scr.initSingles[I->initSingleOff + i];
        if (tid == 0) {
            sm.rngPos = I->rngPos; sm.heapCount = I->heapCount; sm.nodeCount = I->nodeCount; sm.singleCount = I->singleCount;
            sm.underVar = I->underVar; sm.intVar = I->intVar; sm.bestConstant = I->bestConstant;
            sm.sHeapCount = I->heapCount; sm.sSingleCount = I->singleCount; sm.sUnder = I->underVar; sm.sInt = I->intVar;
            sm.nearTies = 0; sm.splits = 0; sm.degenerate = 0; sm.err = RF_DONE; sm.bestSplits = I->bestSplits;
            for (int a = 0; a < 24; a++) sm.clk[a / 12][a % 12] = 0;
        }
        __syncthreads();

        for (;;) {                                                  /* one Clustering::split per iteration */
            if (tid == 0) {
                sm.done = 0; sm.snap = 0;
                if (sm.heapCount + 2 > scr.heapCap || sm.nodeCount + 2 > scr.nodeCap || sm.singleCount + 2 > scr.nodeCap) { sm.err = RF_RESUME_HOST; sm.done = 1; }
                else {
                    const HeapEntry top = heap_pop(heap, sm.heapCount);                  /* popMulti, 581-587 */
                    const ClusterNode cn = nodes[top.id];
                    sm.underVar -= cn.undersamplingVar; sm.intVar -= cn.integrationVar;
                    sm.begin = cn.begin & 0x7fffffffu; sm.end = cn.end; sm.srcBuf = cn.begin >> 31;   /* bit 31: which copy holds the columns */
                    if (sm.end - sm.begin < 2) { sm.err = RF_ERR_SPLIT; sm.done = 1; }
                    sm.nodeKey = alvrl_rng_node_key(key, sm.begin, sm.end);              /* the cluster's own sub-stream (alvrl_rng.h) */
                    sm.u1 = alvrl_rng_uniform(sm.nodeKey, 0); sm.u2 = alvrl_rng_uniform(sm.nodeKey, 1); sm.nodePos = 2;   /* 597-602 */
                    sm.flags = 0; sm.found = 0xffffffffu;
                }
            }
            __syncthreads();
            if (sm.done) break;
            const uint32_t begin = sm.begin, n = sm.end - sm.begin;
            uint32_t *list = ilist + begin;
            /* the columns of this cluster, physically in list order: column p at Xs + p * nrP; the split writes them in sorted
             * order into the other copy, which the two halves inherit */
            const uint32_t srcBuf = sm.srcBuf;
            const float *Xs = (srcBuf ? XB : XA) + I->xOff + (size_t) begin * nrP;
            float *Xd = (srcBuf ? XA : XB) + I->xOff + (size_t) begin * nrP;
            const bool small = n <= RF_SMALL;
            unsigned long long *keys = small ? sm.keys : keysG;
            double *wA = small ? sm.w : wG, *WfA = small ? sm.Wf : WfG, *WrA = small ? sm.Wr : WrG;
            float2 *pairsF = small ? sm.pairs[0] : pairsG, *pairsR = small ? sm.pairs[1] : pairsG + scr.stepStride;
            if (tid == 0) { tPhase = clock64(); sm.clk[small ? 0 : 1][9]++; sm.clk[small ? 0 : 1][10] += n; sm.clk[small ? 0 : 1][11] += (n <= TV) ? 0u : n; }

            uint32_t *listDst = list;                                  /* this kernel sorts the list in place */
            const uint32_t gG = 1u, gMi = 0u; uint32_t gPhase = 0u; uint32_t *const gCnt = nullptr, *const gCursor = nullptr; double *const gCarry = nullptr;   /* no gangs here */
            (void) gPhase; (void) gCnt; (void) gCarry; (void) gCursor;
#include "refine_split.inl"
            if (tid == 0) {
                best = sm.rb[0]; second = sm.rs[0]; bi = sm.ri[0];
                for (uint32_t i = 1; i < RF_WARPS; i++) {
                    const float b2 = sm.rb[i], s2 = sm.rs[i]; const uint32_t i2 = sm.ri[i];
                    if (b2 < best || (b2 == best && i2 < bi)) { second = fminf(s2, best); best = b2; bi = i2; }
                    else second = fminf(second, b2);
                }
                sm.splits++;
                if (bi == 0xffffffffu) { sm.err = RF_ERR_NOBEST; sm.done = 1; }
                else {
                    if (isfinite(second) && fabsf(second - best) <= 1e-6f * fabsf(best)) sm.nearTies++;
                    /* addCluster(begin, split) then addCluster(split, end), 549-572 */
                    const uint32_t split = begin + bi, end = begin + n;
                    for (int half = 0; half < 2; half++) {
                        const uint32_t b = half ? split : begin, e = half ? end : split;
                        const float2 pv = half ? pairsR[n - 1 - bi] : pairsF[bi - 1];
                        if (e == b + 1) {
                            singles[sm.singleCount++] = (uint32_t) (keys[b - begin] & 0xffffffffull);
                            if (pv.x != 0) { sm.err = RF_ERR_SINGLETON_VAR; sm.done = 1; }
                            sm.intVar += pv.y;
                        } else {
                            ClusterNode cn; cn.undersamplingVar = pv.x; cn.integrationVar = pv.y; cn.begin = b | ((srcBuf ^ 1u) << 31); cn.end = e;
                            nodes[sm.nodeCount] = cn;
                            HeapEntry he; he.key = pv.x + pv.y; he.id = sm.nodeCount++;
                            heap_push(heap, sm.heapCount, he);
                            sm.underVar += pv.x; sm.intVar += pv.y;
                        }
                    }
                    const uint32_t numClusters = sm.heapCount + sm.singleCount;
                    if (I->adaptive) {                                              /* refineAdaptively, 436-452 */
                        const float scale = I->numVrlsTotal * I->pixelUndersampling + numClusters;
                        const float curr = scale * (I->tracingVar + sm.underVar + sm.intVar);
                        const float lower = scale * (I->tracingVar + I->unclIntVar);
                        if (!isfinite(curr) || curr <= 0) { sm.err = RF_ERR_CONSTANT; sm.done = 1; }
                        else if (!isfinite(lower) || lower <= 0) { sm.err = RF_ERR_LOWER; sm.done = 1; }
                        else {
                            if (curr < sm.bestConstant) { sm.snap = 1; sm.bestConstant = curr; sm.bestSplits = I->splitsBase + sm.splits; }
                            if (lower >= sm.bestConstant || sm.heapCount == 0) sm.done = 1;
                        }
                    } else if (I->fixedSplits) { if (I->splitsBase + sm.splits >= I->fixedSplits || sm.heapCount == 0) sm.done = 1; }   /* second pass, 455-470 */
                    else if (!(numClusters < I->targetClusters && sm.heapCount > 0)) sm.done = 1;     /* refineFixedDepth, 387-399 */
                }
            }
            __syncthreads();
            RF_TICK(8);
            const bool doSnap = sm.snap != 0, isDone = sm.done != 0;
            if (doSnap) {
                for (uint32_t i = tid; i < sm.heapCount; i += RF_THREADS) snapHeap[i] = heap[i];
                if (tid == 0) { sm.sHeapCount = sm.heapCount; sm.sSingleCount = sm.singleCount; sm.sUnder = sm.underVar; sm.sInt = sm.intVar; }
            }
            __syncthreads();
            if (isDone) break;
        }
        __syncthreads();
        if (tid == 0) {
            sm.begin = atomicAdd(scr.cursors + 1, sm.heapCount + sm.sHeapCount);
            sm.end = atomicAdd(scr.cursors + 2, sm.singleCount);
        }
        __syncthreads();
        for (uint32_t i = tid; i < sm.heapCount; i += RF_THREADS) { ClusterNode cn = nodes[heap[i].id]; cn.begin &= 0x7fffffffu; scr.outNodes[sm.begin + i] = cn; }
        for (uint32_t i = tid; i < sm.sHeapCount; i += RF_THREADS) { ClusterNode cn = nodes[snapHeap[i].id]; cn.begin &= 0x7fffffffu; scr.outNodes[sm.begin + sm.heapCount + i] = cn; }
        for (uint32_t i = tid; i < sm.singleCount; i += RF_THREADS) scr.outSingles[sm.end + i] = singles[i];
        if (tid == 0) {
            I->outNodeOff = sm.begin; I->outSingleOff = sm.end;
            for (int a = 0; a < 24; a++) I->clk[a / 12][a % 12] = sm.clk[a / 12][a % 12];
            I->rngPos = sm.rngPos; I->underVar = sm.underVar; I->intVar = sm.intVar; I->bestConstant = sm.bestConstant;
            I->heapCount = sm.heapCount; I->nodeCount = sm.nodeCount; I->singleCount = sm.singleCount;
            I->sHeapCount = sm.sHeapCount; I->sSingleCount = sm.sSingleCount; I->sUnder = sm.sUnder; I->sInt = sm.sInt;
            I->nearTies = sm.nearTies; I->status = sm.err; I->splits = sm.splits; I->degenerate = sm.degenerate; I->bestSplits = sm.bestSplits;
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * k_refine_mt: the same refinement with MANY CTAs per Clustering object.
 *
 * A split is a pure function of the cluster it splits: its columns, its list range and -- in the counter sample stream --
 * its own sub-stream of uniforms (alvrl_rng_node_key).  What is sequential is only the ORDER in which the refinement consumes
 * the results: the queue pops, the running variance sums, the convergence test, the snapshots.  So the work is cut in two
 * kinds of tasks that persistent CTAs pull from one global ticket queue:
 *   split(o, node)  -- refine_split.inl on one cluster of object o; the result (split index, the variance pairs of the two
 *                      halves, the VRLs of singleton halves) goes into the node record.  The cluster's list range and columns
 *                      are read from copy `src` (bit 31 of node.begin) and written, sorted, to the other copy, so a split
 *                      that the refinement never gets to consume leaves the cluster's own data untouched.
 *   control(o)      -- one CTA replays the refinement of object o as far as the results reach: pop the top of the queue, apply
 *                      its split exactly like k_refine's thread 0 (same heap operations, same fp32 sums, same convergence
 *                      test, snapshots), until the top is a cluster without a result.  It names that cluster in waitNode[o]
 *                      and tops the splits in flight up to MT_K with the unsplit clusters of the largest keys (the ones the
 *                      queue will pop next, unless their own children overtake them).  The split that delivers the awaited
 *                      result schedules the next control(o); ctl[o] keeps at most one pass scheduled or running (a request
 *                      that arrives during a pass makes it run again).  Publication and re-check are ordered by fences on both
 *                      sides (store, fence, load), so a result cannot slip between "not ready" and "waiting for it".
 * Only clusters that are IN the queue are split ahead of time (never the children of a cluster whose split has not been
 * consumed), which is what makes two copies of lists and columns enough: the destination range of such a split holds the dead
 * source of the cluster's parent.  At the end -- once no split of the object is in flight any more -- the leaves whose list
 * sits in copy 1 are copied into copy 0, which reproduces the in-place list of the sequential algorithm (the list inside a
 * cluster is what the consumed splits left, 641-642).
 * No co-residency is assumed: a CTA only waits for a ticket it has drawn, and tickets are filled by running CTAs.
 * ------------------------------------------------------------------------------------------------------------------- */
#define MT_K_MAX 128                /* most splits of one object in flight (sm.sel) */

struct MtNode {                         /* 64 bytes */
    float under, integ; uint32_t begin, end;            /* ClusterNode; bit 31 of begin: which copy holds list range and columns */
    uint32_t state, bi;                                 /* 0 no result, 1 handed out, 2 split; split index relative to begin */
    float2 pvH, pvT;                                    /* (undersampling, integration) variance of the head / tail half */
    uint32_t svH, svT;                                  /* first / last VRL of the sorted order (a singleton half) */
    uint32_t flags;                                     /* bit 0 near tie, bits 8..27 random-direction draws, bits 28..31 RF_ERR_* */
    uint32_t pad[3];
};

struct MtPools {
    MtNode *nodes; HeapEntry *heap, *snap; uint32_t *singles;   /* per object: [nodeCap], [heapCap], [heapCap], [nodeCap] */
    uint32_t *outstanding;                                      /* per object: split tasks in flight */
    uint32_t *ctl;                                              /* per object: 0 no control pass scheduled, 1 scheduled or running, 2 ... and asked to run again */
    uint32_t *waitNode;                                         /* per object: the cluster whose result the last control pass stopped at (MT_NONE / MT_ANY) */
    unsigned long long *slots; uint32_t qmask;                  /* ticket ring: (generation << 40) | (type << 39) | (object << 24) | node */
    uint32_t *ctr;                                              /* [0] next ticket to draw, [1] next ticket to fill, [2] objects not done */
    uint32_t nodeCap, heapCap;
    uint32_t inflight;                                          /* splits of one object kept in flight (<= MT_K_MAX) */
    unsigned long long *clk;                                    /* [grid][32] profile counters */
};

#define MT_NONE 0xffffffffu
#define MT_ANY 0xfffffffeu
__device__ __forceinline__ unsigned long long mt_word(uint32_t ticket, uint32_t qmask, uint32_t type, uint32_t obj, uint32_t node) {
    const unsigned long long gen = (unsigned long long) (ticket / (qmask + 1u)) + 1ull;
    return (gen << 40) | ((unsigned long long) type << 39) | ((unsigned long long) obj << 24) | (unsigned long long) node;
}
/* thread 0, after a __syncthreads(): publishes `cnt` tasks (everything this CTA wrote before becomes visible first) */
__device__ __forceinline__ void mt_push(const MtPools &mp, uint32_t type, uint32_t obj, const uint32_t *nodeIds, uint32_t cnt) {
    __threadfence();
    const uint32_t t0 = atomicAdd(mp.ctr + 1, cnt);
    for (uint32_t j = 0; j < cnt; j++)
        atomicExch(mp.slots + ((t0 + j) & mp.qmask), mt_word(t0 + j, mp.qmask, type, obj, nodeIds ? nodeIds[j] : 0u));
}

/* thread 0: make sure a control pass of object o runs after this point (at most one is scheduled or running at any time) */
__device__ __forceinline__ void mt_request_control(const MtPools &mp, uint32_t o) {
    for (;;) {
        const uint32_t old = atomicCAS(mp.ctl + o, 0u, 1u);
        if (old == 0u) { mt_push(mp, 1u, o, nullptr, 1u); return; }
        if (old == 2u) return;
        if (atomicCAS(mp.ctl + o, 1u, 2u) == 1u) return;
    }
}

__global__ void __launch_bounds__(RF_THREADS, 1) k_refine_mt(float *XA, float *XB, const double *__restrict__ Vcol, RfInst *insts, uint32_t numInst,
                                                             uint32_t *L0, uint32_t *L1, const float *__restrict__ cw, RfScratch scr, MtPools mp) {
    extern __shared__ __align__(16) unsigned char rfRaw[];
    RfShared &sm = *reinterpret_cast<RfShared *>(rfRaw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t scanIt = 0;
    uint32_t ringPhase = 0;
    if (tid == 0) {
        for (int a = 0; a < 4; a++) rf_mbar_init(&sm.mbar[a >> 1][a & 1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int a = 0; a < 24; a++) sm.clk[a / 12][a % 12] = 0;
        for (int a = 0; a < 8; a++) sm.mtClk[a] = 0;
    }
    long long tPhase = 0;
    uint32_t gangCap = 1u; while (gangCap * 2u <= gridDim.x && gangCap * 2u <= min(scr.gangMax, (uint32_t) RF_GANG_MAX)) gangCap <<= 1;   /* a gang never exceeds the grid */

    for (;;) {
        /* ---- draw a ticket and wait for its task ---- */
        __syncthreads();
        if (tid == 0) {
            const long long w0 = clock64();
            const uint32_t ticket = atomicAdd(mp.ctr, 1u);
            const unsigned long long want = (unsigned long long) (ticket / (mp.qmask + 1u)) + 1ull;
            volatile unsigned long long *slot = mp.slots + (ticket & mp.qmask);
            unsigned long long w;
            for (;;) {
                w = *slot;
                if ((w >> 40) == want) break;
                if (*(volatile uint32_t *) (mp.ctr + 2) == 0u) { w = 0; break; }
                __nanosleep(100);
            }
            __threadfence();
            sm.task[0] = (uint32_t) (w & 0xffffffffull); sm.task[1] = (uint32_t) (w >> 32);
            sm.mtClk[1] += (unsigned long long) (clock64() - w0);
        }
        __syncthreads();
        const unsigned long long word = ((unsigned long long) sm.task[1] << 32) | sm.task[0];
        if (word == 0ull) break;
        const uint32_t o = (uint32_t) (word >> 24) & 0x7fffu, taskNode = (uint32_t) word & 0xffffffu;
        const bool isControl = ((word >> 39) & 1ull) != 0;
        RfInst *I = insts + o;
        MtNode *nodes = mp.nodes + (uint64_t) o * mp.nodeCap;
        const uint32_t nr = I->nr;
        const double lw = scr.rowW ? 1.0 : I->lw;                     /* per-row weights enter B_k and Vcol instead */
        const float *icw = cw + I->cwOff;
        const uint32_t nrP = I->nrP, nq = nrP >> 2, tS = (nrP + 31u) & ~31u;
        const double *Vi = Vcol + I->vOff;
        const uint32_t TV = min((uint32_t) RF_SMALL, (uint32_t) RF_TILE_FLOATS / tS);

        if (!isControl) {
            /* ================= split(o, taskNode) ================= */
            MtNode *nd = nodes + taskNode;
            if (tid == 0) {
                const uint32_t b = *(volatile uint32_t *) &nd->begin, e = *(volatile uint32_t *) &nd->end;
                sm.begin = b & 0x7fffffffu; sm.end = e; sm.srcBuf = b >> 31;
                sm.nodeKey = alvrl_rng_node_key(I->rngKey, sm.begin, sm.end);
                sm.u1 = alvrl_rng_uniform(sm.nodeKey, 0); sm.u2 = alvrl_rng_uniform(sm.nodeKey, 1); sm.nodePos = 2;   /* 597-602 */
                sm.flags = 0; sm.found = 0xffffffffu; sm.err = RF_DONE; sm.degenerate = 0;
                sm.mtClk[3]++;
                /* a large cluster arrives as G identical, consecutive tickets: the CTAs that draw them form a gang.  Members are
                 * numbered in the order they arrive (node.pad[2]); the first one leads: its scratch arrays are the gang's */
                const uint32_t G = I->nr <= RF_MAXROWS ? rf_gang_size(sm.end - sm.begin, scr.gangMin, gangCap) : 1u;
                uint32_t mi = 0, leader = blockIdx.x;
                if (G > 1u) {
                    mi = atomicAdd(&nd->pad[2], 1u);
                    if (mi == 0u) { atomicExch(&nd->pad[0], blockIdx.x + 1u); }
                    else { uint32_t l; while ((l = *(volatile uint32_t *) &nd->pad[0]) == 0u) __nanosleep(64); leader = l - 1u; }
                }
                sm.gang[0] = G; sm.gang[1] = mi; sm.gang[2] = leader;
            }
            __syncthreads();
            const uint32_t gG = sm.gang[0], gMi = sm.gang[1], gLead = sm.gang[2];
            uint32_t gPhase = 0u; uint32_t *const gCnt = &nd->pad[1], *const gCursor = &nd->bi;   /* bi is free until the leader reports */
            double *const gCarry = scr.carry + (uint64_t) gLead * 2u * RF_GANG_MAX * RF_MAXROWS;
            unsigned long long *keysG = scr.keys + (uint64_t) gLead * scr.keyStride;
            double *wG = scr.w + (uint64_t) gLead * scr.stepStride, *WfG = scr.Wf + (uint64_t) gLead * scr.stepStride,
                   *WrG = scr.Wr + (uint64_t) gLead * scr.stepStride;
            float2 *pairsG = scr.pairs + (uint64_t) gLead * 2 * scr.stepStride;
            uint32_t *srcG = scr.srcPos + (uint64_t) gLead * scr.stepStride, *posTmp = scr.posTmp + (uint64_t) gLead * scr.stepStride;
            const uint32_t begin = sm.begin, n = sm.end - sm.begin;
            const uint32_t srcBuf = sm.srcBuf;
            const uint32_t *list = (srcBuf ? L1 : L0) + I->listOff + begin;
            uint32_t *listDst = (srcBuf ? L0 : L1) + I->listOff + begin;
            const float *Xs = (srcBuf ? XB : XA) + I->xOff + (size_t) begin * nrP;
            float *Xd = (srcBuf ? XA : XB) + I->xOff + (size_t) begin * nrP;
            const bool small = n <= RF_SMALL;
            unsigned long long *keys = small ? sm.keys : keysG;
            double *wA = small ? sm.w : wG, *WfA = small ? sm.Wf : WfG, *WrA = small ? sm.Wr : WrG;
            float2 *pairsF = small ? sm.pairs[0] : pairsG, *pairsR = small ? sm.pairs[1] : pairsG + scr.stepStride;
            if (tid == 0) { tPhase = clock64(); sm.clk[small ? 0 : 1][9]++; sm.clk[small ? 0 : 1][10] += n; sm.clk[small ? 0 : 1][11] += (n <= TV) ? 0u : n; }
            if (n >= 2) {
                do {
#include "refine_split.inl"
                } while (0);
            } else if (tid == 0) sm.err = RF_ERR_SPLIT;
            __syncthreads();
            if (gMi != 0u) continue;                     /* gang members are done; the leader reports (an error is the same on every member) */
            if (tid == 0) {
                uint32_t flags = 0;
                if (sm.err != RF_DONE) flags = sm.err << 28;
                else {
                    float best = sm.rb[0], second = sm.rs[0]; uint32_t bi = sm.ri[0];
                    for (uint32_t i = 1; i < RF_WARPS; i++) {
                        const float b2 = sm.rb[i], s2 = sm.rs[i]; const uint32_t i2 = sm.ri[i];
                        if (b2 < best || (b2 == best && i2 < bi)) { second = fminf(s2, best); best = b2; bi = i2; }
                        else second = fminf(second, b2);
                    }
                    if (bi == 0xffffffffu) flags = (uint32_t) RF_ERR_NOBEST << 28;
                    else {
                        if (isfinite(second) && fabsf(second - best) <= 1e-6f * fabsf(best)) flags |= 1u;
                        nd->bi = bi; nd->pvH = pairsF[bi - 1]; nd->pvT = pairsR[n - 1 - bi];
                        nd->svH = (uint32_t) (keys[0] & 0xffffffffull); nd->svT = (uint32_t) (keys[n - 1] & 0xffffffffull);
                    }
                }
                flags |= min(sm.degenerate, 0xfffffu) << 8;
                nd->flags = flags;
                __threadfence();
                *(volatile uint32_t *) &nd->state = 2u;
                __threadfence();
                atomicSub(mp.outstanding + o, 1u);
                __threadfence();                                                /* result and counter first, then look at the wait node */
                const uint32_t wn = *(volatile uint32_t *) (mp.waitNode + o);
                if (wn == taskNode || wn == MT_ANY) mt_request_control(mp, o);  /* the refinement is waiting for exactly this result */
            }
            RF_TICK(8);
            continue;
        }

        /* ================= control(o) ================= */
        const long long c0 = clock64();
        HeapEntry *heapG = mp.heap + (uint64_t) o * mp.heapCap, *snapG = mp.snap + (uint64_t) o * mp.heapCap;
        uint32_t *singles = mp.singles + (uint64_t) o * mp.nodeCap;
        SplitHeap heap; heap.lo = sm.heap; heap.hi = heapG + RF_HEAP_CAP; heap.cap = RF_HEAP_CAP;
        uint32_t *ilist0 = L0 + I->listOff;
        const uint32_t *ilist1 = L1 + I->listOff;
        if (!I->mtInit) {                                           /* first pass: the initial queue and singletons */
            for (uint32_t i = tid; i < I->heapCount; i += RF_THREADS) {
                const ClusterNode cn = scr.initNodes[I->initNodeOff + i];
                HeapEntry e; e.key = cn.undersamplingVar + cn.integrationVar; e.id = i;
                MtNode m; m.under = cn.undersamplingVar; m.integ = cn.integrationVar; m.begin = cn.begin; m.end = cn.end; m.state = 0; m.bi = 0;
                m.pvH = make_float2(0, 0); m.pvT = make_float2(0, 0); m.svH = m.svT = m.flags = 0; m.pad[0] = m.pad[1] = m.pad[2] = 0;
                nodes[i] = m; heap[i] = e; snapG[i] = e;
            }
            for (uint32_t i = tid; i < I->singleCount; i += RF_THREADS) singles[i] = scr.initSingles[I->initSingleOff + i];
        } else {
            for (uint32_t i = tid; i < min(I->heapCount, (uint32_t) RF_HEAP_CAP); i += RF_THREADS) sm.heap[i] = heapG[i];
        }
        __syncthreads();
        if (tid == 0) {
            sm.heapCount = I->heapCount; sm.nodeCount = I->nodeCount; sm.singleCount = I->singleCount;
            sm.underVar = I->underVar; sm.intVar = I->intVar; sm.bestConstant = I->bestConstant;
            if (!I->mtInit) {
                sm.sHeapCount = I->heapCount; sm.sSingleCount = I->singleCount; sm.sUnder = I->underVar; sm.sInt = I->intVar;
                sm.nearTies = 0; sm.splits = 0; sm.degenerate = 0;
                sm.bestSplits = I->bestSplits;
            } else {
                sm.bestSplits = I->bestSplits;
                sm.sHeapCount = I->sHeapCount; sm.sSingleCount = I->sSingleCount; sm.sUnder = I->sUnder; sm.sInt = I->sInt;
                sm.nearTies = I->nearTies; sm.splits = I->splits; sm.degenerate = I->degenerate;
            }
            sm.err = RF_DONE; sm.done = 0; sm.stop = 0; sm.selCount = 0;
            if (I->mtInit && I->mtFinishing) { sm.err = I->status; sm.done = 1; }    /* converged earlier: only the splits in flight were awaited */
            *(volatile uint32_t *) (mp.waitNode + o) = MT_NONE;
            sm.mtClk[2]++;
        }
        __syncthreads();
        for (; !sm.done;) {                                         /* consume one split per iteration */
            if (tid == 0) {
                sm.snap = 0;
                if (sm.heapCount == 0) sm.done = 1;
                else if (sm.heapCount + 2 > mp.heapCap || sm.nodeCount + 2 > mp.nodeCap || sm.singleCount + 2 > mp.nodeCap) { sm.err = RF_RESUME_HOST; sm.done = 1; }
                else {
                    const uint32_t topId = heap[0].id;
                    MtNode *nd = nodes + topId;
                    bool ready = *(volatile uint32_t *) &nd->state == 2u;
                    if (!ready) {
                        /* no result yet: name the cluster this object waits for, then look once more -- its split either sees the
                         * name (and schedules the next control pass) or finished before, in which case the replay goes on */
                        *(volatile uint32_t *) (mp.waitNode + o) = topId;
                        __threadfence();
                        ready = *(volatile uint32_t *) &nd->state == 2u;
                        if (ready) *(volatile uint32_t *) (mp.waitNode + o) = MT_NONE;
                    }
                    if (!ready) sm.stop = 1;
                    else {
                        heap_pop(heap, sm.heapCount);                                   /* popMulti, 581-587 */
                        MtNode cn;                                                      /* the result may have landed during this pass: read it from L2 */
                        { const uint4 *src4 = reinterpret_cast<const uint4 *>(nd); uint4 *dst4 = reinterpret_cast<uint4 *>(&cn);
                          for (int q4 = 0; q4 < 4; q4++) dst4[q4] = __ldcg(src4 + q4); }
                        sm.underVar -= cn.under; sm.intVar -= cn.integ;
                        const uint32_t begin = cn.begin & 0x7fffffffu, end = cn.end, srcBuf = cn.begin >> 31;
                        const uint32_t err = cn.flags >> 28;
                        sm.degenerate += (cn.flags >> 8) & 0xfffffu;
                        if (err) { sm.err = err; sm.done = 1; }
                        else {
                            sm.splits++;
                            if (cn.flags & 1u) sm.nearTies++;
                            /* addCluster(begin, split) then addCluster(split, end), 549-572 */
                            const uint32_t split = begin + cn.bi;
                            for (int half = 0; half < 2; half++) {
                                const uint32_t b = half ? split : begin, e = half ? end : split;
                                const float2 pv = half ? cn.pvT : cn.pvH;
                                if (e == b + 1) {
                                    const uint32_t v = half ? cn.svT : cn.svH;
                                    singles[sm.singleCount++] = v;
                                    ilist0[b] = v;
                                    if (pv.x != 0) { sm.err = RF_ERR_SINGLETON_VAR; sm.done = 1; }
                                    sm.intVar += pv.y;
                                } else {
                                    MtNode m; m.under = pv.x; m.integ = pv.y; m.begin = b | ((srcBuf ^ 1u) << 31); m.end = e; m.state = 0; m.bi = 0;
                                    m.pvH = make_float2(0, 0); m.pvT = make_float2(0, 0); m.svH = m.svT = m.flags = 0; m.pad[0] = m.pad[1] = m.pad[2] = 0;
                                    nodes[sm.nodeCount] = m;
                                    HeapEntry he; he.key = pv.x + pv.y; he.id = sm.nodeCount++;
                                    heap_push(heap, sm.heapCount, he);
                                    sm.underVar += pv.x; sm.intVar += pv.y;
                                }
                            }
                            const uint32_t numClusters = sm.heapCount + sm.singleCount;
                            if (I->adaptive) {                                              /* refineAdaptively, 436-452 */
                                const float scale = I->numVrlsTotal * I->pixelUndersampling + numClusters;
                                const float curr = scale * (I->tracingVar + sm.underVar + sm.intVar);
                                const float lower = scale * (I->tracingVar + I->unclIntVar);
                                if (!isfinite(curr) || curr <= 0) { sm.err = RF_ERR_CONSTANT; sm.done = 1; }
                                else if (!isfinite(lower) || lower <= 0) { sm.err = RF_ERR_LOWER; sm.done = 1; }
                                else {
                                    if (curr < sm.bestConstant) { sm.snap = 1; sm.bestConstant = curr; sm.bestSplits = I->splitsBase + sm.splits; }
                                    if (lower >= sm.bestConstant || sm.heapCount == 0) sm.done = 1;
                                }
                            } else if (I->fixedSplits) { if (I->splitsBase + sm.splits >= I->fixedSplits || sm.heapCount == 0) sm.done = 1; }   /* second pass, 455-470 */
                            else if (!(numClusters < I->targetClusters && sm.heapCount > 0)) sm.done = 1;     /* refineFixedDepth, 387-399 */
                        }
                    }
                }
            }
            __syncthreads();
            const bool doSnap = sm.snap != 0, isDone = sm.done != 0, isStop = sm.stop != 0;
            if (doSnap) {
                for (uint32_t i = tid; i < sm.heapCount; i += RF_THREADS) snapG[i] = heap[i];
                if (tid == 0) { sm.sHeapCount = sm.heapCount; sm.sSingleCount = sm.singleCount; sm.sUnder = sm.underVar; sm.sInt = sm.intVar; }
            }
            __syncthreads();
            if (isDone || isStop) break;
        }
        if (!sm.done) {
            /* ---- hand out the unsplit clusters with the largest keys ---- */
            float *cand = sm.tile;
            const uint32_t m = min(sm.heapCount, (uint32_t) RF_TILE_FLOATS);
            for (uint32_t i = tid; i < m; i += RF_THREADS) {
                const HeapEntry e = heap[i];
                cand[i] = (*(volatile uint32_t *) &nodes[e.id].state == 0u) ? e.key : -INFINITY;
            }
            if (tid == 0) {
                /* keep MT_K splits in flight; the top of the queue is handed out in any case (it is what the object waits for) */
                const uint32_t inflight = *(volatile uint32_t *) (mp.outstanding + o);
                sm.selCount = 0;
                sm.pick[0] = inflight < mp.inflight ? mp.inflight - inflight : 0u;
                if (sm.pick[0] == 0u && *(volatile uint32_t *) &nodes[heap[0].id].state == 0u) sm.pick[0] = 1u;
            }
            __syncthreads();
            const uint32_t budget = sm.pick[0];
            for (uint32_t round = 0; round < budget; round++) {
                float bk = -INFINITY; uint32_t bidx = 0xffffffffu;
                for (uint32_t i = tid; i < m; i += RF_THREADS) { const float k = cand[i]; if (k > bk) { bk = k; bidx = i; } }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float k2 = __shfl_down_sync(0xffffffffu, bk, off); const uint32_t i2 = __shfl_down_sync(0xffffffffu, bidx, off);
                    if (k2 > bk || (k2 == bk && i2 < bidx)) { bk = k2; bidx = i2; }
                }
                if (lane == 0) { sm.rb[warp] = bk; sm.ri[warp] = bidx; }
                __syncthreads();
                if (tid == 0) {
                    bk = sm.rb[0]; bidx = sm.ri[0];
                    for (uint32_t i = 1; i < RF_WARPS; i++) { const float k2 = sm.rb[i]; const uint32_t i2 = sm.ri[i]; if (k2 > bk || (k2 == bk && i2 < bidx)) { bk = k2; bidx = i2; } }
                    if (bidx == 0xffffffffu) sm.found = 0xffffffffu;
                    else { sm.found = bidx; cand[bidx] = -INFINITY; sm.sel[sm.selCount++] = heap[bidx].id; }
                }
                __syncthreads();
                if (sm.found == 0xffffffffu) break;
            }
            __syncthreads();
            if (tid < sm.selCount) *(volatile uint32_t *) &nodes[sm.sel[tid]].state = 1u;
            __syncthreads();
        }
        if (sm.done) {
            /* converged (or failed): splits handed out ahead of time may still be running and write into this object's lists; the
             * final list is assembled by the pass that finds none in flight */
            if (tid == 0) {
                sm.stop = 0;
                if (*(volatile uint32_t *) (mp.outstanding + o) != 0u) {
                    *(volatile uint32_t *) (mp.waitNode + o) = MT_ANY;
                    __threadfence();
                    if (*(volatile uint32_t *) (mp.outstanding + o) != 0u) sm.stop = 1;
                    else *(volatile uint32_t *) (mp.waitNode + o) = MT_NONE;
                }
            }
            __syncthreads();
        }
        /* ---- write the state back ---- */
        for (uint32_t i = tid; i < min(sm.heapCount, (uint32_t) RF_HEAP_CAP); i += RF_THREADS) heapG[i] = sm.heap[i];
        const bool finished = sm.done != 0 && sm.stop == 0;
        if (finished) {
            /* the list of the sequential algorithm: every leaf's range in the order its parent's sort left (copy 1 -> copy 0) */
            for (uint32_t i = warp; i < sm.heapCount; i += RF_WARPS) {
                const HeapEntry e = heap[i];
                const uint32_t b = nodes[e.id].begin, en = nodes[e.id].end;
                if (b >> 31) for (uint32_t p = (b & 0x7fffffffu) + lane; p < en; p += 32) ilist0[p] = ilist1[p];
            }
            if (tid == 0) {
                sm.begin = atomicAdd(scr.cursors + 1, sm.heapCount + sm.sHeapCount);
                sm.end = atomicAdd(scr.cursors + 2, sm.singleCount);
            }
            __syncthreads();
            for (uint32_t i = tid; i < sm.heapCount; i += RF_THREADS) {
                const MtNode &m = nodes[heap[i].id];
                ClusterNode cn; cn.undersamplingVar = m.under; cn.integrationVar = m.integ; cn.begin = m.begin & 0x7fffffffu; cn.end = m.end;
                scr.outNodes[sm.begin + i] = cn;
            }
            for (uint32_t i = tid; i < sm.sHeapCount; i += RF_THREADS) {
                const MtNode &m = nodes[snapG[i].id];
                ClusterNode cn; cn.undersamplingVar = m.under; cn.integrationVar = m.integ; cn.begin = m.begin & 0x7fffffffu; cn.end = m.end;
                scr.outNodes[sm.begin + sm.heapCount + i] = cn;
            }
            for (uint32_t i = tid; i < sm.singleCount; i += RF_THREADS) scr.outSingles[sm.end + i] = singles[i];
        }
        if (tid == 0) {
            if (finished) { I->outNodeOff = sm.begin; I->outSingleOff = sm.end; }
            if (sm.done) { I->status = sm.err; I->mtFinishing = 1; }
            I->underVar = sm.underVar; I->intVar = sm.intVar; I->bestConstant = sm.bestConstant;
            I->heapCount = sm.heapCount; I->nodeCount = sm.nodeCount; I->singleCount = sm.singleCount;
            I->sHeapCount = sm.sHeapCount; I->sSingleCount = sm.sSingleCount; I->sUnder = sm.sUnder; I->sInt = sm.sInt;
            I->nearTies = sm.nearTies; I->splits = sm.splits; I->degenerate = sm.degenerate; I->bestSplits = sm.bestSplits;
            I->mtInit = 1; I->mtWaves++;
        }
        __syncthreads();
        if (tid == 0) {
            if (finished) { __threadfence(); atomicSub(mp.ctr + 2, 1u); }        /* the control role is never released: no further pass */
            else {
                if (sm.selCount) {
                    __threadfence();
                    atomicAdd(mp.outstanding + o, sm.selCount);                 /* before the tasks can complete */
                    /* clusters for one CTA go out together; a cluster for a gang goes out as G consecutive tickets of its own
                     * (consecutive: the CTAs draw tickets in order, so no later task can be waited for before a gang is complete) */
                    uint32_t *plain = reinterpret_cast<uint32_t *>(sm.chunkEnd), np = 0;   /* (scratch: the pick buffers are free in a control pass) */
                    for (uint32_t q = 0; q < sm.selCount; q++) {
                        const MtNode &sn = nodes[sm.sel[q]];
                        const uint32_t G = I->nr <= RF_MAXROWS ? rf_gang_size(sn.end - (sn.begin & 0x7fffffffu), scr.gangMin, gangCap) : 1u;
                        if (G > 1u) { uint32_t ids[RF_GANG_MAX]; for (uint32_t a = 0; a < G; a++) ids[a] = sm.sel[q]; mt_push(mp, 0u, o, ids, G); }
                        else plain[np++] = sm.sel[q];
                    }
                    if (np) mt_push(mp, 0u, o, plain, np);
                }
                __threadfence();
                if (atomicCAS(mp.ctl + o, 1u, 0u) != 1u) {                      /* a result the object waits for arrived meanwhile: run again */
                    atomicExch(mp.ctl + o, 1u);
                    mt_push(mp, 1u, o, nullptr, 1u);
                }
            }
            sm.mtClk[0] += (unsigned long long) (clock64() - c0);
        }
    }
    __syncthreads();
    if (tid == 0 && mp.clk) {
        unsigned long long *out = mp.clk + (uint64_t) blockIdx.x * 32;
        for (int a = 0; a < 24; a++) out[a] = sm.clk[a / 12][a % 12];
        for (int a = 0; a < 8; a++) out[24 + a] = sm.mtClk[a];
    }
}

} // namespace alvrl
