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
#include "heap_order.h"

namespace alvrl {

#define RF_THREADS 512
#define RF_WARPS (RF_THREADS / 32)
#define RF_MAXROWS 512
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
    /* state, in and out */
    uint32_t rngPos; float underVar, intVar, bestConstant;
    uint32_t heapCount, nodeCount, singleCount;
    uint32_t sHeapCount, sSingleCount; float sUnder, sInt;
    uint32_t nearTies, status, splits, degenerate;
    /* in: the initial queue (nodes in array order) and singletons (insertion order) in the compact init arrays;
     * out: the final queue followed by the best-so-far snapshot (nodes in array order) and the singletons, compacted */
    uint32_t initNodeOff, initSingleOff, outNodeOff, outSingleOff;
    unsigned long long clk[2][12];      /* cycles per phase, [small | large cluster]: pick, direction, stage, project, sort, weights, sweep, pairs, argmin+queue, count */
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
};

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
    uint32_t inst, begin, end, srcBuf, found, pick[2], flags, done, snap, err;
    float u1, u2, norm[3];
    /* refinement state (thread 0) */
    uint32_t rngPos, heapCount, nodeCount, singleCount, sHeapCount, sSingleCount, nearTies, splits, degenerate;
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
                                                    double *__restrict__ Vcol) {
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
        if (r < I.nr) { e = col[r]; vy += (double) e.y / wc; }
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
        const double lw = I->lw;
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
            sm.nearTies = 0; sm.splits = 0; sm.degenerate = 0; sm.err = RF_DONE;
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
                    sm.u1 = alvrl_rng_uniform(key, sm.rngPos++); sm.u2 = alvrl_rng_uniform(key, sm.rngPos++);   /* 597-602 */
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

            /* ---- weightedSample x 2 (597-602, 1534-1580) ---- */
            const uint32_t numChunks = (n + RF_CHUNK - 1) / RF_CHUNK;
            uint32_t idx[2] = {0, 0};
            {
                uint32_t cT = 0;                                    /* chunk whose running sums acc[] holds */
                for (int draw = 0; draw < 2; draw++) {
                    /* running sums: chunkEnd[c] = sum after chunk c.  Draw 1 resumes in the chunk of the first centre. */
                    uint32_t cFrom = 0;
                    if (draw == 1) {
                        const uint32_t l1 = idx[0] - cT * RF_CHUNK, cnt = min((uint32_t) RF_CHUNK, n - cT * RF_CHUNK);
                        if (tid == 0) {
                            float a = l1 ? sm.acc[l1 - 1] : (cT ? sm.chunkEnd[cT - 1] : 0.0f);
                            a += 0.0f; sm.acc[l1] = a;
                            for (uint32_t i = l1 + 1; i < cnt; i++) { a += sm.sw[0][i]; sm.acc[i] = a; }
                            sm.chunkEnd[cT] = a;
                        }
                        cFrom = cT + 1;
                    }
                    if (draw == 0 && numChunks == 1) {
                        if (tid < n) sm.sw[0][tid] = icw[list[tid]];
                        if (tid + RF_THREADS < n) sm.sw[0][tid + RF_THREADS] = icw[list[tid + RF_THREADS]];
                        __syncthreads();
                        if (tid == 0) { float a = 0.0f; for (uint32_t i = 0; i < n; i++) { a += sm.sw[0][i]; sm.acc[i] = a; } sm.chunkEnd[0] = a; }
                    } else if (cFrom < numChunks) {                 /* double-buffered gather / chain over the remaining chunks */
                        __syncthreads();
                        for (uint32_t i = tid; i < min((uint32_t) RF_CHUNK, n - cFrom * RF_CHUNK); i += RF_THREADS) sm.sw[1][i] = icw[list[cFrom * RF_CHUNK + i]];
                        for (uint32_t c = cFrom; c < numChunks; c++) {
                            const uint32_t b = (c - cFrom + 1) & 1;
                            __syncthreads();
                            if (c + 1 < numChunks)
                                for (uint32_t i = tid; i < min((uint32_t) RF_CHUNK, n - (c + 1) * RF_CHUNK); i += RF_THREADS) sm.sw[b ^ 1][i] = icw[list[(c + 1) * RF_CHUNK + i]];
                            if (tid == 0) {
                                float a = c ? sm.chunkEnd[c - 1] : 0.0f;
                                const uint32_t cnt = min((uint32_t) RF_CHUNK, n - c * RF_CHUNK);
                                for (uint32_t i = 0; i < cnt; i++) a += sm.sw[b][i];
                                sm.chunkEnd[c] = a;
                            }
                        }
                    }
                    __syncthreads();
                    const float weightSum = sm.chunkEnd[numChunks - 1];
                    const float alpha = (draw == 0 ? sm.u1 : sm.u2) * weightSum;
                    if (tid == 0 && !(weightSum > 0)) sm.flags |= 2u;
                    uint32_t cNew = 0;
                    if (numChunks > 1) {                            /* first chunk whose end sum reaches alpha */
                        for (uint32_t c = tid; c < numChunks; c += RF_THREADS) if (sm.chunkEnd[c] >= alpha) atomicMin(&sm.found, c);
                        __syncthreads();
                        cNew = sm.found == 0xffffffffu ? 0u : sm.found;
                        __syncthreads();
                        if (tid == 0) sm.found = 0xffffffffu;
                        if (!(draw == 1 && cNew == cT)) {           /* re-chain that chunk, keeping its running sums */
                            const uint32_t cnt = min((uint32_t) RF_CHUNK, n - cNew * RF_CHUNK);
                            for (uint32_t i = tid; i < cnt; i += RF_THREADS) sm.sw[0][i] = icw[list[cNew * RF_CHUNK + i]];
                            __syncthreads();
                            if (tid == 0) { float a = cNew ? sm.chunkEnd[cNew - 1] : 0.0f; for (uint32_t i = 0; i < cnt; i++) { a += sm.sw[0][i]; sm.acc[i] = a; } }
                        }
                        __syncthreads();
                    }
                    cT = cNew;
                    {
                        const uint32_t cnt = min((uint32_t) RF_CHUNK, n - cT * RF_CHUNK);
                        for (uint32_t i = tid; i < cnt; i += RF_THREADS) if (sm.acc[i] >= alpha) atomicMin(&sm.found, i);
                    }
                    __syncthreads();
                    if (sm.found == 0xffffffffu) { idx[draw] = 0; if (tid == 0) sm.flags |= 2u; }
                    else idx[draw] = cT * RF_CHUNK + sm.found;
                    __syncthreads();
                    if (tid == 0) sm.found = 0xffffffffu;
                }
            }
            __syncthreads();
            if (sm.flags & 2u) { if (tid == 0) sm.err = RF_ERR_WEIGHTS; break; }
            RF_TICK(0);

            /* ---- direction (604-623) ---- */
            for (uint32_t r = tid; r < nr; r += RF_THREADS) { sm.c1[r] = Xs[(size_t) idx[0] * nrP + r]; sm.c2[r] = Xs[(size_t) idx[1] * nrP + r]; }
            __syncthreads();
            if (tid == 0) { float a = 0; for (uint32_t r = 0; r < nr; r++) a += fabsf(sm.c1[r]) * fabsf(sm.c1[r]); sm.norm[0] = sqrtf(a); }
            else if (tid == 32) { float a = 0; for (uint32_t r = 0; r < nr; r++) a += fabsf(sm.c2[r]) * fabsf(sm.c2[r]); sm.norm[1] = sqrtf(a); }
            else if (tid == 64) { float a = 0; for (uint32_t r = 0; r < nr; r++) { const float d = sm.c2[r] - sm.c1[r]; a += fabsf(d) * fabsf(d); } sm.norm[2] = sqrtf(a); }
            __syncthreads();
            if (sm.norm[0] != 0 && sm.norm[1] != 0 && sm.norm[2] != 0) {
                const float dl = sm.norm[2];
                for (uint32_t r = tid; r < nrP; r += RF_THREADS) sm.sdir[r] = r < nr ? (sm.c2[r] - sm.c1[r]) / dl : 0.0f;
            } else {
                /* degenerate centres: direction uniform on the n-sphere, warp::squareToStdNormal(next2D()).x per row (616-622);
                 * log and cos evaluated in double and rounded (pinned transcendental, same on the host path and in the oracle) */
                for (;;) {
                    const uint32_t base = sm.rngPos;
                    for (uint32_t r = tid; r < nr; r += RF_THREADS) {
                        const float s1 = alvrl_rng_uniform(key, base + 2 * r), s2 = alvrl_rng_uniform(key, base + 2 * r + 1);
                        const float rr = sqrtf(-2 * (float) log((double) (1 - s1))), phi = (float) (2 * M_PI * s2);
                        sm.c1[r] = (float) cos((double) phi) * rr;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        sm.rngPos = base + 2 * nr; sm.degenerate++;
                        float a = 0; for (uint32_t r = 0; r < nr; r++) a += fabsf(sm.c1[r]) * fabsf(sm.c1[r]);
                        sm.norm[2] = sqrtf(a);
                    }
                    __syncthreads();
                    if (sm.norm[2] != 0) break;
                }
                const float dl = sm.norm[2];
                for (uint32_t r = tid; r < nrP; r += RF_THREADS) sm.sdir[r] = r < nr ? sm.c1[r] / dl : 0.0f;
            }
            __syncthreads();

            RF_TICK(1);
            /* ---- projections (625-640).  The columns are staged through the tile in chunks with asynchronous 16-byte copies (the
             *      whole chunk is in flight at once), then thread = column sums sequentially in fp32 in row order out of shared
             *      memory (the zero padding of columns and direction adds exact zeros).  A local matrix that fits the tile in one
             *      chunk stays there for the variance sweep. ---- */
            const bool fits = n <= TV;
            if (fits) {
                for (uint32_t i = tid; i < n * nq; i += RF_THREADS) {
                    const uint32_t c = i / nq, q = i - c * nq;
                    rf_cp_async16(sm.tile + c * tS + 4 * (q ^ (c & 7u)), Xs + (size_t) c * nrP + 4 * q);
                }
                rf_cp_commit(); rf_cp_wait<0>();
                __syncthreads();
            }
            RF_TICK(2);
            for (uint32_t c = tid; c < n; c += RF_THREADS) {
                float s = 0, pj = 0;
                if (fits) {
                    const float *x = sm.tile + c * tS;
                    const uint32_t sw = c & 7u;
#pragma unroll 2
                    for (uint32_t q = 0; q < nq; q++) {
                        const float4 e = *reinterpret_cast<const float4 *>(x + 4 * (q ^ sw));
                        float a;
                        a = fabsf(e.x); s += a * a; a = fabsf(e.y); s += a * a; a = fabsf(e.z); s += a * a; a = fabsf(e.w); s += a * a;
                    }
                    const float len = sqrtf(s);
                    if (len != 0) {
#pragma unroll 2
                        for (uint32_t q = 0; q < nq; q++) {
                            const float4 e = *reinterpret_cast<const float4 *>(x + 4 * (q ^ sw));
                            const float4 d = *reinterpret_cast<const float4 *>(sm.sdir + 4 * q);
                            pj += d.x * (e.x / len); pj += d.y * (e.y / len); pj += d.z * (e.z / len); pj += d.w * (e.w / len);
                        }
                    }
                } else {
                    /* too large for the tile: every thread streams its own (contiguous) column from L2, eight loads in flight */
                    const float4 *col4 = reinterpret_cast<const float4 *>(Xs + (size_t) c * nrP);
                    for (uint32_t q0 = 0; q0 < nq; q0 += 8) {
                        float4 e[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) e[u] = (q0 + u < nq) ? col4[q0 + u] : make_float4(0, 0, 0, 0);
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            float a;
                            a = fabsf(e[u].x); s += a * a; a = fabsf(e[u].y); s += a * a; a = fabsf(e[u].z); s += a * a; a = fabsf(e[u].w); s += a * a;
                        }
                    }
                    const float len = sqrtf(s);
                    if (len != 0) {
                        for (uint32_t q0 = 0; q0 < nq; q0 += 8) {
                            float4 e[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) e[u] = (q0 + u < nq) ? col4[q0 + u] : make_float4(0, 0, 0, 0);
#pragma unroll
                            for (int u = 0; u < 8; u++) {
                                if (q0 + u < nq) {
                                    const float4 d = *reinterpret_cast<const float4 *>(sm.sdir + 4 * (q0 + u));
                                    pj += d.x * (e[u].x / len); pj += d.y * (e[u].y / len); pj += d.z * (e[u].z / len); pj += d.w * (e[u].w / len);
                                }
                            }
                        }
                    }
                }
                const float q = pj + 0.0f;                                      /* -0.0 and +0.0 compare equal in the pair order */
                uint32_t b = __float_as_uint(q);
                b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                const uint32_t vid = list[c];
                keys[c] = ((unsigned long long) b << 32) | vid;
                if (small) sm.pos[c] = (uint16_t) c; else posTmp[vid] = c;
            }
            /* ---- std::sort of (projection, vrl) pairs (641): bitonic network on the unique keys, always out of shared memory:
             *      up to RF_SORT_BLOCK keys in one piece (the tile is free when the keys live in global memory), more as
             *      block-local passes plus global steps for the strides that span blocks ---- */
            uint32_t m = 2; while (m < n) m <<= 1;
            for (uint32_t i = n + tid; i < m; i += RF_THREADS) { keys[i] = ~0ull; if (small) sm.pos[i] = 0; }
            __syncthreads();
            RF_TICK(3);
            if (small) {
                for (uint32_t k = 2; k <= m; k <<= 1) rf_sort_steps(sm.keys, m, 0, k, k >> 1, sm.pos);
            } else {
                unsigned long long *sk = reinterpret_cast<unsigned long long *>(sm.tile);
                const uint32_t blkLen = min(m, (uint32_t) RF_SORT_BLOCK);
                for (uint32_t blk = 0; blk < m; blk += blkLen) {                /* every block sorted (direction by global index) */
                    for (uint32_t i = tid; i < blkLen; i += RF_THREADS) sk[i] = keys[blk + i];
                    __syncthreads();
                    for (uint32_t k = 2; k <= blkLen; k <<= 1) rf_sort_steps(sk, blkLen, blk, k, k >> 1, nullptr);
                    for (uint32_t i = tid; i < blkLen; i += RF_THREADS) keys[blk + i] = sk[i];
                    __syncthreads();
                }
                for (uint32_t k = 2 * blkLen; k <= m; k <<= 1) {                /* merges across blocks */
                    for (uint32_t j = k >> 1; j >= blkLen; j >>= 1) {
                        for (uint32_t t = tid; t < (m >> 1); t += RF_THREADS) {
                            const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                            const unsigned long long a = keys[lo], b = keys[hi];
                            if ((a > b) == ((lo & k) == 0)) { keys[lo] = b; keys[hi] = a; }
                        }
                        __syncthreads();
                    }
                    for (uint32_t blk = 0; blk < m; blk += blkLen) {
                        for (uint32_t i = tid; i < blkLen; i += RF_THREADS) sk[i] = keys[blk + i];
                        __syncthreads();
                        rf_sort_steps(sk, blkLen, blk, k, blkLen >> 1, nullptr);
                        for (uint32_t i = tid; i < blkLen; i += RF_THREADS) keys[blk + i] = sk[i];
                        __syncthreads();
                    }
                }
            }
            RF_TICK(4);
            /* ---- sorted list, weights and prefix weights (forward and reverse order) ---- */
            {
                double cW[2] = {0, 0};
                for (uint32_t k0 = 0; k0 < n; k0 += RF_THREADS) {
                    const uint32_t k = k0 + tid, cnt = min((uint32_t) RF_THREADS, n - k0);
                    double v[2] = {0, 0}, tot[2];
                    if (k < n) {
                        const uint32_t vf = (uint32_t) (keys[k] & 0xffffffffull), vr = (uint32_t) (keys[n - 1 - k] & 0xffffffffull);
                        list[k] = vf;
                        if (!small) srcG[k] = posTmp[vf];
                        v[0] = (double) icw[vf]; v[1] = (double) icw[vr];
                        wA[k] = v[0];
                    }
                    rf_scan<2, false>(v, sm.scan[(scanIt++) & 1], (cnt + 31) / 32, n > RF_THREADS, tot);
                    if (k < n) { WfA[k] = cW[0] + v[0]; WrA[k] = cW[1] + v[1]; }
                    cW[0] += tot[0]; cW[1] += tot[1];
                }
            }
            /* the sorted copy: from the tile when the local matrix is resident, else gathered column by column (the variance
             * sweeps then stream it with bulk copies) */
            if (fits) {
                for (uint32_t i = tid; i < n * nq; i += RF_THREADS) {
                    const uint32_t k = i / nq, q = i - k * nq, pc = sm.pos[k];
                    *reinterpret_cast<float4 *>(Xd + (size_t) k * nrP + 4 * q) = *reinterpret_cast<const float4 *>(sm.tile + pc * tS + 4 * (q ^ (pc & 7u)));
                }
            } else {
                for (uint32_t k4 = warp * 4; k4 < n; k4 += RF_WARPS * 4) {     /* warp = column, four columns in flight */
                    const float4 *src[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { const uint32_t k = min(k4 + u, n - 1); src[u] = reinterpret_cast<const float4 *>(Xs + (size_t) (small ? (uint32_t) sm.pos[k] : srcG[k]) * nrP); }
                    for (uint32_t q = lane; q < nq; q += 32) {
                        float4 e[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) e[u] = src[u][q];
#pragma unroll
                        for (int u = 0; u < 4; u++) if (k4 + u < n) reinterpret_cast<float4 *>(Xd + (size_t) (k4 + u) * nrP)[q] = e[u];
                    }
                }
                __threadfence_block();
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            __syncthreads();
            RF_TICK(5);
            /* ---- calculateClusterVariance (1058-1120): the forward sweep on threads 0..255 and the reverse sweep on threads
             *      256..511, thread = row, sequential over the sorted steps (S_r is a running sum, accesses are contiguous across
             *      rows).  A local matrix that is not resident streams from the sorted copy through a two-stage ring in the tile,
             *      one bulk copy (TMA) per KC steps, the next chunk in flight while the current one is computed.  The per-step sums
             *      over rows B_k = sum_r (w_k S_r(k-1) - W_{k-1} x_r(k))^2 are reduced 16 steps at a time by a transposing
             *      shuffle reduction and across warps through a double-buffered stage ---- */
            {
                const uint32_t half = tid >> 8, hr = tid & 255u, hw = hr >> 5;
                double *Bh = reinterpret_cast<double *>(half ? pairsR : pairsF);       /* B_k lives where pairs[k] goes afterwards */
                const double *WA = half ? WrA : WfA;
                double (*part)[32][8] = reinterpret_cast<double (*)[32][8]>(&sm.sw[0][0]) + half * 2;   /* [buffer][step][warp] */
                /* steps per stage (>= 12 for nr <= 512); the step loop runs in groups of 16, so 17..31 steps would pay a
                 * second, mostly empty group per stage */
                uint32_t KC = min(32u, (uint32_t) (RF_TILE_FLOATS / 4) / nrP);
                if (KC > 16u && KC < 32u) KC = 16u;
                float *ring = sm.tile + half * (RF_TILE_FLOATS / 2);
                const uint32_t nch = (n + KC - 1) / KC;
                auto issue = [&](uint32_t c) {                                          /* the columns of chunk c -> ring stage c & 1 */
                    const uint32_t k0 = c * KC, cnt = min(KC, n - k0);
                    if (hr == 0) {
                        asm volatile("fence.proxy.async;" ::: "memory");
                        const uint32_t bytes = cnt * nrP * (uint32_t) sizeof(float);
                        rf_mbar_expect_tx(&sm.mbar[half][c & 1u], bytes);
                        rf_bulk_load(ring + (c & 1u) * KC * nrP, Xd + (size_t) (half ? n - k0 - cnt : k0) * nrP, bytes, &sm.mbar[half][c & 1u]);
                    }
                    if (!small && hr < cnt) {                                           /* w_k, W_{k-1} live in global memory */
                        const uint32_t k = k0 + hr, sp = half ? n - 1 - k : k;
                        rf_cp_async8(&sm.stepW[half][c & 1u][hr], wA + sp);
                        if (k) rf_cp_async8(&sm.stepWp[half][c & 1u][hr], WA + k - 1); else sm.stepWp[half][c & 1u][hr] = 0.0;
                    }
                    rf_cp_commit();
                };
                /* thread = rows hr and hr + RS (nr <= 512): the rows are folded onto the fewest warps, RS = roundup(nr / 2, 32),
                 * so that both row slots of a thread carry a row -- the sweep is issue-bound, and a warp whose second slot is
                 * empty costs as many issue slots as a full one */
                const uint32_t RS = min(256u, (((nr + 1u) >> 1) + 31u) & ~31u);
                const uint32_t rA = hr, rB = hr + RS;
                const bool actA = hr < RS && rA < nr, actB = hr < RS && rB < nr;
                const uint32_t nw = RS >> 5;
                double SA = 0, SB = 0;
                uint32_t buf = 0;
                if (!fits) issue(0);
                for (uint32_t c = 0; c < nch; c++) {
                    const uint32_t k0 = c * KC, cnt = min(KC, n - k0);
                    if (!fits) {
                        if (c + 1 < nch) { issue(c + 1); rf_cp_wait<1>(); } else rf_cp_wait<0>();
                        rf_mbar_wait(&sm.mbar[half][c & 1u], (ringPhase >> (c & 1u)) & 1u);
                        ringPhase ^= 1u << (c & 1u);
                        if (!small) asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
                    }
                    if (hw < nw) {
                        const float *stage = ring + (c & 1u) * KC * nrP;
                        for (uint32_t g = 0; g < cnt; g += 16) {
                            float xa[16], xb[16]; double b[16];
#pragma unroll
                            for (int u = 0; u < 16; u++) {
                                xa[u] = 0; xb[u] = 0;
                                if (g + u < cnt) {
                                    if (fits) {
                                        const uint32_t k = k0 + g + u, pc = sm.pos[half ? n - 1 - k : k];
                                        const float *colp = sm.tile + pc * tS;
                                        if (actA) xa[u] = colp[4 * ((rA >> 2) ^ (pc & 7u)) + (rA & 3u)];
                                        if (actB) xb[u] = colp[4 * ((rB >> 2) ^ (pc & 7u)) + (rB & 3u)];
                                    } else {
                                        const float *colp = stage + (half ? cnt - 1 - (g + u) : g + u) * nrP;
                                        if (actA) xa[u] = colp[rA];
                                        if (actB) xb[u] = colp[rB];
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 16; u++) {
                                b[u] = 0;
                                if (g + u < cnt) {
                                    const uint32_t k = k0 + g + u, sp = half ? n - 1 - k : k;
                                    const double wk = small ? wA[sp] : sm.stepW[half][c & 1u][g + u];
                                    const double Wp = small ? (k ? WA[k - 1] : 0.0) : sm.stepWp[half][c & 1u][g + u];
                                    const double xad = (double) xa[u], xbd = (double) xb[u];
                                    const double ta = wk * SA - Wp * xad, tb = wk * SB - Wp * xbd;
                                    SA += xad; SB += xbd;
                                    b[u] = ta * ta + tb * tb;
                                }
                            }
                            rf_reduce16(b, lane);
                            if (lane < 16 && g + lane < cnt) part[buf][g + lane][hw] = b[0];
                        }
                    }
                    asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
                    if (hr < cnt) {
                        double sum = 0;
                        for (uint32_t w8 = 0; w8 < nw; w8++) sum += part[buf][hr][w8];
                        Bh[k0 + hr] = sum;
                    }
                    buf ^= 1;
                }
            }
            __syncthreads();
            RF_TICK(6);
            /* prefix pairs (1098-1106), thread = step: first = lw W_k Q_k, second = lw W_k SV_k */
            for (uint32_t dir = 0; dir < 2; dir++) {
                const double *WA = dir ? WrA : WfA;
                float2 *pairs = dir ? pairsR : pairsF;
                const double *Bh = reinterpret_cast<const double *>(pairs);
                double cQ = 0, cV = 0;
                for (uint32_t k0 = 0; k0 < n; k0 += RF_THREADS) {
                    const uint32_t cnt = min((uint32_t) RF_THREADS, n - k0), k = k0 + tid;
                    const bool on = tid < cnt;
                    double v2[2] = {0, 0}, tot2[2], Wk = 1.0;
                    if (on) {
                        const uint32_t sp = dir ? n - 1 - k : k;
                        const double wk = wA[sp];
                        Wk = WA[k];
                        if (k) { const double Wp = WA[k - 1]; v2[0] = (1.0 / wk + 1.0 / Wp) * Bh[k] / (Wk * Wk); }
                        v2[1] = Vi[(uint32_t) (keys[sp] & 0xffffffffull)];
                    }
                    rf_scan<2, false>(v2, sm.scan[(scanIt++) & 1], (cnt + 31) / 32, n > RF_THREADS, tot2);
                    if (on) pairs[k] = make_float2(k == 0 ? 0.0f : (float) (lw * (Wk * (cQ + v2[0]))), (float) (lw * ((cV + v2[1]) * Wk)));
                    cQ += tot2[0]; cV += tot2[1];
                }
            }
            __syncthreads();
            RF_TICK(7);
            /* ---- first minimum of head + tail variance (664-675) ---- */
            float best = INFINITY, second = INFINITY; uint32_t bi = 0xffffffffu;
            for (uint32_t k = 1 + tid; k < n; k += RF_THREADS) {
                const float2 h = pairsF[k - 1], tl = pairsR[n - 1 - k];
                const float v = h.x + h.y + tl.x + tl.y;
                if (v < best) { second = best; best = v; bi = k; }
                else if (v < second) second = v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float b2 = __shfl_down_sync(0xffffffffu, best, o), s2 = __shfl_down_sync(0xffffffffu, second, o);
                const uint32_t i2 = __shfl_down_sync(0xffffffffu, bi, o);
                if (b2 < best || (b2 == best && i2 < bi)) { second = fminf(s2, best); best = b2; bi = i2; }
                else second = fminf(second, b2);
            }
            if (lane == 0) { sm.rb[warp] = best; sm.rs[warp] = second; sm.ri[warp] = bi; }
            __syncthreads();
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
                            if (curr < sm.bestConstant) { sm.snap = 1; sm.bestConstant = curr; }
                            if (lower >= sm.bestConstant || sm.heapCount == 0) sm.done = 1;
                        }
                    } else if (!(numClusters < I->targetClusters && sm.heapCount > 0)) sm.done = 1;     /* refineFixedDepth, 387-399 */
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
            I->nearTies = sm.nearTies; I->status = sm.err; I->splits = sm.splits; I->degenerate = sm.degenerate;
        }
    }
}

} // namespace alvrl
