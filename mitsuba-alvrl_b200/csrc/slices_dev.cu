/*
 * slices_dev.cu -- Preprocessor::getSlices / getSlicesPQ (src/integrators/vrl/Preprocessor.cpp:1200-1227,1349-1418) with the
 * per-node work on the device; replaces the sequential host loop of slices.h on the hot path.
 *
 * What the reference does per popped node: a Hoare partition of the node's gather points around the split plane (1368-1393)
 * and the 6-D extrema of the two halves (SliceNode ctor, 1301-1339).  The heap (a std::vector kept with std::push_heap /
 * std::pop_heap, whose array order numbers the slices, 1400-1417) is a hundred entries: it stays on the host.
 *
 * The Hoare partition's outcome is a pure function of the flags L[p] = "component dim of record p is larger than the split"
 * (isLarger, 1420-1430): with nS records not larger, the left part becomes [lo, lo + nS); every record that already sits on
 * its side stays where it is, and the t-th misplaced larger record from the left swaps with the t-th misplaced not-larger
 * record from the right -- exactly the pairs the two scanning indices meet.  So one count, one exclusive scan of the flags
 * and one swap kernel reproduce the reference's record order bit for bit (slices.h keeps the sequential loop; the tests compare
 * both with the oracle).  Degenerate nodes (all records on one side: a split plane that rounds onto an extremum) hand the
 * whole build back to the host loop, which follows the reference's sentinel behaviour literally.
 * The extrema are min / max reductions: order independent, exact.
 *
 * Misses (non-finite gather points) are moved to the front by the reference's own sequential swap loop (1206-1221); it
 * touches 2 records per miss, so the host replays it on the (sorted) list of miss positions and the device applies the
 * resulting sparse permutation.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>
#include "context.h"

namespace alvrl {

namespace {

struct SRec { float4 a, b; };            /* a = pos.xyz, dir.x; b = dir.y, dir.z, bits(pixel index), 0 */

__device__ __forceinline__ float srec_comp(const SRec &r, int dim) {
    switch (dim) { case 0: return r.a.x; case 1: return r.a.y; case 2: return r.a.z; case 3: return r.a.w; case 4: return r.b.x; default: return r.b.y; }
}

__global__ void k_slice_make_recs(const float *__restrict__ pos, const float *__restrict__ dir, uint32_t P, SRec *__restrict__ rec,
                                  uint32_t *__restrict__ missList, uint32_t *__restrict__ missCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    SRec r;
    r.a = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i]);
    r.b = make_float4(dir[3 * i + 1], dir[3 * i + 2], __uint_as_float(i), 0.0f);
    rec[i] = r;
    if (!(isfinite(r.a.x) && isfinite(r.a.y) && isfinite(r.a.z))) missList[atomicAdd(missCount, 1u)] = i;
}
/* the sparse permutation of the miss loop: rec[position] = record of the original index */
__global__ void k_slice_patch(const float *__restrict__ pos, const float *__restrict__ dir, const uint2 *__restrict__ patch, uint32_t n, SRec *__restrict__ rec) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t p = patch[t].x, i = patch[t].y;
    SRec r;
    r.a = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i]);
    r.b = make_float4(dir[3 * i + 1], dir[3 * i + 2], __uint_as_float(i), 0.0f);
    rec[p] = r;
}

struct NodeRes {                         /* what the host needs back per iteration */
    uint32_t nS, k, pad0, pad1;
    float mn[2][6], mx[2][6];
};

/* flags S[p] = !(v > split) and their count per block */
__global__ void k_slice_flags(const SRec *__restrict__ rec, uint32_t lo, uint32_t n, int dim, float split, uint8_t *__restrict__ flagS,
                              uint32_t *__restrict__ blockCnt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    int s = 0;
    if (t < n) { s = !(srec_comp(rec[lo + t], dim) > split); flagS[t] = (uint8_t) s; }
    const int c = __syncthreads_count(s);
    if (threadIdx.x == 0) blockCnt[blockIdx.x] = (uint32_t) c;
}
/* exclusive scan of the block counts (one block); total -> res->nS; res->k = 0 */
__global__ void k_slice_scan(uint32_t *__restrict__ blockCnt, uint32_t nb, NodeRes *res) {
    __shared__ uint32_t warpSum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? blockCnt[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= (uint32_t) o) x += y; }
        if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? warpSum[threadIdx.x] : 0u;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= (uint32_t) o) w += y; }
            warpSum[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t wOff = (threadIdx.x >> 5) ? warpSum[(threadIdx.x >> 5) - 1] : 0u;
        if (i < nb) blockCnt[i] = carry + wOff + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warpSum[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) { res->nS = carry; res->k = 0; }
}
/* positions of the misplaced records: posA[t] = t-th larger record of the left part (from the left), posB[t] = t-th not-larger
 * record of the right part (from the right) */
__global__ void k_slice_ranks(const uint8_t *__restrict__ flagS, const uint32_t *__restrict__ blockOff, uint32_t lo, uint32_t n, NodeRes *res,
                              uint32_t *__restrict__ posA, uint32_t *__restrict__ posB) {
    __shared__ uint32_t warpSum[8];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nS = res->nS;
    const int s = t < n ? flagS[t] : 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, s);
    if (lane == 0) warpSum[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = blockOff[blockIdx.x];
    for (uint32_t w = 0; w < warp; w++) off += warpSum[w];
    const uint32_t cS = off + __popc(bal & ((1u << lane) - 1u));            /* not-larger records before this one */
    int misplacedLeft = 0;
    if (t < n) {
        if (t < nS) { if (!s) { posA[t - cS] = lo + t; misplacedLeft = 1; } }
        else if (s) posB[nS - cS - 1u] = lo + t;
    }
    const int c = __syncthreads_count(misplacedLeft);
    if (threadIdx.x == 0 && c) atomicAdd(&res->k, (uint32_t) c);
}
__global__ void k_slice_swap(SRec *__restrict__ rec, const uint32_t *__restrict__ posA, const uint32_t *__restrict__ posB, const NodeRes *res, uint32_t maxK) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= res->k || t >= maxK) return;
    const uint32_t a = posA[t], b = posB[t];
    const SRec ra = rec[a], rb = rec[b];
    rec[a] = rb; rec[b] = ra;
}
/* 6-D extrema of the two children [lo, lo + nS) and [lo + nS, lo + n): block partials, then one block */
__global__ void k_slice_bbox(const SRec *__restrict__ rec, uint32_t lo, uint32_t n, const NodeRes *res, float *__restrict__ part /* [nb][2][12] */) {
    __shared__ float sh[8][2][12];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nS = res->nS;
    const float inf = INFINITY;
    float mn[2][6], mx[2][6];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int c = 0; c < 6; c++) { mn[h][c] = inf; mx[h][c] = -inf; }
    if (t < n) {
        const SRec r = rec[lo + t];
        const float v[6] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y};
        const int h = t < nS ? 0 : 1;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            if (h == 0) { mn[0][c] = v[c]; mx[0][c] = v[c]; } else { mn[1][c] = v[c]; mx[1][c] = v[c]; }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int c = 0; c < 6; c++) {
            float a = mn[h][c], b = mx[h][c];
            for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
            if (lane == 0) { sh[warp][h][c] = a; sh[warp][h][6 + c] = b; }
        }
    __syncthreads();
    if (threadIdx.x < 24) {
        const int h = threadIdx.x / 12, c = threadIdx.x % 12;
        float a = sh[0][h][c];
        for (uint32_t w = 1; w < (blockDim.x >> 5); w++) a = c < 6 ? fminf(a, sh[w][h][c]) : fmaxf(a, sh[w][h][c]);
        part[(size_t) blockIdx.x * 24 + threadIdx.x] = a;
    }
}
__global__ void k_slice_bbox_final(const float *__restrict__ part, uint32_t nb, NodeRes *res) {
    __shared__ float sh[8][24];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < 24; q++) {
        const bool isMin = (q % 12) < 6;
        float a = isMin ? INFINITY : -INFINITY;
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) { const float v = part[(size_t) i * 24 + q]; a = isMin ? fminf(a, v) : fmaxf(a, v); }
        for (int o = 16; o > 0; o >>= 1) { const float y = __shfl_xor_sync(0xffffffffu, a, o); a = isMin ? fminf(a, y) : fmaxf(a, y); }
        if (lane == 0) sh[warp][q] = a;
    }
    __syncthreads();
    if (threadIdx.x < 24) {
        const int q = threadIdx.x; const bool isMin = (q % 12) < 6;
        float a = sh[0][q];
        for (uint32_t w = 1; w < (blockDim.x >> 5); w++) a = isMin ? fminf(a, sh[w][q]) : fmaxf(a, sh[w][q]);
        const int h = q / 12, c = q % 12;
        if (c < 6) res->mn[h][c] = a; else res->mx[h][c - 6] = a;
    }
}
/* extrema of one range (the root) */
__global__ void k_slice_bbox_root(const SRec *__restrict__ rec, uint32_t lo, uint32_t n, NodeRes *res) {
    /* one block: the root is scanned once per frame */
    __shared__ float sh[32][12];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn[6], mx[6];
    for (int c = 0; c < 6; c++) { mn[c] = INFINITY; mx[c] = -INFINITY; }
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
        const SRec r = rec[lo + t];
        const float v[6] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y};
        for (int c = 0; c < 6; c++) { mn[c] = fminf(mn[c], v[c]); mx[c] = fmaxf(mx[c], v[c]); }
    }
    for (int c = 0; c < 6; c++) {
        float a = mn[c], b = mx[c];
        for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
        if (lane == 0) { sh[warp][c] = a; sh[warp][6 + c] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int c = threadIdx.x;
        float a = sh[0][c];
        for (uint32_t w = 1; w < (blockDim.x >> 5); w++) a = c < 6 ? fminf(a, sh[w][c]) : fmaxf(a, sh[w][c]);
        if (c < 6) res->mn[0][c] = a; else res->mx[0][c - 6] = a;
    }
}
/* pixel -> slice and the permuted pixel ids: slice of position k by binary search over the sorted range starts */
__global__ void k_slice_finish(const SRec *__restrict__ rec, uint32_t first, uint32_t P, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ ids,
                               uint32_t S, uint32_t *__restrict__ recIdx, uint32_t *__restrict__ pixelToSlice) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const uint32_t px = __float_as_uint(rec[k].b.z);
    recIdx[k] = px;
    if (k < first) { pixelToSlice[px] = ALVRL_NO_SLICE; return; }
    uint32_t a = 0, b = S;                          /* last start <= k */
    while (b - a > 1) { const uint32_t m = (a + b) >> 1; if (starts[m] <= k) a = m; else b = m; }
    pixelToSlice[px] = ids[a];
}

struct HNode { uint32_t lo, hi; float diag; unsigned char dim; float split; float centroid[6]; bool operator<(const HNode &o) const { return diag < o.diag; } };

/* findSplitPoint, 1451-1487 (same code as slices.h) */
void splitPoint(const float mn[3], const float mx[3], unsigned char &dim, float &split, float &extent) {
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (dx == 0 && dy == 0 && dz == 0) { extent = 0; dim = 0; split = std::numeric_limits<float>::quiet_NaN(); return; }
    int c;
    if (dx > dy) c = (dx > dz) ? 0 : 2; else c = (dy > dz) ? 1 : 2;
    const float d = c == 0 ? dx : (c == 1 ? dy : dz);
    dim = (unsigned char) c;
    split = (float) ((double) mn[c] + 0.5 * (double) d);
    extent = d;
}
/* SliceNode ctor from the extrema, 1301-1339 */
HNode makeNode(uint32_t lo, uint32_t hi, const float mn[6], const float mx[6]) {
    HNode n; n.lo = lo; n.hi = hi;
    for (int c = 0; c < 6; c++) n.centroid[c] = std::numeric_limits<float>::quiet_NaN();
    if (lo + 1 == hi) { n.diag = 0; n.dim = 0; n.split = std::numeric_limits<float>::quiet_NaN(); return n; }
    float dp, dd;
    { const float a = mn[0] - mx[0], b = mn[1] - mx[1], c = mn[2] - mx[2]; dp = a * a + b * b + c * c; }
    { const float a = mn[3] - mx[3], b = mn[4] - mx[4], c = mn[5] - mx[5]; dd = a * a + b * b + c * c; }
    n.diag = std::sqrt(dp + dd);
    unsigned char dimP, dimD; float splitP, splitD, extP, extD;
    splitPoint(mn, mx, dimP, splitP, extP);
    splitPoint(mn + 3, mx + 3, dimD, splitD, extD);
    if (extP > extD) { n.dim = dimP; n.split = splitP; } else { n.dim = 3 + dimD; n.split = splitD; }
    for (int c = 0; c < 6; c++) n.centroid[c] = mn[c] + 0.5f * (mx[c] - mn[c]);       /* box midpoints, 1337-1338 */
    return n;
}

} // namespace

/* returns false when a degenerate node was met (the caller falls back to the host loop of slices.h) */
bool build_slices_device(alvrl_ctx *c, const float *dPos, const float *dDir, uint32_t P, uint32_t targetNumSlices) {
    cudaStream_t st = c->stream;
    DevBuf<SRec> dRec; dRec.alloc(P);
    DevBuf<uint32_t> dMiss, dCnt, dBlockCnt, dPosA, dPosB;
    DevBuf<uint8_t> dFlag; DevBuf<float> dPart; DevBuf<NodeRes> dRes;
    const uint32_t TB = 256, nbMax = (P + TB - 1) / TB;
    dMiss.alloc(P); dCnt.alloc(1); dBlockCnt.alloc(nbMax); dPosA.alloc(P / 2 + 1); dPosB.alloc(P / 2 + 1); dFlag.alloc(P); dPart.alloc((size_t) nbMax * 24); dRes.alloc(1);
    ALVRL_CUDA(cudaMemsetAsync(dCnt.p, 0, sizeof(uint32_t), st));
    k_slice_make_recs<<<nbMax, TB, 0, st>>>(dPos, dDir, P, dRec.p, dMiss.p, dCnt.p);
    uint32_t launches = 1;
    uint32_t M = 0;
    dCnt.download(&M, 1, st);
    uint32_t firstGood = 0;
    if (M) {
        /* the reference's loop over the misses (1206-1221), replayed on the positions it touches */
        std::vector<uint32_t> miss(M);
        dMiss.download(miss.data(), M, st);
        std::sort(miss.begin(), miss.end());
        while (firstGood < M && miss[firstGood] == firstGood) firstGood++;          /* leading misses are in place already */
        if (firstGood < P) {
            std::unordered_map<uint32_t, uint32_t> cur;                             /* position -> original index (default: itself) */
            auto get = [&](uint32_t p) { auto it = cur.find(p); return it == cur.end() ? p : it->second; };
            for (uint32_t j = firstGood; j < M; j++) {                              /* every miss behind the first good record */
                const uint32_t m = miss[j];
                const uint32_t a = get(m), b = get(firstGood);
                cur[m] = b; cur[firstGood] = a;
                firstGood++;
            }
            if (!cur.empty()) {
                std::vector<uint2> patch; patch.reserve(cur.size());
                for (auto &kv : cur) patch.push_back(make_uint2(kv.first, kv.second));
                DevBuf<uint2> dPatch; dPatch.upload(patch, st);
                k_slice_patch<<<((uint32_t) patch.size() + TB - 1) / TB, TB, 0, st>>>(dPos, dDir, dPatch.p, (uint32_t) patch.size(), dRec.p);
                launches++;
                ALVRL_CUDA(cudaStreamSynchronize(st));
            }
        }
    }
    c->sliceLo.clear(); c->sliceSize.clear(); c->sliceCentroid.clear();
    c->dRecIdx.alloc(P); c->dPixelToSlice.alloc(P);
    std::vector<HNode> heap;
    NodeRes res;
    if (firstGood < P) {
        k_slice_bbox_root<<<1, 1024, 0, st>>>(dRec.p, firstGood, P - firstGood, dRes.p);
        launches++;
        dRes.download(&res, 1, st);
        heap.push_back(makeNode(firstGood, P, res.mn[0], res.mx[0]));
        while (heap.size() < targetNumSlices && heap.front().diag > 0) {                    /* 1364 */
            std::pop_heap(heap.begin(), heap.end());
            const HNode top = heap.back();
            heap.pop_back();
            const uint32_t n = top.hi - top.lo, nb = (n + TB - 1) / TB;
            k_slice_flags<<<nb, TB, 0, st>>>(dRec.p, top.lo, n, (int) top.dim, top.split, dFlag.p, dBlockCnt.p);
            k_slice_scan<<<1, 1024, 0, st>>>(dBlockCnt.p, nb, dRes.p);
            k_slice_ranks<<<nb, TB, 0, st>>>(dFlag.p, dBlockCnt.p, top.lo, n, dRes.p, dPosA.p, dPosB.p);
            k_slice_swap<<<(n / 2 + TB) / TB, TB, 0, st>>>(dRec.p, dPosA.p, dPosB.p, dRes.p, n / 2 + 1);
            k_slice_bbox<<<nb, TB, 0, st>>>(dRec.p, top.lo, n, dRes.p, dPart.p);
            k_slice_bbox_final<<<1, 256, 0, st>>>(dPart.p, nb, dRes.p);
            launches += 6;
            dRes.download(&res, 1, st);
            if (res.nS == 0 || res.nS == n) { c->stats.kernelLaunches += launches; return false; }   /* degenerate: the host loop decides */
            const uint32_t mid = top.lo + res.nS;
            heap.push_back(makeNode(top.lo, mid, res.mn[0], res.mx[0])); std::push_heap(heap.begin(), heap.end());
            heap.push_back(makeNode(mid, top.hi, res.mn[1], res.mx[1])); std::push_heap(heap.begin(), heap.end());
        }
    }
    /* slice id = heap array position, 1400-1417 */
    const uint32_t S = (uint32_t) heap.size();
    std::vector<std::pair<uint32_t, uint32_t>> byStart(S);
    for (uint32_t s = 0; s < S; s++) { c->sliceLo.push_back(heap[s].lo); c->sliceSize.push_back(heap[s].hi - heap[s].lo); c->sliceCentroid.insert(c->sliceCentroid.end(), heap[s].centroid, heap[s].centroid + 6); byStart[s] = std::make_pair(heap[s].lo, s); }
    std::sort(byStart.begin(), byStart.end());
    std::vector<uint32_t> starts(std::max(1u, S)), ids(std::max(1u, S));
    for (uint32_t s = 0; s < S; s++) { starts[s] = byStart[s].first; ids[s] = byStart[s].second; }
    DevBuf<uint32_t> dStarts, dIds; dStarts.upload(starts, st); dIds.upload(ids, st);
    k_slice_finish<<<nbMax, TB, 0, st>>>(dRec.p, S ? firstGood : P, P, dStarts.p, dIds.p, std::max(1u, S), c->dRecIdx.p, c->dPixelToSlice.p);
    launches++;
    ALVRL_CUDA(cudaGetLastError());
    ALVRL_CUDA(cudaStreamSynchronize(st));
    c->stats.kernelLaunches += launches;
    return true;
}

/* ---- consumers of the device-resident slices ------------------------------------------------------------------------ */
/* rowPixel[g] = recIdx[position[g]]: Slice::sampleRepresentativePixels picks positions inside a slice's range, 66-121 */
__global__ void k_slice_gather_rows(const uint32_t *__restrict__ recIdx, const uint32_t *__restrict__ position, uint32_t G, uint32_t *__restrict__ rowPixel) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) rowPixel[g] = recIdx[position[g]];
}
void slice_gather_rows_device(alvrl_ctx *c, const std::vector<uint32_t> &positions, std::vector<uint32_t> &rowPixel) {
    const uint32_t G = (uint32_t) positions.size();
    rowPixel.resize(G);
    if (!G) return;
    DevBuf<uint32_t> dPos; dPos.upload(positions, c->stream);
    c->dRowPixel.alloc(G);
    k_slice_gather_rows<<<(G + 255) / 256, 256, 0, c->stream>>>(c->dRecIdx.p, dPos.p, G, c->dRowPixel.p);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    c->dRowPixel.download(rowPixel.data(), G, c->stream);
}
/* the pixel lists of the render pass, slice by slice: pixels are visited in index order and appended to their slice's list
 * (warp-aggregated cursor bumps: the lists come out nearly sorted, which is all the render kernel wants -- ray coherence;
 * the image does not depend on the order) */
__global__ void k_slice_bucket_pixels(const uint32_t *__restrict__ pixelToSlice, uint32_t P, uint32_t W, uint32_t H, const uint32_t *__restrict__ sliceStart,
                                      uint32_t *__restrict__ cursor, uint32_t *__restrict__ slicePixels) {
    /* a block visits a 16 x 16 tile of the image and a warp an 8 x 4 patch of it, so that 32 consecutive entries of a slice's
     * list are neighbours in both directions (pixel index = y + H x) */
    const uint32_t tilesY = (H + 15u) / 16u;
    const uint32_t tx = blockIdx.x / tilesY, ty = blockIdx.x % tilesY;
    const uint32_t warp = threadIdx.x >> 5, l = threadIdx.x & 31;
    const uint32_t x = tx * 16u + (warp & 1u) * 8u + (l & 7u), y = ty * 16u + (warp >> 1) * 4u + (l >> 3);
    const uint32_t p = (x < W && y < H) ? y + H * x : P;
    const uint32_t s = p < P ? pixelToSlice[p] : ALVRL_NO_SLICE;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t peers = __match_any_sync(0xffffffffu, s);
    if (s == ALVRL_NO_SLICE) return;
    const uint32_t leader = __ffs(peers) - 1, rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor + s, (uint32_t) __popc(peers));
    base = __shfl_sync(peers, base, leader);
    slicePixels[sliceStart[s] + base + rank] = p;
}
void slice_bucket_pixels_device(alvrl_ctx *c, const std::vector<uint32_t> &sliceStart, uint32_t total) {
    const uint32_t P = c->numPixels(), S = (uint32_t) sliceStart.size();
    c->dSlicePixels.alloc(std::max(1u, total));
    if (!S || !total) return;
    DevBuf<uint32_t> dStart, dCursor; dStart.upload(sliceStart, c->stream); dCursor.alloc(S);
    ALVRL_CUDA(cudaMemsetAsync(dCursor.p, 0, S * sizeof(uint32_t), c->stream));
    k_slice_bucket_pixels<<<((c->cam.W + 15u) / 16u) * ((c->cam.H + 15u) / 16u), 256, 0, c->stream>>>(c->dPixelToSlice.p, P, c->cam.W, c->cam.H, dStart.p, dCursor.p, c->dSlicePixels.p);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    ALVRL_CUDA(cudaStreamSynchronize(c->stream));
}

} // namespace alvrl
