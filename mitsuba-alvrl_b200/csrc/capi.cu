/*
 * capi.cu -- the extern "C" entry points of include/alvrl.h and the host control flow of the path:
 * vrlIntegrator::preprocess / prepass / render restated around device kernels
 * (src/integrators/vrl/vrlIntegrator.cpp:237-356,386-599; Preprocessor.cpp:1130-1193,1502-1525).
 * No exception crosses the boundary; there is no CPU implementation of any kernel in this library.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <fstream>
#include <sstream>
#include <algorithm>
#include "context.h"
#include "occluders.h"
#include "kernels.h"
#include "hostio.h"
#include "shapes.h"

using namespace alvrl;

namespace {

/* pair-level culling (occ_query.h): the side of every slab face / plane that a whole VRL lies on, stored in the two spare
 * words of its record (dir.w: plane bits, power.w: slab-side bits) */
__global__ void k_vrl_occ_sides(VrlRec *__restrict__ vrls, uint32_t n, OccDev oc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = vrls[i].s, e = vrls[i].e;
    const uint32_t slabs = occ_slab_sides(oc, s.x, s.y, s.z, e.x, e.y, e.z, oc.cullMargin);
    const uint32_t planes = occ_plane_sides(oc, s.x, s.y, s.z, e.x, e.y, e.z, oc.cullMargin, false);
    vrls[i].dir.w = __uint_as_float(planes);
    vrls[i].power.w = __uint_as_float(slabs);
}
/* representative records of the render lists: rec[k] = vrls[idx[k]] with the cluster weight in e.w */
__global__ void k_gather_reps(const VrlRec *__restrict__ vrls, const uint32_t *__restrict__ idx, const float *__restrict__ w, uint32_t n,
                              VrlRec *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    VrlRec r = vrls[idx[k]];
    r.e.w = w[k];
    out[k] = r;
}

thread_local std::string g_err;

int fail(int code, const std::string &m) { g_err = m; return code; }

#define API_BEGIN try {
#define API_END                                                                    \
    } catch (const alvrl::Error &e) { return fail(e.code, e.what()); }              \
    catch (const std::exception &e) { return fail(ALVRL_ERR_ARG, e.what()); }       \
    return ALVRL_OK;

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void use_device(alvrl_ctx *c) { ALVRL_CUDA(cudaSetDevice(c->device)); }

HostSampler *new_stream(const alvrl_params &P) {
    if (P.rngMode == ALVRL_RNG_MODE_SFMT) return new SfmtStream(P.seed);
    return new CounterStream(P.seed);
}

/* BVH + TriAccel upload; ShapeKDTree / Scene bounding boxes (gkdtree.h:1213-1220, scene.cpp:387-413) */
void ensure_scene(alvrl_ctx *c) {
    if (!c->sceneDirty) return;
    if (!c->haveMesh) throw Error(ALVRL_ERR_STATE, "set_mesh first");
    const double tScene0 = now_ms();
    const uint32_t nt = (uint32_t) c->triMat.size();
    HostBvh bvh;
    BvhBuilder(c->verts.data(), c->tris.data(), nt).build(bvh);
    std::vector<TriRec> recs(nt);
    std::vector<TriFast> fast(nt);
    std::vector<float4> tv(3 * (size_t) nt);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < nt; i++) {
        const uint32_t t = bvh.triOrder[i];
        const float *A = &c->verts[3 * (size_t) c->tris[3 * (size_t) t]], *B = &c->verts[3 * (size_t) c->tris[3 * (size_t) t + 1]],
                    *C = &c->verts[3 * (size_t) c->tris[3 * (size_t) t + 2]];
        recs[i] = makeTriRec(A, B, C, t);
        fast[i] = makeTriFast(A, B, C);
        tv[3 * (size_t) t] = make_float4(A[0], A[1], A[2], 0); tv[3 * (size_t) t + 1] = make_float4(B[0], B[1], B[2], 0);
        tv[3 * (size_t) t + 2] = make_float4(C[0], C[1], C[2], 0);
        for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], std::min(A[k], std::min(B[k], C[k]))); mx[k] = std::max(mx[k], std::max(A[k], std::max(B[k], C[k]))); }
    }
    const float eps = 1e-3f;
    for (int k = 0; k < 3; k++) {
        c->kdMin[k] = mn[k] - ((mx[k] - mn[k]) * eps + eps);
        c->kdMax[k] = mx[k] + ((mx[k] - c->kdMin[k]) * eps + eps);
        c->sceneMin[k] = c->kdMin[k]; c->sceneMax[k] = c->kdMax[k];
    }
    for (size_t i = 0; i + 2 < c->extraBounds.size(); i += 3)
        for (int k = 0; k < 3; k++) { c->sceneMin[k] = std::min(c->sceneMin[k], c->extraBounds[i + k]); c->sceneMax[k] = std::max(c->sceneMax[k], c->extraBounds[i + k]); }
    c->dNodes.upload(bvh.nodes, c->stream);
    std::vector<BvhNode> leaves;
    for (const BvhNode &nd : bvh.nodes) { uint32_t lf; memcpy(&lf, &nd.hi.w, 4); if (lf) leaves.push_back(nd); }
    c->dLeafNodes.upload(leaves, c->stream);
    c->dTris.upload(recs, c->stream);
    c->dTrisFast.upload(fast, c->stream);
    c->dTriVerts.upload(tv, c->stream);
    c->dTriMat.upload(c->triMat, c->stream);
    SceneDev &s = c->sceneDev;
    s.nodes = c->dNodes.p; s.tris = c->dTris.p; s.trisFast = c->dTrisFast.p; s.numNodes = (uint32_t) bvh.nodes.size(); s.numTris = nt; s.leafNodes = c->dLeafNodes.p; s.numLeaves = (uint32_t) leaves.size();
    for (int k = 0; k < 3; k++) { s.kdMin[k] = c->kdMin[k]; s.kdMax[k] = c->kdMax[k]; }
    s.anyHit = c->P.anyHitShadowRays ? 1 : 0;
    /* visibility strategy of the fast flavour for small scenes (ALVRL_VIS=tree|flat|occ overrides the choice) */
    s.occTris = nullptr; s.numOccTris = 0; memset(&c->occHost, 0, sizeof(c->occHost));
    s.visMode = (s.numLeaves <= 32 && nt <= 128) ? 1 : 0;
    const char *visEnv = getenv("ALVRL_VIS");
    if (s.visMode == 1 && !(visEnv && !strcmp(visEnv, "flat"))) {
        const OccluderSet os = compile_occluders(c->verts.data(), c->tris.data(), nt, s.numLeaves);
        if (os.use) {
            c->dOcc.upload(os.tris, c->stream);
            s.occTris = c->dOcc.p; s.numOccTris = os.numTris; s.visMode = 2; c->occHost = os.dev;
        }
    }
    if (visEnv && !strcmp(visEnv, "tree")) s.visMode = 0;
    s.nodes4 = nullptr; s.numNodes4 = 0;
    if (s.visMode == 0) {                                    /* large scenes: 4-wide tree for the fast flavour's any-hit query */
        std::vector<Bvh4Node> n4;
        const uint32_t depth = collapse4(bvh, n4);
        /* a lane's stack holds at most three pending siblings per level (dev_common.cuh::ALVRL_BVH4_STACK = 64); a deeper tree
         * keeps the binary skip-pointer query, which needs no stack */
        if (3 * depth + 4 <= 64 && !getenv("ALVRL_NO_BVH4")) {
            c->dNodes4.upload(n4, c->stream);
            s.nodes4 = c->dNodes4.p; s.numNodes4 = (uint32_t) n4.size();
        }
    }
    c->occHost.cullMargin = 1e-5f * std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
    if (getenv("ALVRL_NO_PAIR_CULL")) c->occHost.cullMargin = INFINITY;      /* experiments: nothing is ever "strictly beyond" */
    c->vrlSidesValid = false;
    c->stats.bvhNodes = s.numNodes; c->stats.visMode = (uint32_t) s.visMode;
    c->stats.msSceneBuild = (float) (now_ms() - tScene0);
    c->sceneDirty = false; c->segsDirty = true;
}

/* per-VRL side bits of the compiled occluder set (needs scene + VRLs; redone when either changes) */
void ensure_vrl_sides(alvrl_ctx *c) {
    ensure_scene(c);
    if (c->vrlSidesValid || !c->haveVrls) return;
    if (c->sceneDev.visMode == 2) {
        const uint32_t n = (uint32_t) c->vrlHost.size();
        k_vrl_occ_sides<<<(n + 255) / 256, 256, 0, c->stream>>>(c->dVrls.p, n, c->occHost);
        c->stats.kernelLaunches++;
        ALVRL_CUDA(cudaGetLastError());
    }
    c->vrlSidesValid = true;
    c->renderListsDirty = true;
}

void ensure_primary(alvrl_ctx *c) {
    ensure_scene(c);
    if (!c->segsDirty && c->havePrimary) return;
    if (!c->haveCam || !c->haveMat) throw Error(ALVRL_ERR_STATE, "scene incomplete: set_camera / set_materials first");
    for (uint32_t m : c->triMat) if (m >= c->matBits.size()) throw Error(ALVRL_ERR_ARG, "triangle material index out of range of set_materials");
    const uint32_t P = c->numPixels();
    c->dPixSegs.alloc(P); c->dHitPrim.alloc(P); c->dHitT.alloc(P);
    launch_primary(c->sceneDev, c->medium, c->cam, c->dTriVerts.p, c->dTriMat.p, c->dMatAlbedo.p, c->dMatBits.p, c->haveMedium,
                   c->dPixSegs.p, c->dHitPrim.p, c->dHitT.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    ALVRL_CUDA(cudaStreamSynchronize(c->stream));
    c->havePrimary = true; c->segsDirty = false; c->chainsValid = false;
}

/* the specular chains below the camera segments (chain.cu): count, scan, write; host mirrors of the grouping */
void ensure_chains(alvrl_ctx *c) {
    ensure_primary(c);
    if (c->chainsValid) return;
    const uint32_t P = c->numPixels();
    c->anyDelta = false;
    for (uint32_t b : c->matBits) if (b & ALVRL_BSDF_DELTA) c->anyDelta = true;
    c->chainOffset.assign(P + 1, 0u); c->chainMeta.clear();
    if (c->anyDelta) {
        if (c->optics.size() != 12 * c->matBits.size()) throw Error(ALVRL_ERR_STATE, "materials with delta components need alvrl_set_material_optics");
        DevBuf<uint32_t> dCount; dCount.alloc(P);
        launch_chain_count(c->sceneDev, c->medium, c->haveMedium, c->dTriVerts.p, c->dTriMat.p, c->dMatAlbedo.p, c->dMatBits.p, c->dMatOptics.p, c->P.seed,
                           c->P.specularForcedRRdepth, c->P.initialSpecularThroughput, c->dPixSegs.p, c->dHitPrim.p, c->dHitT.p, P, dCount.p, c->stream);
        c->stats.kernelLaunches++;
        ALVRL_CUDA(cudaGetLastError());
        std::vector<uint32_t> cnt(P);
        dCount.download(cnt.data(), P, c->stream);
        for (uint32_t i = 0; i < P; i++) c->chainOffset[i + 1] = c->chainOffset[i] + cnt[i];
        const uint32_t total = c->chainOffset[P];
        if (total) {
            DevBuf<uint32_t> dOff; dOff.upload(c->chainOffset, c->stream);
            c->dChainSegs.alloc(total); c->dChainMeta.alloc(total);
            launch_chain_write(c->sceneDev, c->medium, c->haveMedium, c->dTriVerts.p, c->dTriMat.p, c->dMatAlbedo.p, c->dMatBits.p, c->dMatOptics.p, c->P.seed,
                               c->P.specularForcedRRdepth, c->P.initialSpecularThroughput, c->dPixSegs.p, c->dHitPrim.p, c->dHitT.p, P, dOff.p,
                               c->dChainSegs.p, c->dChainMeta.p, c->stream);
            c->stats.kernelLaunches++;
            ALVRL_CUDA(cudaGetLastError());
            c->chainMeta.resize(total);
            c->dChainMeta.download(c->chainMeta.data(), total, c->stream);
            std::vector<uint32_t> key(total);
            for (uint32_t e = 0; e < total; e++) {
                if (c->chainMeta[e].x >= (1u << 24) || c->chainMeta[e].y >= 255u) throw Error(ALVRL_ERR_UNSUPPORTED, "specular chains: more than 2^24 pixels or 254 segments below one pixel");
                key[e] = c->chainMeta[e].x + ((c->chainMeta[e].y + 1u) << 24);
            }
            c->dChainKey.upload(key, c->stream);
            ALVRL_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    c->chainsValid = true;
}

void invalidate_from_slices(alvrl_ctx *c) { c->pixelListsDirty = true; c->haveRows = false; c->haveR = false; c->haveClusters = false; c->haveFallback = false; c->renderListsDirty = true; }

/* host loop of slices.h (reference order, sequential), then the same device-resident representation the device builder leaves */
void finish_slices(alvrl_ctx *c, const std::vector<P3> &pos, const std::vector<P3> &dir) {
    SliceTree tree(pos, dir);
    std::vector<SliceInfo> slices;
    const std::vector<uint32_t> toSlice = tree.build((uint32_t) c->P.targetNumSlices, slices);
    const uint32_t P = (uint32_t) toSlice.size();
    std::vector<uint32_t> recIdx; recIdx.reserve(P);
    c->sliceLo.clear(); c->sliceSize.clear();
    for (const SliceInfo &si : slices) {
        c->sliceLo.push_back((uint32_t) recIdx.size()); c->sliceSize.push_back((uint32_t) si.pixels.size());
        recIdx.insert(recIdx.end(), si.pixels.begin(), si.pixels.end());
    }
    recIdx.resize(P, 0u);
    c->sliceCentroid = tree.centroids;
    c->dRecIdx.upload(recIdx, c->stream); c->dPixelToSlice.upload(toSlice, c->stream);
    c->haveSlices = true;
    invalidate_from_slices(c);
    c->stats.numSlices = c->numSlices();
}

TransportParams make_transport_params(alvrl_ctx *c, uint32_t domain) {
    TransportParams T;
    memset(&T, 0, sizeof(T));
    T.scene = c->sceneDev; T.medium = c->medium; T.occ = c->occHost;
    T.Nvv = c->P.volVolSamples; T.Nvs = c->P.volSurfSamples; T.shortVrls = c->P.shortVrls; T.Rsamples = c->P.Rsamples;
    T.seed = c->P.seed; T.rngDomain = domain;
    T.tape = nullptr; T.tapeK = 0;
    T.numVrls = (uint32_t) c->vrlHost.size();
    T.normalization = (float) (1.0 / (double) c->particleCount);       /* vrlIntegrator.cpp:805 */
    T.invParticleDiv = (float) c->particleCount;                       /* 590: Li /= getParticleCount() */
    T.rowBase = 0;
    return T;
}

/* reference-stream tape for build_R: the prepass loops draw 2*Nvv (+ Nvs when the vol->surf loop runs) uniforms per
 * (row, vrl), rows in slice order, VRLs in index order (vrlIntegrator.cpp:322-333,804-816) */
void make_sfmt_tape(alvrl_ctx *c, std::vector<float> &tape) {
    const uint32_t N = (uint32_t) c->vrlHost.size(), G = (uint32_t) c->rowPixel.size(), K = c->K(), S = c->numSlices();
    const uint64_t need = (uint64_t) G * N * K;
    if (need > (1ull << 31)) throw Error(ALVRL_ERR_UNSUPPORTED, "rngMode=SFMT needs a G*N*(2*Nvv+Nvs) float tape; too large for this configuration");
    tape.assign(need, 0.5f);
    std::vector<SegRec> rows(G);
    c->dRowSegs.download(rows.data(), G, c->stream);
    const int w = std::max(1, c->P.workerCount);
    std::vector<std::unique_ptr<HostSampler>> clones;
    if (w > 1) for (int i = 0; i < w; i++) clones.emplace_back(c->mainSampler->clone());   /* Rbuilder ctor, 1048 */
    const int Nvv2 = 2 * c->P.volVolSamples, Nvs = c->P.volSurfSamples;
    for (int id = 0; id < w; id++) {
        HostSampler *smp = w > 1 ? clones[id].get() : c->mainSampler.get();
        const uint32_t s0 = (uint32_t) (((uint64_t) id * S) / w), s1 = (uint32_t) (((uint64_t) (id + 1) * S) / w);
        for (uint32_t row = c->rowOffset[s0]; row < c->rowOffset[s1]; row++) {
            const SegRec &sg = rows[row];
            uint32_t flags; memcpy(&flags, &sg.dn.w, 4);
            if (!(flags & SEG_VALID)) continue;
            const bool surf = !(sg.tE.x == 0 && sg.tE.y == 0 && sg.tE.z == 0) && (flags & SEG_SMOOTH);
            const int draws = Nvv2 + (surf ? Nvs : 0);
            for (uint32_t v = 0; v < N; v++) {
                float *t = &tape[((uint64_t) row * N + v) * K];
                for (int k = 0; k < draws; k++) t[k] = smp->next1D();
            }
        }
    }
}

void build_render_lists(alvrl_ctx *c) {
    ensure_vrl_sides(c);
    if (!c->renderListsDirty) return;
    const uint32_t S = c->numSlices();
    std::vector<uint32_t> sliceStart(S, 0), repOffset(S + 1, 0), repIdx;
    std::vector<float> repW;
    std::vector<uint4> work;
    uint32_t total = 0;
    for (uint32_t s = 0; s < S; s++) {
        const uint32_t base = total, cnt = c->sliceSize[s];
        sliceStart[s] = base; total += cnt;
        for (uint32_t o = 0; o < cnt; o += ALVRL_CTA_SEGS_HOST)
            work.push_back(make_uint4(s, base + o, std::min<uint32_t>(ALVRL_CTA_SEGS_HOST, cnt - o), 0));
        const std::vector<uint32_t> &vr = c->selectedVrls[s];
        const std::vector<float> &wt = c->clusterWeight[s];
        for (size_t i = 0; i < vr.size(); i++) {
            if (vr[i] >= c->vrlHost.size()) throw Error(ALVRL_ERR_ARG, "cluster representative out of range");
            repIdx.push_back(vr[i]); repW.push_back(wt[i]);
        }
        repOffset[s + 1] = (uint32_t) repIdx.size();
    }
    /* the pixel lists are bucketed on the device from pixelToSlice (nearly sorted inside a slice: ray coherence) */
    if (c->pixelListsDirty) { slice_bucket_pixels_device(c, sliceStart, total); c->pixelListsDirty = false; }
    c->dWork.upload(work, c->stream);
    c->workHost = work;
    c->dRepOffset.upload(repOffset, c->stream);
    /* the representatives' records are gathered on the device (they carry the per-VRL side bits of ensure_vrl_sides) */
    const uint32_t nRep = (uint32_t) repIdx.size();
    c->dRepRecs.alloc(std::max<uint32_t>(1, nRep));
    if (nRep) {
        DevBuf<uint32_t> dIdx; DevBuf<float> dW;
        dIdx.upload(repIdx, c->stream); dW.upload(repW, c->stream);
        k_gather_reps<<<(nRep + 255) / 256, 256, 0, c->stream>>>(c->dVrls.p, dIdx.p, dW.p, nRep, c->dRepRecs.p);
        c->stats.kernelLaunches++;
        ALVRL_CUDA(cudaGetLastError());
        ALVRL_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->numWork = (uint32_t) work.size();
    c->renderListsDirty = false;
}

/* work items restricted to the slice range of this handle (multi-GPU sharding by slice) */
void select_work(alvrl_ctx *c, const uint4 *&work, uint32_t &numWork, std::vector<uint4> &tmp, DevBuf<uint4> &dTmp) {
    const uint32_t S = c->numSlices();
    const uint32_t sb = std::min(c->sliceBegin, S), se = std::min(c->sliceEnd, S);
    if (sb == 0 && se == S) { work = c->dWork.p; numWork = c->numWork; return; }
    tmp.clear();
    for (const uint4 &w : c->workHost) if (w.x >= sb && w.x < se) tmp.push_back(w);
    dTmp.upload(tmp, c->stream);
    work = dTmp.p; numWork = (uint32_t) tmp.size();
}

void render_clustered_into(alvrl_ctx *c, float4 *fb, cudaStream_t st) {
    if (!c->haveClusters) throw Error(ALVRL_ERR_STATE, "build_clusters first");
    ensure_primary(c);
    build_render_lists(c);
    TransportParams T = make_transport_params(c, ALVRL_RNG_RENDER);
    const uint4 *work; uint32_t numWork; std::vector<uint4> tmp; DevBuf<uint4> dTmp;
    select_work(c, work, numWork, tmp, dTmp);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    if (c->mathMode == 1) launch_render_strict(T, true, c->dPixSegs.p, c->dSlicePixels.p, work, numWork, c->dRepRecs.p, c->dRepOffset.p, fb, c->cam.W, c->cam.H, st);
    else launch_render_fast(T, true, c->dPixSegs.p, c->dSlicePixels.p, work, numWork, c->dRepRecs.p, c->dRepOffset.p, fb, c->cam.W, c->cam.H, st);
    cudaEventRecord(e1, st);
    ALVRL_CUDA(cudaGetLastError());
    ALVRL_CUDA(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->stats.msTransportKernelRender = ms;
    c->stats.kernelLaunches++;
    uint64_t pairs = 0;
    const uint32_t S = c->numSlices(), sb = std::min(c->sliceBegin, S), se = std::min(c->sliceEnd, S);
    for (uint32_t s = sb; s < se; s++) pairs += (uint64_t) c->sliceSize[s] * c->selectedVrls[s].size();
    /* LiSpec (462-505): the in-medium segments of the pixels' specular chains, rendered from their own list with the
     * representatives of the slice of the ORIGINAL camera ray (552-566), then added to their pixel, weighted, in chain order */
    ensure_chains(c);
    if (c->anyDelta && c->chainOffset.back()) {
        std::vector<uint32_t> pts(c->numPixels());
        c->dPixelToSlice.download(pts.data(), pts.size(), st);
        std::vector<std::vector<uint32_t>> bySlice(S);
        std::vector<uint32_t> xPix;
        const uint32_t total = c->chainOffset.back();
        for (uint32_t pix = 0; pix < c->numPixels(); pix++) {
            const uint32_t a = c->chainOffset[pix], b = c->chainOffset[pix + 1], sl = pts[pix];
            if (a == b || sl == ALVRL_NO_SLICE || sl < sb || sl >= se) continue;
            bool any = false;
            for (uint32_t e = a; e < b; e++) if (c->chainMeta[e].w & 1u) { bySlice[sl].push_back(e); any = true; }
            if (any) xPix.push_back(pix);
        }
        std::vector<uint32_t> list; std::vector<uint4> xwork;
        for (uint32_t s = sb; s < se; s++) {
            const uint32_t base = (uint32_t) list.size(), cnt = (uint32_t) bySlice[s].size();
            list.insert(list.end(), bySlice[s].begin(), bySlice[s].end());
            for (uint32_t o = 0; o < cnt; o += ALVRL_CTA_SEGS_HOST) xwork.push_back(make_uint4(s, base + o, std::min<uint32_t>(ALVRL_CTA_SEGS_HOST, cnt - o), 0));
            pairs += (uint64_t) cnt * c->selectedVrls[s].size();
        }
        if (!list.empty()) {
            c->dXList.upload(list, st); c->dXWork.upload(xwork, st); c->dXPix.upload(xPix, st);
            c->dSubLi.alloc(total);
            ALVRL_CUDA(cudaMemsetAsync(c->dSubLi.p, 0, (size_t) total * sizeof(float4), st));
            if (c->mathMode == 1) launch_render_strict(T, true, c->dChainSegs.p, c->dXList.p, c->dXWork.p, (uint32_t) xwork.size(), c->dRepRecs.p, c->dRepOffset.p, c->dSubLi.p, c->cam.W, c->cam.H, st, c->dChainKey.p);
            else launch_render_fast(T, true, c->dChainSegs.p, c->dXList.p, c->dXWork.p, (uint32_t) xwork.size(), c->dRepRecs.p, c->dRepOffset.p, c->dSubLi.p, c->cam.W, c->cam.H, st, c->dChainKey.p);
            /* per pixel: first = chainOffset[pix], end = chainOffset[pix + 1] */
            std::vector<uint32_t> firstEnd(2 * xPix.size());
            for (size_t i = 0; i < xPix.size(); i++) { firstEnd[2 * i] = c->chainOffset[xPix[i]]; firstEnd[2 * i + 1] = c->chainOffset[xPix[i] + 1]; }
            c->dXPixFirst.upload(firstEnd, st);
            launch_chain_accumulate(fb, c->cam.W, c->cam.H, c->dSubLi.p, c->dChainSegs.p, c->dXPix.p, c->dXPixFirst.p, (uint32_t) xPix.size(), st);
            c->stats.kernelLaunches += 2;
            ALVRL_CUDA(cudaGetLastError());
            ALVRL_CUDA(cudaStreamSynchronize(st));
        }
    }
    c->stats.pairsRender += pairs;
    c->stats.shadowRays += pairs * (uint64_t) (c->P.volVolSamples + c->P.volSurfSamples);
}

} // namespace

extern "C" {

const char *alvrl_last_error(void) { return g_err.c_str(); }

void alvrl_params_default(alvrl_params *p) {
    memset(p, 0, sizeof(*p));
    p->shortVrls = 1; p->vrlTargetNum = 500; p->maxParticleDepth = -1; p->specularForcedRRdepth = 100;
    p->initialSpecularThroughput = 20; p->volVolSamples = 2; p->volSurfSamples = 2; p->globalCluster = 0;
    p->globalUndersampling = -1; p->localRefinement = 1; p->localUndersampling = -1; p->fallBackUndersampling = 5;
    p->targetNumSlices = 100; p->targetPixelUndersampling = 64; p->sliceCurvatureFactor = 0.5f;
    p->neighbourCount = 0; p->neighbourWeight = 0; p->Rsamples = 1; p->depthCorrection = 1; p->maxPasses = 1;
    p->rngMode = ALVRL_RNG_MODE_COUNTER; p->seed = 0; p->anyHitShadowRays = 1; p->workerCount = 1; p->rrDepth = 5;
}

int alvrl_create(int device, const alvrl_params *p, alvrl_handle *out) {
    if (!p || !out) return fail(ALVRL_ERR_ARG, "null argument");
    /* the reference's constructor checks (vrlIntegrator.cpp:149-156; Preprocessor.cpp:36-38) */
    if (p->volVolSamples != 0 && p->volVolSamples < 2) return fail(ALVRL_ERR_ARG, "Need at least 2 volVolSamples for variance estimate");
    if (p->volSurfSamples != 0 && p->volSurfSamples < 2) return fail(ALVRL_ERR_ARG, "Need at least 2 volSurfSamples for variance estimate");
    if (p->targetNumSlices < 1) return fail(ALVRL_ERR_ARG, "Invalid target number of slices!");
    if (p->neighbourWeight > 0 && p->rngMode != ALVRL_RNG_MODE_COUNTER)
        return fail(ALVRL_ERR_UNSUPPORTED, "neighbourWeight > 0 (neighbour slices in L_i) needs the counter sample stream on the device path");
    if (p->neighbourWeight >= 1) return fail(ALVRL_ERR_ARG, "neighbourWeight must be below 1");
    if (p->Rsamples != 1) return fail(ALVRL_ERR_UNSUPPORTED, "Rsamples != 1 is outside the device path");
    if (p->depthCorrection != 1 && p->rngMode != ALVRL_RNG_MODE_COUNTER)
        return fail(ALVRL_ERR_UNSUPPORTED, "depthCorrection != 1 replays the split decisions: it needs the counter sample stream (rngMode = COUNTER)");
    if (!(p->depthCorrection > 0)) return fail(ALVRL_ERR_ARG, "depthCorrection must be positive");
    if (p->numVrlFalseColor || p->slicesFalseColor || p->convergenceFalseColor)
        return fail(ALVRL_ERR_UNSUPPORTED, "false-colour debug outputs (vrlIntegrator.cpp:199-201) are not implemented on the device path");
    if (p->maxPasses > 1) return fail(ALVRL_ERR_UNSUPPORTED, "maxPasses > 1: the device path renders one pass per VRL set (the caller accumulates passes)");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ALVRL_ERR_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(ALVRL_ERR_ARG, "cuda device index out of range");
    alvrl_ctx *c = nullptr;
    API_BEGIN
    c = new alvrl_ctx();
    c->P = *p; c->device = device;
    memset(&c->stats, 0, sizeof(c->stats));
    memset(&c->medium, 0, sizeof(c->medium));
    memset(&c->cam, 0, sizeof(c->cam));
    memset(&c->sceneDev, 0, sizeof(c->sceneDev));
    ALVRL_CUDA(cudaSetDevice(device));
    ALVRL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->timer.init();
    c->mainSampler.reset(new_stream(c->P));
    DevPool::get().addHandle(device);
    const char *mm = getenv("ALVRL_MATH");
    c->mathMode = (mm && std::string(mm) == "strict") ? 1 : 0;
    *out = c;
    API_END
}

void alvrl_destroy(alvrl_handle c) {
    if (!c) return;
    cudaSetDevice(c->device);
    c->timer.destroy();
    if (c->stream) cudaStreamDestroy(c->stream);
    const int dev = c->device;
    delete c;
    DevPool::get().dropHandle(dev);
}

int alvrl_get_params(alvrl_handle c, alvrl_params *out) { *out = c->P; return ALVRL_OK; }
int alvrl_set_math_mode(alvrl_handle c, int strict) { c->mathMode = strict ? 1 : 0; return ALVRL_OK; }

int alvrl_set_mesh(alvrl_handle c, const float *v, uint32_t nv, const uint32_t *tris, uint32_t nt, const uint32_t *mat) {
    API_BEGIN
    if (!v || !tris || !mat || nt == 0) throw Error(ALVRL_ERR_ARG, "empty mesh");
    for (size_t i = 0; i < 3 * (size_t) nt; i++) if (tris[i] >= nv) throw Error(ALVRL_ERR_ARG, "triangle index out of range");
    c->verts.assign(v, v + 3 * (size_t) nv); c->tris.assign(tris, tris + 3 * (size_t) nt); c->triMat.assign(mat, mat + nt);
    c->haveMesh = true; c->sceneDirty = true; c->havePrimary = false; c->haveSlices = false; invalidate_from_slices(c);
    API_END
}

/* analytic shapes as triangles appended to the host copy of the mesh (shapes.h); the device sees them at the next ensure_scene */
static void appended_triangles(alvrl_ctx *c, size_t trisBefore, uint32_t material, uint32_t *first, uint32_t *count) {
    const size_t nt = c->tris.size() / 3;
    c->triMat.resize(nt, material);
    if (first) *first = (uint32_t) trisBefore;
    if (count) *count = (uint32_t) (nt - trisBefore);
    c->haveMesh = true; c->sceneDirty = true; c->havePrimary = false; c->haveSlices = false; invalidate_from_slices(c);
}

int alvrl_add_rectangle(alvrl_handle c, const float toWorld[16], int flipNormals, uint32_t material, uint32_t *firstTriangle) {
    API_BEGIN
    if (!toWorld) throw Error(ALVRL_ERR_ARG, "alvrl_add_rectangle: null transform");
    for (int i = 0; i < 12; i++) if (!std::isfinite(toWorld[i])) throw Error(ALVRL_ERR_ARG, "alvrl_add_rectangle: 'toWorld' is not finite");
    const size_t before = c->tris.size() / 3;
    tessellate_rectangle(toWorld, flipNormals != 0, c->verts, c->tris);
    appended_triangles(c, before, material, firstTriangle, nullptr);
    API_END
}

int alvrl_add_sphere(alvrl_handle c, const float center[3], float radius, int flipNormals, uint32_t thetaSteps, uint32_t material,
                     uint32_t *firstTriangle, uint32_t *triangleCount) {
    API_BEGIN
    if (!center || !std::isfinite(center[0]) || !std::isfinite(center[1]) || !std::isfinite(center[2]) || !std::isfinite(radius))
        throw Error(ALVRL_ERR_ARG, "alvrl_add_sphere: centre and radius must be finite");
    if (thetaSteps > 2048) throw Error(ALVRL_ERR_ARG, "alvrl_add_sphere: at most 2048 polar steps");
    const size_t before = c->tris.size() / 3;
    tessellate_sphere(center, radius, flipNormals != 0, thetaSteps ? thetaSteps : 64u, c->verts, c->tris);
    appended_triangles(c, before, material, firstTriangle, triangleCount);
    API_END
}

int alvrl_set_materials(alvrl_handle c, const float *albedo, const uint32_t *bits, uint32_t nm) {
    API_BEGIN
    use_device(c);
    c->albedo.assign(albedo, albedo + 3 * (size_t) nm); c->matBits.assign(bits, bits + nm);
    std::vector<float4> a(nm);
    for (uint32_t i = 0; i < nm; i++) a[i] = make_float4(albedo[3 * i], albedo[3 * i + 1], albedo[3 * i + 2], 0);
    c->dMatAlbedo.upload(a, c->stream); c->dMatBits.upload(c->matBits, c->stream);
    c->haveMat = true; c->segsDirty = true;
    API_END
}

int alvrl_set_material_optics(alvrl_handle c, const float *optics, uint32_t nm) {
    API_BEGIN
    use_device(c);
    if (!c->haveMat || nm != c->matBits.size()) throw Error(ALVRL_ERR_ARG, "alvrl_set_material_optics: call alvrl_set_materials first, with the same material count");
    if (!optics) throw Error(ALVRL_ERR_ARG, "alvrl_set_material_optics: null optics");
    for (uint32_t i = 0; i < nm; i++)
        if ((c->matBits[i] & ALVRL_BSDF_DIELECTRIC) && !(optics[12 * (size_t) i] > 0)) throw Error(ALVRL_ERR_ARG, "alvrl_set_material_optics: dielectric eta must be positive");
    c->optics.assign(optics, optics + 12 * (size_t) nm);
    std::vector<float4> o(3 * (size_t) nm);
    for (uint32_t i = 0; i < nm; i++) for (int k = 0; k < 3; k++) o[3 * (size_t) i + k] = make_float4(optics[12 * (size_t) i + 4 * k], optics[12 * (size_t) i + 4 * k + 1], optics[12 * (size_t) i + 4 * k + 2], optics[12 * (size_t) i + 4 * k + 3]);
    c->dMatOptics.upload(o, c->stream);
    c->chainsValid = false; c->haveR = false; c->haveClusters = false;
    API_END
}

int alvrl_get_chain_segments(alvrl_handle c, uint32_t *offset, float *segs) {
    API_BEGIN
    use_device(c);
    if (!offset) throw Error(ALVRL_ERR_ARG, "alvrl_get_chain_segments: null offset");
    ensure_chains(c);
    const uint32_t P = c->numPixels(), total = c->chainOffset[P];
    memcpy(offset, c->chainOffset.data(), (size_t) (P + 1) * sizeof(uint32_t));
    if (segs && total) {
        std::vector<SegRec> s(total);
        c->dChainSegs.download(s.data(), total, c->stream);
        for (uint32_t e = 0; e < total; e++) {
            float *o = segs + 16 * (size_t) e;
            o[0] = s[e].o.x; o[1] = s[e].o.y; o[2] = s[e].o.z; o[3] = s[e].d.x; o[4] = s[e].d.y; o[5] = s[e].d.z;
            o[6] = s[e].p.x; o[7] = s[e].p.y; o[8] = s[e].p.z; o[9] = s[e].o.w;
            o[10] = s[e].p.w; o[11] = s[e].n.w; o[12] = s[e].albedo.w; o[13] = (c->chainMeta[e].w & 1u) ? 1.0f : 0.0f;
            o[14] = (float) c->chainMeta[e].z; o[15] = (float) (c->chainMeta[e].w >> 1);
        }
    }
    API_END
}

int alvrl_set_extra_bounds(alvrl_handle c, const float *p, uint32_t n) {
    c->extraBounds.assign(p, p + 3 * (size_t) n); c->sceneDirty = true;
    return ALVRL_OK;
}

int alvrl_set_medium_homogeneous(alvrl_handle c, const float a[3], const float s[3], float w, int32_t phase, float g) {
    API_BEGIN
    MediumDev &m = c->medium;
    memset(&m, 0, sizeof(m));
    m.type = 0; m.phaseType = phase; m.g = g;
    for (int i = 0; i < 3; i++) { m.sigmaS[i] = s[i]; m.sigmaT[i] = a[i] + s[i]; }
    /* mediumSamplingWeight default, homogeneous.cpp:168-184 */
    float sw = w;
    if (sw < 0) {
        sw = -1;
        for (int i = 0; i < 3; i++) { float alb = m.sigmaS[i] / m.sigmaT[i]; if (alb > sw && m.sigmaT[i] != 0) sw = alb; }
        if (sw > 0) sw = std::max(sw, 0.5f);
    }
    m.samplingWeight = sw;
    m.grey = (m.sigmaT[0] == m.sigmaT[1] && m.sigmaT[1] == m.sigmaT[2]) ? 1 : 0;
    c->haveMedium = true; c->segsDirty = true; c->haveR = false;
    API_END
}

int alvrl_set_medium_grid(alvrl_handle c, const float *density, const int32_t res[3], const float mn[3], const float mx[3], float scale,
                          const float albedo[3], const float sBase[3], int32_t phase, float g) {
    API_BEGIN
    use_device(c);
    if (res[0] < 2 || res[1] < 2 || res[2] < 2) throw Error(ALVRL_ERR_ARG, "grid resolution must be >= 2");
    MediumDev &m = c->medium;
    memset(&m, 0, sizeof(m));
    m.type = 1; m.phaseType = phase; m.g = g; m.scale = scale;
    const size_t n = (size_t) res[0] * res[1] * res[2];
    c->dGrid.upload(density, n, c->stream);
    m.density = c->dGrid.p;
    m.stepSize = INFINITY;
    for (int i = 0; i < 3; i++) {
        m.res[i] = res[i]; m.bmin[i] = mn[i]; m.bmax[i] = mx[i]; m.albedo[i] = albedo[i]; m.sigmaS[i] = sBase[i]; m.sigmaT[i] = 0;
        const float ext = mx[i] - mn[i];
        m.gsc[i] = (float) (res[i] - 1) / ext;                         /* gridvolume.cpp:190-196 */
        m.gtr[i] = m.gsc[i] * (-mn[i]);
        m.stepSize = std::min(m.stepSize, 0.5f * ext / (float) (res[i] - 1));   /* gridvolume.cpp:197-199 */
    }
    c->haveMedium = true; c->segsDirty = true; c->haveR = false;
    API_END
}

int alvrl_set_medium_grid_file(alvrl_handle c, const char *path, const float *mn, const float *mx, float scale,
                               const float albedo[3], const float sBase[3], int32_t phase, float g) {
    VolFile vf;
    try { read_vol_file(path, vf); }
    catch (const HostIoError &e) { return fail(e.code, e.what()); }
    catch (const std::exception &e) { return fail(ALVRL_ERR_IO, e.what()); }
    if ((mn == nullptr) != (mx == nullptr)) return fail(ALVRL_ERR_ARG, "alvrl_set_medium_grid_file: give both bbox_min and bbox_max, or neither");
    return alvrl_set_medium_grid(c, vf.density.data(), vf.res, mn ? mn : vf.bmin, mx ? mx : vf.bmax, scale, albedo, sBase, phase, g);
}

int alvrl_set_camera(alvrl_handle c, const float s2c[16], const float c2w[16], uint32_t W, uint32_t H, float nearClip, float farClip) {
    API_BEGIN
    if (W == 0 || H == 0) throw Error(ALVRL_ERR_ARG, "empty film");
    memcpy(c->cam.s2c, s2c, 64); memcpy(c->cam.c2w, c2w, 64);
    c->cam.W = W; c->cam.H = H; c->cam.nearClip = nearClip; c->cam.farClip = farClip;
    c->cam.invResX = 1.0f / (float) W; c->cam.invResY = 1.0f / (float) H;
    c->haveCam = true; c->segsDirty = true; c->havePrimary = false; c->haveSlices = false; invalidate_from_slices(c);
    API_END
}

int alvrl_set_vrls(alvrl_handle c, const float *s, const float *e, const float *p, uint32_t n, uint64_t particleCount) {
    API_BEGIN
    use_device(c);
    c->vrlHost.clear();
    const bool noScatter = c->haveMedium && c->medium.sigmaS[0] == 0 && c->medium.sigmaS[1] == 0 && c->medium.sigmaS[2] == 0;
    for (uint32_t i = 0; i < n; i++) {
        const float *S = s + 3 * (size_t) i, *E = e + 3 * (size_t) i, *Pw = p + 3 * (size_t) i;
        for (int k = 0; k < 3; k++)
            if (!std::isfinite(Pw[k]) || Pw[k] < 0) throw Error(ALVRL_ERR_ARG, "invalid parsed VRL power");   /* VRL.h:51-53 */
        if (noScatter) continue;                                          /* vrlVector::put, VRL.h:148-158 */
        if (Pw[0] == 0 && Pw[1] == 0 && Pw[2] == 0) continue;
        const float dx = E[0] - S[0], dy = E[1] - S[1], dz = E[2] - S[2];
        const float len = std::sqrt(dx * dx + dy * dy + dz * dz);
        const float sx = S[0] - E[0], sy = S[1] - E[1], sz = S[2] - E[2];
        if (std::sqrt(sx * sx + sy * sy + sz * sz) == 0) continue;
        const float r = 1.0f / len;                                        /* normalize(end - start): recip multiply */
        VrlRec v;
        v.s = make_float4(S[0], S[1], S[2], len);
        v.e = make_float4(E[0], E[1], E[2], 1.0f);
        v.dir = make_float4(dx * r, dy * r, dz * r, 0);
        v.power = make_float4(Pw[0], Pw[1], Pw[2], 0);
        c->vrlHost.push_back(v);
    }
    c->particleCount = particleCount ? particleCount : c->vrlHost.size();
    if (c->vrlHost.empty()) throw Error(ALVRL_ERR_ARG, "no usable VRLs");
    c->dVrls.upload(c->vrlHost, c->stream);
    c->vrlSidesValid = false;
    c->haveVrls = true; c->haveR = false; c->haveClusters = false; c->haveFallback = false; c->renderListsDirty = true;
    c->stats.numVrls = (uint32_t) c->vrlHost.size();
    API_END
}

int alvrl_load_vrl_file(alvrl_handle c, const char *path) {
    std::vector<float> s, e, p;
    try { read_vrl_file(path, s, e, p); }                                   /* hostio.h: VRL.h:43-54, 120-128 */
    catch (const HostIoError &err) { return fail(err.code, err.what()); }
    return alvrl_set_vrls(c, s.data(), e.data(), p.data(), (uint32_t) (s.size() / 3), 0);    /* m_numParticles = size(), VRL.h:128 */
}

int alvrl_set_sample_tape(alvrl_handle c, const float *tape, uint64_t n) {
    API_BEGIN
    if (!tape) { c->userTape.clear(); }
    else c->userTape.assign(tape, tape + n);
    c->haveR = false;
    API_END
}

int alvrl_build_slices(alvrl_handle c) {
    API_BEGIN
    use_device(c);
    double t0 = now_ms();
    ensure_primary(c);
    const uint32_t P = c->numPixels();
    /* directionScale, Preprocessor.cpp:1137 (Scene::getAABB diagonal) */
    const float ax = c->sceneMin[0] - c->sceneMax[0], ay = c->sceneMin[1] - c->sceneMax[1], az = c->sceneMin[2] - c->sceneMax[2];
    const float directionScale = std::sqrt(ax * ax + ay * ay + az * az) / 8 * c->P.sliceCurvatureFactor;
    /* gather point (NaN for misses, Preprocessor.cpp:1172-1177) and scaled normal of every pixel: 24 bytes each come back */
    DevBuf<float> dPos, dDir; dPos.alloc(3 * (size_t) P); dDir.alloc(3 * (size_t) P);
    launch_gather_points(c->dPixSegs.p, P, directionScale, dPos.p, dDir.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    /* the partitions and extrema of getSlicesPQ run on the device (slices_dev.cu); ALVRL_SLICES_HOST=1, or a degenerate node,
     * takes the sequential host loop over the downloaded gather points instead -- same slices either way */
    bool onDevice = false;
    if (!getenv("ALVRL_SLICES_HOST")) {
        onDevice = build_slices_device(c, dPos.p, dDir.p, P, (uint32_t) c->P.targetNumSlices);
        if (onDevice) { c->haveSlices = true; invalidate_from_slices(c); c->stats.numSlices = c->numSlices(); }
    }
    if (!onDevice) {
        std::vector<P3> pos(P), dir(P);
        static_assert(sizeof(P3) == 12, "P3 is three packed floats");
        dPos.download(reinterpret_cast<float *>(pos.data()), 3 * (size_t) P, c->stream);
        dDir.download(reinterpret_cast<float *>(dir.data()), 3 * (size_t) P, c->stream);
        finish_slices(c, pos, dir);
    }
    c->stats.msSlices = (float) (now_ms() - t0);
    API_END
}

int alvrl_build_slices_from_gather(alvrl_handle c, const float *pos_, const float *dir_) {
    API_BEGIN
    use_device(c);
    if (!c->haveCam) throw Error(ALVRL_ERR_STATE, "set_camera first");
    const uint32_t P = c->numPixels();
    std::vector<P3> pos(P), dir(P);
    for (uint32_t i = 0; i < P; i++) { pos[i] = P3{pos_[3 * i], pos_[3 * i + 1], pos_[3 * i + 2]}; dir[i] = P3{dir_[3 * i], dir_[3 * i + 1], dir_[3 * i + 2]}; }
    finish_slices(c, pos, dir);
    API_END
}

int alvrl_sample_slice_mapping(alvrl_handle c) {
    API_BEGIN
    if (!c->haveSlices) throw Error(ALVRL_ERR_STATE, "build_slices first");
    double t0 = now_ms();
    use_device(c);
    const size_t S = c->numSlices();
    c->rowOffset.assign(S + 1, 0); c->sliceUndersampling.resize(S);
    size_t totalPix = 0, totalRep = 0;
    std::vector<uint32_t> positions;                       /* of the chosen pixels in the partition-ordered pixel array */
    for (size_t i = 0; i < S; i++) {
        c->mainSampler->setContext(ALVRL_RNG_SLICEMAP, (uint32_t) i, 0);
        const std::vector<uint32_t> idx = sampleRepresentativeIndices(c->sliceSize[i], c->P.targetPixelUndersampling, c->mainSampler.get());
        for (uint32_t k : idx) positions.push_back(c->sliceLo[i] + k);
        c->rowOffset[i + 1] = (uint32_t) positions.size();
        c->sliceUndersampling[i] = ((float) idx.size()) / c->sliceSize[i];      /* Preprocessor.cpp:1513 */
        totalRep += idx.size(); totalPix += c->sliceSize[i];
    }
    slice_gather_rows_device(c, positions, c->rowPixel);                          /* rowPixel = pixel ids of the positions */
    c->globalPixelUndersampling = ((float) totalRep) / totalPix;                          /* 1519 */
    c->haveRows = true; c->haveR = false; c->haveClusters = false; c->haveFallback = false;
    c->stats.msSliceMapping = (float) (now_ms() - t0);
    c->stats.numRows = (uint32_t) c->rowPixel.size();
    API_END
}

int alvrl_build_R(alvrl_handle c) {
    API_BEGIN
    use_device(c);
    if (!c->haveRows || !c->haveVrls || !c->haveMedium) throw Error(ALVRL_ERR_STATE, "sample_slice_mapping / set_vrls / set_medium first");
    double t0 = now_ms();
    ensure_primary(c);
    ensure_vrl_sides(c);
    const uint32_t N = (uint32_t) c->vrlHost.size(), G = (uint32_t) c->rowPixel.size(), S = c->numSlices();
    c->dRowPixel.upload(c->rowPixel, c->stream);                              /* (set_rep_pixels may have replaced the mirror) */
    c->dRowSegs.alloc(G);
    launch_gather_rows(c->dPixSegs.p, c->dRowPixel.p, G, c->dRowSegs.p, c->stream);
    c->stats.kernelLaunches++;
    /* every entry of the rows this handle owns is written by the kernel (inactive rows store zeros), and nothing reads the
     * others (alvrl_get_R answers zeros for them): no 13 GB memset per frame, and no storage for rows of other ranks */
    {
        const uint32_t sb_ = std::min(c->sliceBegin, S), se_ = std::min(c->sliceEnd, S);
        const uint32_t a = c->rowOffset[sb_] & ~31u, b = c->rowOffset[se_];
        c->rShift = a;
        c->ldR = std::max(32u, (b - a + 31u) & ~31u);
        c->dRstore.alloc((size_t) N * c->ldR);
        c->dR.p = c->dRstore.p - a;
    }
    TransportParams T = make_transport_params(c, ALVRL_RNG_R);
    if (!c->userTape.empty()) {
        if (c->userTape.size() < (uint64_t) G * N * c->K()) throw Error(ALVRL_ERR_ARG, "sample tape too short: need G*N*(2*Nvv+Nvs) floats");
        c->dTape.upload(c->userTape, c->stream);
        T.tape = c->dTape.p; T.tapeK = c->K();
    } else if (c->P.rngMode == ALVRL_RNG_MODE_SFMT) {
        ALVRL_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<float> tape;
        make_sfmt_tape(c, tape);
        c->dTape.upload(tape, c->stream);
        T.tape = c->dTape.p; T.tapeK = c->K();
    }
    const uint32_t sb = std::min(c->sliceBegin, S), se = std::min(c->sliceEnd, S);
    const uint32_t r0 = c->rowOffset[sb], r1 = c->rowOffset[se];
    T.rowBase = r0;
    c->timer.start(c->stream);
    if (c->mathMode == 1) launch_build_R_strict(T, c->dRowSegs.p + r0, r1 - r0, c->dVrls.p, c->dR.p + r0, c->ldR, c->stream);
    else launch_build_R_fast(T, c->dRowSegs.p + r0, r1 - r0, c->dVrls.p, c->dR.p + r0, c->ldR, c->stream);
    c->stats.msTransportKernelR = c->timer.stop(c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    ALVRL_CUDA(cudaStreamSynchronize(c->stream));
    /* rows whose pixel continues as a specular chain: the chain's in-medium segments are rows of a second matrix (their own
     * streams, their weights inside the estimate), added to the row of their pixel in chain order (getVRLContributions 812-813) */
    ensure_chains(c);
    uint64_t chainRows = 0;
    if (c->anyDelta && c->chainOffset.back()) {
        std::vector<uint32_t> xIdx, xKey, xFirst(r1 - r0 + 1, 0u);
        for (uint32_t r = r0; r < r1; r++) {
            const uint32_t pix = c->rowPixel[r];
            for (uint32_t e = c->chainOffset[pix]; e < c->chainOffset[pix + 1]; e++)
                if (c->chainMeta[e].w & 1u) {
                    if (r >= (1u << 24)) throw Error(ALVRL_ERR_UNSUPPORTED, "specular chains: more than 2^24 rows");
                    xIdx.push_back(e); xKey.push_back(r + ((c->chainMeta[e].y + 1u) << 24));
                }
            xFirst[r - r0 + 1] = (uint32_t) xIdx.size();
        }
        const uint32_t nX = (uint32_t) xIdx.size();
        if (nX) {
            if (T.tape) throw Error(ALVRL_ERR_UNSUPPORTED, "specular chains draw from the counter stream (no sample tape / rngMode=SFMT)");
            c->dXIdx.upload(xIdx, c->stream); c->dXKey.upload(xKey, c->stream); c->dXFirst.upload(xFirst, c->stream);
            c->dXSegs.alloc(nX);
            launch_gather_rows(c->dChainSegs.p, c->dXIdx.p, nX, c->dXSegs.p, c->stream);
            const uint32_t ldX = (nX + 31u) & ~31u;
            c->dX.alloc((size_t) N * ldX);
            TransportParams TX = T; TX.rowBase = 0;
            if (c->mathMode == 1) launch_build_R_strict(TX, c->dXSegs.p, nX, c->dVrls.p, c->dX.p, ldX, c->stream, c->dXKey.p, true);
            else launch_build_R_fast(TX, c->dXSegs.p, nX, c->dVrls.p, c->dX.p, ldX, c->stream, c->dXKey.p, true);
            launch_add_chain_rows(c->dR.p, c->ldR, r0, r1 - r0, c->dX.p, ldX, c->dXFirst.p, N, c->stream);
            c->stats.kernelLaunches += 3;
            ALVRL_CUDA(cudaGetLastError());
            ALVRL_CUDA(cudaStreamSynchronize(c->stream));
            chainRows = nX;
        }
    }
    c->stats.pairsPreprocess += chainRows * N;
    c->stats.shadowRays += chainRows * N * (uint64_t) (c->P.volVolSamples + c->P.volSurfSamples);
    c->haveR = true; c->haveClusters = false; c->haveFallback = false;
    c->builtRow0 = r0; c->builtRow1 = r1;
    c->stats.pairsPreprocess += (uint64_t) (r1 - r0) * N;
    c->stats.shadowRays += (uint64_t) (r1 - r0) * N * (uint64_t) (c->P.volVolSamples + c->P.volSurfSamples);
    c->stats.msBuildR = (float) (now_ms() - t0);
    API_END
}

int alvrl_get_column_nonzero(alvrl_handle c, uint8_t *flags) {
    API_BEGIN
    use_device(c);
    if (!c->haveR) throw Error(ALVRL_ERR_STATE, "build_R first");
    std::vector<uint8_t> f;
    column_nonzero_device(c, f);
    memcpy(flags, f.data(), f.size());
    API_END
}
int alvrl_set_column_nonzero(alvrl_handle c, const uint8_t *flags) {
    if (!flags) { c->columnFlagsOverride.clear(); return ALVRL_OK; }
    c->columnFlagsOverride.assign(flags, flags + c->vrlHost.size());
    return ALVRL_OK;
}
int alvrl_measure_fp32_peak(int device, float *tflops) {
    API_BEGIN
    ALVRL_CUDA(cudaSetDevice(device));
    *tflops = measure_fp32_peak_tflops();
    ALVRL_CUDA(cudaGetLastError());
    API_END
}

int alvrl_build_clusters(alvrl_handle c) {
    API_BEGIN
    use_device(c);
    if (!c->haveR) throw Error(ALVRL_ERR_STATE, "build_R first");
    double t0 = now_ms();
    build_clusters_device(c, c->P.rngMode == ALVRL_RNG_MODE_SFMT);
    c->haveClusters = true; c->renderListsDirty = true;
    c->stats.msClusters = (float) (now_ms() - t0);
    API_END
}

int alvrl_prepass(alvrl_handle c) {
    int rc;
    if ((rc = alvrl_sample_slice_mapping(c))) return rc;
    if ((rc = alvrl_build_R(c))) return rc;
    return alvrl_build_clusters(c);
}

int alvrl_render(alvrl_handle c, float *rgb) {
    API_BEGIN
    use_device(c);
    double t0 = now_ms();
    const uint32_t P = c->numPixels();
    c->dFb.alloc(P); c->dRgb.alloc(3 * (size_t) P);
    ALVRL_CUDA(cudaMemsetAsync(c->dFb.p, 0, (size_t) P * sizeof(float4), c->stream));
    render_clustered_into(c, c->dFb.p, c->stream);
    launch_fb_to_rgb(c->dFb.p, c->dRgb.p, P, c->stream);
    c->stats.kernelLaunches++;
    c->dRgb.download(rgb, 3 * (size_t) P, c->stream);
    c->stats.msRender = (float) (now_ms() - t0);
    API_END
}

int alvrl_render_unclustered(alvrl_handle c, float *rgb) {
    API_BEGIN
    use_device(c);
    if (!c->haveVrls || !c->haveMedium) throw Error(ALVRL_ERR_STATE, "set_vrls / set_medium first");
    double t0 = now_ms();
    ensure_primary(c);
    ensure_vrl_sides(c);
    ensure_chains(c);
    if (c->anyDelta && c->chainOffset.back()) throw Error(ALVRL_ERR_UNSUPPORTED, "alvrl_render_unclustered does not follow specular chains (use the clustered render)");
    const uint32_t P = c->numPixels(), N = (uint32_t) c->vrlHost.size();
    std::vector<uint32_t> px(P), off = {0, N};
    std::vector<uint4> work;
    for (uint32_t i = 0; i < P; i++) px[i] = i;
    for (uint32_t o = 0; o < P; o += ALVRL_CTA_SEGS_HOST) work.push_back(make_uint4(0, o, std::min<uint32_t>(ALVRL_CTA_SEGS_HOST, P - o), 0));
    DevBuf<uint32_t> dPx, dOff; DevBuf<uint4> dWork;
    dPx.upload(px, c->stream); dOff.upload(off, c->stream); dWork.upload(work, c->stream);
    c->dFb.alloc(P); c->dRgb.alloc(3 * (size_t) P);
    ALVRL_CUDA(cudaMemsetAsync(c->dFb.p, 0, (size_t) P * sizeof(float4), c->stream));
    TransportParams T = make_transport_params(c, ALVRL_RNG_RENDER);
    c->timer.start(c->stream);
    if (c->mathMode == 1) launch_render_strict(T, false, c->dPixSegs.p, dPx.p, dWork.p, (uint32_t) work.size(), c->dVrls.p, dOff.p, c->dFb.p, c->cam.W, c->cam.H, c->stream);
    else launch_render_fast(T, false, c->dPixSegs.p, dPx.p, dWork.p, (uint32_t) work.size(), c->dVrls.p, dOff.p, c->dFb.p, c->cam.W, c->cam.H, c->stream);
    c->stats.msTransportKernelRender = c->timer.stop(c->stream);
    ALVRL_CUDA(cudaGetLastError());
    launch_fb_to_rgb(c->dFb.p, c->dRgb.p, P, c->stream);
    c->stats.kernelLaunches += 2;
    c->dRgb.download(rgb, 3 * (size_t) P, c->stream);
    c->stats.pairsRender += (uint64_t) P * N;
    c->stats.msRender = (float) (now_ms() - t0);
    API_END
}

int alvrl_set_slice_range(alvrl_handle c, uint32_t b, uint32_t e) {
    if (b > e) return fail(ALVRL_ERR_ARG, "slice range: begin > end");
    if (b != c->sliceBegin || e != c->sliceEnd) {      /* R rows and clusters outside the old range do not exist on this handle */
        c->haveR = false; c->haveClusters = false; c->haveFallback = false; c->renderListsDirty = true;
    }
    c->sliceBegin = b; c->sliceEnd = e;
    return ALVRL_OK;
}

int alvrl_set_seed(alvrl_handle c, uint64_t seed) {
    if (c->P.rngMode == ALVRL_RNG_MODE_SFMT) return fail(ALVRL_ERR_UNSUPPORTED, "alvrl_set_seed: the sequential SFMT stream continues from pass to pass by itself");
    if (seed != c->P.seed) {
        c->P.seed = seed;
        c->mainSampler.reset(new_stream(c->P));
        c->globalStream.reset();
        c->chainsValid = false; c->haveRows = false; c->haveR = false; c->haveClusters = false; c->haveFallback = false; c->renderListsDirty = true;
    }
    return ALVRL_OK;
}

/* ---- VRL tracer (tracer.cu) ------------------------------------------------------------------------ */
int alvrl_set_area_emitter(alvrl_handle c, const uint32_t *tris, uint32_t n, const float radiance[3]) {
    API_BEGIN
    use_device(c);
    if (!c->haveMesh || !tris || !n || !radiance) throw Error(ALVRL_ERR_ARG, "alvrl_set_area_emitter: set_mesh first; at least one triangle");
    const uint32_t nt = (uint32_t) c->triMat.size();
    /* TriMesh::prepareSamplingTable (trimesh.cpp:388-403): DiscreteDistribution::append / normalize (pmf.h:48-52, 101-114) over
     * Triangle::surfaceArea (triangle.cpp:61-67), in float */
    c->emTris.assign(tris, tris + n);
    c->emCdf.assign(1, 0.0f);
    for (uint32_t i = 0; i < n; i++) {
        if (tris[i] >= nt) throw Error(ALVRL_ERR_ARG, "alvrl_set_area_emitter: triangle index out of range");
        const float *p0 = &c->verts[3 * (size_t) c->tris[3 * (size_t) tris[i]]], *p1 = &c->verts[3 * (size_t) c->tris[3 * (size_t) tris[i] + 1]],
                    *p2 = &c->verts[3 * (size_t) c->tris[3 * (size_t) tris[i] + 2]];
        const float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
        c->emCdf.push_back(c->emCdf.back() + 0.5f * std::sqrt(cx * cx + cy * cy + cz * cz));
    }
    const float sum = c->emCdf.back(), normalization = 1.0f / sum;
    if (!(sum > 0)) throw Error(ALVRL_ERR_ARG, "alvrl_set_area_emitter: the emitter has no area");
    for (size_t i = 1; i < c->emCdf.size(); ++i) c->emCdf[i] *= normalization;
    c->emCdf.back() = 1.0f;
    for (int k = 0; k < 3; k++) c->emPower[k] = radiance[k] * (float) M_PI * sum;          /* area.cpp:198 */
    c->dEmTris.upload(c->emTris, c->stream); c->dEmCdf.upload(c->emCdf, c->stream);
    for (int k = 0; k < 3; k++) c->emRadiance[k] = radiance[k];
    c->emInvArea = normalization;                                                           /* TriMesh::m_invSurfaceArea, trimesh.cpp:401 */
    std::vector<uint8_t> isEm(nt, 0);                                                       /* Shape::isEmitter of the triangle's shape */
    for (uint32_t i = 0; i < n; i++) isEm[tris[i]] = 1;
    c->dTriEmitter.upload(isEm, c->stream);
    c->haveEmitter = true;
    API_END
}

int alvrl_trace_vrls(alvrl_handle c, uint32_t target) {
    API_BEGIN
    use_device(c);
    if (!c->haveEmitter || !c->haveMedium || !c->haveMat) throw Error(ALVRL_ERR_STATE, "alvrl_trace_vrls: set_area_emitter / set_medium / set_materials first");
    ensure_scene(c);
    bool anyDelta = false;
    for (uint32_t b : c->matBits) if (b & ALVRL_BSDF_DELTA) anyDelta = true;
    if (anyDelta && c->optics.size() != 12 * c->matBits.size()) throw Error(ALVRL_ERR_STATE, "materials with delta components need alvrl_set_material_optics");
    if (!anyDelta && c->dMatOptics.n == 0) { std::vector<float4> z(3 * std::max<size_t>(1, c->matBits.size()), make_float4(0, 0, 0, 0)); c->dMatOptics.upload(z, c->stream); }
    if (!target) target = (uint32_t) c->P.vrlTargetNum;
    /* randomWalk (vrlTracer.h:29-40) stops after the particle that brings the count to the target: trace batches of independent
     * particles, find that particle on the host */
    std::vector<uint32_t> counts;
    uint64_t have = 0; uint32_t traced = 0, cut = 0;
    uint32_t batch = std::max<uint32_t>(1024u, target / 16u);
    DevBuf<uint32_t> dCnt;
    while (!cut) {
        dCnt.alloc(batch);
        launch_trace_count(c->sceneDev, c->medium, c->dEmTris.p, c->dEmCdf.p, (uint32_t) c->emTris.size(), c->emPower, c->P.seed, c->P.shortVrls,
                           c->P.maxParticleDepth, c->P.rrDepth, c->dTriVerts.p, c->dTriMat.p, c->dMatAlbedo.p, c->dMatBits.p, c->dMatOptics.p, traced, batch, dCnt.p, c->stream);
        c->stats.kernelLaunches++;
        ALVRL_CUDA(cudaGetLastError());
        counts.resize((size_t) traced + batch);
        dCnt.download(counts.data() + traced, batch, c->stream);
        for (uint32_t i = traced; i < traced + batch && !cut; i++) { have += counts[i]; if (have >= target) cut = i + 1; }
        traced += batch;
        if (!cut && traced > (1u << 30)) throw Error(ALVRL_ERR_ARG, "alvrl_trace_vrls: no VRLs are being generated (is the emitter inside the medium?)");
        batch = std::min<uint32_t>(batch * 2u, 1u << 22);
    }
    std::vector<uint32_t> offset((size_t) cut + 1, 0u);
    for (uint32_t i = 0; i < cut; i++) offset[i + 1] = offset[i] + counts[i];
    const uint32_t n = offset[cut];
    DevBuf<uint32_t> dOff; dOff.upload(offset, c->stream);
    DevBuf<float> dOut; dOut.alloc(9 * (size_t) n);
    launch_trace_write(c->sceneDev, c->medium, c->dEmTris.p, c->dEmCdf.p, (uint32_t) c->emTris.size(), c->emPower, c->P.seed, c->P.shortVrls,
                       c->P.maxParticleDepth, c->P.rrDepth, c->dTriVerts.p, c->dTriMat.p, c->dMatAlbedo.p, c->dMatBits.p, c->dMatOptics.p, cut, dOff.p, dOut.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    std::vector<float> rec(9 * (size_t) n);
    dOut.download(rec.data(), rec.size(), c->stream);
    std::vector<float> s(3 * (size_t) n), e(3 * (size_t) n), p(3 * (size_t) n);
    for (uint32_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) { s[3 * (size_t) i + k] = rec[9 * (size_t) i + k]; e[3 * (size_t) i + k] = rec[9 * (size_t) i + 3 + k]; p[3 * (size_t) i + k] = rec[9 * (size_t) i + 6 + k]; }
    const int rc = alvrl_set_vrls(c, s.data(), e.data(), p.data(), n, cut);               /* the put() filter already ran on the device: nothing is dropped */
    if (rc != ALVRL_OK) return rc;
    API_END
}

int alvrl_volpath_render(alvrl_handle c, uint32_t spp, uint32_t internalSamples, uint32_t flags, int32_t maxDepth, float *rgb) {
    API_BEGIN
    use_device(c);
    if (!c->haveEmitter || !c->haveMedium || !c->haveMat || !c->haveCam) throw Error(ALVRL_ERR_STATE, "alvrl_volpath_render: set_area_emitter / set_medium / set_materials / set_camera first");
    if (!spp || !internalSamples || !rgb) throw Error(ALVRL_ERR_ARG, "alvrl_volpath_render: spp and internalSamples must be positive");
    if (flags & ~127u) throw Error(ALVRL_ERR_ARG, "alvrl_volpath_render: unknown flag bits");
    ensure_scene(c);
    bool anyDelta = false;
    for (uint32_t b : c->matBits) if (b & ALVRL_BSDF_DELTA) anyDelta = true;
    if (anyDelta && c->optics.size() != 12 * c->matBits.size()) throw Error(ALVRL_ERR_STATE, "materials with delta components need alvrl_set_material_optics");
    if (!anyDelta && c->dMatOptics.n == 0) { std::vector<float4> z(3 * std::max<size_t>(1, c->matBits.size()), make_float4(0, 0, 0, 0)); c->dMatOptics.upload(z, c->stream); }
    const uint32_t P = c->numPixels();
    c->dVolpathAcc.alloc(P);
    ALVRL_CUDA(cudaMemsetAsync(c->dVolpathAcc.p, 0, (size_t) P * sizeof(float4), c->stream));
    const bool centre = spp == 1 || (flags & ALVRL_VOLPATH_CENTRE_SAMPLES);                /* integrator.cpp:240-246 */
    for (uint32_t j = 0; j < spp; j++) {                                                    /* one outer sample of every pixel per launch: the film's order */
        launch_volpath_sample(c->sceneDev, c->medium, c->cam, c->dEmTris.p, c->dEmCdf.p, (uint32_t) c->emTris.size(), c->emRadiance, c->emInvArea, c->dTriEmitter.p,
                              c->dTriVerts.p, c->dTriMat.p, c->dMatAlbedo.p, c->dMatBits.p, c->dMatOptics.p, c->P.seed, j, (int) internalSamples, flags, centre,
                              maxDepth, c->P.rrDepth, c->dVolpathAcc.p, c->stream);
        c->stats.kernelLaunches++;
    }
    ALVRL_CUDA(cudaGetLastError());
    c->dVolpathRgb.alloc(3 * (size_t) P);
    launch_volpath_develop(c->dVolpathAcc.p, c->cam.W, c->cam.H, c->dVolpathRgb.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    c->dVolpathRgb.download(rgb, 3 * (size_t) P, c->stream);
    API_END
}

int alvrl_get_vrls(alvrl_handle c, float *s, float *e, float *p, uint64_t *pc) {
    API_BEGIN
    if (!c->haveVrls) throw Error(ALVRL_ERR_STATE, "no VRLs");
    const size_t n = c->vrlHost.size();
    for (size_t i = 0; i < n; i++) {
        const VrlRec &v = c->vrlHost[i];
        if (s) { s[3 * i] = v.s.x; s[3 * i + 1] = v.s.y; s[3 * i + 2] = v.s.z; }
        if (e) { e[3 * i] = v.e.x; e[3 * i + 1] = v.e.y; e[3 * i + 2] = v.e.z; }
        if (p) { p[3 * i] = v.power.x; p[3 * i + 1] = v.power.y; p[3 * i + 2] = v.power.z; }
    }
    if (pc) *pc = c->particleCount;
    API_END
}

/* ---- film (film.cu) ------------------------------------------------------------------------------- */
int alvrl_film_configure(alvrl_handle c, int filter, float param) {
    API_BEGIN
    if (!c->haveCam) throw Error(ALVRL_ERR_STATE, "set_camera first");
    use_device(c);
    /* filter radius and profile: box.cpp:38,46-48, tent.cpp:34,42-44, gaussian.cpp:30-35,52-58 */
    float radius, stddev = 0.5f;
    if (filter == ALVRL_FILTER_BOX) radius = (param > 0 ? param : 0.5f) + 1e-5f;
    else if (filter == ALVRL_FILTER_TENT) radius = 1.0f;
    else if (filter == ALVRL_FILTER_GAUSSIAN) { stddev = param > 0 ? param : 0.5f; radius = 4 * stddev; }
    else throw Error(ALVRL_ERR_ARG, "alvrl_film_configure: unknown filter");
    auto eval = [&](float x) -> float {
        if (filter == ALVRL_FILTER_BOX) return std::fabs(x) <= radius ? 1.0f : 0.0f;
        if (filter == ALVRL_FILTER_TENT) return std::max(0.0f, 1.0f - std::fabs(x / radius));
        const float alpha = -1.0f / (2.0f * stddev * stddev);
        return std::max(0.0f, (float) std::exp((double) (alpha * x * x)) - (float) std::exp((double) (alpha * radius * radius)));
    };
    /* ReconstructionFilter::configure, rfilter.cpp:37-55 */
    FilmFilterDev &f = c->film;
    float sum = 0.0f;
    for (int i = 0; i < ALVRL_FILTER_RESOLUTION; i++) { const float v = eval((radius * i) / ALVRL_FILTER_RESOLUTION); f.table[i] = v; sum += v; }
    f.table[ALVRL_FILTER_RESOLUTION] = 0.0f;
    f.scaleFactor = ALVRL_FILTER_RESOLUTION / radius;
    sum *= 2 * radius / ALVRL_FILTER_RESOLUTION;
    const float normalization = 1.0f / sum;
    for (int i = 0; i < ALVRL_FILTER_RESOLUTION; i++) f.table[i] *= normalization;
    f.radius = radius; f.taps = (int) std::floor(radius);
    c->haveFilm = true;
    c->dFilm.alloc(5 * (size_t) c->numPixels());
    ALVRL_CUDA(cudaMemsetAsync(c->dFilm.p, 0, 5 * (size_t) c->numPixels() * sizeof(float), c->stream));
    c->filmPasses = 0;
    API_END
}

int alvrl_film_clear(alvrl_handle c) {
    API_BEGIN
    if (!c->haveFilm) throw Error(ALVRL_ERR_STATE, "alvrl_film_configure first");
    use_device(c);
    ALVRL_CUDA(cudaMemsetAsync(c->dFilm.p, 0, 5 * (size_t) c->numPixels() * sizeof(float), c->stream));
    c->filmPasses = 0;
    API_END
}

int alvrl_film_put(alvrl_handle c, const float *rgb) {
    API_BEGIN
    if (!c->haveFilm) throw Error(ALVRL_ERR_STATE, "alvrl_film_configure first");
    use_device(c);
    const uint32_t P = c->numPixels(), W = c->cam.W, H = c->cam.H;
    if (rgb) {                                           /* a frame from the host: [y][x][c] -> the framebuffer layout */
        std::vector<float4> fb(P);
        for (uint32_t i = 0; i < P; i++) fb[i] = make_float4(rgb[3 * (size_t) i], rgb[3 * (size_t) i + 1], rgb[3 * (size_t) i + 2], 1.0f);
        c->dFb.alloc(P);
        c->dFb.upload(fb, c->stream);
    } else if (c->dFb.n < P) throw Error(ALVRL_ERR_STATE, "alvrl_film_put(NULL): no rendered frame on the device (alvrl_render first)");
    launch_film_splat(c->dFb.p, W, H, c->film, c->dFilm.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    ALVRL_CUDA(cudaStreamSynchronize(c->stream));
    c->filmPasses++;
    API_END
}

int alvrl_film_develop(alvrl_handle c, float *rgb) {
    API_BEGIN
    if (!c->haveFilm) throw Error(ALVRL_ERR_STATE, "alvrl_film_configure first");
    if (!rgb) throw Error(ALVRL_ERR_ARG, "alvrl_film_develop: null output");
    use_device(c);
    const uint32_t P = c->numPixels();
    c->dRgb.alloc(3 * (size_t) P);
    launch_film_develop(c->dFilm.p, P, c->dRgb.p, c->stream);
    c->stats.kernelLaunches++;
    c->dRgb.download(rgb, 3 * (size_t) P, c->stream);
    API_END
}

int alvrl_film_write_npy(alvrl_handle c, const char *path) {
    if (!path) return fail(ALVRL_ERR_ARG, "alvrl_film_write_npy: null file name");
    std::vector<float> rgb(3 * (size_t) c->numPixels());
    const int rc = alvrl_film_develop(c, rgb.data());
    if (rc != ALVRL_OK) return rc;
    try { write_npy_f32(path, rgb.data(), c->cam.H, c->cam.W, 3); }
    catch (const HostIoError &e) { return fail(e.code, e.what()); }
    return ALVRL_OK;
}

int alvrl_render_device(alvrl_handle c, void *rgba, void *stream) {
    API_BEGIN
    use_device(c);
    render_clustered_into(c, (float4 *) rgba, stream ? (cudaStream_t) stream : c->stream);
    API_END
}

/* ---- introspection ----------------------------------------------------------------------------- */
int alvrl_get_stats(alvrl_handle c, alvrl_stats *out) { *out = c->stats; return ALVRL_OK; }
int alvrl_get_num_vrls(alvrl_handle c, uint32_t *n) { *n = (uint32_t) c->vrlHost.size(); return ALVRL_OK; }

int alvrl_get_primary_hits(alvrl_handle c, uint32_t *prim, float *t, float *p, float *n) {
    API_BEGIN
    use_device(c);
    ensure_primary(c);
    const uint32_t P = c->numPixels();
    if (prim) c->dHitPrim.download(prim, P, c->stream);
    if (t) c->dHitT.download(t, P, c->stream);
    if (p || n) {
        std::vector<SegRec> segs(P);
        c->dPixSegs.download(segs.data(), P, c->stream);
        for (uint32_t i = 0; i < P; i++) {
            if (p) { p[3 * i] = segs[i].p.x; p[3 * i + 1] = segs[i].p.y; p[3 * i + 2] = segs[i].p.z; }
            if (n) { n[3 * i] = segs[i].n.x; n[3 * i + 1] = segs[i].n.y; n[3 * i + 2] = segs[i].n.z; }
        }
    }
    API_END
}

int alvrl_get_pixel_to_slice(alvrl_handle c, uint32_t *out) {
    if (!c->haveSlices) return fail(ALVRL_ERR_STATE, "build_slices first");
    API_BEGIN
    use_device(c);
    c->dPixelToSlice.download(out, c->numPixels(), c->stream);
    API_END
}
int alvrl_get_num_slices(alvrl_handle c, uint32_t *ns, uint32_t *nr) {
    *ns = c->numSlices(); *nr = c->haveRows ? (uint32_t) c->rowPixel.size() : 0;
    return ALVRL_OK;
}
int alvrl_get_rep_pixels(alvrl_handle c, uint32_t *off, uint32_t *px) {
    if (!c->haveRows) return fail(ALVRL_ERR_STATE, "sample_slice_mapping first");
    memcpy(off, c->rowOffset.data(), c->rowOffset.size() * 4); memcpy(px, c->rowPixel.data(), c->rowPixel.size() * 4);
    return ALVRL_OK;
}
int alvrl_set_rep_pixels(alvrl_handle c, const uint32_t *off, const uint32_t *px, uint32_t ns) {
    if (!c->haveSlices || ns != c->numSlices()) return fail(ALVRL_ERR_STATE, "slice count mismatch");
    if (off[0] != 0) return fail(ALVRL_ERR_ARG, "set_rep_pixels: sliceRowOffset[0] must be 0");
    for (uint32_t i = 0; i < ns; i++) if (off[i + 1] < off[i]) return fail(ALVRL_ERR_ARG, "set_rep_pixels: sliceRowOffset must be non-decreasing");
    for (uint32_t i = 0; i < off[ns]; i++) if (px[i] >= c->numPixels()) return fail(ALVRL_ERR_ARG, "set_rep_pixels: pixel index out of range");
    c->rowOffset.assign(off, off + ns + 1); c->rowPixel.assign(px, px + off[ns]);
    c->sliceUndersampling.resize(ns);
    size_t totalPix = 0;
    for (uint32_t i = 0; i < ns; i++) {
        c->sliceUndersampling[i] = ((float) (off[i + 1] - off[i])) / c->sliceSize[i];
        totalPix += c->sliceSize[i];
    }
    c->globalPixelUndersampling = ((float) off[ns]) / totalPix;
    c->haveRows = true; c->haveR = false; c->haveClusters = false; c->haveFallback = false;
    c->stats.numRows = off[ns];
    return ALVRL_OK;
}
int alvrl_get_R(alvrl_handle c, uint32_t r0, uint32_t r1, float *mv) {
    API_BEGIN
    use_device(c);
    if (!c->haveR) throw Error(ALVRL_ERR_STATE, "build_R first");
    if (r0 > r1 || r1 > c->rowPixel.size()) throw Error(ALVRL_ERR_ARG, "get_R: row range out of bounds");
    const uint32_t N = (uint32_t) c->vrlHost.size();
    std::vector<float2> col(c->rowPixel.size() + 32);
    /* rows outside the range this handle built (slice sharding) were never written: they read as zeros */
    const uint32_t b0 = std::max(r0, c->builtRow0), b1 = std::min(r1, c->builtRow1);
    for (uint32_t v = 0; v < N; v++) {
        std::fill(col.begin() + r0, col.begin() + r1, make_float2(0, 0));
        if (b0 < b1) c->dRstore.download(col.data() + b0, b1 - b0, c->stream, (size_t) v * c->ldR + b0 - c->rShift);
        for (uint32_t r = r0; r < r1; r++) { mv[((size_t) (r - r0) * N + v) * 2] = col[r].x; mv[((size_t) (r - r0) * N + v) * 2 + 1] = col[r].y; }
    }
    API_END
}
int alvrl_set_R(alvrl_handle c, const float *mv) {
    API_BEGIN
    use_device(c);
    if (!c->haveRows || !c->haveVrls) throw Error(ALVRL_ERR_STATE, "rows / vrls first");
    const uint32_t N = (uint32_t) c->vrlHost.size(), G = (uint32_t) c->rowPixel.size();
    c->ldR = (G + 31u) & ~31u;
    std::vector<float2> t((size_t) N * c->ldR, make_float2(0, 0));
    for (uint32_t r = 0; r < G; r++)
        for (uint32_t v = 0; v < N; v++) t[(size_t) v * c->ldR + r] = make_float2(mv[((size_t) r * N + v) * 2], mv[((size_t) r * N + v) * 2 + 1]);
    c->dRstore.upload(t, c->stream); c->dR.p = c->dRstore.p; c->rShift = 0;
    c->builtRow0 = 0; c->builtRow1 = G;
    c->haveR = true; c->haveClusters = false; c->haveFallback = false;
    API_END
}
int alvrl_get_cluster_counts(alvrl_handle c, uint32_t *off, uint32_t *ng, uint32_t *nf) {
    API_BEGIN
    use_device(c);
    if (!c->haveClusters) throw Error(ALVRL_ERR_STATE, "build_clusters first");
    const uint32_t S_ = c->numSlices();
    const bool ranged = !(std::min(c->sliceBegin, S_) == 0 && std::min(c->sliceEnd, S_) == S_);
    /* lazily: global + fallback lists (they span all rows of R: a handle that owns a slice range reports them empty) */
    if (!c->haveFallback && c->haveR && !ranged) build_clusters_device(c, true);
    off[0] = 0;
    for (size_t i = 0; i < c->selectedVrls.size(); i++) off[i + 1] = off[i] + (uint32_t) c->selectedVrls[i].size();
    *ng = (uint32_t) c->gcVrls.size(); *nf = (uint32_t) c->fallBackVrls.size();
    API_END
}
int alvrl_get_clusters(alvrl_handle c, uint32_t *vrls, float *weights, uint32_t *gv, float *gw, uint32_t *fv, float *fw) {
    if (!c->haveClusters) return fail(ALVRL_ERR_STATE, "build_clusters first");
    size_t o = 0;
    for (size_t i = 0; i < c->selectedVrls.size(); i++)
        for (size_t j = 0; j < c->selectedVrls[i].size(); j++, o++) { vrls[o] = c->selectedVrls[i][j]; weights[o] = c->clusterWeight[i][j]; }
    for (size_t j = 0; j < c->gcVrls.size(); j++) { if (gv) gv[j] = c->gcVrls[j]; if (gw) gw[j] = c->gcWeight[j]; }
    for (size_t j = 0; j < c->fallBackVrls.size(); j++) { if (fv) fv[j] = c->fallBackVrls[j]; if (fw) fw[j] = c->fallBackWeight[j]; }
    return ALVRL_OK;
}
int alvrl_set_clusters(alvrl_handle c, const uint32_t *off, uint32_t ns, const uint32_t *vrls, const float *weights,
                       const uint32_t *fv, const float *fw, uint32_t nf) {
    if (!c->haveSlices || ns != c->numSlices()) return fail(ALVRL_ERR_STATE, "slice count mismatch");
    if (off[0] != 0) return fail(ALVRL_ERR_ARG, "set_clusters: sliceOffset[0] must be 0");
    for (uint32_t i = 0; i < ns; i++) if (off[i + 1] < off[i]) return fail(ALVRL_ERR_ARG, "set_clusters: sliceOffset must be non-decreasing");
    for (uint32_t i = 0; i < off[ns]; i++) if (vrls[i] >= c->vrlHost.size()) return fail(ALVRL_ERR_ARG, "set_clusters: representative out of range");
    for (uint32_t i = 0; i < nf; i++) if (fv[i] >= c->vrlHost.size()) return fail(ALVRL_ERR_ARG, "set_clusters: fallback representative out of range");
    c->selectedVrls.assign(ns, {}); c->clusterWeight.assign(ns, {});
    for (uint32_t i = 0; i < ns; i++) {
        c->selectedVrls[i].assign(vrls + off[i], vrls + off[i + 1]);
        c->clusterWeight[i].assign(weights + off[i], weights + off[i + 1]);
    }
    c->fallBackVrls.assign(fv, fv + nf); c->fallBackWeight.assign(fw, fw + nf);
    c->haveClusters = true; c->haveFallback = true; c->renderListsDirty = true;
    return ALVRL_OK;
}
int alvrl_trace_rays(alvrl_handle c, const float *o, const float *d, const float *mint, const float *maxt, uint32_t n, uint32_t *prim, float *t) {
    API_BEGIN
    use_device(c);
    ensure_scene(c);
    DevBuf<float> dO, dD, dMin, dMax, dT; DevBuf<uint32_t> dP;
    dO.upload(o, 3 * (size_t) n, c->stream); dD.upload(d, 3 * (size_t) n, c->stream);
    dMin.upload(mint, n, c->stream); dMax.upload(maxt, n, c->stream); dT.alloc(n); dP.alloc(n);
    launch_trace_rays(c->sceneDev, dO.p, dD.p, dMin.p, dMax.p, n, dP.p, dT.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    dP.download(prim, n, c->stream);
    if (t) dT.download(t, n, c->stream);
    API_END
}
int alvrl_eval_transmittance(alvrl_handle c, const float *p1, const int32_t *onSurf, const float *p2, uint32_t n, float *T) {
    API_BEGIN
    use_device(c);
    ensure_scene(c);
    if (!c->haveMedium) throw Error(ALVRL_ERR_STATE, "set_medium first");
    DevBuf<float> d1, d2, dT; DevBuf<int32_t> dS;
    d1.upload(p1, 3 * (size_t) n, c->stream); d2.upload(p2, 3 * (size_t) n, c->stream); dT.alloc(3 * (size_t) n);
    if (onSurf) dS.upload(onSurf, n, c->stream);
    launch_eval_transmittance(c->sceneDev, c->medium, d1.p, onSurf ? dS.p : nullptr, d2.p, n, dT.p, c->stream);
    c->stats.kernelLaunches++;
    ALVRL_CUDA(cudaGetLastError());
    dT.download(T, 3 * (size_t) n, c->stream);
    API_END
}

} // extern "C"
