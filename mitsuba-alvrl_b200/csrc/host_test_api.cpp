/*
 * host_test_api.cpp -- exposes the *host-side* logic of the plugin (SFMT stream, slice builder, representative
 * pixel sampling) through a few C entry points so that `-m "not gpu"` tests can compare it with the oracle on a
 * machine without a GPU.  Built as libalvrl_host.so by g++; not part of the product path (libalvrl.so links the
 * same headers directly).
 */
#include "slices.h"
#include "host_sampler.h"
#include "heap_order.h"
#include "occluders.h"
#include "occ_query.h"
#include "sharding.h"
#include "hostio.h"
#include "shapes.h"
#include "bvh.h"
#include <algorithm>
#include <cstring>
using namespace alvrl;

extern "C" {

/* slice range of one rank (group.cu uses the same function) */
int alvrl_host_balanced_range(const uint32_t *sizes, uint32_t S, int world, int rank, uint32_t *b, uint32_t *e) {
    balanced_slice_range(sizes, S, world, rank, *b, *e);
    return 0;
}

/* measured-time balancing of the slice ranges (sharding.h; group.cu runs the same functions after every frame):
 * cost[S] in/out (corrections == 0 and cost[0] < 0: initialised from the pixel counts), weights[S] out = what the NEXT frame
 * is cut on, t[world] = what each rank's range took in the frame cut on weightsIn (NULL: pixel counts + constant, the first
 * frame's cut) */
int alvrl_host_balance_step(const uint32_t *sliceSize, uint32_t S, int world, double *cost, const uint32_t *weightsIn, const float *t,
                            uint32_t corrections, uint32_t *weightsOut) {
    std::vector<double> c(cost, cost + S);
    if (corrections == 0 && S && cost[0] < 0) initial_slice_costs(sliceSize, S, c);
    std::vector<uint32_t> w(S);
    if (weightsIn) w.assign(weightsIn, weightsIn + S);
    else { uint64_t tot = 0; for (uint32_t i = 0; i < S; i++) tot += sliceSize[i]; for (uint32_t i = 0; i < S; i++) w[i] = sliceSize[i] + (uint32_t) (tot / 500u); }
    const bool counted = correct_slice_costs(c, w.data(), world, t, corrections);
    std::vector<uint32_t> next;
    cut_weights(c, next);
    for (uint32_t i = 0; i < S; i++) { cost[i] = c[i]; weightsOut[i] = next[i]; }
    return counted ? 1 : 0;
}

/* hostio.h: the grid volume reader behind alvrl_set_medium_grid_file and the NumPy writer behind alvrl_film_write_npy.
 * header[11] = {type, xres, yres, zres, channels, then the AABB's six floats as bit patterns}; density NULL: header only.
 * Returns 0 or the ABI's error code, with the message in err. */
int alvrl_host_read_vol(const char *path, int32_t *header, float *density, char *err, uint32_t errLen) {
    try {
        VolFile vf;
        read_vol_file(path, vf, density == nullptr);
        header[0] = vf.type; header[1] = vf.res[0]; header[2] = vf.res[1]; header[3] = vf.res[2]; header[4] = vf.channels;
        memcpy(header + 5, vf.bmin, 12); memcpy(header + 8, vf.bmax, 12);
        if (density) memcpy(density, vf.density.data(), vf.density.size() * sizeof(float));
        return 0;
    } catch (const HostIoError &e) {
        if (err && errLen) { strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; }
        return e.code;
    } catch (const std::exception &e) {                     /* (bad_alloc: alvrl_set_medium_grid_file answers ALVRL_ERR_IO too) */
        if (err && errLen) { strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; }
        return -4;
    }
}
/* out: 9 floats per VRL read (start, end, power); out NULL: count only */
int alvrl_host_read_vrl_file(const char *path, float *out, uint32_t *n) {
    try {
        std::vector<float> s, e, p;
        read_vrl_file(path, s, e, p);
        *n = (uint32_t) (s.size() / 3);
        if (out) for (uint32_t i = 0; i < *n; i++) { memcpy(out + 9 * i, &s[3 * i], 12); memcpy(out + 9 * i + 3, &e[3 * i], 12); memcpy(out + 9 * i + 6, &p[3 * i], 12); }
        return 0;
    } catch (const HostIoError &err) { return err.code; }
}
int alvrl_host_write_npy(const char *path, const float *data, uint32_t height, uint32_t width, uint32_t channels, char *err, uint32_t errLen) {
    try { write_npy_f32(path, data, height, width, channels); return 0; }
    catch (const HostIoError &e) {
        if (err && errLen) { strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; }
        return e.code;
    }
}

/* shapes.h: the tessellations behind alvrl_add_rectangle / alvrl_add_sphere.  verts / tris NULL: counts only. */
int alvrl_host_tessellate(int shape, const float *toWorldOrCenter, float radius, int flipNormals, uint32_t thetaSteps,
                          float *verts, uint32_t *nverts, uint32_t *tris, uint32_t *ntris) {
    std::vector<float> v; std::vector<uint32_t> t;
    try {
        if (shape == 0) tessellate_rectangle(toWorldOrCenter, flipNormals != 0, v, t);
        else tessellate_sphere(toWorldOrCenter, radius, flipNormals != 0, thetaSteps, v, t);
    } catch (const std::exception &) { return -1; }
    *nverts = (uint32_t) (v.size() / 3); *ntris = (uint32_t) (t.size() / 3);
    if (verts) memcpy(verts, v.data(), v.size() * sizeof(float));
    if (tris) memcpy(tris, t.data(), t.size() * sizeof(uint32_t));
    return 0;
}

/* bvh.h: the binary threaded BVH and the 4-wide tree collapsed from it.  Structure: preorder, escape indices, leaf ranges that
 * partition the triangle order, boxes that hold their triangles.  Culling: for every ray segment, each triangle a brute-force
 * test (Moeller-Trumbore in double) hits must sit in a leaf that a traversal with an exact slab test on the stored (padded)
 * boxes reaches -- in the stackless binary tree (hit inner -> node + 1, leaf / miss -> escape) and in the 4-wide tree.
 * stats = {structure error (0 = none), hits, hits missed by the binary tree, hits missed by the 4-wide tree, binary node
 * visits, 4-wide node visits, binary nodes, 4-wide nodes} */
int alvrl_host_bvh_check(const float *verts, const uint32_t *tris, uint32_t nt, const float *o, const float *d, const float *tmax,
                         uint32_t nrays, uint64_t *stats) {
    for (int i = 0; i < 8; i++) stats[i] = 0;
    HostBvh bvh;
    BvhBuilder(verts, tris, nt).build(bvh);
    std::vector<Bvh4Node> wide;
    collapse4(bvh, wide);
    auto asU = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    const uint32_t nn = (uint32_t) bvh.nodes.size();
    stats[6] = nn; stats[7] = wide.size();
    /* structure */
    if (bvh.triOrder.size() != nt) { stats[0] = 1; return 0; }
    { std::vector<uint32_t> seen(nt, 0); for (uint32_t t : bvh.triOrder) { if (t >= nt || seen[t]++) { stats[0] = 2; return 0; } } }
    std::vector<uint32_t> covered(nt, 0);
    for (uint32_t k = 0; k < nn; k++) {
        const BvhNode &n = bvh.nodes[k];
        const uint32_t esc = asU(n.lo.w), lf = asU(n.hi.w);
        if (esc <= k || esc > nn) { stats[0] = 3; return 0; }
        if (lf) {
            const uint32_t first = lf >> 4, cnt = lf & 15u;
            if (cnt == 0 || cnt > 4 || first + cnt > nt || esc != k + 1) { stats[0] = 4; return 0; }
            for (uint32_t i = first; i < first + cnt; i++) {
                covered[i]++;
                for (int c = 0; c < 3; c++) {
                    const float *p = verts + 3 * (size_t) tris[3 * (size_t) bvh.triOrder[i] + c];
                    if (p[0] < n.lo.x || p[1] < n.lo.y || p[2] < n.lo.z || p[0] > n.hi.x || p[1] > n.hi.y || p[2] > n.hi.z) { stats[0] = 5; return 0; }
                }
            }
        } else {
            if (k + 1 >= nn) { stats[0] = 6; return 0; }
            const uint32_t second = asU(bvh.nodes[k + 1].lo.w);            /* the first child's escape = the second child */
            if (second >= esc) { stats[0] = 7; return 0; }
            for (uint32_t ch : {k + 1, second}) {                              /* children's boxes inside the parent's */
                const BvhNode &x = bvh.nodes[ch];
                if (x.lo.x < n.lo.x || x.lo.y < n.lo.y || x.lo.z < n.lo.z || x.hi.x > n.hi.x || x.hi.y > n.hi.y || x.hi.z > n.hi.z) { stats[0] = 8; return 0; }
            }
        }
    }
    for (uint32_t i = 0; i < nt; i++) if (covered[i] != 1) { stats[0] = 9; return 0; }
    /* culling */
    auto slab = [](const double lo[3], const double hi[3], const double O[3], const double D[3], double tm) {
        double t0 = 0, t1 = tm;
        for (int c = 0; c < 3; c++) {
            if (D[c] == 0) { if (O[c] < lo[c] || O[c] > hi[c]) return false; continue; }
            double a = (lo[c] - O[c]) / D[c], b = (hi[c] - O[c]) / D[c];
            if (a > b) std::swap(a, b);
            t0 = std::max(t0, a); t1 = std::min(t1, b);
        }
        return t0 <= t1;
    };
    std::vector<uint8_t> vis2(nt), vis4(nt);
    for (uint32_t r = 0; r < nrays; r++) {
        const double O[3] = {o[3 * r], o[3 * r + 1], o[3 * r + 2]}, D[3] = {d[3 * r], d[3 * r + 1], d[3 * r + 2]};
        const double tm = tmax[r];
        std::fill(vis2.begin(), vis2.end(), 0); std::fill(vis4.begin(), vis4.end(), 0);
        for (uint32_t k = 0; k < nn;) {                                       /* the stackless walk */
            const BvhNode &n = bvh.nodes[k];
            stats[4]++;
            const double lo[3] = {n.lo.x, n.lo.y, n.lo.z}, hi[3] = {n.hi.x, n.hi.y, n.hi.z};
            const uint32_t lf = asU(n.hi.w);
            if (!slab(lo, hi, O, D, tm)) { k = asU(n.lo.w); continue; }
            if (lf) { for (uint32_t i = lf >> 4; i < (lf >> 4) + (lf & 15u); i++) vis2[i] = 1; k = asU(n.lo.w); }
            else k++;
        }
        std::vector<int> stack;
        if (!wide.empty()) stack.push_back(0);
        while (!stack.empty()) {
            const Bvh4Node &n = wide[stack.back()]; stack.pop_back();
            stats[5]++;
            const float *lx = &n.lox.x, *ly = &n.loy.x, *lz = &n.loz.x, *hx = &n.hix.x, *hy = &n.hiy.x, *hz = &n.hiz.x;
            const int *ch = &n.child.x;
            for (int k = 0; k < 4; k++) {
                const double lo[3] = {lx[k], ly[k], lz[k]}, hi[3] = {hx[k], hy[k], hz[k]};
                if (!(lo[0] <= hi[0])) continue;                               /* unused slot: empty box */
                if (!slab(lo, hi, O, D, tm)) continue;
                if (ch[k] >= 0) stack.push_back(ch[k]);
                else { const uint32_t lf = (uint32_t) ~ch[k]; for (uint32_t i = lf >> 4; i < (lf >> 4) + (lf & 15u); i++) { if (i >= nt) { stats[0] = 10; return 0; } vis4[i] = 1; } }
            }
        }
        for (uint32_t i = 0; i < nt; i++) {                                    /* brute force, in leaf order */
            const uint32_t t = bvh.triOrder[i];
            const float *A = verts + 3 * (size_t) tris[3 * (size_t) t], *B = verts + 3 * (size_t) tris[3 * (size_t) t + 1], *Cc = verts + 3 * (size_t) tris[3 * (size_t) t + 2];
            const double e1[3] = {(double) B[0] - A[0], (double) B[1] - A[1], (double) B[2] - A[2]}, e2[3] = {(double) Cc[0] - A[0], (double) Cc[1] - A[1], (double) Cc[2] - A[2]};
            const double p[3] = {D[1] * e2[2] - D[2] * e2[1], D[2] * e2[0] - D[0] * e2[2], D[0] * e2[1] - D[1] * e2[0]};
            const double det = p[0] * e1[0] + p[1] * e1[1] + p[2] * e1[2];
            if (std::fabs(det) < 1e-14) continue;
            const double s[3] = {O[0] - A[0], O[1] - A[1], O[2] - A[2]};
            const double u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) / det;
            const double q[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
            const double v = (D[0] * q[0] + D[1] * q[1] + D[2] * q[2]) / det, tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) / det;
            if (u < 0 || v < 0 || u + v > 1 || tt < 0 || tt > tm) continue;
            stats[1]++;
            if (!vis2[i]) stats[2]++;
            if (!vis4[i]) stats[3]++;
        }
    }
    return 0;
}

int alvrl_host_sfmt_ulongs(uint64_t seed, uint32_t cloneDepth, uint32_t skip, uint64_t *out, uint32_t n) {
    Sfmt19937 g(seed);
    if (cloneDepth == 0) { for (uint32_t i = 0; i < skip; i++) g.next64(); for (uint32_t i = 0; i < n; i++) out[i] = g.next64(); return 0; }
    for (uint32_t i = 0; i < skip; i++) g.next64();
    Sfmt19937 child(g);
    for (uint32_t i = 0; i < n; i++) out[i] = child.next64();
    return 0;
}
int alvrl_host_sfmt_floats(uint64_t seed, float *out, uint32_t n) {
    SfmtStream s(seed);
    for (uint32_t i = 0; i < n; i++) out[i] = s.next1D();
    return 0;
}
/* Preprocessor::getSlices + sampleSliceMapping on caller-supplied gather points */
int alvrl_host_slices(const float *pos, const float *dir, uint32_t n, uint32_t targetNumSlices, float targetUndersampling,
                      int rngMode, uint64_t seed, uint32_t *pixelToSlice, uint32_t *numSlices, uint32_t *rowOffset /*target+1*/,
                      uint32_t *rowPixel /*n*/) {
    std::vector<P3> p(n), d(n);
    for (uint32_t i = 0; i < n; i++) { p[i] = P3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}; d[i] = P3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}; }
    std::vector<SliceInfo> slices;
    SliceTree tree(p, d);
    std::vector<uint32_t> map = tree.build(targetNumSlices, slices);
    memcpy(pixelToSlice, map.data(), n * 4);
    *numSlices = (uint32_t) slices.size();
    std::unique_ptr<HostSampler> smp(rngMode == ALVRL_RNG_MODE_SFMT ? (HostSampler *) new SfmtStream(seed) : (HostSampler *) new CounterStream(seed));
    uint32_t rows = 0;
    rowOffset[0] = 0;
    for (size_t i = 0; i < slices.size(); i++) {
        smp->setContext(ALVRL_RNG_SLICEMAP, (uint32_t) i, 0);
        std::vector<uint32_t> px = sampleRepresentativePixels(slices[i], targetUndersampling, smp.get());
        memcpy(rowPixel + rows, px.data(), px.size() * 4);
        rows += (uint32_t) px.size();
        rowOffset[i + 1] = rows;
    }
    return 0;
}

/* heap_order.h against std::push_heap / std::pop_heap (the boost::heap::priority_queue of Clustering): replays a sequence
 * of operations (op[i] != 0: push keys[i]; op[i] == 0: pop) on both and compares the whole array after every step.
 * Returns -1 when identical throughout, else the index of the first differing operation. */
int alvrl_host_heap_check(const float *keys, const uint8_t *op, uint32_t n) {
    struct Node { float key; uint32_t id; bool operator<(const Node &o) const { return key < o.key; } };
    std::vector<Node> ref; std::vector<HeapEntry> mine(n + 1); uint32_t count = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (op[i] || ref.empty()) {
            ref.push_back(Node{keys[i], i}); std::push_heap(ref.begin(), ref.end());
            HeapEntry e; e.key = keys[i]; e.id = i; heap_push(mine.data(), count, e);
        } else {
            std::pop_heap(ref.begin(), ref.end()); const Node top = ref.back(); ref.pop_back();
            const HeapEntry t = heap_pop(mine.data(), count);
            if (t.id != top.id) return (int) i;
        }
        if (count != ref.size()) return (int) i;
        for (uint32_t k = 0; k < count; k++) if (mine[k].id != ref[k].id) return (int) i;
    }
    return -1;
}


/* occluders.h: compile a small mesh; counts = {use, numSlabs, numPlanes, numTris, numPolytopes, numBoxes}; the OccDev block
 * (occ_query.h) is copied to `dev`, the planar groups' triangle records to `tris` (float4, up to maxFloat4) */
int alvrl_host_compile_occluders(const float *verts, const uint32_t *tris, uint32_t nt, uint32_t numLeaves, uint32_t *counts, void *dev,
                                 float *triRecs, uint32_t maxFloat4) {
    const OccluderSet os = compile_occluders(verts, tris, nt, numLeaves);
    counts[0] = os.use; counts[1] = os.numSlabs; counts[2] = os.numPlanes; counts[3] = os.numTris; counts[4] = os.numPolytopes; counts[5] = os.numBoxes;
    if (os.tris.size() > maxFloat4) return -1;
    if (os.use) { memcpy(dev, &os.dev, sizeof(OccDev)); if (!os.tris.empty()) memcpy(triRecs, os.tris.data(), os.tris.size() * sizeof(float4)); }
    return (int) sizeof(OccDev);
}
/* occ_query.h on the host: n segments (origin o, unit direction d, [tmin, tmax]) against a compiled set */
int alvrl_host_occ_query(const void *dev, const float *triRecs, const float *o, const float *d, const float *tmin, const float *tmax,
                         uint32_t n, uint8_t *out) {
    OccDev oc; memcpy(&oc, dev, sizeof(oc));
    const float4 *tr = reinterpret_cast<const float4 *>(triRecs);
    for (uint32_t i = 0; i < n; i++)
        out[i] = occ_query(oc, tr, o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2], tmin[i], tmax[i], true) ? 1 : 0;
    return 0;
}


/* pair-level culling of occ_query.h: for every pair (camera segment [E, Usurf], VRL [S, End]) the side bits of the two ends
 * are ANDed like the transport kernels do (per lane, without the warp vote), `samples` random shadow segments of the pair
 * are queried with and without the resulting active masks, and every disagreement is counted.
 * stats = {mismatches, box tests culled, box tests, plane tests culled, plane tests, occluded segments} */
int alvrl_host_pair_cull_check(const void *dev, const float *triRecs, const float *E, const float *Usurf, const float *S, const float *End,
                               uint32_t nPairs, uint32_t samples, uint64_t seed, float margin, uint64_t *stats) {
    OccDev oc; memcpy(&oc, dev, sizeof(oc));
    oc.cullMargin = margin;
    const float4 *tr = reinterpret_cast<const float4 *>(triRecs);
    uint64_t x = seed * 0x9e3779b97f4a7c15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (float) ((x >> 40) * (1.0 / 16777216.0)); };
    for (int k = 0; k < 6; k++) stats[k] = 0;
    const uint32_t nb = oc.numBoxes;
    const uint32_t boxAll = (1u << nb) - 1u, planeAll = (1u << oc.numPlanes) - 1u;
    for (uint32_t i = 0; i < nPairs; i++) {
        const float *e = E + 3 * i, *u = Usurf + 3 * i, *s = S + 3 * i, *t = End + 3 * i;
        const uint32_t segHull = occ_slab_sides(oc, e[0], e[1], e[2], u[0], u[1], u[2], margin);
        const uint32_t segSurf = occ_slab_sides(oc, u[0], u[1], u[2], u[0], u[1], u[2], margin);
        const uint32_t segPl = occ_plane_sides(oc, e[0], e[1], e[2], u[0], u[1], u[2], margin, true);
        const uint32_t vSl = occ_slab_sides(oc, s[0], s[1], s[2], t[0], t[1], t[2], margin);
        const uint32_t vPl = occ_plane_sides(oc, s[0], s[1], s[2], t[0], t[1], t[2], margin, false);
        uint32_t pc = segPl & vPl; pc = (pc | (pc >> 16)) & 0xffffu;
        const uint32_t planeAct = planeAll & ~pc;
        const uint32_t boxVV = boxAll & ~occ_boxes_culled(segHull & vSl, nb), boxVS = boxAll & ~occ_boxes_culled(segSurf & vSl, nb);
        for (uint32_t q = 0; q < samples; q++) {
            const bool surf = q & 1;
            const float a = surf ? 1.0f : rnd(), b = rnd();
            float o[3], d[3], len2 = 0;
            for (int k = 0; k < 3; k++) { o[k] = surf ? u[k] : e[k] + a * (u[k] - e[k]); d[k] = (s[k] + b * (t[k] - s[k])) - o[k]; len2 += d[k] * d[k]; }
            const float len = std::sqrt(len2);
            if (!(len > 0)) continue;
            for (int k = 0; k < 3; k++) d[k] /= len;
            const float tmin = surf ? 1e-4f * std::max(std::fabs(o[0]), std::max(std::fabs(o[1]), std::fabs(o[2]))) : 0.0f;
            const bool full = occ_query(oc, tr, o[0], o[1], o[2], d[0], d[1], d[2], tmin, len, true);
            const bool culled = occ_query(oc, tr, o[0], o[1], o[2], d[0], d[1], d[2], tmin, len, true, surf ? boxVS : boxVV, planeAct);
            stats[0] += full != culled; stats[5] += full;
            stats[1] += __builtin_popcount(boxAll & ~(surf ? boxVS : boxVV)); stats[2] += nb;
            stats[3] += __builtin_popcount(planeAll & ~planeAct); stats[4] += oc.numPlanes;
        }
    }
    return 0;
}

}
