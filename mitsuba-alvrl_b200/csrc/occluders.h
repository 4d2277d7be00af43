/*
 * occluders.h -- host "occluder compiler" for small scenes (the Cornell-box class of the BASELINE configs).
 *
 * Scene::evalTransmittance (src/librender/scene.cpp:619-679) only asks whether *anything* opaque lies on a shadow
 * segment.  For a scene of a few dozen triangles the answer does not need a tree at all:
 *
 *   - a connected component of the mesh that is CLOSED (every welded edge has exactly two triangles) and CONVEX (every
 *     vertex on or behind every face plane) is the boundary of the intersection of its face half-spaces.  A segment
 *     touches that boundary iff the parametric interval [t_in, t_out] clipped by the half-spaces is non-empty and one
 *     of its ends falls inside [mint, maxt].  Antiparallel face pairs (boxes) share one normal: a SLAB c_lo <= n.x <= c_hi
 *     costs one reciprocal; a lone half-space is a slab with c_lo = -inf.
 *   - every other triangle joins a PLANAR GROUP (coplanar triangles, e.g. the two halves of a wall): one plane
 *     evaluation decides whether the segment crosses the plane inside [mint, maxt]; only then are the group's
 *     triangles tested (plane + barycentric functionals, TriFast).  For two points inside a room no wall is crossed.
 *
 * The result (OccDev, occ_query.h) travels in the kernel parameters of the transport kernels; only the planar groups'
 * triangle records live in shared memory.  It decides exactly what the triangle tests decide, up to grazing rays (the documented tie class).
 * The compiler declines (use = false) when the scene does not fit the budgets or a flat leaf sweep would be cheaper.
 */
#pragma once
#include <vector>
#include <map>
#include <array>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <numeric>
#include <cstdlib>
#include "types.h"
#include "bvh.h"
#include "occ_query.h"

namespace alvrl {

struct OccluderSet {
    bool use = false;
    uint32_t numSlabs = 0, numPlanes = 0, numTris = 0, numPolytopes = 0, numBoxes = 0;
    OccDev dev;                       /* slabs (boxes first) and planes: travels in the kernel parameters */
    std::vector<float4> tris;         /* TriFast p, q, r of the planar groups' triangles: shared memory */
};

namespace occ_detail {
struct Plane { double n[3], c; };

inline bool tri_plane(const float *A, const float *B, const float *C, Plane &p) {
    const double e1[3] = {(double) B[0] - A[0], (double) B[1] - A[1], (double) B[2] - A[2]};
    const double e2[3] = {(double) C[0] - A[0], (double) C[1] - A[1], (double) C[2] - A[2]};
    const double N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double n2 = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    if (!(n2 > 0)) return false;
    const double inv = 1.0 / std::sqrt(n2);
    for (int k = 0; k < 3; k++) p.n[k] = N[k] * inv;
    p.c = p.n[0] * A[0] + p.n[1] * A[1] + p.n[2] * A[2];
    return true;
}
inline double ndist(const double *a, const double *b, double sign) {
    double s = 0;
    for (int k = 0; k < 3; k++) { const double d = a[k] - sign * b[k]; s += d * d; }
    return std::sqrt(s);
}
struct Dsu {
    std::vector<uint32_t> p;
    explicit Dsu(uint32_t n) : p(n) { std::iota(p.begin(), p.end(), 0u); }
    uint32_t find(uint32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void join(uint32_t a, uint32_t b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};
} // namespace occ_detail

/* numLeaves: leaves of the BVH of the same scene (cost of the alternative, the flat leaf sweep) */
inline OccluderSet compile_occluders(const float *verts, const uint32_t *tris, uint32_t nt, uint32_t numLeaves) {
    using namespace occ_detail;
    OccluderSet out;
    if (nt == 0 || nt > ALVRL_OCC_MAX_TRIS) return out;

    /* weld vertices by exact position */
    std::map<std::array<float, 3>, uint32_t> weld;
    std::vector<std::array<float, 3>> wpos;
    std::vector<uint32_t> wv(3 * (size_t) nt);
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < 3 * nt; i++) {
        const float *p = verts + 3 * (size_t) tris[i];
        std::array<float, 3> key = {p[0] + 0.0f, p[1] + 0.0f, p[2] + 0.0f};      /* -0 -> +0 */
        auto it = weld.find(key);
        if (it == weld.end()) { it = weld.emplace(key, (uint32_t) wpos.size()).first; wpos.push_back(key); }
        wv[i] = it->second;
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], (double) p[k]); hi[k] = std::max(hi[k], (double) p[k]); }
    }
    const double extent = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    const double posTol = 1e-5 * std::max(extent, 1e-30), dirTol = 2e-6;

    /* planes; degenerate triangles can never be hit (triaccel.h: zero normal -> rejected) and are dropped */
    std::vector<Plane> pl(nt);
    std::vector<uint8_t> ok(nt);
    for (uint32_t t = 0; t < nt; t++)
        ok[t] = tri_plane(verts + 3 * (size_t) tris[3 * t], verts + 3 * (size_t) tris[3 * t + 1], verts + 3 * (size_t) tris[3 * t + 2], pl[t]) &&
                wv[3 * t] != wv[3 * t + 1] && wv[3 * t + 1] != wv[3 * t + 2] && wv[3 * t] != wv[3 * t + 2];

    /* components over shared welded edges */
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> edges;
    for (uint32_t t = 0; t < nt; t++) {
        if (!ok[t]) continue;
        for (int k = 0; k < 3; k++) {
            uint32_t a = wv[3 * t + k], b = wv[3 * t + (k + 1) % 3];
            if (a > b) std::swap(a, b);
            edges[{a, b}].push_back(t);
        }
    }
    Dsu dsu(nt);
    for (auto &e : edges) for (size_t i = 1; i < e.second.size(); i++) dsu.join(e.second[0], e.second[i]);
    std::vector<uint8_t> closed(nt, 1);                           /* per component root */
    for (auto &e : edges) if (e.second.size() != 2) closed[dsu.find(e.second[0])] = 0;

    struct Solid { std::vector<float4> a; std::vector<float> chi; bool box; };
    std::vector<Solid> solids;
    std::vector<float4> planes, triStream;
    std::vector<uint32_t> planeInfo;
    std::vector<uint8_t> inPolytope(nt, 0);
    for (uint32_t root = 0; root < nt; root++) {
        if (!ok[root] || dsu.find(root) != root || !closed[root]) continue;
        std::vector<uint32_t> comp;
        for (uint32_t t = 0; t < nt; t++) if (ok[t] && dsu.find(t) == root) comp.push_back(t);
        if (comp.size() < 4) continue;
        /* centroid of the component's vertices orients the face normals outwards */
        double cen[3] = {0, 0, 0}; size_t cnt = 0;
        for (uint32_t t : comp) for (int k = 0; k < 3; k++) { const auto &p = wpos[wv[3 * t + k]]; cen[0] += p[0]; cen[1] += p[1]; cen[2] += p[2]; cnt++; }
        for (int k = 0; k < 3; k++) cen[k] /= (double) cnt;
        std::vector<Plane> faces;
        bool convex = true;
        for (uint32_t t : comp) {
            Plane p = pl[t];
            if (p.n[0] * cen[0] + p.n[1] * cen[1] + p.n[2] * cen[2] - p.c > 0) { for (int k = 0; k < 3; k++) p.n[k] = -p.n[k]; p.c = -p.c; }
            for (uint32_t u : comp) for (int k = 0; k < 3 && convex; k++) {
                const auto &q = wpos[wv[3 * u + k]];
                if (p.n[0] * q[0] + p.n[1] * q[1] + p.n[2] * q[2] - p.c > posTol) convex = false;
            }
            if (!convex) break;
            bool dup = false;
            for (const Plane &f : faces) if (ndist(f.n, p.n, 1.0) < dirTol && std::fabs(f.c - p.c) < posTol) { dup = true; break; }
            if (!dup) faces.push_back(p);
        }
        if (!convex || faces.size() < 4) continue;
        /* the centroid must be strictly inside (a flat, doubly covered sheet is not a solid) */
        bool solid = true;
        for (const Plane &f : faces) if (!(f.n[0] * cen[0] + f.n[1] * cen[1] + f.n[2] * cen[2] - f.c < -posTol)) solid = false;
        if (!solid) continue;
        /* pair antiparallel faces into slabs */
        std::vector<uint8_t> used(faces.size(), 0);
        Solid sol; sol.box = true;
        for (size_t i = 0; i < faces.size(); i++) {
            if (used[i]) continue;
            used[i] = 1;
            double clo = -INFINITY;
            for (size_t j = i + 1; j < faces.size(); j++)
                if (!used[j] && ndist(faces[i].n, faces[j].n, -1.0) < dirTol) { used[j] = 1; clo = -faces[j].c; break; }
            sol.a.push_back(make_float4((float) faces[i].n[0], (float) faces[i].n[1], (float) faces[i].n[2], (float) clo));
            sol.chi.push_back((float) faces[i].c);
            if (!std::isfinite(clo)) sol.box = false;
        }
        sol.box = sol.box && sol.a.size() == 3;
        solids.push_back(sol);
        out.numPolytopes++;
        for (uint32_t t : comp) inPolytope[t] = 1;
    }
    std::stable_sort(solids.begin(), solids.end(), [](const Solid &x, const Solid &y) { return x.box > y.box; });

    /* planar groups of the remaining triangles */
    struct Group { Plane p; std::vector<uint32_t> tris; };
    std::vector<Group> groups;
    for (uint32_t t = 0; t < nt; t++) {
        if (!ok[t] || inPolytope[t]) continue;
        Group *g = nullptr;
        for (Group &c : groups) {
            const bool same = ndist(c.p.n, pl[t].n, 1.0) < dirTol && std::fabs(c.p.c - pl[t].c) < posTol;
            const bool flip = ndist(c.p.n, pl[t].n, -1.0) < dirTol && std::fabs(c.p.c + pl[t].c) < posTol;
            if ((same || flip) && c.tris.size() < 255) { g = &c; break; }
        }
        if (!g) { groups.push_back(Group{pl[t], {}}); g = &groups.back(); }
        g->tris.push_back(t);
    }
    for (const Group &g : groups) {
        const uint32_t first = (uint32_t) (triStream.size() / 3), info = (first << 8) | (uint32_t) g.tris.size();
        planes.push_back(make_float4((float) g.p.n[0], (float) g.p.n[1], (float) g.p.n[2], (float) g.p.c));
        planeInfo.push_back(info);
        for (uint32_t t : g.tris) {
            const TriFast f = makeTriFast(verts + 3 * (size_t) tris[3 * t], verts + 3 * (size_t) tris[3 * t + 1], verts + 3 * (size_t) tris[3 * t + 2]);
            triStream.push_back(f.p); triStream.push_back(f.q); triStream.push_back(f.r);
        }
    }

    for (const Solid &so : solids) { out.numSlabs += (uint32_t) so.a.size(); out.numBoxes += so.box ? 1 : 0; }
    out.numPlanes = (uint32_t) planes.size(); out.numTris = (uint32_t) (triStream.size() / 3);
    if (out.numSlabs > ALVRL_OCC_MAX_SLABS || out.numPlanes > ALVRL_OCC_MAX_PLANES || out.numTris > ALVRL_OCC_MAX_TRIS) return out;
    /* instruction estimates per shadow ray: 15 per slab, 13 per plane; flat sweep: 22 per leaf box + ~100 of triangle tests */
    const uint32_t costOcc = 15 * out.numSlabs + 13 * out.numPlanes, costFlat = 22 * numLeaves + 100;
    if (costOcc > costFlat) return out;
    memset(&out.dev, 0, sizeof(out.dev));
    uint32_t k = 0;
    for (const Solid &so : solids)
        for (size_t i = 0; i < so.a.size(); i++, k++) {
            out.dev.slabA[k] = so.a[i];
            out.dev.slabB[k] = make_float2(so.chi[i], 0.0f);
            if (i + 1 == so.a.size()) { const uint32_t one = 1u; memcpy(&out.dev.slabB[k].y, &one, 4); }
        }
    for (size_t i = 0; i < planes.size(); i++) { out.dev.planes[i] = planes[i]; out.dev.planeInfo[i] = planeInfo[i]; }
    out.dev.numBoxes = out.numBoxes; out.dev.numSlabs = out.numSlabs; out.dev.numPlanes = out.numPlanes; out.dev.numTris = out.numTris;
    out.tris = triStream;
    out.use = true;
    return out;
}

} // namespace alvrl
