/*
 * bvh.h -- host builder of the threaded (stackless) BVH that replaces the reference's SAH kd-tree
 * (include/mitsuba/render/sahkdtree3.h, gkdtree.h) on the device.
 *
 * Binned-SAH binary BVH, leaves of <= 4 triangles, nodes stored in depth-first preorder.  Every node
 * carries an *escape index* (= index of the first node after its subtree), so traversal needs no stack:
 *     hit inner  -> node + 1          hit leaf / miss -> escape
 * The traversal order is fixed (near-to-far along the split axis for rays with positive direction).
 * Semantics, not structure, must match the reference: closest hit inside [mint, maxt] with the
 * TriAccel arithmetic; boxes are padded so that a triangle the TriAccel test accepts is never culled.
 */
#pragma once
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>
#include "types.h"

namespace alvrl {

struct HostBvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> triOrder;   /* leaf order -> original triangle index */
    uint32_t maxLeaf = 0;
};

struct Box3 {
    float lo[3], hi[3];
    Box3() { for (int i = 0; i < 3; i++) { lo[i] = INFINITY; hi[i] = -INFINITY; } }
    void grow(const float *p) { for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], p[i]); hi[i] = std::max(hi[i], p[i]); } }
    void grow(const Box3 &b) { for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], b.lo[i]); hi[i] = std::max(hi[i], b.hi[i]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0)) return 0;
        return 2 * (dx * dy + dy * dz + dz * dx);
    }
};

class BvhBuilder {
public:
    BvhBuilder(const float *verts, const uint32_t *tris, uint32_t nt) : m_nt(nt) {
        m_box.resize(nt); m_cen.resize(3 * (size_t) nt); m_idx.resize(nt);
        float maxAbs = 0;
        for (uint32_t i = 0; i < nt; i++) {
            for (int k = 0; k < 3; k++) {
                const float *p = verts + 3 * (size_t) tris[3 * (size_t) i + k];
                m_box[i].grow(p);
                for (int c = 0; c < 3; c++) maxAbs = std::max(maxAbs, std::fabs(p[c]));
            }
            for (int c = 0; c < 3; c++) m_cen[3 * (size_t) i + c] = 0.5f * (m_box[i].lo[c] + m_box[i].hi[c]);
            m_idx[i] = i;
        }
        m_pad = 2e-5f * std::max(1.0f, maxAbs);
    }
    void build(HostBvh &out) {
        out.nodes.clear(); out.nodes.reserve(m_nt / 2 + 16);
        if (m_nt) recurse(out, 0, m_nt);
        out.triOrder = m_idx;
    }
private:
    static const int kBins = 16, kLeaf = 4;      /* leaf count is stored in 4 bits */
    uint32_t m_nt; float m_pad;
    std::vector<Box3> m_box; std::vector<float> m_cen; std::vector<uint32_t> m_idx;

    static float asFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

    uint32_t recurse(HostBvh &out, uint32_t b, uint32_t e) {
        uint32_t me = (uint32_t) out.nodes.size();
        out.nodes.push_back(BvhNode());
        Box3 bounds, cb;
        for (uint32_t i = b; i < e; i++) { bounds.grow(m_box[m_idx[i]]); cb.grow(&m_cen[3 * (size_t) m_idx[i]]); }
        uint32_t n = e - b;
        uint32_t mid = 0;
        bool leaf = n == 1;
        float bestCost = INFINITY;
        if (!leaf) {
            int axis = 0; float ext = -1;
            for (int c = 0; c < 3; c++) if (cb.hi[c] - cb.lo[c] > ext) { ext = cb.hi[c] - cb.lo[c]; axis = c; }
            if (!(ext > 0)) {
                mid = b + n / 2;              /* all centroids coincide: split the list */
            } else {
                Box3 binBox[kBins]; uint32_t binCnt[kBins] = {0};
                float scale = kBins * (1 - 1e-6f) / ext;
                for (uint32_t i = b; i < e; i++) {
                    int k = std::min(kBins - 1, std::max(0, (int) ((m_cen[3 * (size_t) m_idx[i] + axis] - cb.lo[axis]) * scale)));
                    binBox[k].grow(m_box[m_idx[i]]); binCnt[k]++;
                }
                float rightArea[kBins]; uint32_t rightCnt[kBins];
                Box3 acc; uint32_t cnt = 0;
                for (int k = kBins - 1; k > 0; k--) { acc.grow(binBox[k]); cnt += binCnt[k]; rightArea[k] = acc.area(); rightCnt[k] = cnt; }
                Box3 accL; uint32_t cntL = 0; float best = INFINITY; int bestK = -1;
                for (int k = 0; k < kBins - 1; k++) {
                    accL.grow(binBox[k]); cntL += binCnt[k];
                    if (cntL == 0 || rightCnt[k + 1] == 0) continue;
                    float cost = accL.area() * cntL + rightArea[k + 1] * rightCnt[k + 1];
                    if (cost < best) { best = cost; bestK = k; }
                }
                bestCost = best;
                if (bestK < 0) mid = b + n / 2;
                else {
                    auto it = std::partition(m_idx.begin() + b, m_idx.begin() + e, [&](uint32_t t) {
                        int k = std::min(kBins - 1, std::max(0, (int) ((m_cen[3 * (size_t) t + axis] - cb.lo[axis]) * scale)));
                        return k <= bestK;
                    });
                    mid = (uint32_t) (it - m_idx.begin());
                    if (mid == b || mid == e) mid = b + n / 2;
                }
            }
        }
        /* SAH termination: keep <= kLeaf triangles together only when splitting does not pay (cost of a box test = 1
         * triangle test): two walls of a room in one leaf would give a box that every ray enters */
        if (!leaf && n <= kLeaf) {
            const float area = bounds.area();
            const float leafCost = area * n, splitCost = area * 1.0f + bestCost;
            if (!(splitCost < leafCost)) leaf = true;
        }
        if (leaf) {
            out.maxLeaf = std::max(out.maxLeaf, n);
        } else {
            recurse(out, b, mid);
            recurse(out, mid, e);
        }
        uint32_t escape = (uint32_t) out.nodes.size();
        BvhNode &nd = out.nodes[me];
        nd.lo = make_float4(bounds.lo[0] - m_pad, bounds.lo[1] - m_pad, bounds.lo[2] - m_pad, asFloat(escape));
        nd.hi = make_float4(bounds.hi[0] + m_pad, bounds.hi[1] + m_pad, bounds.hi[2] + m_pad, asFloat(leaf ? ((b << 4) | n) : 0u));
        return me;
    }
};

/*
 * 4-wide tree for the fast flavour's any-hit query (types.h::Bvh4Node), collapsed from the binary tree: the two children of
 * a node, then repeatedly the inner child with the largest surface area replaced by its own two children until four slots
 * are taken.  Same leaves, same (padded) boxes, so the set of triangles a ray gets to test can only shrink to those whose
 * leaf box it enters -- as in the binary tree.  Measured on C4 (998 k triangles, tools/micro/bvh_visits.cpp): 20.8 node visits
 * per shadow ray instead of 81, 44 instead of 173 for the slowest ray of a warp.
 */
inline uint32_t collapse4(const HostBvh &b, std::vector<Bvh4Node> &out) {        /* returns the depth of the 4-wide tree */
    out.clear();
    if (b.nodes.empty()) return 0;
    auto asU = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    auto area = [&](uint32_t k) { const BvhNode &n = b.nodes[k]; const float dx = n.hi.x - n.lo.x, dy = n.hi.y - n.lo.y, dz = n.hi.z - n.lo.z; return dx * dy + dy * dz + dz * dx; };
    struct Item { uint32_t bin; uint32_t me; uint32_t depth; };
    uint32_t maxDepth = 1, curDepth = 1;
    std::vector<Item> todo;
    auto emit = [&](const uint32_t *kids, int n, uint32_t me) {
        Bvh4Node nd;
        float lo[3][4], hi[3][4]; int ch[4];
        for (int k = 0; k < 4; k++) { for (int c = 0; c < 3; c++) { lo[c][k] = INFINITY; hi[c][k] = -INFINITY; } ch[k] = (int) 0x80000000u; }
        for (int k = 0; k < n; k++) {
            const BvhNode &x = b.nodes[kids[k]];
            lo[0][k] = x.lo.x; lo[1][k] = x.lo.y; lo[2][k] = x.lo.z; hi[0][k] = x.hi.x; hi[1][k] = x.hi.y; hi[2][k] = x.hi.z;
            const uint32_t lf = asU(x.hi.w);
            if (lf) ch[k] = ~(int) lf;
            else { ch[k] = (int) out.size(); out.push_back(Bvh4Node()); todo.push_back({kids[k], (uint32_t) ch[k], curDepth + 1}); }
        }
        nd.lox = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]); nd.loy = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
        nd.loz = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]); nd.hix = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
        nd.hiy = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]); nd.hiz = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
        nd.child = make_int4(ch[0], ch[1], ch[2], ch[3]); nd.pad = make_int4(0, 0, 0, 0);
        out[me] = nd;
    };
    out.push_back(Bvh4Node());
    if (asU(b.nodes[0].hi.w)) { const uint32_t k = 0; emit(&k, 1, 0); return 1; }   /* the root is a leaf */
    todo.push_back({0, 0, 1});
    while (!todo.empty()) {                                  /* breadth-first-ish: children of a node sit close together */
        const Item it = todo.back(); todo.pop_back();
        curDepth = it.depth; maxDepth = std::max(maxDepth, curDepth);
        uint32_t kids[4]; int n = 2;
        kids[0] = it.bin + 1; kids[1] = asU(b.nodes[it.bin + 1].lo.w);           /* second child = the first one's escape index */
        while (n < 4) {
            int best = -1; float ba = -1;
            for (int k = 0; k < n; k++) if (!asU(b.nodes[kids[k]].hi.w) && area(kids[k]) > ba) { ba = area(kids[k]); best = k; }
            if (best < 0) break;
            const uint32_t x = kids[best];
            kids[best] = x + 1; kids[n++] = asU(b.nodes[x + 1].lo.w);
        }
        emit(kids, n, it.me);
    }
    return maxDepth;
}

/* Wald TriAccel precomputation, include/mitsuba/render/triaccel.h:61-95 (IEEE fp32, same operation order) */
inline TriRec makeTriRec(const float *A, const float *B, const float *C, uint32_t origIndex) {
    static const int waldModulo[4] = {1, 2, 0, 1};
    float b[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]}, c[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    float N[3] = {c[1] * b[2] - c[2] * b[1], c[2] * b[0] - c[0] * b[2], c[0] * b[1] - c[1] * b[0]};
    uint32_t k = 0;
    for (int j = 0; j < 3; j++) if (std::fabs(N[j]) > std::fabs(N[k])) k = j;
    uint32_t u = waldModulo[k], v = waldModulo[k + 1];
    const float n_k = N[k], denom = b[u] * c[v] - b[v] * c[u];
    TriRec r;
    auto asF = [](uint32_t x) { float f; memcpy(&f, &x, 4); return f; };
    if (denom == 0) {
        r.a = make_float4(asF(3u), 0, 0, 0); r.b = make_float4(0, 0, 0, 0); r.c = make_float4(0, 0, asF(origIndex), 0);
        return r;
    }
    float n_u = N[u] / n_k, n_v = N[v] / n_k;
    float n_d = (A[0] * N[0] + A[1] * N[1] + A[2] * N[2]) / n_k;
    float b_nu = b[u] / denom, b_nv = -b[v] / denom, a_u = A[u], a_v = A[v], c_nu = c[v] / denom, c_nv = -c[u] / denom;
    r.a = make_float4(asF(k), n_u, n_v, n_d);
    r.b = make_float4(a_u, a_v, b_nu, b_nv);
    r.c = make_float4(c_nu, c_nv, asF(origIndex), 0);
    return r;
}

/* plane + barycentric functionals of a triangle (double precision set-up, stored as float) */
inline TriFast makeTriFast(const float *A, const float *B, const float *C) {
    const double e1[3] = {(double) B[0] - A[0], (double) B[1] - A[1], (double) B[2] - A[2]};
    const double e2[3] = {(double) C[0] - A[0], (double) C[1] - A[1], (double) C[2] - A[2]};
    const double N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double n2 = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    TriFast t;
    if (!(n2 > 0)) { t.p = make_float4(0, 0, 0, 1); t.q = make_float4(0, 0, 0, -1); t.r = make_float4(0, 0, 0, -1); return t; }
    const double inv = 1.0 / std::sqrt(n2);
    const double n[3] = {N[0] * inv, N[1] * inv, N[2] * inv};
    /* u (weight of B) = dot(P - A, e2 x N) / |N|^2, v (weight of C) = dot(P - A, N x e1) / |N|^2 */
    const double eu[3] = {(e2[1] * N[2] - e2[2] * N[1]) / n2, (e2[2] * N[0] - e2[0] * N[2]) / n2, (e2[0] * N[1] - e2[1] * N[0]) / n2};
    const double ev[3] = {(N[1] * e1[2] - N[2] * e1[1]) / n2, (N[2] * e1[0] - N[0] * e1[2]) / n2, (N[0] * e1[1] - N[1] * e1[0]) / n2};
    t.p = make_float4((float) n[0], (float) n[1], (float) n[2], (float) (n[0] * A[0] + n[1] * A[1] + n[2] * A[2]));
    t.q = make_float4((float) eu[0], (float) eu[1], (float) eu[2], (float) -(eu[0] * A[0] + eu[1] * A[1] + eu[2] * A[2]));
    t.r = make_float4((float) ev[0], (float) ev[1], (float) ev[2], (float) -(ev[0] * A[0] + ev[1] * A[1] + ev[2] * A[2]));
    return t;
}

} // namespace alvrl
