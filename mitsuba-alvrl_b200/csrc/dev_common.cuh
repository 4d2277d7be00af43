/*
 * dev_common.cuh -- device helpers shared by every kernel of the path.
 *
 * Two kinds of arithmetic live here:
 *   (1) "exact" helpers (x* prefix) built from __fmul_rn/__fadd_rn/... which ptxas never contracts into
 *       FMAs.  They restate, operation by operation, the fp32 arithmetic the reference performs for
 *       decisions that must be bit-exact: ray set-up, AABB clipping, the TriAccel test, the intersection
 *       record (triaccel.h:97-158, aabb.h:308-338, skdtree.cpp:144-204, skdtree.h:343-428).
 *   (2) plain operators on F3, whose contraction is decided per translation unit (-fmad=false for the
 *       strict flavour, default for the fast flavour) -- used by the transport estimator.
 */
#pragma once
#include "types.h"
#include "../../include/alvrl_rng.h"

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 f3(const float4 &v) { return f3(v.x, v.y, v.z); }
__device__ __forceinline__ F3 operator+(const F3 &a, const F3 &b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(const F3 &a, const F3 &b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator-(const F3 &a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ F3 operator*(float f, const F3 &a) { return f3(f * a.x, f * a.y, f * a.z); }
__device__ __forceinline__ F3 operator*(const F3 &a, float f) { return f3(a.x * f, a.y * f, a.z * f); }
__device__ __forceinline__ float dot(const F3 &a, const F3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float len2(const F3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

/* ---- exact (never contracted) arithmetic ------------------------------------------------------- */
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float xdot(const F3 &a, const F3 &b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
__device__ __forceinline__ F3 xsub3(const F3 &a, const F3 &b) { return f3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
__device__ __forceinline__ F3 xadd3(const F3 &a, const F3 &b) { return f3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
__device__ __forceinline__ F3 xscale(const F3 &a, float f) { return f3(xmul(a.x, f), xmul(a.y, f), xmul(a.z, f)); }
__device__ __forceinline__ float xlen(const F3 &a) { return xsqrt(xdot(a, a)); }
/* TVector3::operator/(f): recip = 1/f, then multiply (include/mitsuba/core/vector.h) */
__device__ __forceinline__ F3 xdivv(const F3 &a, float f) { float r = xdiv(1.0f, f); return xscale(a, r); }
__device__ __forceinline__ F3 xnormalize(const F3 &a) { return xdivv(a, xlen(a)); }
__device__ __forceinline__ F3 xcross(const F3 &a, const F3 &b) {
    return f3(xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)), xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}

/* ---- AABB::rayIntersect, include/mitsuba/core/aabb.h:308-338 ---------------------------------- */
__device__ __forceinline__ bool aabb_clip(const float *bmin, const float *bmax, const F3 &o, const F3 &d, const F3 &dRcp,
                                          float &nearT, float &farT) {
    nearT = -INFINITY; farT = INFINITY;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, rr[3] = {dRcp.x, dRcp.y, dRcp.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (dd[i] == 0) {
            if (oo[i] < bmin[i] || oo[i] > bmax[i]) return false;
        } else {
            float t1 = xmul(xsub(bmin[i], oo[i]), rr[i]);
            float t2 = xmul(xsub(bmax[i], oo[i]), rr[i]);
            if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
            nearT = fmaxf(t1, nearT);
            farT = fminf(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return true;
}

/* ---- TriAccel::rayIntersect, include/mitsuba/render/triaccel.h:97-158 ------------------------- */
__device__ __forceinline__ bool tri_test(const TriRec *__restrict__ tris, uint32_t i, const F3 &o, const F3 &d,
                                         float mint, float maxt, float &t, float &u, float &v, uint32_t &orig) {
    const float4 a = __ldg(&tris[i].a), b = __ldg(&tris[i].b), c = __ldg(&tris[i].c);
    const uint32_t k = __float_as_uint(a.x);
    float o_u, o_v, o_k, d_u, d_v, d_k;
    if (k == 0) { o_u = o.y; o_v = o.z; o_k = o.x; d_u = d.y; d_v = d.z; d_k = d.x; }
    else if (k == 1) { o_u = o.z; o_v = o.x; o_k = o.y; d_u = d.z; d_v = d.x; d_k = d.y; }
    else if (k == 2) { o_u = o.x; o_v = o.y; o_k = o.z; d_u = d.x; d_v = d.y; d_k = d.z; }
    else return false;
    const float n_u = a.y, n_v = a.z, n_d = a.w;
    t = xdiv(xsub(xsub(xsub(n_d, xmul(o_u, n_u)), xmul(o_v, n_v)), o_k),
             xadd(xadd(xmul(d_u, n_u), xmul(d_v, n_v)), d_k));
    if (t < mint || t > maxt) return false;
    const float hu = xsub(xadd(o_u, xmul(t, d_u)), b.x);
    const float hv = xsub(xadd(o_v, xmul(t, d_v)), b.y);
    u = xadd(xmul(hv, b.z), xmul(hu, b.w));
    v = xadd(xmul(hu, c.x), xmul(hv, c.y));
    orig = __float_as_uint(c.z);
    return u >= 0 && v >= 0 && xadd(u, v) <= 1.0f;
}

/*
 * Stackless traversal of the threaded BVH.  [mint, maxt] is the interval ShapeKDTree::rayIntersect hands to
 * the traversal (already clipped to the tree AABB).  ANY: stop at the first accepted triangle (occlusion).
 * Closest-hit tie rule (quirk B13): lowest original triangle index among the minimal t.
 */
template <bool ANY>
__device__ __forceinline__ bool bvh_trace(const SceneDev &sc, const F3 &o, const F3 &d, const F3 &dRcp, float mint, float maxt,
                                          float &tHit, uint32_t &prim, float &uHit, float &vHit) {
    bool found = false;
    tHit = INFINITY; prim = ALVRL_NO_HIT;
    const float lo_t = mint - 1e-5f * fabsf(mint);
    float hi_t = maxt + 1e-5f * fabsf(maxt);
    uint32_t node = 0;
    const uint32_t numNodes = sc.numNodes;
    while (node < numNodes) {
        const float4 lo = __ldg(&sc.nodes[node].lo), hi = __ldg(&sc.nodes[node].hi);
        const float tx1 = (lo.x - o.x) * dRcp.x, tx2 = (hi.x - o.x) * dRcp.x;
        const float ty1 = (lo.y - o.y) * dRcp.y, ty2 = (hi.y - o.y) * dRcp.y;
        const float tz1 = (lo.z - o.z) * dRcp.z, tz2 = (hi.z - o.z) * dRcp.z;
        const float tn = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), lo_t));
        const float tf = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), hi_t));
        if (tn <= tf * 1.00001f) {
            const uint32_t leaf = __float_as_uint(hi.w);
            if (leaf) {
                const uint32_t first = leaf >> 4, cnt = leaf & 15u;
                for (uint32_t i = 0; i < cnt; i++) {
                    float t, u, v; uint32_t orig;
                    if (tri_test(sc.tris, first + i, o, d, mint, maxt, t, u, v, orig)) {
                        if (ANY) { tHit = t; prim = orig; return true; }
                        if (!found || t < tHit || (t == tHit && orig < prim)) {
                            tHit = t; prim = orig; uHit = u; vHit = v; found = true;
                            hi_t = t + 1e-5f * fabsf(t);
                        }
                    }
                }
                node = __float_as_uint(lo.w);
            } else node++;
        } else node = __float_as_uint(lo.w);
    }
    return found;
}

/*
 * ShapeKDTree::rayIntersect(ray, t, shape, n, uv) front end, skdtree.cpp:144-204: clip to the tree AABB, adaptive
 * epsilon *without* the Epsilon floor (154-157), then traverse.  floorEps selects the other overload (126-129).
 */
template <bool ANY>
__device__ __forceinline__ bool scene_intersect(const SceneDev &sc, const F3 &o, const F3 &d, float rayMint, float rayMaxt, bool floorEps,
                                                float &tHit, uint32_t &prim, float &u, float &v) {
    const F3 dRcp = f3(xdiv(1.0f, d.x), xdiv(1.0f, d.y), xdiv(1.0f, d.z));
    float mint, maxt;
    tHit = INFINITY; prim = ALVRL_NO_HIT;
    if (!aabb_clip(sc.kdMin, sc.kdMax, o, d, dRcp, mint, maxt)) return false;
    float rayMinT = rayMint;
    if (rayMinT == ALVRL_EPSILON) {
        float m = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
        if (floorEps) m = fmaxf(m, ALVRL_EPSILON);
        rayMinT = xmul(rayMinT, m);
    }
    if (rayMinT > mint) mint = rayMinT;
    if (rayMaxt < maxt) maxt = rayMaxt;
    if (!(maxt > mint)) return false;
    return bvh_trace<ANY>(sc, o, d, dRcp, mint, maxt, tHit, prim, u, v);
}

/*
 * Visibility part of Scene::evalTransmittance (scene.cpp:619-642) for scenes without ENull surfaces:
 * returns true when the open segment p1 -> p2 is blocked.  dir / remaining are returned for the medium term.
 */
__device__ __forceinline__ bool segment_occluded(const SceneDev &sc, const F3 &p1, bool p1OnSurface, const F3 &p2,
                                                 F3 &dir, float &remaining, uint32_t *hitPrim = nullptr) {
    F3 d = xsub3(p2, p1);
    remaining = xlen(d);
    dir = xdivv(d, remaining);
    if (hitPrim) *hitPrim = ALVRL_NO_HIT;
    if (!(remaining > 0)) return false;
    float t, u, v; uint32_t prim;
    bool hit;
    if (sc.anyHit && !hitPrim) hit = scene_intersect<true>(sc, p1, dir, p1OnSurface ? ALVRL_EPSILON : 0.0f, remaining, false, t, prim, u, v);
    else hit = scene_intersect<false>(sc, p1, dir, p1OnSurface ? ALVRL_EPSILON : 0.0f, remaining, false, t, prim, u, v);
    if (hitPrim) *hitPrim = prim;
    return hit;
}

/*
 * Fast-flavour occlusion query: same semantics (is the open segment blocked by any triangle within [mint, maxt]?), but
 * FMA slab tests on precomputed o/d, plane + barycentric triangle records without the axis switch, MUFU reciprocals,
 * and a while-while loop (lanes look for their next leaf together, then test triangles together) to keep warps
 * converged.  The clip against the tree AABB is dropped: every triangle lies inside it, so it cannot remove a hit.
 * Decisions can differ from the exact path only for rays grazing an edge within ~1e-6 (documented).
 */
__device__ __forceinline__ bool bvh_occluded_fast(const SceneDev &sc, const F3 &o, const F3 &d, float mint, float maxt) {
    const F3 inv = f3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));
    const F3 oi = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    const float lo_t = mint, hi_t = maxt * 1.00001f;
    const uint32_t numNodes = sc.numNodes;
    uint32_t node = 0;
    for (;;) {
        uint32_t leaf = 0;
        while (node < numNodes) {
            const float4 lo = __ldg(&sc.nodes[node].lo), hi = __ldg(&sc.nodes[node].hi);
            const float tx1 = fmaf(lo.x, inv.x, oi.x), tx2 = fmaf(hi.x, inv.x, oi.x);
            const float ty1 = fmaf(lo.y, inv.y, oi.y), ty2 = fmaf(hi.y, inv.y, oi.y);
            const float tz1 = fmaf(lo.z, inv.z, oi.z), tz2 = fmaf(hi.z, inv.z, oi.z);
            const float tn = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), lo_t));
            const float tf = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), hi_t));
            const uint32_t esc = __float_as_uint(lo.w);
            if (tn <= tf) {
                const uint32_t lf = __float_as_uint(hi.w);
                if (lf) { leaf = lf; node = esc; break; }
                node++;
            } else node = esc;
        }
        if (!leaf) return false;
        const uint32_t first = leaf >> 4, cnt = leaf & 15u;
        for (uint32_t i = 0; i < cnt; i++) {
            const float4 p = __ldg(&sc.trisFast[first + i].p);
            const float den = p.x * d.x + p.y * d.y + p.z * d.z;
            const float num = p.w - (p.x * o.x + p.y * o.y + p.z * o.z);
            const float t = __fdividef(num, den);
            if (!(t >= mint && t <= maxt)) continue;
            const float4 q = __ldg(&sc.trisFast[first + i].q), r = __ldg(&sc.trisFast[first + i].r);
            const F3 P = f3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
            const float u = q.x * P.x + q.y * P.y + q.z * P.z + q.w;
            const float v = r.x * P.x + r.y * P.y + r.z * P.z + r.w;
            if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) return true;
        }
    }
}

/*
 * The same query, called by ALL 32 lanes of a warp together (need = false: this lane has no ray).  The per-lane loop above
 * leaves the reconvergence of lanes that left the node search for a leaf to the hardware scheduler, and on a large tree
 * (C4: 1.2 M nodes, ~80 node visits per ray) the warp falls apart into single lanes: ncu counted 1.8 active lanes per
 * instruction in that loop (profiles/r2_c4_bvh_v0.txt).  Here both loops are controlled by warp votes, so every branch
 * that depends on a lane is a short forward `if`: lanes search for their next leaf together (a lane that found one, or
 * finished, idles until the vote ends the search), then test the triangles of their leaves together.
 */
__device__ __forceinline__ bool bvh_occluded_fast_warp(const SceneDev &sc, const F3 &o, const F3 &d, float mint, float maxt, bool need) {
    const F3 inv = f3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));
    const F3 oi = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    const float lo_t = mint, hi_t = maxt * 1.00001f;
    const uint32_t numNodes = sc.numNodes;
    uint32_t node = (need && maxt > mint) ? 0u : numNodes;
    bool hit = false;
    __syncwarp();
    while (__any_sync(0xffffffffu, node < numNodes)) {
        uint32_t leaf = 0;
        while (__any_sync(0xffffffffu, node < numNodes && !leaf)) {
            if (node < numNodes && !leaf) {
                const float4 lo = __ldg(&sc.nodes[node].lo), hi = __ldg(&sc.nodes[node].hi);
                const float tx1 = fmaf(lo.x, inv.x, oi.x), tx2 = fmaf(hi.x, inv.x, oi.x);
                const float ty1 = fmaf(lo.y, inv.y, oi.y), ty2 = fmaf(hi.y, inv.y, oi.y);
                const float tz1 = fmaf(lo.z, inv.z, oi.z), tz2 = fmaf(hi.z, inv.z, oi.z);
                const float tn = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), lo_t));
                const float tf = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), hi_t));
                const uint32_t esc = __float_as_uint(lo.w), lf = __float_as_uint(hi.w);
                const bool in = tn <= tf;
                leaf = in ? lf : 0u;
                node = (in && !lf) ? node + 1u : esc;
            }
        }
        if (leaf) {
            const uint32_t first = leaf >> 4, cnt = leaf & 15u;
            for (uint32_t i = 0; i < cnt; i++) {
                const float4 p = __ldg(&sc.trisFast[first + i].p);
                const float den = p.x * d.x + p.y * d.y + p.z * d.z;
                const float num = p.w - (p.x * o.x + p.y * o.y + p.z * o.z);
                const float t = __fdividef(num, den);
                if (!(t >= mint && t <= maxt)) continue;
                const float4 q = __ldg(&sc.trisFast[first + i].q), r = __ldg(&sc.trisFast[first + i].r);
                const F3 P = f3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
                const float u = q.x * P.x + q.y * P.y + q.z * P.z + q.w;
                const float v = r.x * P.x + r.y * P.y + r.z * P.z + r.w;
                if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) { hit = true; node = numNodes; break; }
            }
        }
    }
    return hit;
}

/*
 * Any-hit query on the 4-wide tree (types.h::Bvh4Node, bvh.h::collapse4), called by ALL 32 lanes of a warp together.  A
 * lane keeps its pending subtrees and leaves on a small stack in local memory; one round of the loop expands the inner
 * node a lane holds (four slab tests on one 128-byte line, hits pushed) or tests the triangles of the leaf it holds, then
 * pops.  The loop and both phases are controlled by warp votes.  A quarter of the dependent node fetches of the binary
 * tree (C4: 21 instead of 81 per ray), which is what bounds a traversal that waits for L2 on every step.
 */
#define ALVRL_BVH4_STACK 64
__device__ __forceinline__ bool bvh4_occluded_warp(const SceneDev &sc, const F3 &o, const F3 &d, float mint, float maxt, bool need) {
    const F3 inv = f3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));
    const F3 oi = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    const float lo_t = mint, hi_t = maxt * 1.00001f;
    const int DONE = (int) 0x80000000u;
    int stack[ALVRL_BVH4_STACK];
    int sp = 0;
    int cur = (need && maxt > mint) ? 0 : DONE;
    bool hit = false;
    __syncwarp();
    while (__any_sync(0xffffffffu, cur != DONE)) {
        if (cur >= 0) {                                                    /* inner node: test the four children */
            const Bvh4Node *nd = sc.nodes4 + cur;
            const float4 lox = __ldg(&nd->lox), hix = __ldg(&nd->hix), loy = __ldg(&nd->loy), hiy = __ldg(&nd->hiy),
                         loz = __ldg(&nd->loz), hiz = __ldg(&nd->hiz);
            const int4 ch = __ldg(&nd->child);
#define ALVRL_B4_CHILD(C, K)                                                                                              \
            {                                                                                                             \
                const float tx1 = fmaf(lox.C, inv.x, oi.x), tx2 = fmaf(hix.C, inv.x, oi.x);                              \
                const float ty1 = fmaf(loy.C, inv.y, oi.y), ty2 = fmaf(hiy.C, inv.y, oi.y);                              \
                const float tz1 = fmaf(loz.C, inv.z, oi.z), tz2 = fmaf(hiz.C, inv.z, oi.z);                              \
                const float tn = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), lo_t));            \
                const float tf = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), hi_t));            \
                if (tn <= tf && K != DONE && sp < ALVRL_BVH4_STACK) stack[sp++] = K;                                      \
            }
            ALVRL_B4_CHILD(x, ch.x) ALVRL_B4_CHILD(y, ch.y) ALVRL_B4_CHILD(z, ch.z) ALVRL_B4_CHILD(w, ch.w)
#undef ALVRL_B4_CHILD
            cur = sp ? stack[--sp] : DONE;
        }
        if (__any_sync(0xffffffffu, cur < 0 && cur != DONE)) {
            if (cur < 0 && cur != DONE) {                                  /* leaf: ~((first << 4) | count) */
                const uint32_t lf = (uint32_t) ~cur;
                const uint32_t first = lf >> 4, cnt = lf & 15u;
                for (uint32_t i = 0; i < cnt; i++) {
                    const float4 p = __ldg(&sc.trisFast[first + i].p);
                    const float den = p.x * d.x + p.y * d.y + p.z * d.z;
                    const float num = p.w - (p.x * o.x + p.y * o.y + p.z * o.z);
                    const float t = __fdividef(num, den);
                    if (!(t >= mint && t <= maxt)) continue;
                    const float4 q = __ldg(&sc.trisFast[first + i].q), r = __ldg(&sc.trisFast[first + i].r);
                    const F3 P = f3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
                    const float u = q.x * P.x + q.y * P.y + q.z * P.z + q.w;
                    const float v = r.x * P.x + r.y * P.y + r.z * P.z + r.w;
                    if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) { hit = true; break; }
                }
                cur = (hit || !sp) ? DONE : stack[--sp];
            }
        }
    }
    return hit;
}

/* ---- media (exact flavour: used by the primary kernel and the strict transport flavour) ---------- */
__device__ __forceinline__ float exp_ref(float x) { return (float) exp((double) x); }   /* math::fastexp, math.h:185-187 */

/* GridDataSource::lookupFloat (gridvolume.cpp:337-388) */
__device__ __forceinline__ float grid_lookup(const MediumDev &m, const F3 &p_) {
    const float px = xadd(xmul(m.gsc[0], p_.x), m.gtr[0]), py = xadd(xmul(m.gsc[1], p_.y), m.gtr[1]), pz = xadd(xmul(m.gsc[2], p_.z), m.gtr[2]);
    const int x1 = (int) floorf(px), y1 = (int) floorf(py), z1 = (int) floorf(pz);
    const int x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
    if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= m.res[0] || y2 >= m.res[1] || z2 >= m.res[2]) return 0.0f;
    const float fx = xsub(px, (float) x1), fy = xsub(py, (float) y1), fz = xsub(pz, (float) z1);
    const float _fx = xsub(1.0f, fx), _fy = xsub(1.0f, fy), _fz = xsub(1.0f, fz);
    const float *d = m.density;
    const size_t rx = m.res[0], ry = m.res[1];
    const float d000 = __ldg(&d[(z1 * ry + y1) * rx + x1]), d001 = __ldg(&d[(z1 * ry + y1) * rx + x2]),
                d010 = __ldg(&d[(z1 * ry + y2) * rx + x1]), d011 = __ldg(&d[(z1 * ry + y2) * rx + x2]),
                d100 = __ldg(&d[(z2 * ry + y1) * rx + x1]), d101 = __ldg(&d[(z2 * ry + y1) * rx + x2]),
                d110 = __ldg(&d[(z2 * ry + y2) * rx + x1]), d111 = __ldg(&d[(z2 * ry + y2) * rx + x2]);
    const float a = xmul(xadd(xmul(xadd(xmul(d000, _fx), xmul(d001, fx)), _fy), xmul(xadd(xmul(d010, _fx), xmul(d011, fx)), fy)), _fz);
    const float b = xmul(xadd(xmul(xadd(xmul(d100, _fx), xmul(d101, fx)), _fy), xmul(xadd(xmul(d110, _fx), xmul(d111, fx)), fy)), fz);
    return xadd(a, b);
}

/* HeterogeneousMedium::integrateDensity (heterogeneous.cpp:301-376), composite Simpson; includes m_scale */
__device__ __forceinline__ float grid_optical_depth(const MediumDev &m, const F3 &o, const F3 &d, float rmint, float rmaxt) {
    const F3 dRcp = f3(xdiv(1.0f, d.x), xdiv(1.0f, d.y), xdiv(1.0f, d.z));
    float mint, maxt;
    if (!aabb_clip(m.bmin, m.bmax, o, d, dRcp, mint, maxt)) return 0.0f;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    const float length = xsub(maxt, mint);
    F3 p = xadd3(o, xscale(d, mint)), pLast = xadd3(o, xscale(d, maxt));
    float maxComp = 0;
    maxComp = fmaxf(fmaxf(maxComp, fabsf(p.x)), fabsf(pLast.x));
    maxComp = fmaxf(fmaxf(maxComp, fabsf(p.y)), fabsf(pLast.y));
    maxComp = fmaxf(fmaxf(maxComp, fabsf(p.z)), fabsf(pLast.z));
    if (length < xmul(1e-6f, maxComp)) return 0.0f;
    uint32_t nSteps = (uint32_t) ceilf(xdiv(length, m.stepSize));
    nSteps += nSteps % 2;
    const float stepSz = xdiv(length, (float) nSteps);
    const F3 inc = xscale(d, stepSz);
    float integrated = xadd(grid_lookup(m, p), grid_lookup(m, pLast));
    /* HETVOL_EARLY_EXIT (heterogeneous.cpp:31, 336-340, 353-360): past -log(Epsilon) of optical depth the march stops with an
     * infinite optical depth (transmittance exactly 0); -(float) log((double) 1e-4f) = 9.21034f */
    const float stopValue = xdiv(xmul(9.21034049987793f, 3.0f), xmul(stepSz, m.scale));
    p = xadd3(p, inc);
    float mm = 4;
    for (uint32_t i = 1; i < nSteps; ++i) {
        integrated = xadd(integrated, xmul(mm, grid_lookup(m, p)));
        mm = 6 - mm;
        if (integrated > stopValue) return INFINITY;
        F3 next = xadd3(p, inc);
        if (p.x == next.x && p.y == next.y && p.z == next.z) break;
        p = next;
    }
    return xmul(xmul(xmul(integrated, m.scale), stepSz), 1.0f / 3.0f);
}

/* Medium::eval -> transmittance only (homogeneous.cpp:387,394-395; heterogeneous.cpp:667,679), exact flavour */
__device__ __forceinline__ void medium_transmittance_exact(const MediumDev &m, const F3 &o, const F3 &d, float dist, float T[3]) {
    if (m.type == 0) {
        float mx = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) { T[i] = exp_ref(xmul(m.sigmaT[i], -dist)); mx = fmaxf(mx, T[i]); }
        if (mx < 1e-20f) T[0] = T[1] = T[2] = 0.0f;
    } else {
        const float e = exp_ref(-grid_optical_depth(m, o, d, 0.0f, dist));
        T[0] = T[1] = T[2] = e;
    }
}
