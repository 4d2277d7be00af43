/*
 * occ_query.h -- any-hit query against the compiled occluder set of occluders.h; one source for the transport kernels
 * and for the host-side unit test (libalvrl_host.so, IEEE division instead of MUFU.RCP).
 *
 * The set travels in the kernel parameters (constant bank): the loops below are warp-uniform, so slab and plane records
 * arrive through the uniform datapath (LDCU) instead of the LSU.  Boxes -- solids of exactly three finite slabs -- come
 * first and are clipped by an unrolled body without flag tests; other convex solids follow as flagged slab runs; every
 * plane of a planar group contributes one bit "crossed inside [tmin, tmax]", and triangle records (shared memory) are
 * only touched for crossed planes.  On the device all lanes of a warp call it together (need = false: the lane has no
 * ray) so that the loops stay converged.
 */
#pragma once
#include <stdint.h>
#include <math.h>
#include <cuda_runtime.h>

#define ALVRL_OCC_MAX_SLABS 24
#define ALVRL_OCC_MAX_PLANES 24
#define ALVRL_OCC_MAX_TRIS 128

struct OccDev {
    float4 slabA[ALVRL_OCC_MAX_SLABS];       /* n.xyz, c_lo  (c_lo <= n.x <= c_hi; a lone half-space has c_lo = -inf) */
    float2 slabB[ALVRL_OCC_MAX_SLABS];       /* c_hi, bits != 0: last slab of its solid */
    float4 planes[ALVRL_OCC_MAX_PLANES];     /* n.xyz, c */
    uint32_t planeInfo[ALVRL_OCC_MAX_PLANES];/* first << 8 | count into the triangle records */
    uint32_t numBoxes, numSlabs, numPlanes, numTris;   /* slabs [0, 3 numBoxes) are the boxes */
    float cullMargin;                                  /* 1e-5 x scene extent: strictness of the pair-level side tests */
};

#ifdef __CUDACC__
__device__ __forceinline__ float alvrl_occ_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
#ifdef __CUDA_ARCH__
#define ALVRL_OCC_HD __device__ __forceinline__
#define ALVRL_OCC_RCP(x) alvrl_occ_rcp(x)
#define ALVRL_OCC_FFS(m) (__ffs(m) - 1)
#define ALVRL_OCC_BITS(f) __float_as_uint(f)
#else
#define ALVRL_OCC_HD inline
#define ALVRL_OCC_RCP(x) (1.0f / (x))
#define ALVRL_OCC_FFS(m) (__builtin_ctz(m))
static inline uint32_t alvrl_occ_bits(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }
#define ALVRL_OCC_BITS(f) alvrl_occ_bits(f)
#endif

/* clip the parametric interval [tn, tf] of the line o + t d by the slab c_lo <= n.x <= c_hi */
#define ALVRL_OCC_CLIP(a, chi)                                                     \
    do {                                                                           \
        const float den_ = fmaf((a).x, dx, fmaf((a).y, dy, (a).z * dz));           \
        const float no_ = fmaf((a).x, ox, fmaf((a).y, oy, (a).z * oz));            \
        const float r_ = ALVRL_OCC_RCP(den_);                                      \
        const float t1_ = ((a).w - no_) * r_, t2_ = ((chi) - no_) * r_;            \
        tn = fmaxf(tn, fminf(t1_, t2_));                                           \
        tf = fminf(tf, fmaxf(t1_, t2_));                                           \
    } while (0)
/* the segment touches the solid's boundary iff the clipped line is non-empty and enters or leaves it inside [tmin, tmax] */
#define ALVRL_OCC_TOUCH() (tn <= tf && ((tn >= tmin && tn <= tmax) || (tf >= tmin && tf <= tmax)))

/*
 * Pair-level conservative culling.  Every shadow segment of one (camera segment, VRL) pair joins a point of the camera
 * segment [E, Usurf] (vol->vol) or Usurf itself (vol->surf) to a point of the VRL [S, End]: it stays inside the convex hull
 * of those points.  If all of them lie strictly beyond one face of a slab, no segment of the pair can touch that solid; if
 * all of them lie strictly on one side of a plane, none can cross it.  "Beyond" is recorded once per camera segment and once
 * per VRL as bit masks (margin = 1e-5 x scene extent, >> the fp32 error of the per-ray tests, so a culled test could not
 * have reported a hit), and a pair's culled set is the AND of the two.  A camera segment ENDS on a surface: Usurf only has
 * to be on the plane's side up to the margin (the ray towards a point strictly on that side leaves the surface).
 *   slab side bits : slab k < 16 -> bit 2k "below c_lo", bit 2k+1 "above c_hi"
 *   plane bits     : plane i < 16 -> bit i "positive side", bit 16+i "negative side"
 * Solids / planes beyond those indices are never culled.  boxActive / planeActive of occ_query are warp-uniform masks of
 * what still has to be tested (bit b: box b; bit i: plane i).
 */
#define ALVRL_OCC_MASK_SLABS 16
#define ALVRL_OCC_MASK_PLANES 16
/* side bits of one point set {p0, p1}; loose1: p1 may sit on a plane (within the margin) */
ALVRL_OCC_HD uint32_t occ_slab_sides(const OccDev &oc, float ax, float ay, float az, float bx, float by, float bz, float m) {
    uint32_t bits = 0;
    const uint32_t ns = oc.numSlabs < ALVRL_OCC_MASK_SLABS ? oc.numSlabs : ALVRL_OCC_MASK_SLABS;
    for (uint32_t k = 0; k < ns; k++) {
        const float4 a = oc.slabA[k];
        const float chi = oc.slabB[k].x;
        const float da = fmaf(a.x, ax, fmaf(a.y, ay, a.z * az)), db = fmaf(a.x, bx, fmaf(a.y, by, a.z * bz));
        bits |= ((da < a.w - m && db < a.w - m) ? 1u : 0u) << (2 * k);
        bits |= ((da > chi + m && db > chi + m) ? 1u : 0u) << (2 * k + 1);
    }
    return bits;
}
ALVRL_OCC_HD uint32_t occ_plane_sides(const OccDev &oc, float ax, float ay, float az, float bx, float by, float bz, float m, bool looseB) {
    uint32_t bits = 0;
    const uint32_t np = oc.numPlanes < ALVRL_OCC_MASK_PLANES ? oc.numPlanes : ALVRL_OCC_MASK_PLANES;
    const float mb = looseB ? -m : m;
    for (uint32_t i = 0; i < np; i++) {
        const float4 p = oc.planes[i];
        const float da = fmaf(p.x, ax, fmaf(p.y, ay, p.z * az)) - p.w, db = fmaf(p.x, bx, fmaf(p.y, by, p.z * bz)) - p.w;
        bits |= ((da > m && db > mb) ? 1u : 0u) << i;
        bits |= ((da < -m && db < -mb) ? 1u : 0u) << (16 + i);
    }
    return bits;
}
/* boxes of a lane that the AND of a segment's and a VRL's slab-side bits culls: bit b = some face separates box b */
ALVRL_OCC_HD uint32_t occ_boxes_culled(uint32_t sideBits, uint32_t numBoxes) {
    uint32_t c = 0;
    for (uint32_t b = 0; b < numBoxes && 6 * b < 32; b++) c |= (((sideBits >> (6 * b)) & 63u) ? 1u : 0u) << b;
    return c;
}

ALVRL_OCC_HD bool occ_query(const OccDev &oc, const float4 *tris, float ox, float oy, float oz, float dx, float dy, float dz, float tmin,
                            float tmax, bool need, uint32_t boxActive = 0xffffffffu, uint32_t planeActive = 0xffffffffu) {
    const float INF = INFINITY;
    bool hit = false;
    /* the first two boxes (the Cornell class has exactly two) are tested by straight-line code at fixed parameter offsets; a
     * loop takes the rest.  Nothing left to test: neither runs */
    const uint32_t nbAll = oc.numBoxes;
#define ALVRL_OCC_BOX(B)                                                          \
    do {                                                                          \
        float tn = -INF, tf = INF;                                                \
        ALVRL_OCC_CLIP(oc.slabA[3 * (B)], oc.slabB[3 * (B)].x);                   \
        ALVRL_OCC_CLIP(oc.slabA[3 * (B) + 1], oc.slabB[3 * (B) + 1].x);           \
        ALVRL_OCC_CLIP(oc.slabA[3 * (B) + 2], oc.slabB[3 * (B) + 2].x);           \
        hit |= ALVRL_OCC_TOUCH();                                                 \
    } while (0)
    if ((boxActive & 1u) && nbAll > 0u) ALVRL_OCC_BOX(0);
    if ((boxActive & 2u) && nbAll > 1u) ALVRL_OCC_BOX(1);
    if (boxActive >> 2)
        for (uint32_t b = 2; b < nbAll; b++) {
            if (!((boxActive >> b) & 1u)) continue;
            ALVRL_OCC_BOX(b);
        }
#undef ALVRL_OCC_BOX
    {
        float tn = -INF, tf = INF;
        for (uint32_t i = 3 * oc.numBoxes; i < oc.numSlabs; i++) {
            ALVRL_OCC_CLIP(oc.slabA[i], oc.slabB[i].x);
            if (ALVRL_OCC_BITS(oc.slabB[i].y)) {
                hit |= ALVRL_OCC_TOUCH();
                tn = -INF; tf = INF;
            }
        }
    }
    uint32_t mask = 0;
    const uint32_t np = planeActive ? oc.numPlanes : 0u;
    for (uint32_t i = 0; i < np; i++) {
        if (!((planeActive >> i) & 1u)) continue;
        const float4 p = oc.planes[i];
        const float den = fmaf(p.x, dx, fmaf(p.y, dy, p.z * dz));
        const float no = fmaf(p.x, ox, fmaf(p.y, oy, p.z * oz));
        const float t = (p.w - no) * ALVRL_OCC_RCP(den);
        mask |= (t >= tmin && t <= tmax ? 1u : 0u) << i;
    }
    if (!need || !(tmax > tmin)) return false;
    if (hit) return true;
    while (mask) {
        const uint32_t i = ALVRL_OCC_FFS(mask);
        mask &= mask - 1;
        const float4 p = oc.planes[i];
        const float den = fmaf(p.x, dx, fmaf(p.y, dy, p.z * dz));
        const float no = fmaf(p.x, ox, fmaf(p.y, oy, p.z * oz));
        const float t = (p.w - no) * ALVRL_OCC_RCP(den);
        const float Px = fmaf(t, dx, ox), Py = fmaf(t, dy, oy), Pz = fmaf(t, dz, oz);
        const uint32_t info = oc.planeInfo[i], first = info >> 8, cnt = info & 255u;
        for (uint32_t k = 0; k < cnt; k++) {
            const float4 q = tris[3 * (first + k) + 1], w = tris[3 * (first + k) + 2];
            const float u = fmaf(q.x, Px, fmaf(q.y, Py, fmaf(q.z, Pz, q.w)));
            const float v = fmaf(w.x, Px, fmaf(w.y, Py, fmaf(w.z, Pz, w.w)));
            if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) return true;
        }
    }
    return false;
}
