/*
 * occ_query.h -- any-hit query against the compiled occluder set of occluders.h; one source for the transport kernels
 * (shared-memory arrays, MUFU reciprocal) and for the host-side unit test (libalvrl_host.so, IEEE division).
 *
 * Uniform loops: every slab clips the parametric interval of its solid, every plane contributes one bit "crossed
 * inside [tmin, tmax]"; triangle records are only touched for crossed planes.  On the device all lanes of a warp call
 * it together (need = false: the lane has no ray) so that the loops stay converged and the loads are broadcasts.
 */
#pragma once
#include <stdint.h>
#include <math.h>
#include <cuda_runtime.h>

#ifdef __CUDA_ARCH__
#define ALVRL_OCC_HD __device__ __forceinline__
#define ALVRL_OCC_RCP(x) alvrl_occ_rcp(x)
__device__ __forceinline__ float alvrl_occ_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define ALVRL_OCC_FFS(m) (__ffs(m) - 1)
#define ALVRL_OCC_BITS(f) __float_as_uint(f)
#else
#define ALVRL_OCC_HD inline
#define ALVRL_OCC_RCP(x) (1.0f / (x))
#define ALVRL_OCC_FFS(m) (__builtin_ctz(m))
static inline uint32_t alvrl_occ_bits(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }
#define ALVRL_OCC_BITS(f) alvrl_occ_bits(f)
#endif

ALVRL_OCC_HD bool occ_query(const float4 *slabA, const float2 *slabB, uint32_t numSlabs, const float4 *planes, const uint32_t *planeInfo,
                            uint32_t numPlanes, const float4 *tris, float ox, float oy, float oz, float dx, float dy, float dz, float tmin,
                            float tmax, bool need) {
    const float INF = INFINITY;
    bool hit = false;
    float tn = -INF, tf = INF;
#ifdef __CUDA_ARCH__
#pragma unroll 3
#endif
    for (uint32_t i = 0; i < numSlabs; i++) {
        const float4 a = slabA[i];
        const float2 b = slabB[i];
        const float den = fmaf(a.x, dx, fmaf(a.y, dy, a.z * dz));
        const float no = fmaf(a.x, ox, fmaf(a.y, oy, a.z * oz));
        const float r = ALVRL_OCC_RCP(den);
        const float t1 = (a.w - no) * r, t2 = (b.x - no) * r;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
        if (ALVRL_OCC_BITS(b.y)) {
            /* last slab of a solid: the segment touches its boundary iff the clipped line is non-empty and enters or
             * leaves the solid inside [tmin, tmax] */
            hit |= tn <= tf && ((tn >= tmin && tn <= tmax) || (tf >= tmin && tf <= tmax));
            tn = -INF; tf = INF;
        }
    }
    uint32_t mask = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 5
#endif
    for (uint32_t i = 0; i < numPlanes; i++) {
        const float4 p = planes[i];
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
        const float4 p = planes[i];
        const float den = fmaf(p.x, dx, fmaf(p.y, dy, p.z * dz));
        const float no = fmaf(p.x, ox, fmaf(p.y, oy, p.z * oz));
        const float t = (p.w - no) * ALVRL_OCC_RCP(den);
        const float Px = fmaf(t, dx, ox), Py = fmaf(t, dy, oy), Pz = fmaf(t, dz, oz);
        const uint32_t info = planeInfo[i], first = info >> 8, cnt = info & 255u;
        for (uint32_t k = 0; k < cnt; k++) {
            const float4 q = tris[3 * (first + k) + 1], w = tris[3 * (first + k) + 2];
            const float u = fmaf(q.x, Px, fmaf(q.y, Py, fmaf(q.z, Pz, q.w)));
            const float v = fmaf(w.x, Px, fmaf(w.y, Py, fmaf(w.z, Pz, w.w)));
            if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) return true;
        }
    }
    return false;
}
