/*
 * walk_common.cuh -- what the random walks on the device share (tracer.cu: light particles, vrlTracer.h; volpath.cu: camera
 * paths, volpath.cpp): the addressed sample stream, the warps of src/libcore/warp.cpp, frames, Fresnel terms -- all in exact
 * arithmetic (single operations with IEEE rounding, sin / cos / log through double as the oracle pins them), so that a walk
 * is the oracle's walk bit for bit.
 */
#pragma once
#include "dev_common.cuh"
#include "../../include/alvrl_rng.h"

namespace alvrl {

struct TracerStream {
    uint32_t key, k;
    __device__ __forceinline__ float next() { return alvrl_rng_uniform(key, k++); }
};

__device__ __forceinline__ float safe_sqrt_t(float v) { return xsqrt(fmaxf(0.0f, v)); }
__device__ __forceinline__ void sincos_t(float theta, float &s, float &c) { s = (float) sin((double) theta); c = (float) cos((double) theta); }
#define ALVRL_PI_D 3.14159265358979323846

/* warp::squareToUniformSphere, warp.cpp:25-31 */
__device__ __forceinline__ F3 square_to_uniform_sphere(float sx, float sy) {
    const float z = xsub(1.0f, xmul(2.0f, sy));
    const float r = safe_sqrt_t(xsub(1.0f, xmul(z, z)));
    float sinPhi, cosPhi;
    sincos_t((float) ((double) 2.0f * ALVRL_PI_D * (double) sx), sinPhi, cosPhi);
    return f3(xmul(r, cosPhi), xmul(r, sinPhi), z);
}
/* warp::squareToCosineHemisphere over squareToUniformDiskConcentric, warp.cpp:43-52, 81-102 */
__device__ __forceinline__ F3 square_to_cosine_hemisphere(float sx, float sy) {
    const float r1 = xsub(xmul(2.0f, sx), 1.0f), r2 = xsub(xmul(2.0f, sy), 1.0f);
    float phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (xmul(r1, r1) > xmul(r2, r2)) { r = r1; phi = (float) ((ALVRL_PI_D / (double) 4.0f) * (double) xdiv(r2, r1)); }
    else { r = r2; phi = (float) ((ALVRL_PI_D / (double) 2.0f) - (double) xdiv(r1, r2) * (ALVRL_PI_D / (double) 4.0f)); }
    float cosPhi, sinPhi;
    sincos_t(phi, sinPhi, cosPhi);
    const float px = xmul(r, cosPhi), py = xmul(r, sinPhi);
    float z = safe_sqrt_t(xsub(xsub(1.0f, xmul(px, px)), xmul(py, py)));
    if (z == 0) z = 1e-10f;
    return f3(px, py, z);
}
/* coordinateSystem, util.cpp:592-601, and Frame(n).toWorld(v) */
__device__ __forceinline__ F3 frame_to_world(const F3 &a, const F3 &v) {
    F3 c;
    if (fabsf(a.x) > fabsf(a.y)) { const float invLen = xdiv(1.0f, xsqrt(xadd(xmul(a.x, a.x), xmul(a.z, a.z)))); c = f3(xmul(a.z, invLen), 0.0f, xmul(-a.x, invLen)); }
    else { const float invLen = xdiv(1.0f, xsqrt(xadd(xmul(a.y, a.y), xmul(a.z, a.z)))); c = f3(0.0f, xmul(a.z, invLen), xmul(-a.y, invLen)); }
    const F3 b = xcross(c, a);
    return xadd3(xadd3(xscale(b, v.x), xscale(c, v.y)), xscale(a, v.z));
}
__device__ __forceinline__ float fresnel_dielectric_ext_t(float cosThetaI_, float &cosThetaT_, float eta) {       /* util.cpp:651-681 */
    if (eta == 1.0f) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    const float scale = (cosThetaI_ > 0) ? xdiv(1.0f, eta) : eta;
    const float cosThetaTSqr = xsub(1.0f, xmul(xsub(1.0f, xmul(cosThetaI_, cosThetaI_)), xmul(scale, scale)));
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    const float cosThetaI = fabsf(cosThetaI_);
    const float cosThetaT = xsqrt(cosThetaTSqr);
    const float Rs = xdiv(xsub(cosThetaI, xmul(eta, cosThetaT)), xadd(cosThetaI, xmul(eta, cosThetaT)));
    const float Rp = xdiv(xsub(xmul(eta, cosThetaI), cosThetaT), xadd(xmul(eta, cosThetaI), cosThetaT));
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return xmul(0.5f, xadd(xmul(Rs, Rs), xmul(Rp, Rp)));
}
__device__ __forceinline__ float fresnel_conductor_exact_t(float cosThetaI, float eta, float k) {                   /* util.cpp:739-761 */
    const float cosThetaI2 = xmul(cosThetaI, cosThetaI), sinThetaI2 = xsub(1.0f, cosThetaI2), sinThetaI4 = xmul(sinThetaI2, sinThetaI2);
    const float temp1 = xsub(xsub(xmul(eta, eta), xmul(k, k)), sinThetaI2);
    const float a2pb2 = safe_sqrt_t(xadd(xmul(temp1, temp1), xmul(xmul(xmul(xmul(k, k), eta), eta), 4.0f)));
    const float a = safe_sqrt_t(xmul(xadd(a2pb2, temp1), 0.5f));
    const float term1 = xadd(a2pb2, cosThetaI2), term2 = xmul(a, xmul(2.0f, cosThetaI));
    const float Rs2 = xdiv(xsub(term1, term2), xadd(term1, term2));
    const float term3 = xadd(xmul(a2pb2, cosThetaI2), sinThetaI4), term4 = xmul(term2, sinThetaI2);
    const float Rp2 = xdiv(xmul(Rs2, xsub(term3, term4)), xadd(term3, term4));
    return xmul(0.5f, xadd(Rp2, Rs2));
}

struct EmitterDev { const uint32_t *tris; const float *cdf; uint32_t n; float power[3]; };

/* HeterogeneousMedium::invertDensityIntegral, heterogeneous.cpp:422-545 (the oracle's Medium::invertDensityIntegral operation for
 * operation): composite Simpson march from ray.mint until the optical depth reaches desiredDensity, then Newton-bisection on the
 * quadratic through the last three lookups.  Returns false when [rmint, rmaxt] holds less than desiredDensity; integratedDensity
 * then is what the segment holds. */
__device__ __forceinline__ bool grid_invert_density_integral(const MediumDev &m, const F3 &o, const F3 &d, float rmint, float rmaxt, float desiredDensity,
                                                             float &integratedDensity, float &t, float &densityAtT) {
    integratedDensity = 0.0f; densityAtT = 0.0f; t = 0.0f;
    const F3 dRcp = f3(xdiv(1.0f, d.x), xdiv(1.0f, d.y), xdiv(1.0f, d.z));
    float mint, maxt;
    if (!aabb_clip(m.bmin, m.bmax, o, d, dRcp, mint, maxt)) return false;
    mint = fmaxf(mint, rmint);
    maxt = fminf(maxt, rmaxt);
    const float length = xsub(maxt, mint);
    F3 p = xadd3(o, xscale(d, mint));
    const F3 pLast = xadd3(o, xscale(d, maxt));
    float maxComp = 0;
    maxComp = fmaxf(fmaxf(maxComp, fabsf(p.x)), fabsf(pLast.x));
    maxComp = fmaxf(fmaxf(maxComp, fabsf(p.y)), fabsf(pLast.y));
    maxComp = fmaxf(fmaxf(maxComp, fabsf(p.z)), fabsf(pLast.z));
    if (length < xmul(1e-6f, maxComp)) return false;
    const uint32_t nSteps = (uint32_t) ceilf(xdiv(length, xmul(2.0f, m.stepSize)));
    const float stepSz = xdiv(length, (float) nSteps), multiplier = xmul(xmul(1.0f / 6.0f, stepSz), m.scale);
    const F3 fullStep = xscale(d, stepSz), halfStep = xscale(fullStep, 0.5f);
    float node1 = grid_lookup(m, p);
    for (uint32_t i = 0; i < nSteps; ++i) {
        const float node2 = grid_lookup(m, xadd3(p, halfStep)), node3 = grid_lookup(m, xadd3(p, fullStep));
        const float newDensity = xadd(integratedDensity, xmul(multiplier, xadd(xadd(node1, xmul(node2, 4.0f)), node3)));
        if (newDensity >= desiredDensity) {
            float a = 0, b = stepSz, x = a, fx = xsub(integratedDensity, desiredDensity);
            const float stepSizeSqr = xmul(stepSz, stepSz), temp = xdiv(m.scale, stepSizeSqr);
            /* the coefficients of the Lagrange polynomial: (3 node1 - 4 node2 + node3), (node1 - 2 node2 + node3) */
            const float c1 = xadd(xsub(xmul(3.0f, node1), xmul(4.0f, node2)), node3), c2 = xadd(xsub(node1, xmul(2.0f, node2)), node3);
            int it = 1;
            while (true) {
                const float dfx = xmul(temp, xadd(xsub(xmul(node1, stepSizeSqr), xmul(xmul(c1, stepSz), x)), xmul(xmul(xmul(2.0f, c2), x), x)));
                x = xsub(x, xdiv(fx, dfx));
                if (x <= a || x >= b || dfx == 0) x = xmul(0.5f, xadd(b, a));
                const float poly = xadd(xsub(xmul(xmul(6.0f, node1), stepSizeSqr), xmul(xmul(xmul(3.0f, c1), stepSz), x)), xmul(xmul(xmul(4.0f, c2), x), x));
                const float intval = xadd(integratedDensity, xmul(xmul(temp, 1.0f / 6.0f), xmul(x, poly)));
                fx = xsub(intval, desiredDensity);
                if (fabsf(fx) < 1e-6f) {
                    t = xadd(xadd(mint, xmul(stepSz, (float) i)), x);
                    integratedDensity = intval;
                    densityAtT = xmul(temp, xadd(xsub(xmul(node1, stepSizeSqr), xmul(xmul(c1, stepSz), x)), xmul(xmul(xmul(2.0f, c2), x), x)));
                    return true;
                } else if (++it > 30) return false;
                if (fx > 0) b = x;
                else a = x;
            }
        }
        const F3 next = xadd3(p, fullStep);
        if (p.x == next.x && p.y == next.y && p.z == next.z) break;
        integratedDensity = newDensity;
        node1 = node3;
        p = next;
    }
    return false;
}

/* its.p (barycentric) and the face normal of a triangle hit, ShapeKDTree::rayIntersect + fillIntersectionRecord<true>
 * (skdtree.h:343-428): what Scene::rayIntersect(ray, its) leaves in its.p / its.geoFrame.n = its.shFrame.n (no vertex normals) */
__device__ __forceinline__ void hit_point_normal(const float4 *__restrict__ triVerts, uint32_t prim, float u, float v, F3 &hp, F3 &fn, F3 &dpdu) {
    const F3 p0 = f3(__ldg(&triVerts[3 * (size_t) prim])), p1 = f3(__ldg(&triVerts[3 * (size_t) prim + 1])), p2 = f3(__ldg(&triVerts[3 * (size_t) prim + 2]));
    hp = xadd3(xadd3(xscale(p0, xsub(xsub(1.0f, u), v)), xscale(p1, u)), xscale(p2, v));
    fn = xcross(xsub3(p1, p0), xsub3(p2, p0));
    const float length = xlen(fn);
    if (!(fn.x == 0 && fn.y == 0 && fn.z == 0)) fn = xdivv(fn, length);
    dpdu = xsub3(p1, p0);
}

} // namespace alvrl
