/*
 * primary.cu -- camera-segment generation and the ray-query test entry points, in exact arithmetic.
 *
 * k_primary restates, per pixel centre, PerspectiveCamera::sampleRay (src/sensors/perspective.cpp:247-269),
 * ShapeKDTree::rayIntersect(ray, its) + fillIntersectionRecord<true> (src/librender/skdtree.cpp:112-142,
 * include/mitsuba/render/skdtree.h:343-428) and the eye->surface transmittance integrateVRL precomputes
 * (vrlIntegrator.cpp:711-719).  The same rays serve Preprocessor::buildSlices (Preprocessor.cpp:1140-1179),
 * the rows of R (vrlIntegrator.cpp:324-331) and the render pass (integrator.cpp:232-264 with spp = 1).
 * Pixel index = y + H*x.
 */
#include "dev_common.cuh"
#include "kernels.h"

namespace alvrl {

__device__ __forceinline__ F3 xform_point(const float *m, const F3 &p) {           /* transform.h:108-125 */
    const float x = xadd(xadd(xadd(xmul(m[0], p.x), xmul(m[1], p.y)), xmul(m[2], p.z)), m[3]);
    const float y = xadd(xadd(xadd(xmul(m[4], p.x), xmul(m[5], p.y)), xmul(m[6], p.z)), m[7]);
    const float z = xadd(xadd(xadd(xmul(m[8], p.x), xmul(m[9], p.y)), xmul(m[10], p.z)), m[11]);
    const float w = xadd(xadd(xadd(xmul(m[12], p.x), xmul(m[13], p.y)), xmul(m[14], p.z)), m[15]);
    if (w == 1.0f) return f3(x, y, z);
    return xdivv(f3(x, y, z), w);
}
__device__ __forceinline__ F3 xform_affine(const float *m, const F3 &p) {          /* transform.h:128-136 */
    return f3(xadd(xadd(xadd(xmul(m[0], p.x), xmul(m[1], p.y)), xmul(m[2], p.z)), m[3]),
              xadd(xadd(xadd(xmul(m[4], p.x), xmul(m[5], p.y)), xmul(m[6], p.z)), m[7]),
              xadd(xadd(xadd(xmul(m[8], p.x), xmul(m[9], p.y)), xmul(m[10], p.z)), m[11]));
}
__device__ __forceinline__ F3 xform_vector(const float *m, const F3 &v) {          /* transform.h:175-183 */
    return f3(xadd(xadd(xmul(m[0], v.x), xmul(m[1], v.y)), xmul(m[2], v.z)),
              xadd(xadd(xmul(m[4], v.x), xmul(m[5], v.y)), xmul(m[6], v.z)),
              xadd(xadd(xmul(m[8], v.x), xmul(m[9], v.y)), xmul(m[10], v.z)));
}

__global__ void __launch_bounds__(128) k_primary(SceneDev sc, MediumDev med, CameraDev cam, const float4 *__restrict__ triVerts,
                                                 const uint32_t *__restrict__ triMat, const float4 *__restrict__ matAlbedo,
                                                 const uint32_t *__restrict__ matBits, int haveMedium, SegRec *__restrict__ pixSegs,
                                                 uint32_t *__restrict__ hitPrim, float *__restrict__ hitT) {
    const uint32_t P = cam.W * cam.H;
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const uint32_t x = pix / cam.H, y = pix % cam.H;
    const float px = (float) x + 0.5f, py = (float) y + 0.5f;
    const F3 nearP = xform_point(cam.s2c, f3(xmul(px, cam.invResX), xmul(py, cam.invResY), 0.0f));
    const F3 dl = xnormalize(nearP);
    const float invZ = xdiv(1.0f, dl.z);
    const float mint = xmul(cam.nearClip, invZ), maxt = xmul(cam.farClip, invZ);
    const F3 o = xform_affine(cam.c2w, f3(0.0f, 0.0f, 0.0f));
    const F3 d = xform_vector(cam.c2w, dl);

    float t, u, v; uint32_t prim;
    const bool hit = scene_intersect<false>(sc, o, d, mint, maxt, true, t, prim, u, v);
    SegRec s;
    const F3 dn = xnormalize(d);
    s.o = make_float4(o.x, o.y, o.z, 0.0f);
    s.d = make_float4(d.x, d.y, d.z, 0.0f);
    s.dn = make_float4(dn.x, dn.y, dn.z, __uint_as_float(0u));
    s.p = make_float4(NAN, NAN, NAN, 1.0f);
    s.n = make_float4(NAN, NAN, NAN, 1.0f);
    s.albedo = make_float4(0, 0, 0, 1.0f);
    s.tE = make_float4(0, 0, 0, 0);
    if (hit) {
        const F3 p0 = f3(__ldg(&triVerts[3 * (size_t) prim])), p1 = f3(__ldg(&triVerts[3 * (size_t) prim + 1])),
                 p2 = f3(__ldg(&triVerts[3 * (size_t) prim + 2]));
        const float b0 = xsub(xsub(1.0f, u), v);
        const F3 p = xadd3(xadd3(xscale(p0, b0), xscale(p1, u)), xscale(p2, v));      /* skdtree.h:362-363 */
        F3 fn = xcross(xsub3(p1, p0), xsub3(p2, p0));
        const float length = xlen(fn);
        if (!(fn.x == 0 && fn.y == 0 && fn.z == 0)) fn = xdivv(fn, length);
        const float wiz = xdot(f3(-d.x, -d.y, -d.z), fn);
        const float dist = xlen(xsub3(p, o));
        const uint32_t mat = triMat[prim];
        uint32_t flags = SEG_VALID | ((matBits[mat] & ALVRL_BSDF_SMOOTH) ? SEG_SMOOTH : 0u) | ((matBits[mat] & ALVRL_BSDF_DELTA) ? SEG_DELTA : 0u);
        s.o.w = dist; s.d.w = wiz;
        s.dn.w = __uint_as_float(flags);
        s.p = make_float4(p.x, p.y, p.z, 1.0f);
        s.n = make_float4(fn.x, fn.y, fn.z, 1.0f);
        s.albedo = matAlbedo[mat]; s.albedo.w = 1.0f;
        if (haveMedium && dist != 0) {                                              /* vrlIntegrator.cpp:711-719 */
            float T[3];
            medium_transmittance_exact(med, o, d, dist, T);
            s.tE = make_float4(T[0], T[1], T[2], 0.0f);
        }
    }
    pixSegs[pix] = s;
    hitPrim[pix] = hit ? prim : ALVRL_NO_HIT;
    hitT[pix] = hit ? t : INFINITY;
}

__global__ void k_gather_rows(const SegRec *__restrict__ pixSegs, const uint32_t *__restrict__ rowPixel, uint32_t numRows, SegRec *__restrict__ rowSegs) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < numRows) rowSegs[r] = pixSegs[rowPixel[r]];
}

/* the slice builder's view of a camera segment: gather point and scaled normal (Preprocessor.cpp:1137-1177), 24 bytes per
 * pixel instead of the whole segment record */
__global__ void k_gather_points(const SegRec *__restrict__ pixSegs, uint32_t P, float directionScale, float *__restrict__ pos, float *__restrict__ dir) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 p = pixSegs[i].p, n = pixSegs[i].n;
    pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
    dir[3 * i] = __fmul_rn(directionScale, n.x); dir[3 * i + 1] = __fmul_rn(directionScale, n.y); dir[3 * i + 2] = __fmul_rn(directionScale, n.z);
}

__global__ void k_trace_rays(SceneDev sc, const float *__restrict__ o, const float *__restrict__ d, const float *__restrict__ mint,
                             const float *__restrict__ maxt, uint32_t n, uint32_t *__restrict__ prim, float *__restrict__ tOut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t, u, v; uint32_t p;
    scene_intersect<false>(sc, f3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), f3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i], false, t, p, u, v);
    prim[i] = p; tOut[i] = t;
}

__global__ void k_eval_transmittance(SceneDev sc, MediumDev med, const float *__restrict__ p1, const int32_t *__restrict__ onSurf,
                                     const float *__restrict__ p2, uint32_t n, float *__restrict__ T) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F3 a = f3(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]), b = f3(p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]);
    F3 dir; float remaining;
    float out[3] = {0, 0, 0};
    if (!segment_occluded(sc, a, onSurf && onSurf[i], b, dir, remaining)) {
        if (!(remaining > 0)) out[0] = out[1] = out[2] = 1.0f;
        else if (med.type == 0) {
            for (int c = 0; c < 3; c++) out[c] = med.sigmaT[c] != 0 ? exp_ref(xmul(med.sigmaT[c], xsub(0.0f, remaining))) : 1.0f;
        } else out[0] = out[1] = out[2] = exp_ref(-grid_optical_depth(med, a, dir, 0.0f, remaining));
    }
    T[3 * i] = out[0]; T[3 * i + 1] = out[1]; T[3 * i + 2] = out[2];
}

__global__ void k_fb_to_rgb(const float4 *__restrict__ fb, float *__restrict__ rgb, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = fb[i];
    rgb[3 * i] = v.x; rgb[3 * i + 1] = v.y; rgb[3 * i + 2] = v.z;
}

void launch_primary(const SceneDev &sc, const MediumDev &med, const CameraDev &cam, const float4 *triVerts, const uint32_t *triMat,
                    const float4 *matAlbedo, const uint32_t *matBits, bool haveMedium, SegRec *pixSegs, uint32_t *hitPrim, float *hitT,
                    cudaStream_t st) {
    const uint32_t P = cam.W * cam.H;
    k_primary<<<(P + 127) / 128, 128, 0, st>>>(sc, med, cam, triVerts, triMat, matAlbedo, matBits, haveMedium ? 1 : 0, pixSegs, hitPrim, hitT);
}
void launch_gather_rows(const SegRec *pixSegs, const uint32_t *rowPixel, uint32_t numRows, SegRec *rowSegs, cudaStream_t st) {
    if (numRows) k_gather_rows<<<(numRows + 127) / 128, 128, 0, st>>>(pixSegs, rowPixel, numRows, rowSegs);
}
void launch_gather_points(const SegRec *pixSegs, uint32_t P, float directionScale, float *pos, float *dir, cudaStream_t st) {
    if (P) k_gather_points<<<(P + 255) / 256, 256, 0, st>>>(pixSegs, P, directionScale, pos, dir);
}
void launch_trace_rays(const SceneDev &sc, const float *o, const float *d, const float *mint, const float *maxt, uint32_t n,
                       uint32_t *prim, float *t, cudaStream_t st) {
    if (n) k_trace_rays<<<(n + 127) / 128, 128, 0, st>>>(sc, o, d, mint, maxt, n, prim, t);
}
void launch_eval_transmittance(const SceneDev &sc, const MediumDev &med, const float *p1, const int32_t *onSurf, const float *p2,
                               uint32_t n, float *T, cudaStream_t st) {
    if (n) k_eval_transmittance<<<(n + 127) / 128, 128, 0, st>>>(sc, med, p1, onSurf, p2, n, T);
}
void launch_fb_to_rgb(const float4 *fb, float *rgb, uint32_t n, cudaStream_t st) {
    if (n) k_fb_to_rgb<<<(n + 127) / 128, 128, 0, st>>>(fb, rgb, n);
}

} // namespace alvrl
