/*
 * volpath.cu -- the ground truth of the path: VolumetricPathTracer::Li / Li_original with `onlyVRLpaths`
 * (src/integrators/path/volpath.cpp:76-460) and rayIntersectAndLookForEmitter (484-535) on the device, in exact arithmetic.
 *
 * The reference authors validate the VRL integrator against this estimator: an unbiased path tracer restricted to the light
 * paths VRLs represent (first vertex in the volume or on a diffuse surface inside the medium, second vertex in the volume,
 * initial specular vertices ignored).  Driven like SamplingIntegrator::renderBlock (src/librender/integrator.cpp:210-268):
 * outer sample j of a pixel (the pixel centre when spp == 1, jittered otherwise) is the mean of `internalSamples` walks
 * (volpath.cpp:111-120).  A walk alternates HomogeneousMedium::sampleDistance (homogeneous.cpp:275-352), emitter sampling
 * (Scene::sampleAttenuatedEmitterDirect, scene.cpp:854-899; AreaLight::sampleDirect, area.cpp:158-174; Shape::sampleDirect,
 * shape.cpp:102-115; TriMesh::samplePosition, trimesh.cpp:412-423) combined with phase-function / BSDF sampling by the power
 * heuristic (volpath.cpp:537-540), and Russian roulette from rrDepth on.
 *
 * thread = pixel, one launch per outer sample: launch j adds sample j of every pixel to a {rgb, weight} accumulator, so the
 * sum runs in the reference's order and the developed image equals the oracle's (oracle_capi.cpp::orc_volpath_render) bit for
 * bit.  Outer sample j of pixel p draws from the counter stream of (ALVRL_RNG_VOLPATH, p, j).  Scope: one homogeneous medium
 * the sensor sits in, one area emitter, diffuse / smooth dielectric / smooth conductor surfaces, no ENull surfaces (so the
 * emitter search after a sampled direction is one intersection), no environment emitter.
 */
#include "walk_common.cuh"
#include "kernels.h"

namespace alvrl {

struct VolpathParams {
    uint64_t seed; uint32_t sample; int internalSamples, maxDepth, rrDepth;
    int only, volToVol, volToSurf, singleScatter, strictNormals, hideEmitters, centre;
    float emRadiance[3], emInvArea;
};
struct VpScene {
    SceneDev sc; MediumDev med; CameraDev cam; EmitterDev em;
    const float4 *triVerts; const uint32_t *triMat; const float4 *matAlbedo; const uint32_t *matBits; const float4 *matOptics; const uint8_t *triEmitter;
};
struct DirectRec { F3 ref, refN, p, n, d; float dist, pdf; };
struct VpHit { bool valid; float t; uint32_t prim; F3 p, n, dpdu; };

__device__ __forceinline__ F3 vp_xform_point(const float *m, const F3 &p) {          /* transform.h:108-125 */
    const float x = xadd(xadd(xadd(xmul(m[0], p.x), xmul(m[1], p.y)), xmul(m[2], p.z)), m[3]);
    const float y = xadd(xadd(xadd(xmul(m[4], p.x), xmul(m[5], p.y)), xmul(m[6], p.z)), m[7]);
    const float z = xadd(xadd(xadd(xmul(m[8], p.x), xmul(m[9], p.y)), xmul(m[10], p.z)), m[11]);
    const float w = xadd(xadd(xadd(xmul(m[12], p.x), xmul(m[13], p.y)), xmul(m[14], p.z)), m[15]);
    if (w == 1.0f) return f3(x, y, z);
    return xdivv(f3(x, y, z), w);
}
__device__ __forceinline__ float mi_weight(float pdfA, float pdfB) { pdfA = xmul(pdfA, pdfA); pdfB = xmul(pdfB, pdfB); return xdiv(pdfA, xadd(pdfA, pdfB)); }
__device__ __forceinline__ bool is_zero3(const float v[3]) { return v[0] == 0 && v[1] == 0 && v[2] == 0; }

/* Scene::rayIntersect(ray, its): closest hit, its.p, its.shFrame.n */
__device__ __forceinline__ void vp_intersect(const VpScene &S, const F3 &o, const F3 &d, float mint, float maxt, VpHit &h) {
    float u, v;
    h.valid = scene_intersect<false>(S.sc, o, d, mint, maxt, true, h.t, h.prim, u, v);
    if (h.valid) hit_point_normal(S.triVerts, h.prim, u, v, h.p, h.n, h.dpdu);
    else h.t = INFINITY;
}

/* PhaseFunction::eval: isotropic.cpp:76-78, hg.cpp:107-110 */
__device__ __forceinline__ float vp_phase_eval(const MediumDev &m, const F3 &wi, const F3 &wo) {
    if (m.phaseType == ALVRL_PHASE_ISOTROPIC) return ALVRL_INV_FOURPI;
    const float g = m.g;
    const float temp = xadd(xadd(1.0f, xmul(g, g)), xmul(xmul(2.0f, g), xdot(wi, wo)));
    return xdiv(xmul(ALVRL_INV_FOURPI, xsub(1.0f, xmul(g, g))), xmul(temp, xsqrt(temp)));
}

/* emitter->sampleDirect(dRec, sample): returns radiance / pdf in value, fills dRec (pdf == 0: rejected) */
__device__ __forceinline__ void vp_emitter_sample_direct(const VpScene &S, const VolpathParams &vp, DirectRec &dRec, float sx, float sy, float value[3]) {
    const EmitterDev &em = S.em;
    uint32_t index;
    {   /* DiscreteDistribution::sampleReuse over the triangles' areas (pmf.h:123-135, 164-169) */
        uint32_t lo = 0, hi = em.n + 1;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (em.cdf[mid] < sy) lo = mid + 1; else hi = mid; }
        const int e1 = (int) lo - 1;
        index = (uint32_t) min((int) em.n - 1, max(0, e1));
        while (index < em.n && xsub(em.cdf[index + 1], em.cdf[index]) == 0) ++index;
        if (index >= em.n) index = em.n - 1;
        sy = xdiv(xsub(sy, em.cdf[index]), xsub(em.cdf[index + 1], em.cdf[index]));
    }
    const uint32_t tri = em.tris[index];
    const F3 q0 = f3(__ldg(&S.triVerts[3 * (size_t) tri])), q1 = f3(__ldg(&S.triVerts[3 * (size_t) tri + 1])), q2 = f3(__ldg(&S.triVerts[3 * (size_t) tri + 2]));
    const float a = safe_sqrt_t(xsub(1.0f, sx));                            /* squareToUniformTriangle, warp.cpp:76-79 */
    const float bx = xsub(1.0f, a), by = xmul(a, sy);
    const F3 sideA = xsub3(q1, q0), sideB = xsub3(q2, q0);
    dRec.p = xadd3(xadd3(q0, xscale(sideA, bx)), xscale(sideB, by));
    dRec.n = xnormalize(xcross(sideA, sideB));
    dRec.pdf = vp.emInvArea;
    dRec.d = xsub3(dRec.p, dRec.ref);                                       /* Shape::sampleDirect, shape.cpp:107-114 */
    const float distSquared = xdot(dRec.d, dRec.d);
    dRec.dist = xsqrt(distSquared);
    dRec.d = xdivv(dRec.d, dRec.dist);
    const float dp = fabsf(xdot(dRec.d, dRec.n));
    dRec.pdf = xmul(dRec.pdf, dp != 0 ? xdiv(distSquared, dp) : 0.0f);
    if (xdot(dRec.d, dRec.refN) >= 0 && xdot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) {            /* area.cpp:168-174 */
        const float r = xdiv(1.0f, dRec.pdf);
        for (int i = 0; i < 3; i++) value[i] = xmul(vp.emRadiance[i], r);
    } else { dRec.pdf = 0.0f; value[0] = value[1] = value[2] = 0.0f; }
}
/* Scene::evalTransmittance(ref, refOnSurface, p, true, ...), scene.cpp:619-679 without ENull surfaces, in the medium or in vacuum */
__device__ __forceinline__ void vp_shadow_transmittance(const VpScene &S, const F3 &ref, bool refOnSurface, const F3 &p, bool inMedium, float T[3]) {
    const F3 dd = xsub3(p, ref);
    const float remaining = xlen(dd);
    const F3 dir = xdivv(dd, remaining);
    T[0] = T[1] = T[2] = 1.0f;
    if (!(remaining > 0)) return;
    float t, u, v; uint32_t prim;
    if (scene_intersect<false>(S.sc, ref, dir, refOnSurface ? ALVRL_EPSILON : 0.0f, xmul(remaining, xsub(1.0f, ALVRL_SHADOW_EPSILON)), false, t, prim, u, v)) { T[0] = T[1] = T[2] = 0.0f; return; }
    if (!inMedium) return;
    if (S.med.type == 1) { T[0] = T[1] = T[2] = exp_ref(-grid_optical_depth(S.med, ref, dir, 0.0f, remaining)); return; }                          /* heterogeneous.cpp:546-548 */
    for (int c = 0; c < 3; c++) T[c] = S.med.sigmaT[c] != 0 ? exp_ref(xmul(S.med.sigmaT[c], xsub(0.0f, remaining))) : 1.0f;      /* homogeneous.cpp:266-273 */
}
/* Scene::pdfEmitterDirect, scene.cpp:949-952; AreaLight::pdfDirect, area.cpp:176-183; Shape::pdfDirect, shape.cpp:117-121 */
__device__ __forceinline__ float vp_pdf_emitter_direct(const VolpathParams &vp, const DirectRec &dRec) {
    if (xdot(dRec.d, dRec.refN) >= 0 && xdot(dRec.d, dRec.n) < 0) return xdiv(xmul(vp.emInvArea, xmul(dRec.dist, dRec.dist)), fabsf(xdot(dRec.d, dRec.n)));
    return 0.0f;
}

/* Li_original, volpath.cpp:121-457 */
__device__ void volpath_li_original(const VpScene &S, const VolpathParams &vp, TracerStream &smp, const F3 &ro0, const F3 &rd0, float mint0, float maxt0, float Li[3]) {
    const MediumDev &med = S.med;
    F3 ro = ro0, rd = rd0;
    Li[0] = Li[1] = Li[2] = 0.0f;
    float eta = 1.0f;
    bool vrlFirstVertexOK = false, vrlSecondVertexOK = false, prevWasDiffuseSurface = false, prevWasVolume = false;
    bool inMedium = true;                                                   /* sensor->getMedium() */
    int depth = 1;
    bool emitted = true, indirectMedium = true;                              /* rRec.type: ERadiance, ERadianceNoEmission after the first scattering */
    VpHit its;
    vp_intersect(S, ro, rd, mint0, maxt0, its);
    float thr[3] = {1.0f, 1.0f, 1.0f};
    bool scattered = false;
    while (depth <= vp.maxDepth || vp.maxDepth < 0) {
        if (vp.only && depth > 2 && !(vrlFirstVertexOK && vrlSecondVertexOK)) break;
        /* (!rRec.depth == 2 || ...) of the reference is ((!depth) == 2 || ...): its first operand is never true */
#define VP_DIRECT_OK() (!vp.only || (depth != 1 && ((prevWasVolume || prevWasDiffuseSurface) && (!prevWasDiffuseSurface || vp.volToSurf) && (!prevWasVolume || vp.volToVol))))
        bool success = false;
        float mT[3] = {1.0f, 1.0f, 1.0f}, pdfFailure = 1.0f, pdfSuccess = 1.0f; F3 mP = ro;
        float mSigmaS[3] = {med.sigmaS[0], med.sigmaS[1], med.sigmaS[2]};
        if (inMedium && med.type == 1) {                                     /* sampleDistance(Ray(ray, 0, its.t)), heterogeneous.cpp:589-616 (simpson) */
            const float desiredDensity = -((float) log((double) xsub(1.0f, smp.next())));
            float integratedDensity, tt, densityAtT;
            if (grid_invert_density_integral(med, ro, rd, 0.0f, its.t, desiredDensity, integratedDensity, tt, densityAtT)) {
                mP = xadd3(ro, xscale(rd, tt));
                success = true;
                for (int i = 0; i < 3; ++i) mSigmaS[i] = xmul(med.albedo[i], densityAtT);
            }
            const float expVal = exp_ref(-integratedDensity);
            pdfFailure = expVal; pdfSuccess = xmul(expVal, densityAtT); mT[0] = mT[1] = mT[2] = expVal;
            success = success && pdfSuccess > 0;
        } else if (inMedium) {                                               /* sampleDistance(Ray(ray, 0, its.t)), homogeneous.cpp:275-352 */
            float rnd = smp.next(), sampledDistance;
            if (rnd < med.samplingWeight) {
                rnd = xdiv(rnd, med.samplingWeight);
                const int channel = min((int) xmul(smp.next(), 3.0f), 2);
                const float samplingDensity = med.sigmaT[channel];
                sampledDistance = xdiv(-((float) log((double) xsub(1.0f, rnd))), samplingDensity);
            } else sampledDistance = INFINITY;
            const float distSurf = xsub(its.t, 0.0f);
            success = true;
            if (sampledDistance < distSurf) {
                const float tt = xadd(sampledDistance, 0.0f);
                mP = xadd3(ro, xscale(rd, tt));
                if (mP.x == ro.x && mP.y == ro.y && mP.z == ro.z) success = false;
            } else { sampledDistance = distSurf; success = false; }
            pdfFailure = 0; pdfSuccess = 0;
            for (int i = 0; i < 3; ++i) { const float tmp = exp_ref(xmul(-med.sigmaT[i], sampledDistance)); pdfFailure = xadd(pdfFailure, tmp); pdfSuccess = xadd(pdfSuccess, xmul(med.sigmaT[i], tmp)); }
            pdfFailure = xdiv(pdfFailure, 3.0f); pdfSuccess = xdiv(pdfSuccess, 3.0f);
            float mx = 0;
            for (int i = 0; i < 3; ++i) { mT[i] = exp_ref(xmul(med.sigmaT[i], -sampledDistance)); mx = fmaxf(mx, mT[i]); }
            pdfSuccess = xmul(pdfSuccess, med.samplingWeight);
            pdfFailure = xadd(xmul(med.samplingWeight, pdfFailure), xsub(1.0f, med.samplingWeight));
            if ((double) mx < 1e-20) mT[0] = mT[1] = mT[2] = 0.0f;
        }
        if (inMedium && success) {
            if (vp.singleScatter) indirectMedium = false;
            if (depth == 1) { if (vp.volToVol) vrlFirstVertexOK = true; }
            if (depth == 2) vrlSecondVertexOK = true;
            if (depth >= vp.maxDepth && vp.maxDepth != -1) break;
            {
                const float rps = xdiv(1.0f, pdfSuccess);
                for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], xmul(xmul(mSigmaS[i], mT[i]), rps));
            }
            DirectRec dRec; dRec.ref = mP; dRec.refN = f3(0.0f, 0.0f, 0.0f); dRec.dist = 0; dRec.pdf = 0;
            const F3 wi = f3(-rd.x, -rd.y, -rd.z);
            if (VP_DIRECT_OK()) {
                const float sx = smp.next(), sy = smp.next();
                float value[3];
                vp_emitter_sample_direct(S, vp, dRec, sx, sy, value);
                if (dRec.pdf != 0) { float T[3]; vp_shadow_transmittance(S, dRec.ref, false, dRec.p, inMedium, T); for (int i = 0; i < 3; i++) value[i] = xmul(value[i], T[i]); }
                if (!is_zero3(value)) {
                    const float phaseVal = vp_phase_eval(med, wi, dRec.d);
                    if (phaseVal != 0) {
                        const float weight = mi_weight(dRec.pdf, phaseVal);  /* PhaseFunction::pdf = eval */
                        for (int i = 0; i < 3; i++) Li[i] = xadd(Li[i], xmul(xmul(xmul(thr[i], value[i]), phaseVal), weight));
                    }
                }
            }
            /* phase function sampling: sample(pRec, pdf, sampler) returns 1 */
            float phasePdf;
            const float px = smp.next(), py = smp.next();
            F3 wo;
            if (med.phaseType == ALVRL_PHASE_ISOTROPIC) { wo = square_to_uniform_sphere(px, py); phasePdf = ALVRL_INV_FOURPI; }
            else {                                                            /* hg.cpp:74-104 */
                float cosTheta;
                const float g = med.g;
                if (fabsf(g) < ALVRL_EPSILON) cosTheta = xsub(1.0f, xmul(2.0f, px));
                else {
                    const float sqrTerm = xdiv(xsub(1.0f, xmul(g, g)), xadd(xsub(1.0f, g), xmul(xmul(2.0f, g), px)));
                    cosTheta = xdiv(xsub(xadd(1.0f, xmul(g, g)), xmul(sqrTerm, sqrTerm)), xmul(2.0f, g));
                }
                const float sinTheta = safe_sqrt_t(xsub(1.0f, xmul(cosTheta, cosTheta)));
                float sinPhi, cosPhi;
                sincos_t((float) (2.0 * ALVRL_PI_D * (double) py), sinPhi, cosPhi);
                wo = frame_to_world(rd, f3(xmul(sinTheta, cosPhi), xmul(sinTheta, sinPhi), cosTheta));
                phasePdf = vp_phase_eval(med, wi, wo);
            }
            ro = mP; rd = wo;
            float value[3] = {0, 0, 0};
            vp_intersect(S, ro, rd, 0.0f, INFINITY, its);                     /* rayIntersectAndLookForEmitter, no ENull surfaces */
            if (its.valid && S.triEmitter[its.prim]) {
                dRec.p = its.p; dRec.n = its.n; dRec.d = rd; dRec.dist = its.t;                   /* setQuery, records.inl:170-178 */
                if (!(xdot(its.n, f3(-rd.x, -rd.y, -rd.z)) <= 0)) { value[0] = vp.emRadiance[0]; value[1] = vp.emRadiance[1]; value[2] = vp.emRadiance[2]; }
            }
            if (!is_zero3(value) && VP_DIRECT_OK()) {
                const float emitterPdf = vp_pdf_emitter_direct(vp, dRec);
                const float w = mi_weight(phasePdf, emitterPdf);
                for (int i = 0; i < 3; i++) Li[i] = xadd(Li[i], xmul(xmul(thr[i], value[i]), w));
            }
            if (!indirectMedium) break;
            emitted = false;
            prevWasVolume = true; prevWasDiffuseSurface = false;
        } else {
            if (inMedium) { const float rpf = xdiv(1.0f, pdfFailure); for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], xmul(mT[i], rpf)); }
            if (!its.valid) break;                                            /* no environment emitter */
            const F3 md = f3(-rd.x, -rd.y, -rd.z);
            if (S.triEmitter[its.prim] && emitted && (!vp.hideEmitters || scattered) && (!vp.only || (vrlFirstVertexOK && vrlSecondVertexOK))) {
                if (!(xdot(its.n, md) <= 0)) for (int i = 0; i < 3; i++) Li[i] = xadd(Li[i], xmul(thr[i], vp.emRadiance[i]));
                else for (int i = 0; i < 3; i++) Li[i] = xadd(Li[i], xmul(thr[i], 0.0f));
            }
            if (depth >= vp.maxDepth && vp.maxDepth != -1) break;
            const VpHit here = its;
            const uint32_t mat = S.triMat[here.prim];
            const uint32_t bits = S.matBits[mat];
            const F3 fn = here.n;
            const F3 fs = xnormalize(xsub3(here.dpdu, xscale(fn, xdot(fn, here.dpdu))));
            const F3 ft = xcross(fn, fs);
            const F3 wi = f3(xdot(md, fs), xdot(md, ft), xdot(md, fn));
            const float wiDotGeoN = -xdot(fn, rd), wiDotShN = wi.z;
            if (xmul(wiDotGeoN, wiDotShN) < 0 && vp.strictNormals) break;
            const bool smooth = !(bits & ALVRL_BSDF_DELTA) && (bits & ALVRL_BSDF_SMOOTH);
            DirectRec dRec; dRec.ref = here.p; dRec.refN = (bits & ALVRL_BSDF_DIELECTRIC) ? f3(0.0f, 0.0f, 0.0f) : fn; dRec.dist = 0; dRec.pdf = 0;      /* records.inl:160-164 */
            dRec.p = dRec.n = dRec.d = f3(0.0f, 0.0f, 0.0f);
            const float4 al = S.matAlbedo[mat];
            const float albedo[3] = {al.x, al.y, al.z};
            if (smooth && (!vp.only || (vrlFirstVertexOK && vrlSecondVertexOK))) {
                const float sx = smp.next(), sy = smp.next();
                float value[3];
                vp_emitter_sample_direct(S, vp, dRec, sx, sy, value);
                if (dRec.pdf != 0) {
                    bool shadowMedium = inMedium;                            /* its.getTargetMedium(dRec.d), scene.cpp:888-889 */
                    if (bits & ALVRL_MAT_TRANSITION) shadowMedium = xdot(dRec.d, fn) > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
                    float T[3];
                    vp_shadow_transmittance(S, here.p, true, dRec.p, shadowMedium, T);
                    for (int i = 0; i < 3; i++) value[i] = xmul(value[i], T[i]);
                }
                if (!is_zero3(value)) {
                    const F3 woL = f3(xdot(dRec.d, fs), xdot(dRec.d, ft), xdot(dRec.d, fn));
                    float bsdfVal[3] = {0, 0, 0};                            /* diffuse.cpp:110-118 */
                    if (!(wi.z <= 0 || woL.z <= 0)) { const float f = xmul(ALVRL_INV_PI, woL.z); for (int i = 0; i < 3; i++) bsdfVal[i] = xmul(albedo[i], f); }
                    const float woDotGeoN = xdot(fn, dRec.d);
                    if (!is_zero3(bsdfVal) && (!vp.strictNormals || xmul(woDotGeoN, woL.z) > 0)) {
                        const float bsdfPdf = (wi.z <= 0 || woL.z <= 0) ? 0.0f : xmul(ALVRL_INV_PI, woL.z);
                        const float weight = mi_weight(dRec.pdf, bsdfPdf);
                        for (int i = 0; i < 3; i++) Li[i] = xadd(Li[i], xmul(xmul(xmul(thr[i], value[i]), bsdfVal[i]), weight));
                    }
                }
            }
            /* BSDF sampling: sample(bRec, pdf, nextSample2D()), mode = ERadiance */
            const float bsx = smp.next(), bsy = smp.next();
            F3 woL = f3(0, 0, 0); float bEta = 1.0f, bsdfPdf = 0.0f; float bw[3] = {0, 0, 0}; bool delta = false;
            if (bits & ALVRL_BSDF_DIELECTRIC) {                              /* dielectric.cpp:281-332 */
                const float4 o0 = __ldg(&S.matOptics[3 * mat]), o1 = __ldg(&S.matOptics[3 * mat + 1]), o2 = __ldg(&S.matOptics[3 * mat + 2]);
                const float e = o0.x, invE = xdiv(1.0f, e);
                float cosThetaT;
                const float F = fresnel_dielectric_ext_t(wi.z, cosThetaT, e);
                delta = true;
                if (bsx <= F) { woL = f3(-wi.x, -wi.y, wi.z); bEta = 1.0f; bsdfPdf = F; bw[0] = o1.z; bw[1] = o1.w; bw[2] = o2.x; }
                else {
                    const float scale = -(cosThetaT < 0 ? invE : e);
                    woL = f3(xmul(scale, wi.x), xmul(scale, wi.y), cosThetaT);
                    bEta = cosThetaT < 0 ? e : invE;
                    bsdfPdf = xsub(1.0f, F);
                    const float factor = cosThetaT < 0 ? invE : e;
                    const float f2 = xmul(factor, factor);
                    bw[0] = xmul(o2.y, f2); bw[1] = xmul(o2.z, f2); bw[2] = xmul(o2.w, f2);
                }
            } else if (bits & ALVRL_BSDF_CONDUCTOR) {                        /* conductor.cpp:268-283 */
                if (wi.z > 0) {
                    const float4 o0 = __ldg(&S.matOptics[3 * mat]), o1 = __ldg(&S.matOptics[3 * mat + 1]), o2 = __ldg(&S.matOptics[3 * mat + 2]);
                    delta = true;
                    woL = f3(-wi.x, -wi.y, wi.z); bsdfPdf = 1.0f;
                    bw[0] = xmul(o1.z, fresnel_conductor_exact_t(wi.z, o0.x, o0.w));
                    bw[1] = xmul(o1.w, fresnel_conductor_exact_t(wi.z, o0.y, o1.x));
                    bw[2] = xmul(o2.x, fresnel_conductor_exact_t(wi.z, o0.z, o1.y));
                }
            } else if ((bits & ALVRL_BSDF_SMOOTH) && wi.z > 0) {             /* diffuse.cpp:139-148 */
                woL = square_to_cosine_hemisphere(bsx, bsy);
                bsdfPdf = xmul(ALVRL_INV_PI, woL.z);
                bw[0] = albedo[0]; bw[1] = albedo[1]; bw[2] = albedo[2];
            }
            if (is_zero3(bw)) break;
            const F3 woW = xadd3(xadd3(xscale(fs, woL.x), xscale(ft, woL.y)), xscale(fn, woL.z));
            const float woDotGeoN = xdot(fn, woW);
            if (xmul(woDotGeoN, woL.z) <= 0 && vp.strictNormals) break;
            if (depth == 1 && delta) depth--;                                 /* 'undo' initial specular vertices */
            if (vp.volToSurf) { if (depth == 1 && inMedium && !delta) vrlFirstVertexOK = true; }
            prevWasVolume = false;
            prevWasDiffuseSurface = !delta;
            ro = here.p; rd = woW;
            for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], bw[i]);
            eta = xmul(eta, bEta);
            if (bits & ALVRL_MAT_TRANSITION) inMedium = xdot(fn, rd) > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
            float value[3] = {0, 0, 0};
            vp_intersect(S, ro, rd, ALVRL_EPSILON, INFINITY, its);
            if (its.valid && S.triEmitter[its.prim]) {
                dRec.p = its.p; dRec.n = its.n; dRec.d = rd; dRec.dist = its.t;
                if (!(xdot(its.n, f3(-rd.x, -rd.y, -rd.z)) <= 0)) { value[0] = vp.emRadiance[0]; value[1] = vp.emRadiance[1]; value[2] = vp.emRadiance[2]; }
            }
            if (!is_zero3(value) && (!vp.only || (vrlFirstVertexOK && vrlSecondVertexOK))) {
                const float emitterPdf = !delta ? vp_pdf_emitter_direct(vp, dRec) : 0.0f;
                const float w = mi_weight(bsdfPdf, emitterPdf);
                for (int i = 0; i < 3; i++) Li[i] = xadd(Li[i], xmul(xmul(thr[i], value[i]), w));
            }
            emitted = false;
        }
        if (depth++ >= vp.rrDepth) {
            const float q = fminf(xmul(xmul(fmaxf(fmaxf(thr[0], thr[1]), thr[2]), eta), eta), 0.95f);
            if (smp.next() >= q) break;
            const float rq = xdiv(1.0f, q);
            for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], rq);
        }
        scattered = true;
#undef VP_DIRECT_OK
    }
    if (vp.only && !(vrlFirstVertexOK && vrlSecondVertexOK)) for (int i = 0; i < 3; i++) Li[i] = xmul(Li[i], 0.0f);
}

/* one outer sample of every pixel: renderBlock's body (integrator.cpp:236-262) + VolumetricPathTracer::Li (volpath.cpp:111-120)
 * + the film's accumulation; acc = {sum rgb, sum of weights} per pixel index y + H * x */
__global__ void __launch_bounds__(64) k_volpath_sample(VpScene S, VolpathParams vp, float4 *__restrict__ acc) {
    const uint32_t P = S.cam.W * S.cam.H;
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const uint32_t x = pix / S.cam.H, y = pix % S.cam.H;
    TracerStream smp; smp.key = alvrl_rng_key(vp.seed, ALVRL_RNG_VOLPATH, pix, vp.sample); smp.k = 0;
    float ox = 0.5f, oy = 0.5f;
    if (!vp.centre) { ox = smp.next(); oy = smp.next(); }
    const float px = xadd((float) x, ox), py = xadd((float) y, oy);
    const F3 nearP = vp_xform_point(S.cam.s2c, f3(xmul(px, S.cam.invResX), xmul(py, S.cam.invResY), 0.0f));       /* perspective.cpp:247-269 */
    const F3 dl = xnormalize(nearP);
    const float invZ = xdiv(1.0f, dl.z);
    const float mint = xmul(S.cam.nearClip, invZ), maxt = xmul(S.cam.farClip, invZ);
    const float *m = S.cam.c2w;
    const F3 o = f3(xadd(xadd(xadd(xmul(m[0], 0.0f), xmul(m[1], 0.0f)), xmul(m[2], 0.0f)), m[3]),
                    xadd(xadd(xadd(xmul(m[4], 0.0f), xmul(m[5], 0.0f)), xmul(m[6], 0.0f)), m[7]),
                    xadd(xadd(xadd(xmul(m[8], 0.0f), xmul(m[9], 0.0f)), xmul(m[10], 0.0f)), m[11]));
    const F3 d = f3(xadd(xadd(xmul(m[0], dl.x), xmul(m[1], dl.y)), xmul(m[2], dl.z)),
                    xadd(xadd(xmul(m[4], dl.x), xmul(m[5], dl.y)), xmul(m[6], dl.z)),
                    xadd(xadd(xmul(m[8], dl.x), xmul(m[9], dl.y)), xmul(m[10], dl.z)));
    float Li[3] = {0, 0, 0};
    for (int i = 0; i < vp.internalSamples; i++) {
        float one[3];
        volpath_li_original(S, vp, smp, o, d, mint, maxt, one);
        for (int k = 0; k < 3; k++) Li[k] = xadd(Li[k], one[k]);
    }
    const float r = xdiv(1.0f, (float) vp.internalSamples);
    for (int k = 0; k < 3; k++) Li[k] = xmul(Li[k], r);
    bool valid = true;                                                       /* ImageBlock::put rejects NaN / negative samples, imageblock.h:147-151 */
    for (int k = 0; k < 3; k++) if (!isfinite(Li[k]) || Li[k] < 0.0f) valid = false;
    if (!valid) return;
    float4 a = acc[pix];
    a.x = xadd(a.x, Li[0]); a.y = xadd(a.y, Li[1]); a.z = xadd(a.z, Li[2]); a.w = xadd(a.w, 1.0f);
    acc[pix] = a;
}
/* development: value * (1 / weight), bitmap.cpp:1617-1624; rgb in image order [y][x][c] */
__global__ void k_volpath_develop(const float4 *__restrict__ acc, uint32_t W, uint32_t H, float *__restrict__ rgb) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= W * H) return;
    const uint32_t x = pix / H, y = pix % H;
    const float4 a = acc[pix];
    const float invWeight = a.w == 0 ? 0.0f : xdiv(1.0f, a.w);
    float *o = rgb + 3 * ((size_t) y * W + x);
    o[0] = xmul(a.x, invWeight); o[1] = xmul(a.y, invWeight); o[2] = xmul(a.z, invWeight);
}

void launch_volpath_sample(const SceneDev &sc, const MediumDev &med, const CameraDev &cam, const uint32_t *emTris, const float *emCdf, uint32_t emN,
                           const float emRadiance[3], float emInvArea, const uint8_t *triEmitter, const float4 *triVerts, const uint32_t *triMat,
                           const float4 *matAlbedo, const uint32_t *matBits, const float4 *matOptics, uint64_t seed, uint32_t sample, int internalSamples,
                           uint32_t flags, bool centre, int maxDepth, int rrDepth, float4 *acc, cudaStream_t st) {
    VpScene S; S.sc = sc; S.med = med; S.cam = cam;
    S.em.tris = emTris; S.em.cdf = emCdf; S.em.n = emN; S.em.power[0] = S.em.power[1] = S.em.power[2] = 0.0f;
    S.triVerts = triVerts; S.triMat = triMat; S.matAlbedo = matAlbedo; S.matBits = matBits; S.matOptics = matOptics; S.triEmitter = triEmitter;
    VolpathParams vp; vp.seed = seed; vp.sample = sample; vp.internalSamples = internalSamples; vp.maxDepth = maxDepth; vp.rrDepth = rrDepth;
    vp.only = (flags & ALVRL_VOLPATH_ONLY_VRL_PATHS) != 0; vp.volToVol = (flags & ALVRL_VOLPATH_VOL_TO_VOL) != 0; vp.volToSurf = (flags & ALVRL_VOLPATH_VOL_TO_SURF) != 0;
    vp.singleScatter = (flags & ALVRL_VOLPATH_SINGLE_SCATTER) != 0; vp.strictNormals = (flags & ALVRL_VOLPATH_STRICT_NORMALS) != 0; vp.hideEmitters = (flags & ALVRL_VOLPATH_HIDE_EMITTERS) != 0;
    vp.centre = centre ? 1 : 0;
    for (int k = 0; k < 3; k++) vp.emRadiance[k] = emRadiance[k];
    vp.emInvArea = emInvArea;
    const uint32_t P = cam.W * cam.H;
    if (P) k_volpath_sample<<<(P + 63) / 64, 64, 0, st>>>(S, vp, acc);
}
void launch_volpath_develop(const float4 *acc, uint32_t W, uint32_t H, float *rgb, cudaStream_t st) {
    if (W && H) k_volpath_develop<<<(W * H + 127) / 128, 128, 0, st>>>(acc, W, H, rgb);
}

} // namespace alvrl
