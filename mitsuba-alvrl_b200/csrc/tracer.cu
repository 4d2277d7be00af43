/*
 * tracer.cu -- the step before the path: vrlTracer::randomWalk / traceOneParticle (src/integrators/vrl/vrlTracer.h:14-58,
 * 91-230) on the device, in exact arithmetic.
 *
 * A light particle leaves an area emitter (Scene::sampleEmitterPosition, scene.cpp:958-974; TriMesh::samplePosition,
 * trimesh.cpp:412-423; Triangle::sample, triangle.cpp:24-45; AreaLight::sampleDirection, area.cpp:115-123), and alternates
 * medium interactions (HomogeneousMedium::sampleDistance, homogeneous.cpp:275-352, strategy = balance; phase function
 * sampling, isotropic.cpp:62-67 / hg.cpp:74-98) and surface interactions (diffuse.cpp:129-138, dielectric.cpp:335-364,
 * conductor.cpp:254-268; medium transitions, records.inl:88-93) with Russian roulette from rrDepth on; every segment it
 * travels inside the scattering medium is stored as a VRL (endCurrentVrl + vrlVector::put, VRL.h:148-158).
 *
 * The reference traces particles one after the other until vrlTargetNum VRLs exist.  Here particle i draws from the counter
 * stream of (ALVRL_RNG_TRACER, i), so particles are independent: thread = particle, two passes of the same walk (count the
 * VRLs; after the host has found the first particle count that reaches the target and scanned the counts, write them).
 * Same operations in the same order as the oracle (oracle_capi.cpp::traceOneParticle, -fmad=false, sin / cos / log through
 * double like the oracle pins them): the VRL sets are compared bit for bit (tests/test_tracer.py).
 */
#include "walk_common.cuh"
#include "kernels.h"

namespace alvrl {

struct TracerParams { uint64_t seed; int shortVrls, maxDepth, rrDepth; };

/* one particle; WRITE = false counts the VRLs it stores, WRITE = true writes them to out[0 .. count) as {start, end, power} */
template <bool WRITE>
__device__ __forceinline__ uint32_t trace_particle(const SceneDev &sc, const MediumDev &med, const EmitterDev &em, const TracerParams &tp,
                                                   const float4 *__restrict__ triVerts, const uint32_t *__restrict__ triMat, const float4 *__restrict__ matAlbedo,
                                                   const uint32_t *__restrict__ matBits, const float4 *__restrict__ matOptics, uint32_t particle, float *__restrict__ out) {
    TracerStream smp; smp.key = alvrl_rng_key(tp.seed, ALVRL_RNG_TRACER, particle, 0u); smp.k = 0;
    uint32_t count = 0;
    const bool sigmaSZero = med.sigmaS[0] == 0 && med.sigmaS[1] == 0 && med.sigmaS[2] == 0;
    /* emission */
    float sx = smp.next(), sy = smp.next();
    uint32_t index;
    {   /* DiscreteDistribution::sample (std::lower_bound on the cdf, pmf.h:123-135) + sampleReuse (164-169) */
        uint32_t lo = 0, hi = em.n + 1;                                    /* first entry >= sy among cdf[0 .. n] */
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (em.cdf[mid] < sy) lo = mid + 1; else hi = mid; }
        const int e1 = (int) lo - 1;
        index = (uint32_t) min((int) em.n - 1, max(0, e1));
        while (index < em.n && xsub(em.cdf[index + 1], em.cdf[index]) == 0) ++index;
        if (index >= em.n) index = em.n - 1;
        sy = xdiv(xsub(sy, em.cdf[index]), xsub(em.cdf[index + 1], em.cdf[index]));
    }
    const uint32_t tri = em.tris[index];
    const F3 q0 = f3(__ldg(&triVerts[3 * (size_t) tri])), q1 = f3(__ldg(&triVerts[3 * (size_t) tri + 1])), q2 = f3(__ldg(&triVerts[3 * (size_t) tri + 2]));
    const float a = safe_sqrt_t(xsub(1.0f, sx));                            /* squareToUniformTriangle, warp.cpp:76-79 */
    const float bx = xsub(1.0f, a), by = xmul(a, sy);
    const F3 sideA = xsub3(q1, q0), sideB = xsub3(q2, q0);
    const F3 p = xadd3(xadd3(q0, xscale(sideA, bx)), xscale(sideB, by));
    const F3 n = xnormalize(xcross(sideA, sideB));
    const float power[3] = {em.power[0], em.power[1], em.power[2]};
    const float dx = smp.next(), dy = smp.next();
    const F3 d0 = frame_to_world(n, square_to_cosine_hemisphere(dx, dy));
    if (power[0] == 0 && power[1] == 0 && power[2] == 0) return 0;
    bool inMedium = true;
    float curPow[3] = {power[0], power[1], power[2]};
    F3 curStart = p; bool curMedium = true;
    auto endCurrent = [&](const F3 &q) {                                     /* endCurrentVrl + vrlVector::put */
        if (xlen(xsub3(curStart, q)) == 0) return;
        if (!curMedium || sigmaSZero) return;
        if (curPow[0] == 0 && curPow[1] == 0 && curPow[2] == 0) return;
        if (WRITE) { float *o = out + 9 * (size_t) count; o[0] = curStart.x; o[1] = curStart.y; o[2] = curStart.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = curPow[0]; o[7] = curPow[1]; o[8] = curPow[2]; }
        count++;
    };
    F3 ro = p, rd = d0; float rmint = ALVRL_EPSILON;
    int depth = 1;
    float thr[3] = {1.0f, 1.0f, 1.0f};
    float eta = 1.0f;
    while (!(thr[0] == 0 && thr[1] == 0 && thr[2] == 0) && (depth <= tp.maxDepth || tp.maxDepth < 0)) {
        float t, u, v; uint32_t prim;
        const bool hit = scene_intersect<false>(sc, ro, rd, rmint, INFINITY, true, t, prim, u, v);
        const float itsT = hit ? t : INFINITY;
        bool scattered = false;
        float mT[3] = {1.0f, 1.0f, 1.0f}, pdfFailure = 1.0f, pdfSuccess = 1.0f; F3 mP = ro;
        float mSigmaS[3] = {med.sigmaS[0], med.sigmaS[1], med.sigmaS[2]};
        if (inMedium && med.type == 1) {                                      /* sampleDistance, heterogeneous.cpp:589-616 (simpson) */
            const float desiredDensity = -((float) log((double) xsub(1.0f, smp.next())));
            float integratedDensity, tt, densityAtT;
            bool success = false;
            if (grid_invert_density_integral(med, ro, rd, 0.0f, itsT, desiredDensity, integratedDensity, tt, densityAtT)) {
                mP = xadd3(ro, xscale(rd, tt));
                success = true;
                for (int i = 0; i < 3; ++i) mSigmaS[i] = xmul(med.albedo[i], densityAtT);
            }
            const float expVal = exp_ref(-integratedDensity);
            pdfFailure = expVal; pdfSuccess = xmul(expVal, densityAtT); mT[0] = mT[1] = mT[2] = expVal;
            scattered = success && pdfSuccess > 0;
        } else if (inMedium) {                                                /* sampleDistance, homogeneous.cpp:275-352 */
            float rnd = smp.next(), sampledDistance;
            if (rnd < med.samplingWeight) {
                rnd = xdiv(rnd, med.samplingWeight);
                const int channel = min((int) xmul(smp.next(), 3.0f), 2);
                const float samplingDensity = med.sigmaT[channel];
                sampledDistance = xdiv(-((float) log((double) xsub(1.0f, rnd))), samplingDensity);
            } else sampledDistance = INFINITY;
            const float distSurf = xsub(itsT, 0.0f);
            bool success = true;
            if (sampledDistance < distSurf) {
                const float tt = xadd(sampledDistance, 0.0f);
                mP = xadd3(ro, xscale(rd, tt));
                if (mP.x == ro.x && mP.y == ro.y && mP.z == ro.z) success = false;
            } else { sampledDistance = distSurf; success = false; }
            pdfFailure = 0; pdfSuccess = 0;
            for (int i = 0; i < 3; ++i) { const float tmp = exp_ref(xmul(-med.sigmaT[i], sampledDistance)); pdfFailure = xadd(pdfFailure, tmp); pdfSuccess = xadd(pdfSuccess, xmul(med.sigmaT[i], tmp)); }
            pdfFailure = xdiv(pdfFailure, 3.0f); pdfSuccess = xdiv(pdfSuccess, 3.0f);          /* /= SPECTRUM_SAMPLES: a true division */
            float mx = 0;
            for (int i = 0; i < 3; ++i) { mT[i] = exp_ref(xmul(med.sigmaT[i], -sampledDistance)); mx = fmaxf(mx, mT[i]); }
            pdfSuccess = xmul(pdfSuccess, med.samplingWeight);
            pdfFailure = xadd(xmul(med.samplingWeight, pdfFailure), xsub(1.0f, med.samplingWeight));
            if ((double) mx < 1e-20) mT[0] = mT[1] = mT[2] = 0.0f;
            scattered = success;
        }
        if (inMedium && scattered) {
            const float rps = xdiv(1.0f, pdfSuccess);
            for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], xmul(xmul(mT[i], mSigmaS[i]), rps));
            const float px = smp.next(), py = smp.next();
            F3 wo;
            if (med.phaseType == ALVRL_PHASE_ISOTROPIC) wo = square_to_uniform_sphere(px, py);
            else {                                                            /* hg.cpp:74-98 */
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
            }
            F3 endPoint = mP;
            if (!tp.shortVrls) {
                if (!hit) break;
                /* its.p: barycentric */
                const F3 p0 = f3(__ldg(&triVerts[3 * (size_t) prim])), p1 = f3(__ldg(&triVerts[3 * (size_t) prim + 1])), p2 = f3(__ldg(&triVerts[3 * (size_t) prim + 2]));
                endPoint = xadd3(xadd3(xscale(p0, xsub(xsub(1.0f, u), v)), xscale(p1, u)), xscale(p2, v));
            }
            endCurrent(endPoint);                                             /* handleMediumScattering */
            for (int i = 0; i < 3; ++i) curPow[i] = xmul(thr[i], power[i]);
            curStart = mP; curMedium = inMedium;
            ro = mP; rd = wo; rmint = 0.0f;
        } else if (hit) {
            if (inMedium) { const float rpf = xdiv(1.0f, pdfFailure); for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], xmul(mT[i], rpf)); }
            const uint32_t mat = triMat[prim];
            const uint32_t bits = matBits[mat];
            const F3 p0 = f3(__ldg(&triVerts[3 * (size_t) prim])), p1 = f3(__ldg(&triVerts[3 * (size_t) prim + 1])), p2 = f3(__ldg(&triVerts[3 * (size_t) prim + 2]));
            const F3 hp = xadd3(xadd3(xscale(p0, xsub(xsub(1.0f, u), v)), xscale(p1, u)), xscale(p2, v));
            F3 fn = xcross(xsub3(p1, p0), xsub3(p2, p0));
            const float length = xlen(fn);
            if (!(fn.x == 0 && fn.y == 0 && fn.z == 0)) fn = xdivv(fn, length);
            const F3 dpdu = xsub3(p1, p0);
            const F3 fs = xnormalize(xsub3(dpdu, xscale(fn, xdot(fn, dpdu))));
            const F3 ft = xcross(fn, fs);
            const F3 md = f3(-rd.x, -rd.y, -rd.z);
            const F3 wi = f3(xdot(md, fs), xdot(md, ft), xdot(md, fn));
            const float bsx = smp.next(), bsy = smp.next();
            F3 woL = f3(0, 0, 0); float bEta = 1.0f; float bw[3] = {0, 0, 0};
            if (bits & ALVRL_BSDF_DIELECTRIC) {                              /* dielectric.cpp:335-364, mode = EImportance */
                const float4 o0 = __ldg(&matOptics[3 * mat]), o1 = __ldg(&matOptics[3 * mat + 1]), o2 = __ldg(&matOptics[3 * mat + 2]);
                const float e = o0.x, invE = xdiv(1.0f, e);
                float cosThetaT;
                const float F = fresnel_dielectric_ext_t(wi.z, cosThetaT, e);
                if (bsx <= F) { woL = f3(-wi.x, -wi.y, wi.z); bEta = 1.0f; bw[0] = o1.z; bw[1] = o1.w; bw[2] = o2.x; }
                else {
                    const float scale = -(cosThetaT < 0 ? invE : e);
                    woL = f3(xmul(scale, wi.x), xmul(scale, wi.y), cosThetaT);
                    bEta = cosThetaT < 0 ? e : invE;
                    bw[0] = xmul(o2.y, 1.0f); bw[1] = xmul(o2.z, 1.0f); bw[2] = xmul(o2.w, 1.0f);
                }
            } else if (bits & ALVRL_BSDF_CONDUCTOR) {                        /* conductor.cpp:254-268 */
                if (wi.z > 0) {
                    const float4 o0 = __ldg(&matOptics[3 * mat]), o1 = __ldg(&matOptics[3 * mat + 1]), o2 = __ldg(&matOptics[3 * mat + 2]);
                    woL = f3(-wi.x, -wi.y, wi.z);
                    bw[0] = xmul(o1.z, fresnel_conductor_exact_t(wi.z, o0.x, o0.w));
                    bw[1] = xmul(o1.w, fresnel_conductor_exact_t(wi.z, o0.y, o1.x));
                    bw[2] = xmul(o2.x, fresnel_conductor_exact_t(wi.z, o0.z, o1.y));
                }
            } else if ((bits & ALVRL_BSDF_SMOOTH) && wi.z > 0) {             /* diffuse.cpp:129-138 */
                woL = square_to_cosine_hemisphere(bsx, bsy);
                const float4 al = matAlbedo[mat];
                bw[0] = al.x; bw[1] = al.y; bw[2] = al.z;
            }
            if (bw[0] == 0 && bw[1] == 0 && bw[2] == 0) { endCurrent(hp); break; }
            const F3 woW = xadd3(xadd3(xscale(fs, woL.x), xscale(ft, woL.y)), xscale(fn, woL.z));
            const float wiDotGeoN = xdot(fn, md), woDotGeoN = xdot(fn, woW);
            if (xmul(wiDotGeoN, wi.z) <= 0 || xmul(woDotGeoN, woL.z) <= 0) { endCurrent(hp); break; }
            for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], bw[i]);
            eta = xmul(eta, bEta);
            if (bits & ALVRL_MAT_TRANSITION) inMedium = woDotGeoN > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
            endCurrent(hp);                                                   /* handleSurfaceScattering */
            for (int i = 0; i < 3; ++i) curPow[i] = xmul(thr[i], power[i]);
            curStart = hp; curMedium = inMedium;
            ro = hp; rd = woW; rmint = ALVRL_EPSILON;
        } else break;
        if (depth++ >= tp.rrDepth) {
            const float q = fminf(xmul(xmul(fmaxf(fmaxf(thr[0], thr[1]), thr[2]), eta), eta), 0.95f);
            if (smp.next() >= q) break;
            const float rq = xdiv(1.0f, q);
            for (int i = 0; i < 3; ++i) thr[i] = xmul(thr[i], rq);
        }
    }
    return count;
}

__global__ void __launch_bounds__(64) k_trace_count(SceneDev sc, MediumDev med, EmitterDev em, TracerParams tp, const float4 *__restrict__ triVerts,
                                                    const uint32_t *__restrict__ triMat, const float4 *__restrict__ matAlbedo, const uint32_t *__restrict__ matBits,
                                                    const float4 *__restrict__ matOptics, uint32_t first, uint32_t n, uint32_t *__restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    counts[i] = trace_particle<false>(sc, med, em, tp, triVerts, triMat, matAlbedo, matBits, matOptics, first + i, nullptr);
}
__global__ void __launch_bounds__(64) k_trace_write(SceneDev sc, MediumDev med, EmitterDev em, TracerParams tp, const float4 *__restrict__ triVerts,
                                                    const uint32_t *__restrict__ triMat, const float4 *__restrict__ matAlbedo, const uint32_t *__restrict__ matBits,
                                                    const float4 *__restrict__ matOptics, uint32_t n, const uint32_t *__restrict__ offset, float *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || offset[i + 1] == offset[i]) return;
    trace_particle<true>(sc, med, em, tp, triVerts, triMat, matAlbedo, matBits, matOptics, i, out + 9 * (size_t) offset[i]);
}

void launch_trace_count(const SceneDev &sc, const MediumDev &med, const uint32_t *emTris, const float *emCdf, uint32_t emN, const float emPower[3],
                        uint64_t seed, int shortVrls, int maxDepth, int rrDepth, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint32_t first, uint32_t n, uint32_t *counts, cudaStream_t st) {
    EmitterDev em; em.tris = emTris; em.cdf = emCdf; em.n = emN; em.power[0] = emPower[0]; em.power[1] = emPower[1]; em.power[2] = emPower[2];
    TracerParams tp; tp.seed = seed; tp.shortVrls = shortVrls; tp.maxDepth = maxDepth; tp.rrDepth = rrDepth;
    if (n) k_trace_count<<<(n + 63) / 64, 64, 0, st>>>(sc, med, em, tp, triVerts, triMat, matAlbedo, matBits, matOptics, first, n, counts);
}
void launch_trace_write(const SceneDev &sc, const MediumDev &med, const uint32_t *emTris, const float *emCdf, uint32_t emN, const float emPower[3],
                        uint64_t seed, int shortVrls, int maxDepth, int rrDepth, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint32_t n, const uint32_t *offset, float *out, cudaStream_t st) {
    EmitterDev em; em.tris = emTris; em.cdf = emCdf; em.n = emN; em.power[0] = emPower[0]; em.power[1] = emPower[1]; em.power[2] = emPower[2];
    TracerParams tp; tp.seed = seed; tp.shortVrls = shortVrls; tp.maxDepth = maxDepth; tp.rrDepth = rrDepth;
    if (n) k_trace_write<<<(n + 63) / 64, 64, 0, st>>>(sc, med, em, tp, triVerts, triMat, matAlbedo, matBits, matOptics, n, offset, out);
}

} // namespace alvrl
