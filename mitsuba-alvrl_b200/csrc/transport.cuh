/*
 * transport.cuh -- the VRL <-> camera-segment transport estimator and its two kernels.
 *
 * integrate_pair() restates vrlIntegrator::integrateVRL (src/integrators/vrl/vrlIntegrator.cpp:603-785)
 * with sampleUVKulla / sampleVtoDistance / getClosestPoints / KullaSampling (860-1032) inlined, per-pair
 * invariants hoisted out of the sample loops and Scene::evalTransmittance (scene.cpp:619-679) fused in as a
 * stackless BVH query.  This file is compiled twice:
 *
 *   transport_strict.cu  (-fmad=false, ALVRL_FAST undefined): the same fp32 operations in the same order as
 *                        the reference, exp through double (math.h:185-187).  Parity flavour.
 *   transport_fast.cu    (FMA contraction, ALVRL_FAST defined): MUFU-based exp/rcp/rsqrt, polynomial atan,
 *                        sinh/asinh through ex2/lg2, merged transmittance exponentials.  Speed flavour; agrees
 *                        with the strict one to ~1e-5 relative (tests state the tolerance).
 *
 * Work decomposition (both kernels): one thread owns one camera segment (its record lives in registers), a
 * CTA owns 128 segments, and the VRL records are staged through shared memory in 64-record tiles by TMA bulk
 * copies (cp.async.bulk + mbarrier, double buffered), so every VRL record is read once per CTA and broadcast
 * to all 128 segments.  k_build_R writes R(mean,var) column-major (one column per VRL, rows contiguous), so a
 * warp's 32 rows store 256 contiguous bytes and the clustering kernels read whole columns coalesced;
 * k_render accumulates RGB in registers and stores once per pixel.
 */
#pragma once
#include "dev_common.cuh"

#ifndef ALVRL_FLAVOR
#error "define ALVRL_FLAVOR (strict|fast) before including transport.cuh"
#endif

/* VRL records per TMA tile and resident CTAs per SM the kernels are compiled for.  Fast flavour: 8 CTAs (32 warps) per
 * SM at 64 registers, 23 KB of shared memory each -- measured best of {5, 6, 7, 8, 10, 12} x {64, 128, 256} on the C2
 * R build (profiles/README.md); the strict flavour keeps its registers. */
#ifndef ALVRL_TILE_VRLS
#ifdef ALVRL_FAST
#define ALVRL_TILE_VRLS 128
#else
#define ALVRL_TILE_VRLS 256
#endif
#endif
#ifndef ALVRL_MIN_CTAS
#ifdef ALVRL_FAST
#define ALVRL_MIN_CTAS 8
#else
#define ALVRL_MIN_CTAS 1
#endif
#endif
#define ALVRL_CTA_SEGS 128

/* ---- flavoured math ---------------------------------------------------------------------------- */
#if defined(ALVRL_FAST) && !defined(ALVRL_DIAG_EXACT_M)
__device__ __forceinline__ float m_exp(float x) { return __expf(x); }
__device__ __forceinline__ float m_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float m_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float m_sqrt(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ float m_atan(float x) { return atanf(x); }
__device__ __forceinline__ float m_tan(float x) { return tanf(x); }
__device__ __forceinline__ float m_sinh(float x) { return sinhf(x); }
__device__ __forceinline__ float m_asinh(float x) { return asinhf(x); }
#else
__device__ __forceinline__ float m_exp(float x) { return exp_ref(x); }
__device__ __forceinline__ float m_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float m_div(float a, float b) { return a / b; }
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
/* correctly rounded through double, like the oracle pins them (libm version of the reference is unpinned) */
__device__ __forceinline__ float m_atan(float x) { return (float) atan((double) x); }
__device__ __forceinline__ float m_tan(float x) { return (float) tan((double) x); }
__device__ __forceinline__ float m_sinh(float x) { return (float) sinh((double) x); }
__device__ __forceinline__ float m_asinh(float x) { return (float) asinh((double) x); }
#endif

__device__ __forceinline__ float m_len(const F3 &a) { return m_sqrt(len2(a)); }
__device__ __forceinline__ float m_dist(const F3 &a, const F3 &b) { return m_len(a - b); }
__device__ __forceinline__ F3 m_normalize(const F3 &a) { float r = m_rcp(m_len(a)); return a * r; }

struct Rng {
    uint32_t key, k;
    const float *tape;
    __device__ __forceinline__ float next() {
        if (tape) return __ldg(&tape[k++]);
        return alvrl_rng_uniform(key, k++);
    }
};

/* Medium::eval restricted to what integrateVRL consumes: transmittance, sigmaS, pdfFailure
 * (homogeneous.cpp:354-396 with strategy=balance; heterogeneous.cpp:665-691 with method=simpson). */
template <int MED>
__device__ __forceinline__ void medium_eval(const MediumDev &m, const F3 &o, const F3 &d, float dist,
                                            float T[3], float sS[3], float &pdfFailure) {
    if (MED == 0) {
        float pf = 0, mx = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float temp = m_exp(-m.sigmaT[i] * dist);
            pf += temp;
            T[i] = temp;         /* (sigmaT * (-distance)).exp(): the same value as temp */
            mx = fmaxf(mx, temp);
            sS[i] = m.sigmaS[i];
        }
        pf = m_div(pf, 3.0f);
        pdfFailure = pf * m.samplingWeight + (1 - m.samplingWeight);
        if ((double) mx < 1e-20) T[0] = T[1] = T[2] = 0.0f;
    } else {
        const float e = m_exp(-grid_optical_depth(m, o, d, 0.0f, dist));
        const F3 p = xadd3(o, xscale(d, dist));
        const float dens = xmul(grid_lookup(m, p), m.scale);
#pragma unroll
        for (int i = 0; i < 3; i++) { T[i] = e; sS[i] = m.albedo[i] * dens; }
        pdfFailure = e;
    }
}

/* Scene::evalTransmittance (scene.cpp:619-679), opaque scenes */
template <int MED>
__device__ __forceinline__ bool eval_transmittance(const TransportParams &P, const F3 &p1, bool onSurf, const F3 &p2, float T[3]) {
    F3 dir; float remaining;
#ifdef ALVRL_FAST
    {
        const F3 dd = p2 - p1;
        const float l2 = len2(dd);
        const float il = rsqrtf(l2);
        remaining = l2 * il;
        dir = dd * il;
        if (!(remaining > 0)) { T[0] = T[1] = T[2] = 1; return true; }
        /* adaptive epsilon of the shadow-ray overload, skdtree.cpp:154-157 */
        const float mint = onSurf ? ALVRL_EPSILON * fmaxf(fmaxf(fabsf(p1.x), fabsf(p1.y)), fabsf(p1.z)) : 0.0f;
        if (remaining > mint && bvh_occluded_fast(P.scene, p1, dir, mint, remaining)) { T[0] = T[1] = T[2] = 0; return false; }
    }
#else
    if (segment_occluded(P.scene, p1, onSurf, p2, dir, remaining)) { T[0] = T[1] = T[2] = 0; return false; }
    if (!(remaining > 0)) { T[0] = T[1] = T[2] = 1; return true; }
#endif
    if (MED == 0) {
        const float negLength = 0.0f - remaining;                          /* homogeneous.cpp:266-273 */
#pragma unroll
        for (int i = 0; i < 3; i++) T[i] = P.medium.sigmaT[i] != 0 ? m_exp(P.medium.sigmaT[i] * negLength) : 1.0f;
    } else {
        const float e = m_exp(-grid_optical_depth(P.medium, p1, dir, 0.0f, remaining));
        T[0] = T[1] = T[2] = e;
    }
    return true;
}

__device__ __forceinline__ float phase_eval(const MediumDev &m, float cosWiWo) {
    if (m.phaseType == ALVRL_PHASE_ISOTROPIC) return ALVRL_INV_FOURPI;    /* isotropic.cpp:76-78 */
    const float temp = 1.0f + m.g * m.g + 2.0f * m.g * cosWiWo;            /* hg.cpp:107-110 */
    return m_div(ALVRL_INV_FOURPI * (1 - m.g * m.g), temp * m_sqrt(temp));
}

__device__ __forceinline__ bool spec_valid(const float c[3]) {            /* spectrum.h:467-472 */
    return isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]) && c[0] >= 0.0f && c[1] >= 0.0f && c[2] >= 0.0f;
}

/* vrlIntegrator::getClosestPoints (962-1032): closest distance h and the closest point on the VRL.
 * D, sN, tN are differences of nearly equal products when the two segments are close to parallel (relative error
 * eps / sin^2): they are evaluated with explicitly rounded operations in the reference's order in BOTH flavours, so that
 * FMA contraction in the fast flavour cannot move h and Vh of such pairs away from the oracle's. */
__device__ __forceinline__ float closest_points(const F3 &S1P0, const F3 &S1P1, const F3 &S2P0, const F3 &S2P1, F3 &S2h) {
    const F3 u = S1P1 - S1P0, v = S2P1 - S2P0, w = S1P0 - S2P0;
    const float a = xdot(u, u), b = xdot(u, v), c = xdot(v, v), d = xdot(u, w), e = xdot(v, w);
    const float D = xsub(xmul(a, c), xmul(b, b));
    float sN, sD = D, tN, tD = D;
    if (D < xmul(xmul(ALVRL_EPSILON, a), c)) { sN = 0.0f; sD = 1.0f; tN = e; tD = c; }
    else {
        sN = xsub(xmul(b, e), xmul(c, d));
        tN = xsub(xmul(a, e), xmul(b, d));
        if (sN < 0.0f) { sN = 0.0f; tN = e; tD = c; }
        else if (sN > sD) { sN = sD; tN = xadd(e, b); tD = c; }
    }
    if (tN < 0.0f) {
        tN = 0.0f;
        if (-d < 0.0f) sN = 0.0f;
        else if (-d > a) sN = sD;
        else { sN = -d; sD = a; }
    } else if (tN > tD) {
        tN = tD;
        const float db = xadd(-d, b);
        if (db < 0.0f) sN = 0;
        else if (db > a) sN = sD;
        else { sN = db; sD = a; }
    }
    const float sc = m_div(sN, sD), tc = m_div(tN, tD);
    const F3 dP = xsub3(xadd3(w, xscale(u, sc)), xscale(v, tc));       /* the distance of two nearly touching segments cancels too */
    S2h = xadd3(S2P0, xscale(v, tc));
    return m_sqrt(xdot(dP, dP));
}
/* cos / sin of the angle between the camera ray and the VRL (sampleVtoDistance, 925-927): 1 - cos^2 cancels for nearly
 * parallel pairs, so both flavours round it exactly like the reference does */
__device__ __forceinline__ void cos_sin_theta(const F3 &dn, const F3 &SV, float &cosTheta, float &sinTheta) {
    cosTheta = xdot(dn, SV);
    sinTheta = m_sqrt(fmaxf(0.0f, xsub(1.0f, xmul(cosTheta, cosTheta))));
}

/*
 * integrateVRL for one (segment, VRL) pair.  WANT_RGB: accumulate the RGB estimate (render); WANT_STAT: luminance
 * mean and variance of the mean (R entry).  Consumes 2*Nvv + Nvs uniforms in the reference's order.
 */
template <int MED, bool WANT_RGB, bool WANT_STAT>
__device__ __forceinline__ void integrate_pair(const TransportParams &P, const SegRec &seg, const float4 vS, const float4 vE,
                                               const float4 vDir, const float4 vPow, Rng &rng,
                                               float rgb[3], float &outMean, float &outVar) {
    const F3 S = f3(vS), End = f3(vE), SV = f3(vDir);
    const float vlen = vS.w;
    const F3 E = f3(seg.o), EU = f3(seg.d), Usurf = f3(seg.p);
    const float edist = seg.o.w;
    const int Nvv = P.Nvv, Nvs = P.Nvs;
    /* LiInternal's `weight` of the segment (1 for a camera segment; specular chains, 500): inside the estimate for the rows of R
     * (getVRLContributions passes it on, 811); the render pass multiplies the pixel's sum afterwards (598) */
    const float wgt[3] = {seg.p.w, seg.n.w, seg.albedo.w};
    (void) wgt;
    if (WANT_RGB) rgb[0] = rgb[1] = rgb[2] = 0;
    outMean = 0; outVar = 0;

    /* ---- volume to volume, L (V|D|S)* V V S* E (646-703) ---- */
    if (Nvv > 0) {
        /* per-pair part of sampleVtoDistance (916-953) */
        float cosTheta, sinTheta;
        cos_sin_theta(f3(seg.dn), SV, cosTheta, sinTheta);
        const bool parallel = sinTheta < ALVRL_EPSILON;
        float h = 0, A0 = 0, A1 = 0, dVhS = 0;
        if (!parallel) {
            F3 Vh;
            h = closest_points(E, Usurf, S, End, Vh);
            const float V0c = -1 * m_dist(Vh, S);
            const float V1c = m_dist(Vh, End);
            A0 = m_asinh(m_div(V0c, h) * sinTheta);
            A1 = m_asinh(m_div(V1c, h) * sinTheta);
            dVhS = m_dist(Vh, S);
        }
        /* per-pair part of KullaSampling(A = E, B = E + dist*d, D = V) (889-896) */
        const F3 B = E + (edist * EU);
        const F3 dirE = m_normalize(B - E);
        const float dAB = m_dist(E, B);
        const float invNvv = m_rcp((float) Nvv);

        float mean = 0, M2 = 0;
        for (int s = 0; s < Nvv; s++) {
            float lum = 0;
            const float u1 = rng.next();
            F3 V; float pdf;
            if (parallel) {
                V = S + u1 * (End - S);
                pdf = m_div(1.0f, vlen);
            } else {
                float newV = h * m_sinh(A0 + (u1 * (A1 - A0)));
                newV = m_div(newV, sinTheta);
                const float result = m_div(1.0f, m_sqrt(h * h + newV * newV * sinTheta * sinTheta));
                const float denom = m_div(A1 - A0, sinTheta);
                newV += dVhS;
                V = S + newV * SV;
                pdf = m_div(result, denom);
            }
            /* KullaSampling along the eye segment w.r.t. V (889-914) */
            const float u2 = rng.next();
            F3 U;
            {
                const float dotPr = dot(dirE, V - E);
                const F3 I = E + (dotPr * dirE);
                const float Dis = m_dist(V, I);
                const float dAI = m_dist(E, I);
                float angle_a = m_atan(m_div(dAI, Dis));
                float angle_b = m_atan(m_div(m_dist(I, B), Dis));
                if (dotPr > 0) {
                    angle_a *= -1;
                    if (dAI > dAB) angle_b *= -1;
                }
                const float t = Dis * m_tan(((1.0f - u2) * angle_a) + (u2 * angle_b));
                const float pdfU = m_div(Dis, (angle_b - angle_a) * (Dis * Dis + t * t));
                U = I + (t * dirE);
                pdf *= pdfU;
            }
            const F3 UV = U - V;
            const float d2 = len2(UV);
            const float dUV = m_sqrt(d2);
            if (dUV != 0) {
                const F3 VU = UV * m_rcp(dUV);
                float Tuv[3];
                if (eval_transmittance<MED>(P, U, false, V, Tuv) && !(Tuv[0] == 0 && Tuv[1] == 0 && Tuv[2] == 0)) {
                    float Te[3], sSe[3], pfE, Tv[3], sSv[3], pfV;
                    medium_eval<MED>(P.medium, E, EU, m_dist(E, U), Te, sSe, pfE);
                    medium_eval<MED>(P.medium, S, SV, m_dist(S, V), Tv, sSv, pfV);
                    const float rpdf = m_rcp(pdf), rd2 = m_div(1.0f, d2), rpf = m_rcp(pfV);
                    const float phU = phase_eval(P.medium, dot(-VU, -EU));
                    const float phV = phase_eval(P.medium, dot(-SV, VU));
                    const float pw[3] = {vPow.x, vPow.y, vPow.z};
                    float c[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        float x = WANT_STAT ? wgt[i] * pw[i] : pw[i];           /* 668-669: contribution = weight; *= power */
                        x *= (sSv[i] * sSe[i]) * rpdf;
                        x *= rd2;
                        x *= Tv[i];
                        x *= Tuv[i];
                        x *= Te[i];
                        if (P.shortVrls) x *= rpf;
                        x *= phU;
                        x *= phV;
                        c[i] = x;
                    }
                    if (spec_valid(c)) {
                        if (WANT_RGB) { rgb[0] += c[0] * invNvv; rgb[1] += c[1] * invNvv; rgb[2] += c[2] * invNvv; }
                        lum = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f;
                    }
                }
            }
            if (WANT_STAT) {                                               /* 693-699 */
                const float delta = lum - mean;
                mean += m_div(delta, (float) (s + 1));
                M2 += delta * (lum - mean);
            }
        }
        if (WANT_STAT) { outMean += mean; outVar += m_div(M2, (float) ((Nvv - 1) * Nvv)); }
    }

    /* ---- volume to surface, L (V|D|S)* V D S* E (706-782) ---- */
    if (Nvs > 0) {
        const float tE[3] = {seg.tE.x, seg.tE.y, seg.tE.z};
        const uint32_t flags = __float_as_uint(seg.dn.w);
        float mean = 0, M2 = 0;
        if (!(tE[0] == 0 && tE[1] == 0 && tE[2] == 0) && (flags & SEG_SMOOTH)) {
            /* per-pair part of KullaSampling(A = S, B = End, D = Usurf) */
            const float dotPr = dot(SV, Usurf - S);
            const F3 I = S + (dotPr * SV);
            const float Dis = m_dist(Usurf, I);
            const float dAI = m_dist(S, I);
            float angle_a = m_atan(m_div(dAI, Dis));
            float angle_b = m_atan(m_div(m_dist(I, End), Dis));
            if (dotPr > 0) {
                angle_a *= -1;
                if (dAI > vlen) angle_b *= -1;
            }
            const float invNvs = m_rcp((float) Nvs);
            const F3 nrm = f3(seg.n);
            const float wiz = seg.d.w;
            for (int s = 0; s < Nvs; s++) {
                float lum = 0;
                const float u = rng.next();
                const float t = Dis * m_tan(((1.0f - u) * angle_a) + (u * angle_b));
                const float pdf = m_div(Dis, (angle_b - angle_a) * (Dis * Dis + t * t));
                const F3 V = I + (t * SV);
                const F3 UV = Usurf - V;
                const float d2 = len2(UV);
                const float dUV = m_sqrt(d2);
                if (dUV != 0) {
                    const F3 VU = UV * m_rcp(dUV);
                    float Tuv[3], Tv[3], sSv[3], pfV;
                    eval_transmittance<MED>(P, Usurf, true, V, Tuv);
                    medium_eval<MED>(P.medium, S, SV, m_dist(S, V), Tv, sSv, pfV);
                    const float rpdf = m_rcp(pdf), rd2 = m_div(1.0f, d2), rpf = m_rcp(pfV);
                    const float phV = phase_eval(P.medium, dot(-SV, VU));
                    const float cosWo = dot(-VU, nrm);                       /* diffuse.cpp:110-118 */
                    const bool front = !(wiz <= 0 || cosWo <= 0);
                    const float al[3] = {seg.albedo.x, seg.albedo.y, seg.albedo.z};
                    const float pw[3] = {vPow.x, vPow.y, vPow.z};
                    float c[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        float x = WANT_STAT ? wgt[i] * pw[i] : pw[i];           /* 745-746 */
                        x *= P.medium.sigmaS[i] * rpdf;                      /* base-class getSigmaS(), quirk B2 */
                        x *= rd2;
                        x *= Tv[i];
                        x *= Tuv[i];
                        x *= tE[i];
                        if (P.shortVrls) x *= rpf;
                        x *= phV;
                        x *= front ? al[i] * (ALVRL_INV_PI * cosWo) : 0.0f;
                        c[i] = x;
                    }
                    if (spec_valid(c)) {
                        if (WANT_RGB) { rgb[0] += c[0] * invNvs; rgb[1] += c[1] * invNvs; rgb[2] += c[2] * invNvs; }
                        lum = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f;
                    }
                }
                if (WANT_STAT) {
                    const float delta = lum - mean;
                    mean += m_div(delta, (float) (s + 1));
                    M2 += delta * (lum - mean);
                }
            }
        } else {
            /* the loop is skipped without consuming uniforms (727) */
        }
        if (WANT_STAT) { outMean += mean; outVar += m_div(M2, (float) ((Nvs - 1) * Nvs)); }
    }
}

#ifdef ALVRL_FAST
#include "transport_fast_impl.cuh"
#endif

/* ---- TMA bulk-copy tile pipeline ------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TileSmem {
    VrlRec tile[2][ALVRL_TILE_VRLS];
    uint64_t full[2];
    uint64_t pad[14];              /* keep the structures that follow in dynamic shared memory 128-byte aligned */
};

#define ALVRL_CAT2(a, b) a##_##b
#define ALVRL_CAT(a, b) ALVRL_CAT2(a, b)
#define ALVRL_NAME(base) ALVRL_CAT(base, ALVRL_FLAVOR)

/*
 * "Building R" (vrlIntegrator.cpp:302-337,792-825): rows = representative-pixel segments, columns = VRLs.
 * grid.x = row blocks of 128, grid.y = VRL chunks of vrlsPerCta (multiple of the tile size).
 */
/* (the grid-medium kernels wait for memory: measured with 5 CTAs per SM and 96 registers they lose 60 % against 8 CTAs with
 * spills -- the march needs warps in flight more than registers) */
#define ALVRL_MIN_CTAS_MED(MED) ALVRL_MIN_CTAS
template <int MED, int SMALL, bool WEIGHTED>
__global__ void __launch_bounds__(ALVRL_CTA_SEGS, ALVRL_MIN_CTAS_MED(MED)) ALVRL_NAME(k_build_R)(TransportParams P, const SegRec *__restrict__ rowSegs, uint32_t numRows,
                                                                       const VrlRec *__restrict__ vrls, float2 *__restrict__ R, uint32_t ldR,
                                                                       uint32_t vrlsPerCta, const uint32_t *__restrict__ rowKey) {
#ifdef ALVRL_FAST
    extern __shared__ __align__(128) unsigned char dynSmem[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(dynSmem);
    BvhSmem &sbvh = *reinterpret_cast<BvhSmem *>(dynSmem + sizeof(TileSmem));
    if (SMALL) stage_bvh<SMALL>(sbvh, P.scene);
#else
    __shared__ __align__(128) TileSmem sm;
#endif
    const uint32_t tid = threadIdx.x;
    /* grid medium: the VRL chunk is the fast grid index, so that the CTAs in flight share a few row blocks -- their camera-
     * segment marches (E -> U) then hit in L2 instead of streaming the 512 MB grid from DRAM once per VRL chunk */
    const uint32_t rowBlock = MED == 1 ? blockIdx.y : blockIdx.x, chunk = MED == 1 ? blockIdx.x : blockIdx.y;
    const uint32_t row = rowBlock * ALVRL_CTA_SEGS + tid;
    const uint32_t N = P.numVrls;
    const uint32_t vBegin = chunk * vrlsPerCta;
    const uint32_t vEnd = min(N, vBegin + vrlsPerCta);
    if (vBegin >= vEnd) return;
    const uint32_t numTiles = (vEnd - vBegin + ALVRL_TILE_VRLS - 1) / ALVRL_TILE_VRLS;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t t) {
        const uint32_t v0 = vBegin + t * ALVRL_TILE_VRLS;
        const uint32_t cnt = min((uint32_t) ALVRL_TILE_VRLS, vEnd - v0);
        const uint32_t bytes = cnt * (uint32_t) sizeof(VrlRec);
        mbar_expect_tx(&sm.full[t & 1], bytes);
        tma_load_1d(&sm.tile[t & 1][0], vrls + v0, bytes, &sm.full[t & 1]);
    };
    if (tid == 0) issue(0);

    SegRec seg;
    memset(&seg, 0, sizeof(seg));
    bool active = row < numRows;
    if (active) {
        seg = rowSegs[row];
        active = (__float_as_uint(seg.dn.w) & SEG_VALID) != 0;
    }
    const bool scattering = !(P.medium.sigmaS[0] == 0 && P.medium.sigmaS[1] == 0 && P.medium.sigmaS[2] == 0);
    /* stream of the row: (row, vrl); the segments of a specular chain are rows of their own matrix and carry the key
     * (row of their pixel) + (ordinal in the chain << 24) */
    const uint32_t keyRow = rowKey ? (row < numRows ? rowKey[row] : 0u) : P.rowBase + row;
#ifdef ALVRL_FAST
    SegSides sides; sides.slabHull = sides.slabSurf = sides.planes = 0;
    if (SMALL == 2) sides = seg_sides(P.occ, seg, active && scattering);
#endif

    for (uint32_t t = 0; t < numTiles; t++) {
        if (tid == 0 && t + 1 < numTiles) issue(t + 1);
        mbar_wait(&sm.full[t & 1], (t >> 1) & 1);
        const uint32_t v0 = vBegin + t * ALVRL_TILE_VRLS;
        const uint32_t cnt = min((uint32_t) ALVRL_TILE_VRLS, vEnd - v0);
        {
#pragma unroll 1
            for (uint32_t j = 0; j < cnt; j++) {
                float mean = 0, var = 0;
                const uint32_t v = v0 + j;
#if defined(ALVRL_FAST) && !defined(ALVRL_DIAG_GENERIC)
                {
                    /* every lane of the warp takes part: the visibility queries and the grid marches are warp-collective */
                    const VrlRec &vr = sm.tile[t & 1][j];
                    Rng rng;
                    rng.tape = P.tape ? P.tape + ((size_t) (P.rowBase + row) * N + v) * P.tapeK : nullptr;
                    rng.key = alvrl_rng_key(P.seed, P.rngDomain, keyRow, v);
                    rng.k = 0;
                    float rgb[3], m, s2;
                    PairCull cull; cull.boxVV = cull.boxVS = cull.planes = 0xffffffffu;
                    if (SMALL == 2) cull = pair_cull(P.occ, sides, vr.dir, vr.power);
                    if constexpr (MED == 1) integrate_pair_grid_fast<false, true, SMALL>(P, &sbvh, seg, vr.s, vr.e, vr.dir, vr.power, rng, rgb, m, s2, active && scattering, cull);
                    else integrate_pair_fast<MED, false, true, SMALL, WEIGHTED>(P, &sbvh, seg, vr.s, vr.e, vr.dir, vr.power, rng, rgb, m, s2, active && scattering, cull);
                    mean = m * P.normalization;
                    var = s2 * P.normalization * P.normalization;
                }
#else
                if (active && scattering) {
                    const VrlRec &vr = sm.tile[t & 1][j];
                    Rng rng;
                    rng.tape = P.tape ? P.tape + ((size_t) (P.rowBase + row) * N + v) * P.tapeK : nullptr;
                    rng.key = alvrl_rng_key(P.seed, P.rngDomain, keyRow, v);
                    rng.k = 0;
                    float rgb[3], m, s2;
                    integrate_pair<(MED == 2 ? 0 : MED), false, true>(P, seg, vr.s, vr.e, vr.dir, vr.power, rng, rgb, m, s2);
                    mean = m * P.normalization;                              /* vrlIntegrator.cpp:812-813 */
                    var = s2 * P.normalization * P.normalization;
                }
#endif
                if (row < numRows) R[(size_t) v * ldR + row] = make_float2(mean, var);   /* lanes = consecutive rows: coalesced */
            }
        }
        __syncthreads();
    }
}

/*
 * Render pass, getClusteredVrlContributions (542-599) / getVRLContributions (792-825) for pixel centres.
 * One CTA = up to 128 pixels of one slice; the slice's representative VRLs (records gathered per slice, cluster
 * weight in e.w) stream through the same TMA tile pipeline.  work[cta] = {slice, firstPixel, pixelCount, 0}.
 */
template <int MED, bool CLUSTERED, int SMALL>
__global__ void __launch_bounds__(ALVRL_CTA_SEGS, ALVRL_MIN_CTAS_MED(MED)) ALVRL_NAME(k_render)(TransportParams P, const SegRec *__restrict__ pixSegs,
                                                                      const uint32_t *__restrict__ slicePixels, const uint4 *__restrict__ work,
                                                                      const VrlRec *__restrict__ repRecs, const uint32_t *__restrict__ repOffset,
                                                                      float4 *__restrict__ fb, uint32_t W, uint32_t H, const uint32_t *__restrict__ segKey) {
#ifdef ALVRL_FAST
    extern __shared__ __align__(128) unsigned char dynSmem[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(dynSmem);
    BvhSmem &sbvh = *reinterpret_cast<BvhSmem *>(dynSmem + sizeof(TileSmem));
    if (SMALL) stage_bvh<SMALL>(sbvh, P.scene);
#else
    __shared__ __align__(128) TileSmem sm;
#endif
    const uint32_t tid = threadIdx.x;
    const uint4 wk = work[blockIdx.x];
    const uint32_t vBegin = repOffset[wk.x], vEnd = repOffset[wk.x + 1];
    const uint32_t numTiles = (vEnd - vBegin + ALVRL_TILE_VRLS - 1) / ALVRL_TILE_VRLS;
    if (tid == 0) {
        mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t t) {
        const uint32_t v0 = vBegin + t * ALVRL_TILE_VRLS;
        const uint32_t cnt = min((uint32_t) ALVRL_TILE_VRLS, vEnd - v0);
        const uint32_t bytes = cnt * (uint32_t) sizeof(VrlRec);
        mbar_expect_tx(&sm.full[t & 1], bytes);
        tma_load_1d(&sm.tile[t & 1][0], repRecs + v0, bytes, &sm.full[t & 1]);
    };
    if (tid == 0 && numTiles) issue(0);

    const bool inRange = tid < wk.z;
    uint32_t pixel = 0;
    SegRec seg;
    memset(&seg, 0, sizeof(seg));
    bool active = false;
    if (inRange) {
        pixel = slicePixels[wk.y + tid];
        seg = pixSegs[pixel];
        active = (__float_as_uint(seg.dn.w) & SEG_VALID) != 0;
    }
    const bool scattering = !(P.medium.sigmaS[0] == 0 && P.medium.sigmaS[1] == 0 && P.medium.sigmaS[2] == 0);
    float Li[3] = {0, 0, 0};
    /* segments of specular chains are rendered from their own list: slicePixels indexes that list, segKey gives the stream key
     * (pixel + (ordinal << 24)) and the result goes to fb[list index] (k_chain_accumulate adds it to the pixel, weighted) */
    const uint32_t keyPix = segKey ? (inRange ? segKey[pixel] : 0u) : pixel;
#ifdef ALVRL_FAST
    SegSides sides; sides.slabHull = sides.slabSurf = sides.planes = 0;
    if (SMALL == 2) sides = seg_sides(P.occ, seg, active && scattering);
#endif
    for (uint32_t t = 0; t < numTiles; t++) {
        if (tid == 0 && t + 1 < numTiles) issue(t + 1);
        mbar_wait(&sm.full[t & 1], (t >> 1) & 1);
        const uint32_t v0 = vBegin + t * ALVRL_TILE_VRLS;
        const uint32_t cnt = min((uint32_t) ALVRL_TILE_VRLS, vEnd - v0);
        {
#pragma unroll 1
            for (uint32_t j = 0; j < cnt; j++) {
                const VrlRec &vr = sm.tile[t & 1][j];
                Rng rng;
                rng.tape = nullptr;
                rng.key = alvrl_rng_key(P.seed, ALVRL_RNG_RENDER, keyPix, v0 + j - vBegin);
                rng.k = 0;
                float rgb[3] = {0, 0, 0}, m, s2;
#if defined(ALVRL_FAST) && !defined(ALVRL_DIAG_GENERIC)
                {
                    PairCull cull; cull.boxVV = cull.boxVS = cull.planes = 0xffffffffu;
                    if (SMALL == 2) cull = pair_cull(P.occ, sides, vr.dir, vr.power);
                    if constexpr (MED == 1) integrate_pair_grid_fast<true, false, SMALL>(P, &sbvh, seg, vr.s, vr.e, vr.dir, vr.power, rng, rgb, m, s2, active && scattering, cull);
                    else integrate_pair_fast<MED, true, false, SMALL, false>(P, &sbvh, seg, vr.s, vr.e, vr.dir, vr.power, rng, rgb, m, s2, active && scattering, cull);
                }
#else
                if (active && scattering) integrate_pair<(MED == 2 ? 0 : MED), true, false>(P, seg, vr.s, vr.e, vr.dir, vr.power, rng, rgb, m, s2);
#endif
                if (CLUSTERED) {                                             /* 587-589: Li += weight_k * integrateVRL */
                    const float w = vr.e.w;
                    Li[0] += w * rgb[0]; Li[1] += w * rgb[1]; Li[2] += w * rgb[2];
                } else {                                                     /* 810,815: vrlContribution *= normalization */
                    Li[0] += rgb[0] * P.normalization; Li[1] += rgb[1] * P.normalization; Li[2] += rgb[2] * P.normalization;
                }
            }
        }
        __syncthreads();
    }
    if (inRange) {
        if (CLUSTERED) {                                                     /* 590: Li /= particleCount (recip multiply) */
            const float r = m_div(1.0f, P.invParticleDiv);
            Li[0] *= r; Li[1] *= r; Li[2] *= r;
        }
        const uint32_t x = pixel / H, y = pixel % H;
        if (segKey) fb[pixel] = make_float4(Li[0], Li[1], Li[2], 1.0f);
        else fb[(size_t) y * W + x] = make_float4(Li[0], Li[1], Li[2], 1.0f);
    }
}
