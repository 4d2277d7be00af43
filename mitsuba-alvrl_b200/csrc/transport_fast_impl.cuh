/*
 * transport_fast_impl.cuh -- speed flavour of integrateVRL (included by transport.cuh when ALVRL_FAST is defined).
 *
 * Same estimator and the same uniforms as the strict flavour (vrlIntegrator.cpp:603-785), restructured for the SM:
 *   - the three transmittances of a sample are one exponential per channel, exp(-sigma_t (|SV| + |VU| + |UE|)), plus
 *     one for pdfFailure; a grey medium (sigma_t equal in all channels, the BASELINE configs) needs two MUFU.EX2 in all;
 *   - |A-I|, |I-B| of KullaSampling follow from the projection parameter (I = A + dotPr*dir), so a sample needs three
 *     square roots instead of seven; the signed-angle form replaces the sign fix-up branches (889-903);
 *   - sinh / asinh / tan go through MUFU (ex2, lg2, sin, cos, rcp); atan keeps the polynomial atanf;
 *   - visibility: scenes with <= 32 BVH leaves keep leaf boxes + triangles in shared memory and answer a shadow ray
 *     with a flat, warp-uniform sweep over the leaf boxes (broadcast LDS, no inner nodes) followed by triangle tests
 *     of the entered leaves only; larger scenes use the per-lane stackless traversal of dev_common.cuh.
 * Agreement with the strict flavour / the oracle: median 1.6e-7 relative per R entry; tests/test_c2_parity_gpu.py holds it to 1e-4
 * per entry on C2 itself, with a stated and asserted fraction of outliers (1.5e-4 of the entries, oracle-flagged grazing ties apart).
 */
#pragma once
#include "occ_query.h"

#define ALVRL_SMALL_LEAVES 32
#define ALVRL_SMALL_TRIS 128

/* small scenes: VIS 1 = all leaf boxes + triangle records in shared memory; VIS 2 = the compiled occluder set
 * (occluders.h): slabs and planes come from the kernel parameters, shared memory only holds the planar groups' triangles */
struct BvhSmem {
    float4 leaves[2 * ALVRL_SMALL_LEAVES];
    float4 tris[3 * ALVRL_SMALL_TRIS];
};

template <int VIS>
__device__ __forceinline__ void stage_bvh(BvhSmem &sb, const SceneDev &sc) {
    if (VIS == 1) {
        const float4 *gn = reinterpret_cast<const float4 *>(sc.leafNodes);
        const float4 *gt = reinterpret_cast<const float4 *>(sc.trisFast);
        for (uint32_t i = threadIdx.x; i < 2 * sc.numLeaves; i += blockDim.x) sb.leaves[i] = __ldg(&gn[i]);
        for (uint32_t i = threadIdx.x; i < 3 * sc.numTris; i += blockDim.x) sb.tris[i] = __ldg(&gt[i]);
    } else if (VIS == 2) {
        for (uint32_t i = threadIdx.x; i < 3 * sc.numOccTris; i += blockDim.x) sb.tris[i] = __ldg(&sc.occTris[i]);
    }
}

/* MUFU.RCP without the IEEE fix-up sequence of __frcp_rn (1 instruction instead of ~8; 1 ulp) */
#ifdef ALVRL_IEEE_RCP
__device__ __forceinline__ float f_rcp(float x) { return __frcp_rn(x); }
#else
__device__ __forceinline__ float f_rcp(float x) { return alvrl_occ_rcp(x); }
#endif

/*
 * Any-hit query for small scenes (<= 32 leaves): a flat sweep over the leaf boxes -- a uniform loop, every lane reads
 * the same box (broadcast LDS), no inner nodes, no dependent control flow -- builds a per-lane bit mask of entered
 * leaves; only those leaves' triangles are tested.  Called by all lanes of the warp together (need = false: no ray).
 */
__device__ __forceinline__ bool occluded_flat(const BvhSmem &sb, uint32_t numLeaves, const F3 &o, const F3 &d, float tmin, float tmax, bool need) {
    const F3 inv = f3(f_rcp(d.x), f_rcp(d.y), f_rcp(d.z));
    const F3 oi = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    const float hi_t = tmax * 1.00001f;
    uint32_t mask = 0;
#pragma unroll 2
    for (uint32_t i = 0; i < numLeaves; i++) {
        const float4 lo = sb.leaves[2 * i], hi = sb.leaves[2 * i + 1];
        const float tx1 = fmaf(lo.x, inv.x, oi.x), tx2 = fmaf(hi.x, inv.x, oi.x);
        const float ty1 = fmaf(lo.y, inv.y, oi.y), ty2 = fmaf(hi.y, inv.y, oi.y);
        const float tz1 = fmaf(lo.z, inv.z, oi.z), tz2 = fmaf(hi.z, inv.z, oi.z);
        const float tn = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), tmin));
        const float tf = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), hi_t));
        mask |= (tn <= tf ? 1u : 0u) << i;
    }
    if (!need || !(tmax > tmin)) mask = 0;
    while (mask) {
        const uint32_t i = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t lf = __float_as_uint(sb.leaves[2 * i + 1].w);
        const uint32_t first = lf >> 4, cnt = lf & 15u;
        for (uint32_t k = 0; k < cnt; k++) {
            const float4 p = sb.tris[3 * (first + k)];
            const float den = p.x * d.x + p.y * d.y + p.z * d.z;
            const float num = p.w - (p.x * o.x + p.y * o.y + p.z * o.z);
            const float t = __fdividef(num, den);
            if (!(t >= tmin && t <= tmax)) continue;
            const float4 q = sb.tris[3 * (first + k) + 1], w = sb.tris[3 * (first + k) + 2];
            const F3 Pt = f3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
            const float u = q.x * Pt.x + q.y * Pt.y + q.z * Pt.z + q.w;
            const float v = w.x * Pt.x + w.y * Pt.y + w.z * Pt.z + w.w;
            if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) return true;
        }
    }
    return false;
}

/* what a pair still has to test of the compiled occluder set (warp-uniform; occ_query.h "pair-level culling") */
struct PairCull { uint32_t boxVV, boxVS, planes; };

/* the camera segment's side of every slab face / plane: hull = {E, Usurf} (vol->vol), surf = {Usurf} (vol->surf).  A lane
 * without a segment culls everything, so that it never keeps a test alive for its warp */
struct SegSides { uint32_t slabHull, slabSurf, planes; };
__device__ __forceinline__ SegSides seg_sides(const OccDev &oc, const SegRec &seg, bool active) {
    SegSides s;
    if (!active) { s.slabHull = s.slabSurf = s.planes = 0xffffffffu; return s; }
    const float m = oc.cullMargin;
    s.slabHull = occ_slab_sides(oc, seg.o.x, seg.o.y, seg.o.z, seg.p.x, seg.p.y, seg.p.z, m);
    s.slabSurf = occ_slab_sides(oc, seg.p.x, seg.p.y, seg.p.z, seg.p.x, seg.p.y, seg.p.z, m);
    s.planes = occ_plane_sides(oc, seg.o.x, seg.o.y, seg.o.z, seg.p.x, seg.p.y, seg.p.z, m, true);
    return s;
}
/* called by all lanes of the warp together */
__device__ __forceinline__ PairCull pair_cull(const OccDev &oc, const SegSides &sg, const float4 vDir, const float4 vPow) {
    const uint32_t vPl = __float_as_uint(vDir.w), vSl = __float_as_uint(vPow.w);
    const uint32_t nb = oc.numBoxes;
    const uint32_t boxAll = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u), planeAll = oc.numPlanes >= 32 ? 0xffffffffu : ((1u << oc.numPlanes) - 1u);
    uint32_t pc = sg.planes & vPl;
    pc = (pc | (pc >> 16)) & 0xffffu;
    PairCull c;
    c.planes = planeAll & ~__reduce_and_sync(0xffffffffu, pc);
    c.boxVV = boxAll & ~__reduce_and_sync(0xffffffffu, occ_boxes_culled(sg.slabHull & vSl, nb));
    c.boxVS = boxAll & ~__reduce_and_sync(0xffffffffu, occ_boxes_culled(sg.slabSurf & vSl, nb));
    return c;
}

template <int SMALL>
__device__ __forceinline__ bool occluded_fast(const TransportParams &P, const BvhSmem *sb, const F3 &p1, bool onSurf, const F3 &dir, float remaining, bool need,
                                              uint32_t boxActive = 0xffffffffu, uint32_t planeActive = 0xffffffffu) {
    /* adaptive epsilon of the shadow-ray overload, skdtree.cpp:154-157 */
    const float mint = onSurf ? ALVRL_EPSILON * fmaxf(fmaxf(fabsf(p1.x), fabsf(p1.y)), fabsf(p1.z)) : 0.0f;
    if (SMALL == 2) return occ_query(P.occ, sb->tris, p1.x, p1.y, p1.z, dir.x, dir.y, dir.z, mint, remaining, need, boxActive, planeActive);
    if (SMALL == 1) return occluded_flat(*sb, P.scene.numLeaves, p1, dir, mint, remaining, need);
    /* all lanes of the warp call these together */
    if (P.scene.nodes4) return bvh4_occluded_warp(P.scene, p1, dir, mint, remaining, need);
    return bvh_occluded_fast_warp(P.scene, p1, dir, mint, remaining, need);
}

/* diagnostic variants (tools/build_variant.sh -DALVRL_DIAG_*): accurate library functions in place of the MUFU forms, to
 * attribute the deviation from the oracle to its sources */
#ifdef ALVRL_DIAG_EXP
__device__ __forceinline__ float f_exp(float x) { return expf(x); }
#else
__device__ __forceinline__ float f_exp(float x) { return __expf(x); }
#endif
#ifdef ALVRL_DIAG_SINH
__device__ __forceinline__ float f_sinh(float a) { return sinhf(a); }
__device__ __forceinline__ float f_asinh(float x) { return asinhf(x); }
#elif defined(ALVRL_SINH_MUFU_ONLY)
__device__ __forceinline__ float f_sinh(float a) { const float e = __expf(a); return 0.5f * (e - f_rcp(e)); }
__device__ __forceinline__ float f_asinh(float x) { const float a = fabsf(x); return copysignf(__logf(a + sqrtf(fmaf(a, a, 1.0f))), x); }
#else
/* sinh / asinh through MUFU (ex2, lg2) away from zero, odd polynomials near zero: (e^a - e^-a) / 2 and log(a + sqrt(a^2 + 1))
 * lose their leading digits there (relative error eps / |a|), and V is sampled AROUND the closest point of the VRL, i.e.
 * around a = 0; measured on C2 (tools/probe_parity.py) the MUFU-only forms put 2.6e-4 of the R entries beyond 1e-4 */
__device__ __forceinline__ float f_sinh(float a) {
    const float e = __expf(a);
    const float big = 0.5f * (e - f_rcp(e));
    const float a2 = a * a;
    const float small = a * fmaf(a2, fmaf(a2, fmaf(a2, 1.0f / 5040.0f, 1.0f / 120.0f), 1.0f / 6.0f), 1.0f);   /* |a| < 0.5: 3e-9 relative */
    return fabsf(a) < 0.5f ? small : big;
}
__device__ __forceinline__ float f_asinh(float x) {
    const float a = fabsf(x), a2 = a * a;
    const float big = __logf(a + sqrtf(fmaf(a, a, 1.0f)));
    const float small = a * fmaf(a2, fmaf(a2, fmaf(a2, -15.0f / 336.0f, 3.0f / 40.0f), -1.0f / 6.0f), 1.0f);   /* a < 0.125: 4e-9 relative */
    return copysignf(a < 0.125f ? small : big, x);
}
#endif
#ifdef ALVRL_DIAG_TAN
__device__ __forceinline__ float f_tan(float x) { return tanf(x); }
#else
__device__ __forceinline__ float f_tan(float x) { return __tanf(x); }
#endif
#ifdef ALVRL_DIAG_ATAN
__device__ __forceinline__ float f_atan(float x) { return (float) atan((double) x); }
#elif defined(ALVRL_ATAN_LIB)
__device__ __forceinline__ float f_atan(float x) { return atanf(x); }
#else
/* atan by one MUFU reciprocal (|x| > 1: pi/2 - atan(1/|x|)) and an 8-term odd minimax polynomial on [0, 1]: absolute error
 * 1.1e-7 (the library atanf: ~28 instructions for the same error class) */
__device__ __forceinline__ float f_atan(float x) {
    const float a = fabsf(x);
    const bool big = a > 1.0f;
    const float r = big ? f_rcp(a) : a, q = r * r;
    float p = -0.004105130676180124f;
    p = fmaf(p, q, 0.022052252665162086f); p = fmaf(p, q, -0.05619571730494499f); p = fmaf(p, q, 0.09663795679807663f);
    p = fmaf(p, q, -0.1391744166612625f); p = fmaf(p, q, 0.19948409497737885f); p = fmaf(p, q, -0.33330032229423523f);
    p = fmaf(p, q, 0.9999994039535522f);
    p *= r;
    return copysignf(big ? 1.57079632679489662f - p : p, x);
}
#endif
#ifdef ALVRL_DIAG_DIV
#define f_div(a, b) ((a) / (b))
#else
#define f_div(a, b) __fdividef(a, b)
#endif
__device__ __forceinline__ float f_len(const F3 &a, float &l2) { l2 = len2(a); return l2 * rsqrtf(fmaxf(l2, 1e-38f)); }

/* homogeneous media only (MED 0: RGB sigma_t, MED 2: grey sigma_t) */
template <int MED, bool WANT_RGB, bool WANT_STAT, int SMALL, bool WEIGHTED>
__device__ __forceinline__ void integrate_pair_fast(const TransportParams &P, const BvhSmem *sb, const SegRec &seg,
                                                    const float4 vS, const float4 vE, const float4 vDir, const float4 vPow, Rng &rng,
                                                    float rgb[3], float &outMean, float &outVar, const bool laneOn, const PairCull cull) {
    const F3 S = f3(vS), End = f3(vE), SV = f3(vDir);
    const float vlen = vS.w;
    const F3 E = f3(seg.o), EU = f3(seg.d), Usurf = f3(seg.p);
    const float edist = seg.o.w;
    const int Nvv = P.Nvv, Nvs = P.Nvs;
    const MediumDev &M = P.medium;
    if (WANT_RGB) rgb[0] = rgb[1] = rgb[2] = 0;
    outMean = 0; outVar = 0;
    const float sT0 = M.sigmaT[0], sT1 = M.sigmaT[1], sT2 = M.sigmaT[2];
    const float sTmin = fminf(sT0, fminf(sT1, sT2));
    const float cutoff = 46.0517f;     /* -ln(1e-20): Medium::eval zeroes a transmittance whose max is below 1e-20 (homogeneous.cpp:394-395) */
    const float wS = M.samplingWeight, wF = 1.0f - M.samplingWeight;
    const float lw0 = 0.212671f, lw1 = 0.715160f, lw2 = 0.072169f;

    /* ---- volume to volume (646-703) ---- */
    if (Nvv > 0) {
        /* power * sigma_s(V) * sigma_s(U): constant per pair in a homogeneous medium */
        /* WEIGHTED: the segment of a specular chain carries LiInternal's weight (p.w, n.w, albedo.w); camera segments weigh 1 and
         * their kernels do not spend the registers */
        float k0 = vPow.x * M.sigmaS[0] * M.sigmaS[0], k1 = vPow.y * M.sigmaS[1] * M.sigmaS[1], k2 = vPow.z * M.sigmaS[2] * M.sigmaS[2];
        if (WEIGHTED) { k0 *= seg.p.w; k1 *= seg.n.w; k2 *= seg.albedo.w; }
        float cosTheta, sinTheta;
        cos_sin_theta(f3(seg.dn), SV, cosTheta, sinTheta);
        const bool parallel = sinTheta < ALVRL_EPSILON;
        float h = 0, A0 = 0, dA = 0, dVhS = 0, rSin = 0, pdfVc = 0;
        if (!parallel) {
            F3 Vh;
            h = closest_points(E, Usurf, S, End, Vh);
            float l2;
            dVhS = f_len(Vh - S, l2);
            const float V1c = f_len(Vh - End, l2);
            rSin = f_rcp(sinTheta);
            const float sh = f_div(sinTheta, h);
            A0 = f_asinh(-dVhS * sh);
            dA = f_asinh(V1c * sh) - A0;
            pdfVc = f_div(sinTheta, dA);                              /* 1 / denom, denom = (A1 - A0) / sinTheta */
        }
        const float invNvv = f_rcp((float) Nvv);
        float mean = 0, M2 = 0;
        for (int k = 0; k < Nvv; k++) {
            const float u1 = rng.next();
            F3 V; float pdf, dSV;                                           /* dSV = |V - S|: V = S + s SV with a unit SV */
            if (parallel) {
                V = S + u1 * (End - S);
                pdf = f_rcp(vlen);
                dSV = u1 * vlen;
            } else {
                const float nv = h * f_sinh(fmaf(u1, dA, A0)) * rSin;
                pdf = rsqrtf(fmaf(nv * nv, sinTheta * sinTheta, h * h)) * pdfVc;
                V = S + (nv + dVhS) * SV;
                dSV = fabsf(nv + dVhS);
            }
            const float u2 = rng.next();
            /* KullaSampling along the eye segment w.r.t. V (889-914), signed-angle form; dir = EU, A = E, |AB| = edist */
            const float dotPr = dot(EU, V - E);
            const F3 I = E + dotPr * EU;
            float l2;
            const float Dis = f_len(V - I, l2);
            const float rDis = f_rcp(Dis);
            const float th_a = f_atan(-dotPr * rDis), th_b = f_atan((edist - dotPr) * rDis);
            const float t = Dis * f_tan(fmaf(u2, th_b - th_a, th_a));
            pdf *= f_div(Dis, (th_b - th_a) * fmaf(t, t, l2));
            const F3 U = I + t * EU;
            const F3 UV = U - V;
            float d2;
            const float dUV = f_len(UV, d2);
            const F3 VU = UV * f_rcp(dUV);
            const float dEU = fabsf(dotPr + t);                             /* |U - E|, U = E + (dotPr + t) EU */
            const bool ok = laneOn && d2 > 0.0f && dEU * sTmin <= cutoff && dSV * sTmin <= cutoff;
            const bool occ = occluded_fast<SMALL>(P, sb, U, false, -VU, dUV, ok, cull.boxVV, cull.planes);
            float lum = 0;
            if (ok && !occ) {
                const float path = dSV + dUV + dEU;
                float T0, T1, T2, pf;
                if (MED == 2) {
                    T0 = T1 = T2 = f_exp(-sT0 * path);
                    pf = fmaf(f_exp(-sT0 * dSV), wS, wF);
                } else {
                    T0 = f_exp(-sT0 * path); T1 = f_exp(-sT1 * path); T2 = f_exp(-sT2 * path);
                    pf = fmaf((f_exp(-sT0 * dSV) + f_exp(-sT1 * dSV) + f_exp(-sT2 * dSV)) * (1.0f / 3.0f), wS, wF);
                }
                float common = f_div(1.0f, pdf * d2);
                if (P.shortVrls) common = f_div(common, pf);
                common *= phase_eval(M, dot(VU, EU)) * phase_eval(M, -dot(SV, VU));
                const float c0 = k0 * T0 * common, c1 = k1 * T1 * common, c2 = k2 * T2 * common;
                /* isValid(), 686: every channel finite and >= 0.  k and T are finite and >= 0, so that is: common is a finite
                 * non-negative number and the largest channel did not overflow */
                if (common >= 0.0f && fmaxf(common, fmaxf(c0, fmaxf(c1, c2))) < INFINITY) {
                    if (WANT_RGB) { rgb[0] = fmaf(c0, invNvv, rgb[0]); rgb[1] = fmaf(c1, invNvv, rgb[1]); rgb[2] = fmaf(c2, invNvv, rgb[2]); }
                    lum = c0 * lw0 + c1 * lw1 + c2 * lw2;
                }
            }
            if (WANT_STAT) {                                                 /* 693-699 */
                const float delta = lum - mean;
                mean += f_div(delta, (float) (k + 1));
                M2 = fmaf(delta, lum - mean, M2);
            }
        }
        if (WANT_STAT) { outMean += mean; outVar += f_div(M2, (float) ((Nvv - 1) * Nvv)); }
    }

    /* ---- volume to surface (706-782) ---- */
    if (Nvs > 0) {
        const float tE0 = seg.tE.x, tE1 = seg.tE.y, tE2 = seg.tE.z;
        const uint32_t flags = __float_as_uint(seg.dn.w);
        const bool surf = laneOn && !(tE0 == 0 && tE1 == 0 && tE2 == 0) && (flags & SEG_SMOOTH);
        float mean = 0, M2 = 0;
        /* lanes without a vol->surf term (727) still walk the loop -- drawing nothing, contributing nothing -- so that the
         * warp stays converged through the visibility queries (every strategy is called by the 32 lanes together) */
        if (__any_sync(0xffffffffu, surf)) {
            /* per-pair part of KullaSampling(A = S, B = End, D = Usurf) */
            const float dotPr = dot(SV, Usurf - S);
            const F3 I = S + dotPr * SV;
            float l2;
            const float Dis = f_len(Usurf - I, l2);
            const float rDis = f_rcp(Dis);
            const float th_a = f_atan(-dotPr * rDis), th_b = f_atan((vlen - dotPr) * rDis);
            const float pdfC = f_div(Dis, th_b - th_a);
            const float invNvs = f_rcp((float) Nvs);
            const F3 nrm = f3(seg.n);
            const bool frontI = seg.d.w > 0;
            float k0 = vPow.x * M.sigmaS[0] * seg.albedo.x * tE0, k1 = vPow.y * M.sigmaS[1] * seg.albedo.y * tE1,
                  k2 = vPow.z * M.sigmaS[2] * seg.albedo.z * tE2;
            if (WEIGHTED) { k0 *= seg.p.w; k1 *= seg.n.w; k2 *= seg.albedo.w; }
            for (int k = 0; k < Nvs; k++) {
                const float u = surf ? rng.next() : 0.5f;
                const float t = Dis * f_tan(fmaf(u, th_b - th_a, th_a));
                const float pdf = f_div(pdfC, fmaf(t, t, l2));
                const float sv = dotPr + t;                                  /* V = S + sv * SV */
                const F3 V = S + sv * SV;
                const F3 UV = Usurf - V;
                float d2;
                const float dUV = f_len(UV, d2);
                const F3 VU = UV * f_rcp(dUV);
                const float dSV = fabsf(sv);
                const float cosWo = -dot(VU, nrm);                           /* diffuse.cpp:110-118 */
                const bool ok = surf && d2 > 0.0f && dSV * sTmin <= cutoff && frontI && cosWo > 0;
                const bool occ = occluded_fast<SMALL>(P, sb, Usurf, true, -VU, dUV, ok, cull.boxVS, cull.planes);
                float lum = 0;
                if (ok && !occ) {
                    const float path = dSV + dUV;
                    float T0, T1, T2, pf;
                    if (MED == 2) {
                        T0 = T1 = T2 = f_exp(-sT0 * path);
                        pf = fmaf(f_exp(-sT0 * dSV), wS, wF);
                    } else {
                        T0 = f_exp(-sT0 * path); T1 = f_exp(-sT1 * path); T2 = f_exp(-sT2 * path);
                        pf = fmaf((f_exp(-sT0 * dSV) + f_exp(-sT1 * dSV) + f_exp(-sT2 * dSV)) * (1.0f / 3.0f), wS, wF);
                    }
                    float common = f_div(ALVRL_INV_PI * cosWo, pdf * d2);
                    if (P.shortVrls) common = f_div(common, pf);
                    common *= phase_eval(M, -dot(SV, VU));
                    const float c0 = k0 * T0 * common, c1 = k1 * T1 * common, c2 = k2 * T2 * common;
                    if (common >= 0.0f && fmaxf(common, fmaxf(c0, fmaxf(c1, c2))) < INFINITY) {
                        if (WANT_RGB) { rgb[0] = fmaf(c0, invNvs, rgb[0]); rgb[1] = fmaf(c1, invNvs, rgb[1]); rgb[2] = fmaf(c2, invNvs, rgb[2]); }
                        lum = c0 * lw0 + c1 * lw1 + c2 * lw2;
                    }
                }
                if (WANT_STAT) {
                    const float delta = lum - mean;
                    mean += f_div(delta, (float) (k + 1));
                    M2 = fmaf(delta, lum - mean, M2);
                }
            }
        }
        if (WANT_STAT) { outMean += mean; outVar += f_div(M2, (float) ((Nvs - 1) * Nvs)); }
    }
}

#include "transport_grid_fast.cuh"
