/*
 * transport_grid_fast.cuh -- speed flavour of integrateVRL in a heterogeneous (grid) medium, method = simpson
 * (heterogeneous.cpp:301-376, 546-548, 665-691; gridvolume.cpp:337-388).  Included by transport_fast_impl.cuh.
 *
 * What a contribution costs here is the ray marching: every sample evaluates three optical depths (E->U, S->V, U->V; two in
 * the vol->surf term), each a composite Simpson sum over ceil(length / stepSize) trilinear lookups -- about 10^3 per unit
 * length of a 512^3 grid.  A thread marching its own ray (the strict flavour, and the first version of this one: ncu counted
 * 9 active lanes per instruction, every lane of a warp loading from its own cache lines, profiles/r2_c3_march_v0.txt) leaves
 * the warp waiting for its longest ray.  Here the 32 lanes of a warp march ONE ray together: lane l evaluates the Simpson nodes
 * l, l + 32, ... and a shuffle reduction adds the weighted densities.  Consecutive nodes are half a voxel apart, so one load
 * instruction of the warp touches a few neighbouring sectors instead of 32 unrelated ones, and the lanes stay busy whatever
 * the lengths of the 32 rays.  The quadrature is the reference's -- the same nodes (the reference's running fp32 positions, reproduced in
 * closed form for the long rays: CoordRun below), the same weights, the same early exit: the running sum of non-negative densities is
 * monotone, so "passed the threshold at some node" == "the complete sum passes it" -- summed in a different order (relative
 * 1e-6 on the optical depth); tests/test_c2_parity_gpu.py::test_other_config_shapes_R_vs_oracle_1e4[C3-fast] holds it to the
 * oracle at 1e-4 per R entry.
 */
#pragma once

#ifndef ALVRL_DRIFT_MIN_STEPS
#define ALVRL_DRIFT_MIN_STEPS 320u      /* rays with fewer Simpson nodes use o + i * inc: their drift is below 10^-5 */
#endif

/* GridDataSource::lookupFloat in grid coordinates (g = gsc * p + gtr already applied), FMA lerps, one 64-bit base address */
__device__ __forceinline__ float grid_lookup_g(const MediumDev &m, float gx, float gy, float gz) {
    const float fx0 = floorf(gx), fy0 = floorf(gy), fz0 = floorf(gz);
    const int x1 = (int) fx0, y1 = (int) fy0, z1 = (int) fz0;
    /* x1 < 0 || x1 + 1 >= res, one unsigned compare per axis */
    if ((unsigned) x1 >= (unsigned) (m.res[0] - 1) || (unsigned) y1 >= (unsigned) (m.res[1] - 1) || (unsigned) z1 >= (unsigned) (m.res[2] - 1)) return 0.0f;
    const float fx = gx - fx0, fy = gy - fy0, fz = gz - fz0;
    const uint32_t rx = (uint32_t) m.res[0], sxy = rx * (uint32_t) m.res[1];
    const float *q = m.density + ((size_t) z1 * sxy + (uint32_t) y1 * rx + (uint32_t) x1);
    const float d000 = __ldg(q), d001 = __ldg(q + 1), d010 = __ldg(q + rx), d011 = __ldg(q + rx + 1);
    const float d100 = __ldg(q + sxy), d101 = __ldg(q + sxy + 1), d110 = __ldg(q + sxy + rx), d111 = __ldg(q + sxy + rx + 1);
    const float a0 = fmaf(fx, d001 - d000, d000), a1 = fmaf(fx, d011 - d010, d010);
    const float b0 = fmaf(fx, d101 - d100, d100), b1 = fmaf(fx, d111 - d110, d110);
    const float a = fmaf(fy, a1 - a0, a0), b = fmaf(fy, b1 - b0, b0);
    return fmaf(fz, b - a, a);
}
__device__ __forceinline__ float grid_lookup_fast(const MediumDev &m, const F3 &p) {
    return grid_lookup_g(m, fmaf(m.gsc[0], p.x, m.gtr[0]), fmaf(m.gsc[1], p.y, m.gtr[1]), fmaf(m.gsc[2], p.z, m.gtr[2]));
}

/*
 * The reference marches with a RUNNING position, p += increment in fp32 (heterogeneous.cpp:343-369).  With ~10^3 steps of
 * ~10^-3 added to coordinates of ~0.5, each add drops the bits of the increment below ulp(p): the positions drift from
 * o + i * inc by up to n * ulp / 2 (1.5 % of a voxel at 512^3), systematically, and the optical depth moves by a relative
 * 10^-5 .. 6 * 10^-4 -- more than the tolerance of an R entry on fine grids (measured with exact positions o + i * inc: 4.7 % of
 * the entries beyond 1e-4 at 384^3, 2.5 % in the 512^3 bench sample; with the running positions below: 0.36 %, the same
 * near-singular pairs at every resolution).  A lane that evaluates node i directly therefore has to know the reference's
 * p_i.  It can: while a coordinate stays in one binade, p is a multiple of u = ulp(p) and fl(p + inc) = p + k u with
 * k = round(inc / u) the same at every step (a tie rounds to even: after one step the mantissa is even and k is the even
 * neighbour), so p_j = p_s + (j - s) k u exactly; the add that leaves the binade is done as a true fp32 add.  CoordRun holds
 * the current run of one coordinate; tools/micro/running_sum.py checks the construction against sequential float32
 * accumulation (12 007 random rays, none differ).
 */
struct CoordRun {
    float pS, u, inc; int iS, iE, k;          /* positions p_j = pS + ((j - iS) * k) * u for j in [iS, iE] */
    __device__ __forceinline__ void start(float p, int i, int n) {
        pS = p; iS = i;
        const uint32_t bits = __float_as_uint(p), eb = (bits >> 23) & 255u;
        if (eb < 67u) { k = 0; u = 0.0f; iE = i; return; }                    /* |p| < 2^-60: the next add gives the increment itself */
        u = __uint_as_float((eb - 23u) << 23);                                 /* ulp of the binade of |p| */
        const float q = inc * __uint_as_float((277u - eb) << 23);              /* inc / u, exact */
        if (!(fabsf(q) < 8388608.0f)) { k = 0; iE = i; return; }               /* the step leaves the binade at once */
        const float fl = floorf(q), frac = q - fl;
        int kk;
        if (frac == 0.5f) {
            if (bits & 1u) { k = 0; iE = i; return; }                           /* odd mantissa: one true add first */
            kk = (int) fl; kk += kk & 1;                                         /* the even one of {fl, fl + 1} */
        } else kk = (int) fl + (frac > 0.5f ? 1 : 0);
        k = kk;
        const int M = (int) ((bits & 0x7fffffu) | 0x800000u);                   /* |p| / u */
        const int along = (bits >> 31) ? -kk : kk;                               /* the step along |p| */
        int m;
        if (along > 0) m = (0xffffff - M) / along; else if (along < 0) m = (M - 0x800000) / (-along); else m = n;
        iE = i + min(m, n - i);
    }
    __device__ __forceinline__ float at(int j) {                                 /* j >= iS, called with non-decreasing j */
        while (j > iE) {
            const float pe = pS + (float) ((iE - iS) * k) * u;                   /* exact */
            const int ni = iE + 1;
            start(__fadd_rn(pe, inc), ni, 0x3fffffff);
        }
        return pS + (float) ((j - iS) * k) * u;
    }
};

/*
 * Optical depth of the segment o + t d, t in [0, dist], for every lane's ray; called by the 32 lanes of a warp together
 * (need = false: this lane has no ray).  The set-up of a ray (clip to the density AABB, degenerate-segment test, step count)
 * follows integrateDensity 301-325 with the reference's operations, evaluated redundantly by all lanes on shuffled values.
 */
__device__ __forceinline__ float warp_grid_optical_depth(const MediumDev &m, const F3 &o_, const F3 &d_, float dist_, bool need) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t todo = __ballot_sync(0xffffffffu, need);
    float mine = 0.0f;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const F3 o = f3(__shfl_sync(0xffffffffu, o_.x, src), __shfl_sync(0xffffffffu, o_.y, src), __shfl_sync(0xffffffffu, o_.z, src));
        const F3 d = f3(__shfl_sync(0xffffffffu, d_.x, src), __shfl_sync(0xffffffffu, d_.y, src), __shfl_sync(0xffffffffu, d_.z, src));
        const float dist = __shfl_sync(0xffffffffu, dist_, src);
        float tau = 0.0f;
        const F3 dRcp = f3(xdiv(1.0f, d.x), xdiv(1.0f, d.y), xdiv(1.0f, d.z));
        float mint, maxt;
        if (aabb_clip(m.bmin, m.bmax, o, d, dRcp, mint, maxt)) {
            mint = fmaxf(mint, 0.0f);
            maxt = fminf(maxt, dist);
            const float length = xsub(maxt, mint);
            const F3 p0 = xadd3(o, xscale(d, mint)), pLast = xadd3(o, xscale(d, maxt));
            float maxComp = fmaxf(fmaxf(fabsf(p0.x), fabsf(pLast.x)), fmaxf(fmaxf(fabsf(p0.y), fabsf(pLast.y)), fmaxf(fabsf(p0.z), fabsf(pLast.z))));
            if (!(length < xmul(1e-6f, maxComp))) {
                uint32_t nSteps = (uint32_t) ceilf(xdiv(length, m.stepSize));
                nSteps += nSteps & 1u;
                const float stepSz = xdiv(length, (float) nSteps);
                float sum = 0.0f;
                if (nSteps >= ALVRL_DRIFT_MIN_STEPS) {
                    /* node i sits at the reference's running position p_i (CoordRun); node nSteps is ray(maxt) itself (319, 332-333) */
                    CoordRun rx, ry, rz;
                    rx.inc = xmul(d.x, stepSz); ry.inc = xmul(d.y, stepSz); rz.inc = xmul(d.z, stepSz);
                    rx.start(p0.x, 0, (int) nSteps); ry.start(p0.y, 0, (int) nSteps); rz.start(p0.z, 0, (int) nSteps);
                    for (uint32_t i = lane; i < nSteps; i += 32u) {
                        const float w = i == 0u ? 1.0f : ((i & 1u) ? 4.0f : 2.0f);
                        const float px = rx.at((int) i), py = ry.at((int) i), pz = rz.at((int) i);
                        sum = fmaf(w, grid_lookup_g(m, xadd(xmul(m.gsc[0], px), m.gtr[0]), xadd(xmul(m.gsc[1], py), m.gtr[1]), xadd(xmul(m.gsc[2], pz), m.gtr[2])), sum);
                    }
                } else {
                    /* a short ray: its running positions stay within n * ulp / 2 < 10^-5 of o + i * inc (a 200th of a voxel at 512^3) */
                    const float g0x = fmaf(m.gsc[0], p0.x, m.gtr[0]), g0y = fmaf(m.gsc[1], p0.y, m.gtr[1]), g0z = fmaf(m.gsc[2], p0.z, m.gtr[2]);
                    const float gix = m.gsc[0] * d.x * stepSz, giy = m.gsc[1] * d.y * stepSz, giz = m.gsc[2] * d.z * stepSz;
#pragma unroll 2
                    for (uint32_t i = lane; i < nSteps; i += 32u) {
                        const float fi = (float) i;
                        const float w = i == 0u ? 1.0f : ((i & 1u) ? 4.0f : 2.0f);
                        sum = fmaf(w, grid_lookup_g(m, fmaf(fi, gix, g0x), fmaf(fi, giy, g0y), fmaf(fi, giz, g0z)), sum);
                    }
                }
                if (lane == (nSteps & 31u)) sum += grid_lookup_fast(m, pLast);
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
                /* HETVOL_EARLY_EXIT (31, 336-340, 353-360): -log(Epsilon) of optical depth reached -> +infinity */
                const float stopValue = __fdividef(9.21034049987793f * 3.0f, stepSz * m.scale);
                tau = sum > stopValue ? INFINITY : sum * m.scale * stepSz * (1.0f / 3.0f);
            }
        }
        if ((int) lane == src) mine = tau;
    }
    return mine;
}

/* integrateVRL in a grid medium; called by all lanes of the warp together (laneOn = false: no segment in this lane) */
template <bool WANT_RGB, bool WANT_STAT, int SMALL>
__device__ __forceinline__ void integrate_pair_grid_fast(const TransportParams &P, const BvhSmem *sb, const SegRec &seg,
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
    const float lw0 = 0.212671f, lw1 = 0.715160f, lw2 = 0.072169f;
    /* LiInternal's weight of a specular-chain segment: inside the estimate for the rows of R only (the render pass multiplies the
     * pixel's sum afterwards) */
    const float wg0 = WANT_STAT ? seg.p.w : 1.0f, wg1 = WANT_STAT ? seg.n.w : 1.0f, wg2 = WANT_STAT ? seg.albedo.w : 1.0f;

    /* ---- volume to volume (646-703) ---- */
    if (Nvv > 0) {
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
            pdfVc = f_div(sinTheta, dA);
        }
        const float invNvv = f_rcp((float) Nvv);
        float mean = 0, M2 = 0;
        for (int k = 0; k < Nvv; k++) {
            const float u1 = rng.next();
            F3 V; float pdf, dSV;
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
            const float dEU = fabsf(dotPr + t);
            const bool ok = laneOn && d2 > 0.0f && d2 < INFINITY;
            const bool occ = occluded_fast<SMALL>(P, sb, U, false, -VU, dUV, ok, cull.boxVV, cull.planes);
            const bool go = ok && !occ;
            /* evalTransmittance(U -> V) (scene.cpp:619-679), Medium::eval along E -> U and S -> V (665-691) */
            const float tauUV = warp_grid_optical_depth(M, U, -VU, dUV, go);
            const float tauEU = warp_grid_optical_depth(M, E, EU, dEU, go);
            const float tauSV = warp_grid_optical_depth(M, S, SV, dSV, go);
            float lum = 0;
            if (go && tauUV < INFINITY) {                                        /* transmittanceUV.isZero() -> continue (663-665) */
                const float densU = grid_lookup_fast(M, E + dEU * EU) * M.scale, densV = grid_lookup_fast(M, S + dSV * SV) * M.scale;
                const float T = f_exp(-(tauUV + tauEU + tauSV));
                const float pf = f_exp(-tauSV);                                  /* pdfFailure = expVal (690) */
                float common = f_div(densU * densV, pdf * d2) * T;
                if (P.shortVrls) common = f_div(common, pf);                     /* 0 / 0 = NaN past the early exit: invalid, dropped */
                common *= phase_eval(M, dot(VU, EU)) * phase_eval(M, -dot(SV, VU));
                const float c0 = vPow.x * M.albedo[0] * M.albedo[0] * common * wg0, c1 = vPow.y * M.albedo[1] * M.albedo[1] * common * wg1,
                            c2 = vPow.z * M.albedo[2] * M.albedo[2] * common * wg2;
                if (common >= 0.0f && fmaxf(common, fmaxf(c0, fmaxf(c1, c2))) < INFINITY) {
                    if (WANT_RGB) { rgb[0] = fmaf(c0, invNvv, rgb[0]); rgb[1] = fmaf(c1, invNvv, rgb[1]); rgb[2] = fmaf(c2, invNvv, rgb[2]); }
                    lum = c0 * lw0 + c1 * lw1 + c2 * lw2;
                }
            }
            if (WANT_STAT) {
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
        if (__any_sync(0xffffffffu, surf)) {
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
            /* vrlMedium->getSigmaS() is the base sigma_s of the grid medium (quirk B2), 748 */
            const float k0 = vPow.x * M.sigmaS[0] * seg.albedo.x * tE0 * wg0, k1 = vPow.y * M.sigmaS[1] * seg.albedo.y * tE1 * wg1,
                        k2 = vPow.z * M.sigmaS[2] * seg.albedo.z * tE2 * wg2;
            for (int k = 0; k < Nvs; k++) {
                const float u = surf ? rng.next() : 0.5f;
                const float t = Dis * f_tan(fmaf(u, th_b - th_a, th_a));
                const float pdf = f_div(pdfC, fmaf(t, t, l2));
                const float sv = dotPr + t;
                const F3 V = S + sv * SV;
                const F3 UV = Usurf - V;
                float d2;
                const float dUV = f_len(UV, d2);
                const F3 VU = UV * f_rcp(dUV);
                const float dSV = fabsf(sv);
                const float cosWo = -dot(VU, nrm);
                const bool ok = surf && d2 > 0.0f && d2 < INFINITY;
                /* the reference evaluates the transmittance before the BSDF; a back-facing sample contributes zero either way */
                const bool lit = ok && frontI && cosWo > 0;
                const bool occ = occluded_fast<SMALL>(P, sb, Usurf, true, -VU, dUV, lit, cull.boxVS, cull.planes);
                const bool go = lit && !occ;
                const float tauUV = warp_grid_optical_depth(M, Usurf, -VU, dUV, go);
                const float tauSV = warp_grid_optical_depth(M, S, SV, dSV, go);
                float lum = 0;
                if (go) {
                    const float T = f_exp(-(tauUV + tauSV));
                    const float pf = f_exp(-tauSV);
                    float common = f_div(ALVRL_INV_PI * cosWo, pdf * d2) * T;
                    if (P.shortVrls) common = f_div(common, pf);
                    common *= phase_eval(M, -dot(SV, VU));
                    const float c0 = k0 * common, c1 = k1 * common, c2 = k2 * common;
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
