/*
 * chain.cu -- the specular chains of LiInternal (vrlIntegrator.cpp:445-511) on the device, in exact arithmetic.
 *
 * A camera segment that ends on a surface with delta components (smooth dielectric: reflection + transmission,
 * src/bsdfs/dielectric.cpp:335-387; smooth conductor: reflection, src/bsdfs/conductor.cpp:254-268) continues: for each
 * component, BSDF::sample(bRec, Point2(0.5)) gives a direction and a weight, a Russian roulette on the throughput
 * (initialSpecularThroughput, 0.98 from specularForcedRRdepth on) keeps or drops the branch, the medium follows the
 * surface's interior / exterior media (records.inl:81-86), and LiInternal recurses along the new ray.  Every segment of
 * that tree adds its VRL contributions, multiplied by its `weight`, to the pixel (render pass) or to the row of R
 * (getLiLuminanceVrlContributions, 514-526), looked up in the slice of the original camera ray.
 *
 * k_chain: thread = pixel.  The recursion becomes a stack of pending rays popped in the reference's order (component 0
 * and everything below it before component 1), so the segments of a pixel come out in the order LiInternal visits them.
 * Two passes of the same walk: count, then -- after an exclusive scan of the counts -- write.  The roulette draw of a
 * branch comes from the counter stream of (pixel, path code), path code = 2 * parent's + component, camera segment = 1;
 * the oracle (oracle_capi.cpp::chainBelow) draws the same way.  Same operations in the same order as the oracle
 * (-fmad=false): the segments are compared bit for bit (tests/test_chain.py).
 */
#include "dev_common.cuh"
#include "kernels.h"
#include "../../include/alvrl_rng.h"

namespace alvrl {

#define ALVRL_CHAIN_MAX_DEPTH 30

struct ChainRay { F3 o, d; float thr[3], w[3]; uint32_t code; int depth; int inMedium; };

/* fresnelDielectricExt, src/libcore/util.cpp:651-681 */
__device__ __forceinline__ float fresnel_dielectric_ext(float cosThetaI_, float &cosThetaT_, float eta) {
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
__device__ __forceinline__ float safe_sqrt_x(float v) { return xsqrt(fmaxf(0.0f, v)); }
/* fresnelConductorExact, util.cpp:739-761, one channel */
__device__ __forceinline__ float fresnel_conductor_exact(float cosThetaI, float eta, float k) {
    const float cosThetaI2 = xmul(cosThetaI, cosThetaI), sinThetaI2 = xsub(1.0f, cosThetaI2), sinThetaI4 = xmul(sinThetaI2, sinThetaI2);
    const float temp1 = xsub(xsub(xmul(eta, eta), xmul(k, k)), sinThetaI2);
    const float a2pb2 = safe_sqrt_x(xadd(xmul(temp1, temp1), xmul(xmul(xmul(xmul(k, k), eta), eta), 4.0f)));    /* k*k*eta*eta*4, left to right */
    const float a = safe_sqrt_x(xmul(xadd(a2pb2, temp1), 0.5f));
    const float term1 = xadd(a2pb2, cosThetaI2), term2 = xmul(a, xmul(2.0f, cosThetaI));
    const float Rs2 = xdiv(xsub(term1, term2), xadd(term1, term2));
    const float term3 = xadd(xmul(a2pb2, cosThetaI2), sinThetaI4), term4 = xmul(term2, sinThetaI2);
    const float Rp2 = xdiv(xmul(Rs2, xsub(term3, term4)), xadd(term3, term4));
    return xmul(0.5f, xadd(Rp2, Rs2));
}

/* the segment record of a chain ray that hit triangle prim at (t, u, v): what k_primary stores for a camera segment, plus the
 * weight of the segment (p.w, n.w, albedo.w) */
__device__ __forceinline__ SegRec chain_segment(const MediumDev &med, int haveMedium, const float4 *__restrict__ triVerts, const uint32_t *__restrict__ triMat,
                                                const float4 *__restrict__ matAlbedo, const uint32_t *__restrict__ matBits, const F3 &o, const F3 &d,
                                                uint32_t prim, float u, float v, const float w[3], bool inMedium, F3 &pOut) {
    SegRec s;
    const F3 p0 = f3(__ldg(&triVerts[3 * (size_t) prim])), p1 = f3(__ldg(&triVerts[3 * (size_t) prim + 1])), p2 = f3(__ldg(&triVerts[3 * (size_t) prim + 2]));
    const float b0 = xsub(xsub(1.0f, u), v);
    const F3 p = xadd3(xadd3(xscale(p0, b0), xscale(p1, u)), xscale(p2, v));          /* skdtree.h:362-363 */
    F3 fn = xcross(xsub3(p1, p0), xsub3(p2, p0));
    const float length = xlen(fn);
    if (!(fn.x == 0 && fn.y == 0 && fn.z == 0)) fn = xdivv(fn, length);
    const float wiz = xdot(f3(-d.x, -d.y, -d.z), fn);
    const float dist = xlen(xsub3(p, o));
    const uint32_t mat = triMat[prim];
    const uint32_t flags = (inMedium ? SEG_VALID : 0u) | ((matBits[mat] & ALVRL_BSDF_SMOOTH) ? SEG_SMOOTH : 0u) | ((matBits[mat] & ALVRL_BSDF_DELTA) ? SEG_DELTA : 0u);
    const F3 dn = xnormalize(d);
    s.o = make_float4(o.x, o.y, o.z, dist);
    s.d = make_float4(d.x, d.y, d.z, wiz);
    s.dn = make_float4(dn.x, dn.y, dn.z, __uint_as_float(flags));
    s.p = make_float4(p.x, p.y, p.z, w[0]);
    s.n = make_float4(fn.x, fn.y, fn.z, w[1]);
    const float4 al = matAlbedo[mat];
    s.albedo = make_float4(al.x, al.y, al.z, w[2]);
    s.tE = make_float4(0, 0, 0, 0);
    if (haveMedium && inMedium && dist != 0) {                                          /* vrlIntegrator.cpp:711-719 */
        float T[3];
        medium_transmittance_exact(med, o, d, dist, T);
        s.tE = make_float4(T[0], T[1], T[2], 0.0f);
    }
    pOut = p;
    return s;
}

/* one walk over the chain below a camera segment; WRITE = false counts, WRITE = true stores the segments at out[0 .. count) */
template <bool WRITE>
__device__ __forceinline__ uint32_t chain_walk(const SceneDev &sc, const MediumDev &med, int haveMedium, const float4 *__restrict__ triVerts,
                                               const uint32_t *__restrict__ triMat, const float4 *__restrict__ matAlbedo,
                                               const uint32_t *__restrict__ matBits, const float4 *__restrict__ matOptics,
                                               uint64_t seed, int specRRdepth, float initialThroughput, uint32_t pixel,
                                               F3 o, F3 d, float tHit, uint32_t prim, F3 hitP, SegRec *__restrict__ out, uint4 *__restrict__ meta) {
    ChainRay stack[ALVRL_CHAIN_MAX_DEPTH + 2];
    int sp = 0;
    uint32_t count = 0;
    /* the node being expanded: ray (o, d) that hit triangle prim at parameter tHit, point hitP */
    float thr[3] = {initialThroughput, initialThroughput, initialThroughput}, w[3] = {1.0f, 1.0f, 1.0f};
    uint32_t code = 1u; int depth = 1; int inMedium = 1;                                /* Li: newQuery(ESensorRay, sensor medium) */
    for (;;) {
        const uint32_t mat = triMat[prim];
        const uint32_t bits = matBits[mat];
        if ((bits & ALVRL_BSDF_DELTA) && depth <= ALVRL_CHAIN_MAX_DEPTH) {              /* 447-448 */
            float T[3] = {1.0f, 1.0f, 1.0f};                                            /* 450-459 */
            if (inMedium && haveMedium) medium_transmittance_exact(med, o, d, tHit, T);
            if (!(T[0] == 0 && T[1] == 0 && T[2] == 0)) {
                /* face normal and shading frame (skdtree.h:367-378,395-396,426; util.cpp:603-608) */
                const F3 p0 = f3(__ldg(&triVerts[3 * (size_t) prim])), p1 = f3(__ldg(&triVerts[3 * (size_t) prim + 1])), p2 = f3(__ldg(&triVerts[3 * (size_t) prim + 2]));
                F3 fn = xcross(xsub3(p1, p0), xsub3(p2, p0));
                const float length = xlen(fn);
                if (!(fn.x == 0 && fn.y == 0 && fn.z == 0)) fn = xdivv(fn, length);
                const F3 dpdu = xsub3(p1, p0);
                const F3 fs = xnormalize(xsub3(dpdu, xscale(fn, xdot(fn, dpdu))));
                const F3 ft = xcross(fn, fs);
                const F3 md = f3(-d.x, -d.y, -d.z);
                const F3 wi = f3(xdot(md, fs), xdot(md, ft), xdot(md, fn));              /* its.wi = toLocal(-ray.d) */
                const float4 o0 = __ldg(&matOptics[3 * mat]), o1 = __ldg(&matOptics[3 * mat + 1]), o2 = __ldg(&matOptics[3 * mat + 2]);
                const float specR[3] = {o1.z, o1.w, o2.x}, specT[3] = {o2.y, o2.z, o2.w};
                const int compCount = (bits & ALVRL_BSDF_DIELECTRIC) ? 2 : 1;
                ChainRay kids[2]; int nk = 0;
                for (int i = 0; i < compCount; i++) {                                   /* 467-504 */
                    F3 wo; float eta = 1.0f; float bw[3];
                    if (bits & ALVRL_BSDF_DIELECTRIC) {                                  /* dielectric.cpp:365-384, 218-226 */
                        const float etaM = o0.x, invEta = xdiv(1.0f, etaM);
                        float cosThetaT;
                        const float F = fresnel_dielectric_ext(wi.z, cosThetaT, etaM);
                        if (i == 0) { wo = f3(-wi.x, -wi.y, wi.z); eta = 1.0f; for (int q = 0; q < 3; q++) bw[q] = xmul(specR[q], F); }
                        else {
                            const float scale = -(cosThetaT < 0 ? invEta : etaM);
                            wo = f3(xmul(scale, wi.x), xmul(scale, wi.y), cosThetaT);
                            eta = cosThetaT < 0 ? etaM : invEta;
                            const float factor = cosThetaT < 0 ? invEta : etaM;
                            const float m = xmul(xmul(factor, factor), xsub(1.0f, F));
                            for (int q = 0; q < 3; q++) bw[q] = xmul(specT[q], m);
                        }
                    } else {                                                             /* conductor.cpp:254-268 */
                        if (wi.z <= 0) continue;
                        wo = f3(-wi.x, -wi.y, wi.z); eta = 1.0f;
                        bw[0] = xmul(specR[0], fresnel_conductor_exact(wi.z, o0.x, o0.w));
                        bw[1] = xmul(specR[1], fresnel_conductor_exact(wi.z, o0.y, o1.x));
                        bw[2] = xmul(specR[2], fresnel_conductor_exact(wi.z, o0.z, o1.y));
                    }
                    if (bw[0] == 0 && bw[1] == 0 && bw[2] == 0) continue;               /* 477-478 */
                    const float eta2 = xmul(eta, eta);
                    float thr2[3];
                    for (int q = 0; q < 3; q++) thr2[q] = xmul(xmul(xmul(thr[q], T[q]), bw[q]), eta2);      /* 480 */
                    const float maxRR = depth >= specRRdepth ? 0.98f : 1.0f;
                    const float rrProb = fminf(maxRR, fmaxf(fmaxf(thr2[0], thr2[1]), thr2[2]));
                    const uint32_t childCode = code * 2u + (uint32_t) i;
                    if (rrProb <= 0) continue;
                    if (rrProb < 1) {                                                    /* 485: rRec.nextSample1D() */
                        const float uu = alvrl_rng_uniform(alvrl_rng_key(seed, ALVRL_RNG_CHAIN, pixel, childCode), 0);
                        if (uu > rrProb) continue;
                    }
                    const float rInv = xdiv(1.0f, rrProb);
                    ChainRay &k = kids[nk++];
                    for (int q = 0; q < 3; q++) { k.thr[q] = xmul(thr2[q], rInv); k.w[q] = xmul(xmul(xmul(w[q], T[q]), bw[q]), rInv); }   /* 488, 500 */
                    k.o = hitP;                                                          /* 490: RayDifferential(its.p, its.toWorld(wo)) */
                    k.d = xadd3(xadd3(xscale(fs, wo.x), xscale(ft, wo.y)), xscale(fn, wo.z));
                    k.code = childCode; k.depth = depth + 1;
                    k.inMedium = inMedium;
                    if (bits & ALVRL_MAT_TRANSITION)                                     /* 491-493, records.inl:81-86 */
                        k.inMedium = xdot(k.d, fn) > 0 ? ((bits & ALVRL_MAT_EXTERIOR_MEDIUM) ? 1 : 0) : ((bits & ALVRL_MAT_INTERIOR_MEDIUM) ? 1 : 0);
                }
                for (int i = nk - 1; i >= 0; i--) stack[sp++] = kids[i];                 /* component 0 is popped first */
            }
        }
        /* next ray whose LiInternal finds a surface (416-423: an infinite segment returns nothing and ends its branch) */
        bool found = false;
        while (sp && !found) {
            const ChainRay r = stack[--sp];
            float t, u, v; uint32_t pr;
            if (!scene_intersect<false>(sc, r.o, r.d, ALVRL_EPSILON, INFINITY, true, t, pr, u, v)) continue;
            F3 p;
            const SegRec s = chain_segment(med, haveMedium, triVerts, triMat, matAlbedo, matBits, r.o, r.d, pr, u, v, r.w, r.inMedium != 0, p);
            if (WRITE) { out[count] = s; meta[count] = make_uint4(pixel, count, r.code, (triMat[pr] << 1) | (r.inMedium ? 1u : 0u)); }
            count++;
            o = r.o; d = r.d; tHit = t; prim = pr; hitP = p;
            for (int q = 0; q < 3; q++) { thr[q] = r.thr[q]; w[q] = r.w[q]; }
            code = r.code; depth = r.depth; inMedium = r.inMedium;
            found = true;
        }
        if (!found) break;
    }
    return count;
}

/* pass 1: counts[pix] = number of chain segments below the camera segment of pixel pix (0 unless it ends on a delta surface) */
__global__ void __launch_bounds__(64) k_chain_count(SceneDev sc, MediumDev med, int haveMedium, const float4 *__restrict__ triVerts, const uint32_t *__restrict__ triMat,
                                                    const float4 *__restrict__ matAlbedo, const uint32_t *__restrict__ matBits, const float4 *__restrict__ matOptics,
                                                    uint64_t seed, int specRRdepth, float initialThroughput, const SegRec *__restrict__ pixSegs,
                                                    const uint32_t *__restrict__ hitPrim, const float *__restrict__ hitT, uint32_t P, uint32_t *__restrict__ counts) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    uint32_t n = 0;
    const SegRec s = pixSegs[pix];
    if (__float_as_uint(s.dn.w) & SEG_DELTA)
        n = chain_walk<false>(sc, med, haveMedium, triVerts, triMat, matAlbedo, matBits, matOptics, seed, specRRdepth, initialThroughput, pix,
                              f3(s.o), f3(s.d), hitT[pix], hitPrim[pix], f3(s.p), nullptr, nullptr);
    counts[pix] = n;
}
/* pass 2: the segments of pixel pix at out[offset[pix] ...), meta = {pixel, ordinal in the pixel's chain, path code, (material << 1) | in-medium} */
__global__ void __launch_bounds__(64) k_chain_write(SceneDev sc, MediumDev med, int haveMedium, const float4 *__restrict__ triVerts, const uint32_t *__restrict__ triMat,
                                                    const float4 *__restrict__ matAlbedo, const uint32_t *__restrict__ matBits, const float4 *__restrict__ matOptics,
                                                    uint64_t seed, int specRRdepth, float initialThroughput, const SegRec *__restrict__ pixSegs,
                                                    const uint32_t *__restrict__ hitPrim, const float *__restrict__ hitT, uint32_t P, const uint32_t *__restrict__ offset,
                                                    SegRec *__restrict__ out, uint4 *__restrict__ meta) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const SegRec s = pixSegs[pix];
    if (!(__float_as_uint(s.dn.w) & SEG_DELTA) || offset[pix + 1] == offset[pix]) return;
    chain_walk<true>(sc, med, haveMedium, triVerts, triMat, matAlbedo, matBits, matOptics, seed, specRRdepth, initialThroughput, pix,
                     f3(s.o), f3(s.d), hitT[pix], hitPrim[pix], f3(s.p), out + offset[pix], meta + offset[pix]);
}

/* R rows of a slice range: the contributions of the chain segments are added to the row of their pixel, in chain order
 * (getVRLContributions 812-813 accumulates into vrlContributions across the recursion).  X: the chain segments' own matrix,
 * column-major like R; xFirst[row] .. xFirst[row + 1]: the extra rows of a row */
__global__ void k_add_chain_rows(float2 *__restrict__ R, uint32_t ldR, uint32_t rowBegin, uint32_t numRows, const float2 *__restrict__ X, uint32_t ldX,
                                 const uint32_t *__restrict__ xFirst, uint32_t N) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (r >= numRows || v >= N) return;
    const uint32_t a = xFirst[r], b = xFirst[r + 1];
    if (a == b) return;
    float2 acc = R[(size_t) v * ldR + rowBegin + r];
    for (uint32_t e = a; e < b; e++) { const float2 x = X[(size_t) v * ldX + e]; acc.x = xadd(acc.x, x.x); acc.y = xadd(acc.y, x.y); }
    R[(size_t) v * ldR + rowBegin + r] = acc;
}

/* render pass: fb[pixel] += sum over the pixel's in-medium chain segments, in chain order, of Li_segment * weight (598: return Li * weight;
 * 507: Li = LiDirect + LiSpec).  pixList[i]: a pixel with chain segments, first[2 i] .. first[2 i + 1]: its entries in subLi / segs */
__global__ void k_chain_accumulate(float4 *__restrict__ fb, uint32_t W, uint32_t H, const float4 *__restrict__ subLi, const SegRec *__restrict__ segs,
                                   const uint32_t *__restrict__ pixList, const uint32_t *__restrict__ first, uint32_t nPix) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPix) return;
    const uint32_t pixel = pixList[i], x = pixel / H, y = pixel % H;
    float4 acc = fb[(size_t) y * W + x];
    for (uint32_t e = first[2 * i]; e < first[2 * i + 1]; e++) {
        if (!(__float_as_uint(segs[e].dn.w) & SEG_VALID)) continue;          /* a segment outside the medium adds nothing (545-547) */
        const float4 li = subLi[e];
        acc.x = xadd(acc.x, xmul(li.x, segs[e].p.w)); acc.y = xadd(acc.y, xmul(li.y, segs[e].n.w)); acc.z = xadd(acc.z, xmul(li.z, segs[e].albedo.w));
    }
    acc.w = 1.0f;
    fb[(size_t) y * W + x] = acc;
}
void launch_chain_accumulate(float4 *fb, uint32_t W, uint32_t H, const float4 *subLi, const SegRec *segs, const uint32_t *pixList, const uint32_t *first, uint32_t nPix, cudaStream_t st) {
    if (nPix) k_chain_accumulate<<<(nPix + 127) / 128, 128, 0, st>>>(fb, W, H, subLi, segs, pixList, first, nPix);
}

void launch_chain_count(const SceneDev &sc, const MediumDev &med, bool haveMedium, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint64_t seed, int specRRdepth, float initialThroughput, const SegRec *pixSegs,
                        const uint32_t *hitPrim, const float *hitT, uint32_t P, uint32_t *counts, cudaStream_t st) {
    k_chain_count<<<(P + 63) / 64, 64, 0, st>>>(sc, med, haveMedium ? 1 : 0, triVerts, triMat, matAlbedo, matBits, matOptics, seed, specRRdepth, initialThroughput,
                                                 pixSegs, hitPrim, hitT, P, counts);
}
void launch_chain_write(const SceneDev &sc, const MediumDev &med, bool haveMedium, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint64_t seed, int specRRdepth, float initialThroughput, const SegRec *pixSegs,
                        const uint32_t *hitPrim, const float *hitT, uint32_t P, const uint32_t *offset, SegRec *out, uint4 *meta, cudaStream_t st) {
    k_chain_write<<<(P + 63) / 64, 64, 0, st>>>(sc, med, haveMedium ? 1 : 0, triVerts, triMat, matAlbedo, matBits, matOptics, seed, specRRdepth, initialThroughput,
                                                 pixSegs, hitPrim, hitT, P, offset, out, meta);
}
void launch_add_chain_rows(float2 *R, uint32_t ldR, uint32_t rowBegin, uint32_t numRows, const float2 *X, uint32_t ldX, const uint32_t *xFirst, uint32_t N, cudaStream_t st) {
    if (!numRows || !N) return;
    dim3 g((numRows + 127) / 128, N);
    k_add_chain_rows<<<g, 128, 0, st>>>(R, ldR, rowBegin, numRows, X, ldX, xFirst, N);
}

} // namespace alvrl
