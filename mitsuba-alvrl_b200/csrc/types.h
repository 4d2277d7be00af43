/*
 * types.h -- device data layout of the VRL hot path (shared by host code and kernels).
 *
 * Everything a kernel reads is packed into 16-byte float4 records so that one thread fetches a whole
 * record with LDG.128 / LDS.128:
 *
 *   VrlRec  (64 B)  one VRL of vrlVector (VRL.h:17-100), with the per-VRL constants integrateVRL
 *                   recomputes for every pair (normalize(end-start), |end-start|) hoisted.
 *   SegRec (112 B)  one camera-ray segment: what LiInternal/integrateVRL read from `ray` and `rRec.its`
 *                   (vrlIntegrator.cpp:398-423,626-631,707-719) for one pixel.
 *   BvhNode (32 B)  threaded (stackless) BVH node: AABB + escape index + leaf range.
 *   TriRec  (48 B)  Wald TriAccel (triaccel.h:37-59) + original triangle index.
 *   R              float2 {mean, var} per (row, vrl), stored COLUMN-major: R[vrl * ldR + row], ldR = roundup(G, 32).
 *                   One column = one VRL over all representative pixels, rows grouped by slice, so the
 *                   transport kernel (thread = row) stores coalesced and the clustering kernels
 *                   (thread = row, loop over the VRLs of a cluster) load coalesced.
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/alvrl.h"
#include "occ_query.h"

#define ALVRL_EPSILON 1e-4f          /* include/mitsuba/core/constants.h:32 */
#define ALVRL_SHADOW_EPSILON 1e-3f   /* constants.h:33 */
#define ALVRL_INV_PI 0.31830988618379067154f
#define ALVRL_INV_FOURPI 0.07957747154594766788f
#define ALVRL_CTA_SEGS_HOST 128u     /* segments per CTA of the transport kernels */

struct VrlRec {
    float4 s;      /* start.xyz, |end - start| */
    float4 e;      /* end.xyz, cluster weight (render lists) */
    float4 dir;    /* normalize(end - start).xyz, 0 */
    float4 power;  /* rgb, 0 */
};

#define SEG_VALID  1u   /* primary ray hit geometry (vrlIntegrator.cpp:418-423) */
#define SEG_SMOOTH 2u   /* BSDF has an ESmooth component (vrlIntegrator.cpp:726-727) */
#define SEG_DELTA  4u   /* BSDF has delta components: the segment continues as a specular chain (445-448) */

struct SegRec {
    float4 o;      /* ray.o.xyz, distance(its.p, ray.o) */
    float4 d;      /* ray.d.xyz, cosTheta(its.wi) */
    float4 dn;     /* normalize(ray.d).xyz (sampleVtoDistance re-normalises, 925), flags as int bits */
    float4 p;      /* its.p.xyz (barycentric), weight.r of the segment (LiInternal's `weight`: 1 for a camera segment) */
    float4 n;      /* shading normal.xyz, weight.g */
    float4 albedo; /* diffuse reflectance rgb, weight.b */
    float4 tE;     /* transmittance eye -> surface rgb (vrlIntegrator.cpp:711-719), 0 */
};

struct BvhNode {
    float4 lo;     /* bmin.xyz, __int_as_float(escape index) */
    float4 hi;     /* bmax.xyz, __int_as_float(leaf: (first << 4) | count, inner: 0) */
};
#define BVH_END 0x7fffffff

/* 4-wide node of the fast flavour's any-hit query on large scenes (bvh.h::collapse4): the boxes of up to four children in
 * SoA form -- one 128-byte line per node, seven 16-byte loads -- and what each child is: >= 0 an inner node (index into the
 * Bvh4Node array), < 0 a leaf, ~((first << 4) | count) into the triangle arrays; an unused slot has an empty box
 * (lo = +inf, hi = -inf) that no ray enters */
struct __align__(128) Bvh4Node {
    float4 lox, loy, loz, hix, hiy, hiz;
    int4 child;
    int4 pad;
};

struct TriRec {
    float4 a;      /* __int_as_float(k), n_u, n_v, n_d */
    float4 b;      /* a_u, a_v, b_nu, b_nv */
    float4 c;      /* c_nu, c_nv, __int_as_float(original triangle index), 0 */
};

/* fast-flavour triangle: plane + barycentric functionals, no axis switch.
 *   t = (p.w - dot(p.xyz, o)) / dot(p.xyz, d);  u = dot(q.xyz, P) + q.w;  v = dot(r.xyz, P) + r.w;  P = o + t d */
struct TriFast { float4 p, q, r; };

struct MediumDev {
    int type;               /* 0 homogeneous, 1 grid (simpson) */
    int phaseType; float g;
    float sigmaS[3], sigmaT[3];   /* Medium::getSigmaS()/sigmaT; grid: sigmaS = sigmaS_base (quirk B2) */
    float samplingWeight;
    int grey;               /* sigmaT equal in all channels (homogeneous fast path) */
    /* grid */
    const float *density; int res[3]; float bmin[3], bmax[3]; float scale, stepSize; float albedo[3];
    float gsc[3], gtr[3];   /* worldToGrid: g = gsc * p + gtr (gridvolume.cpp:188-196) */
};

/* pre-rasterised reconstruction filter (include/mitsuba/core/rfilter.h:28,76-77,96-98) */
#define ALVRL_FILTER_RESOLUTION 31
struct FilmFilterDev {
    float table[ALVRL_FILTER_RESOLUTION + 1];
    float scaleFactor;      /* MTS_FILTER_RESOLUTION / radius */
    float radius;
    int taps;               /* floor(radius): the integer offsets a sample at a pixel centre reaches */
};

struct CameraDev {
    float s2c[16], c2w[16];
    uint32_t W, H; float nearClip, farClip; float invResX, invResY;
};

struct SceneDev {
    const BvhNode *nodes; const TriRec *tris; const TriFast *trisFast; const BvhNode *leafNodes; uint32_t numNodes, numTris, numLeaves;
    float kdMin[3], kdMax[3];     /* ShapeKDTree AABB incl. the 1e-3 enlargement (gkdtree.h:1213-1220) */
    int anyHit;
    /* small scenes, fast flavour (occluders.h): visMode 0 = tree traversal, 1 = flat leaf sweep, 2 = compiled occluder set */
    const float4 *occTris; uint32_t numOccTris; int visMode;
    const Bvh4Node *nodes4; uint32_t numNodes4;   /* visMode 0 */
};

struct TransportParams {
    SceneDev scene; MediumDev medium;
    int Nvv, Nvs, shortVrls, Rsamples;
    uint64_t seed; uint32_t rngDomain;
    const float *tape; uint32_t tapeK;     /* parity mode: u = tape[(row*N + vrl)*tapeK + k] */
    uint32_t numVrls;                      /* N */
    float normalization;                   /* (float)(1.0 / particleCount), vrlIntegrator.cpp:805 */
    float invParticleDiv;                  /* particleCount as float (Li /= particleCount, 590) */
    uint32_t rowBase;                      /* global index of the first row handed to k_build_R (slice sharding) */
    OccDev occ;                            /* compiled occluder set (scene.visMode == 2), read through the constant bank */
};
