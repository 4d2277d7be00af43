/*
 * alvrl_rng.h -- the counter-based sample stream of the B200 VRL path.
 *
 * The reference draws every uniform from one sequential SFMT-19937 stream per
 * worker thread (src/samplers/independent.cpp:95-103, src/libcore/random.cpp:630-639),
 * which cannot be reproduced by a massively parallel device.  The GPU path therefore
 * addresses its uniforms by *what they are used for*:
 *
 *     u = alvrl_rng_uniform(alvrl_rng_key(seed, domain, a, b), k)
 *
 * domain / a / b / k:
 *     ALVRL_RNG_R        a = row of R (representative pixel), b = VRL index,
 *                        k = draw number inside integrateVRL (vrlIntegrator.cpp:603-785):
 *                        k = 2*s, 2*s+1 for vol->vol sample s, then 2*Nvv + s for vol->surf
 *     ALVRL_RNG_RENDER   a = pixel index (y + H*x, the reference's m_slices index),
 *                        b = position in the slice's representative list, k as above
 *     ALVRL_RNG_SLICEMAP a = slice, b = 0, k = draw number in sampleRepresentativePixels
 *     ALVRL_RNG_CLUSTER  a = clustering id (slice, or ALVRL_RNG_GLOBAL_ID), b = 0,
 *                        k = draw number inside that Clustering object (sampleRepresentatives);
 *                        the draws of Clustering::split (Preprocessor.cpp:590-684) are addressed by the
 *                        cluster they split: key' = alvrl_rng_node_key(key, begin, end), k = 0, 1 for the
 *                        two weightedSample draws, k = 2.. for the random direction of degenerate centres.
 *                        A split is thus a pure function of the cluster's list range and columns, whatever
 *                        the order in which the refinement visits the clusters -- which is what lets the
 *                        device split many clusters of one Clustering object concurrently.
 *
 * The float conversion is the reference's own ((x >> 9) | 0x3f800000) - 1.0f
 * (random.cpp:630-639), so u is in [0, 1).  Integer-only => bit-identical on host and device.
 * The CPU oracle includes this header so that both sides can be compared on the same samples.
 */
#ifndef ALVRL_RNG_H
#define ALVRL_RNG_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ALVRL_HD __host__ __device__ __forceinline__
#else
#define ALVRL_HD static inline
#endif

enum {
    ALVRL_RNG_R        = 1,
    ALVRL_RNG_RENDER   = 2,
    ALVRL_RNG_SLICEMAP = 3,
    ALVRL_RNG_CLUSTER  = 4,
    ALVRL_RNG_CHAIN    = 5,    /* a = pixel index, b = path code of the branch: the roulette draw of LiInternal (485) */
    ALVRL_RNG_TRACER   = 6,    /* a = particle index, b = 0: the draws of vrlTracer::traceOneParticle in their order */
    ALVRL_RNG_VOLPATH  = 7     /* a = pixel index, b = outer sample index: the draws of renderBlock + VolumetricPathTracer::Li */
};
#define ALVRL_RNG_GLOBAL_ID 0xfffffffeu

/* 32-bit finaliser (two multiply-xorshift rounds) */
ALVRL_HD uint32_t alvrl_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

ALVRL_HD uint32_t alvrl_rng_key(uint64_t seed, uint32_t domain, uint32_t a, uint32_t b) {
    uint32_t h = alvrl_mix32((uint32_t) seed ^ (domain * 0x9e3779b9u));
    h = alvrl_mix32(h ^ (uint32_t) (seed >> 32));
    h = alvrl_mix32(h + a * 0x85ebca6bu + 0x165667b1u);
    h = alvrl_mix32(h ^ (b * 0xc2b2ae35u + 0x27d4eb2fu));
    return h;
}

/* sub-stream of one cluster [begin, end) of a Clustering object's VRL list */
ALVRL_HD uint32_t alvrl_rng_node_key(uint32_t key, uint32_t begin, uint32_t end) {
    uint32_t h = alvrl_mix32(key ^ (begin * 0x85ebca6bu + 0x2545f491u));
    return alvrl_mix32(h + end * 0xc2b2ae35u + 0x68e31da4u);
}

ALVRL_HD uint32_t alvrl_rng_bits(uint32_t key, uint32_t k) {
    return alvrl_mix32(key + k * 0x9e3779b9u);
}

ALVRL_HD float alvrl_bits_to_float(uint32_t bits) {
    uint32_t u = (bits >> 9) | 0x3f800000u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u) - 1.0f;
#else
    float f; memcpy(&f, &u, 4);
    return f - 1.0f;
#endif
}

ALVRL_HD float alvrl_rng_uniform(uint32_t key, uint32_t k) {
    return alvrl_bits_to_float(alvrl_rng_bits(key, k));
}

#endif /* ALVRL_RNG_H */
