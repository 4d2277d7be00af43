/* kernels.h -- host-callable launchers of the CUDA kernels (one per translation unit / flavour). */
#pragma once
#include "types.h"

namespace alvrl {

/* primary.cu (exact arithmetic) */
void launch_primary(const SceneDev &sc, const MediumDev &med, const CameraDev &cam, const float4 *triVerts,
                    const uint32_t *triMat, const float4 *matAlbedo, const uint32_t *matBits, bool haveMedium,
                    SegRec *pixSegs, uint32_t *hitPrim, float *hitT, cudaStream_t st);
void launch_gather_rows(const SegRec *pixSegs, const uint32_t *rowPixel, uint32_t numRows, SegRec *rowSegs, cudaStream_t st);
void launch_gather_points(const SegRec *pixSegs, uint32_t P, float directionScale, float *pos, float *dir, cudaStream_t st);
void launch_trace_rays(const SceneDev &sc, const float *o, const float *d, const float *mint, const float *maxt, uint32_t n,
                       uint32_t *prim, float *t, cudaStream_t st);
void launch_eval_transmittance(const SceneDev &sc, const MediumDev &med, const float *p1, const int32_t *onSurf, const float *p2,
                               uint32_t n, float *T, cudaStream_t st);
void launch_fb_to_rgb(const float4 *fb, float *rgb, uint32_t n, cudaStream_t st);

/* chain.cu (exact arithmetic): specular chains */
void launch_chain_count(const SceneDev &sc, const MediumDev &med, bool haveMedium, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint64_t seed, int specRRdepth, float initialThroughput, const SegRec *pixSegs,
                        const uint32_t *hitPrim, const float *hitT, uint32_t P, uint32_t *counts, cudaStream_t st);
void launch_chain_write(const SceneDev &sc, const MediumDev &med, bool haveMedium, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint64_t seed, int specRRdepth, float initialThroughput, const SegRec *pixSegs,
                        const uint32_t *hitPrim, const float *hitT, uint32_t P, const uint32_t *offset, SegRec *out, uint4 *meta, cudaStream_t st);
void launch_add_chain_rows(float2 *R, uint32_t ldR, uint32_t rowBegin, uint32_t numRows, const float2 *X, uint32_t ldX, const uint32_t *xFirst, uint32_t N, cudaStream_t st);
void launch_chain_accumulate(float4 *fb, uint32_t W, uint32_t H, const float4 *subLi, const SegRec *segs, const uint32_t *pixList, const uint32_t *first, uint32_t nPix, cudaStream_t st);

/* tracer.cu (exact arithmetic): the VRL tracer */
void launch_trace_count(const SceneDev &sc, const MediumDev &med, const uint32_t *emTris, const float *emCdf, uint32_t emN, const float emPower[3],
                        uint64_t seed, int shortVrls, int maxDepth, int rrDepth, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint32_t first, uint32_t n, uint32_t *counts, cudaStream_t st);
void launch_trace_write(const SceneDev &sc, const MediumDev &med, const uint32_t *emTris, const float *emCdf, uint32_t emN, const float emPower[3],
                        uint64_t seed, int shortVrls, int maxDepth, int rrDepth, const float4 *triVerts, const uint32_t *triMat, const float4 *matAlbedo,
                        const uint32_t *matBits, const float4 *matOptics, uint32_t n, const uint32_t *offset, float *out, cudaStream_t st);

/* volpath.cu (exact arithmetic): VolumetricPathTracer with onlyVRLpaths, one outer sample of every pixel per launch */
void launch_volpath_sample(const SceneDev &sc, const MediumDev &med, const CameraDev &cam, const uint32_t *emTris, const float *emCdf, uint32_t emN,
                           const float emRadiance[3], float emInvArea, const uint8_t *triEmitter, const float4 *triVerts, const uint32_t *triMat,
                           const float4 *matAlbedo, const uint32_t *matBits, const float4 *matOptics, uint64_t seed, uint32_t sample, int internalSamples,
                           uint32_t flags, bool centre, int maxDepth, int rrDepth, float4 *acc, cudaStream_t st);
void launch_volpath_develop(const float4 *acc, uint32_t W, uint32_t H, float *rgb, cudaStream_t st);

/* film.cu (exact arithmetic) */
void launch_film_splat(const float4 *fb, uint32_t W, uint32_t H, const FilmFilterDev &f, float *acc, cudaStream_t st);
void launch_film_develop(const float *acc, uint32_t n, float *rgb, cudaStream_t st);

/* transport_strict.cu / transport_fast.cu */
/* rowKey (optional): the stream key of every row (default: P.rowBase + row); weighted: the segments carry LiInternal's weight */
void launch_build_R_strict(const TransportParams &P, const SegRec *rowSegs, uint32_t numRows, const VrlRec *vrls, float2 *R,
                           uint32_t ldR, cudaStream_t st, const uint32_t *rowKey = nullptr, bool weighted = false);
void launch_build_R_fast(const TransportParams &P, const SegRec *rowSegs, uint32_t numRows, const VrlRec *vrls, float2 *R,
                         uint32_t ldR, cudaStream_t st, const uint32_t *rowKey = nullptr, bool weighted = false);
void launch_render_strict(const TransportParams &P, bool clustered, const SegRec *pixSegs, const uint32_t *slicePixels,
                          const uint4 *work, uint32_t numWork, const VrlRec *repRecs, const uint32_t *repOffset, float4 *fb,
                          uint32_t W, uint32_t H, cudaStream_t st, const uint32_t *segKey = nullptr);
void launch_render_fast(const TransportParams &P, bool clustered, const SegRec *pixSegs, const uint32_t *slicePixels,
                        const uint4 *work, uint32_t numWork, const VrlRec *repRecs, const uint32_t *repOffset, float4 *fb,
                        uint32_t W, uint32_t H, cudaStream_t st, const uint32_t *segKey = nullptr);

} // namespace alvrl
