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

/* transport_strict.cu / transport_fast.cu */
void launch_build_R_strict(const TransportParams &P, const SegRec *rowSegs, uint32_t numRows, const VrlRec *vrls, float2 *R,
                           uint32_t ldR, cudaStream_t st);
void launch_build_R_fast(const TransportParams &P, const SegRec *rowSegs, uint32_t numRows, const VrlRec *vrls, float2 *R,
                         uint32_t ldR, cudaStream_t st);
void launch_render_strict(const TransportParams &P, bool clustered, const SegRec *pixSegs, const uint32_t *slicePixels,
                          const uint4 *work, uint32_t numWork, const VrlRec *repRecs, const uint32_t *repOffset, float4 *fb,
                          uint32_t W, uint32_t H, cudaStream_t st);
void launch_render_fast(const TransportParams &P, bool clustered, const SegRec *pixSegs, const uint32_t *slicePixels,
                        const uint4 *work, uint32_t numWork, const VrlRec *repRecs, const uint32_t *repOffset, float4 *fb,
                        uint32_t W, uint32_t H, cudaStream_t st);

} // namespace alvrl
