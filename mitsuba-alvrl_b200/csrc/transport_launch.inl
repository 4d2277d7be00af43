/* transport_launch.inl -- launchers shared by the two flavours (included after transport.cuh). */
#include <cstdlib>
#include <algorithm>
namespace alvrl {

#ifdef ALVRL_FAST
template <typename K> static void set_dyn_smem(K kernel, size_t bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}
#define ALVRL_SMEM_R (sizeof(TileSmem) + sizeof(BvhSmem))
#define ALVRL_SMEM_RENDER (sizeof(TileSmem) + sizeof(BvhSmem))
#define ALVRL_GO_R1(MED, SM, WT) do { set_dyn_smem(ALVRL_NAME(k_build_R)<MED, SM, WT>, ALVRL_SMEM_R); \
        ALVRL_NAME(k_build_R)<MED, SM, WT><<<grid, ALVRL_CTA_SEGS, ALVRL_SMEM_R, st>>>(P, rowSegs, numRows, vrls, R, ldR, per, rowKey); } while (0)
/* the weighted variant exists for the homogeneous kernels only (segments of specular chains); the grid kernel always reads the weight */
#define ALVRL_GO_R(MED, SM) do { if (MED != 1 && weighted) ALVRL_GO_R1(MED, SM, (MED != 1)); else ALVRL_GO_R1(MED, SM, false); } while (0)
#endif

void ALVRL_NAME(launch_build_R)(const TransportParams &P, const SegRec *rowSegs, uint32_t numRows, const VrlRec *vrls, float2 *R,
                                uint32_t ldR, cudaStream_t st, const uint32_t *rowKey, bool weighted) {
    if (numRows == 0 || P.numVrls == 0) return;
    const uint32_t rowBlocks = (numRows + ALVRL_CTA_SEGS - 1) / ALVRL_CTA_SEGS;
    /* CTAs run for milliseconds and ~5 fit on an SM: with a few thousand of them the last, partly filled wave costs up to
     * a fifth of the kernel (2 480 CTAs = 3.35 waves at C2).  Aim at ~128 CTAs per SM (148 SMs), i.e. >= 20 waves, VRL chunks
     * being multiples of the tile (ALVRL_R_CTAS_PER_SM overrides, for experiments) */
    uint32_t perSm = 128u;
    if (const char *e = getenv("ALVRL_R_CTAS_PER_SM")) perSm = (uint32_t) std::max(1, atoi(e));
    const uint32_t target = 148u * perSm;
    uint32_t chunks = (target + rowBlocks - 1) / rowBlocks;
    const uint32_t maxChunks = (P.numVrls + ALVRL_TILE_VRLS - 1) / ALVRL_TILE_VRLS;
    if (chunks > maxChunks) chunks = maxChunks;
    if (chunks < 1) chunks = 1;
    uint32_t per = (P.numVrls + chunks - 1) / chunks;
    per = ((per + ALVRL_TILE_VRLS - 1) / ALVRL_TILE_VRLS) * ALVRL_TILE_VRLS;
    chunks = (P.numVrls + per - 1) / per;
    dim3 grid(rowBlocks, chunks);
    if (P.medium.type == 1) grid = dim3(chunks, rowBlocks);             /* see k_build_R: chunk-major order for the grid medium */
#ifdef ALVRL_FAST
    const int vis = P.scene.visMode;
    if (P.medium.type == 1) { if (vis == 2) ALVRL_GO_R(1, 2); else if (vis == 1) ALVRL_GO_R(1, 1); else ALVRL_GO_R(1, 0); }
    else if (P.medium.grey) { if (vis == 2) ALVRL_GO_R(2, 2); else if (vis == 1) ALVRL_GO_R(2, 1); else ALVRL_GO_R(2, 0); }
    else { if (vis == 2) ALVRL_GO_R(0, 2); else if (vis == 1) ALVRL_GO_R(0, 1); else ALVRL_GO_R(0, 0); }
#else
    (void) weighted;                                                    /* the strict flavour always applies the segment's weight */
    if (P.medium.type == 0) ALVRL_NAME(k_build_R)<0, 0, false><<<grid, ALVRL_CTA_SEGS, 0, st>>>(P, rowSegs, numRows, vrls, R, ldR, per, rowKey);
    else ALVRL_NAME(k_build_R)<1, 0, false><<<grid, ALVRL_CTA_SEGS, 0, st>>>(P, rowSegs, numRows, vrls, R, ldR, per, rowKey);
#endif
}

void ALVRL_NAME(launch_render)(const TransportParams &P, bool clustered, const SegRec *pixSegs, const uint32_t *slicePixels,
                               const uint4 *work, uint32_t numWork, const VrlRec *repRecs, const uint32_t *repOffset, float4 *fb,
                               uint32_t W, uint32_t H, cudaStream_t st, const uint32_t *segKey) {
    if (numWork == 0) return;
#ifdef ALVRL_FAST
#define ALVRL_RENDER_SMEM ALVRL_SMEM_RENDER
#define ALVRL_RENDER_ATTR(MED, CL, SM) set_dyn_smem(ALVRL_NAME(k_render)<MED, CL, SM>, ALVRL_SMEM_RENDER)
#else
#define ALVRL_RENDER_SMEM 0
#define ALVRL_RENDER_ATTR(MED, CL, SM) (void) 0
#endif
#define ALVRL_LAUNCH_RENDER(MED, SM)                                                                                                         \
    do {                                                                                                                                     \
        if (clustered) { ALVRL_RENDER_ATTR(MED, true, SM);                                                                                   \
            ALVRL_NAME(k_render)<MED, true, SM><<<numWork, ALVRL_CTA_SEGS, ALVRL_RENDER_SMEM, st>>>(P, pixSegs, slicePixels, work, repRecs, repOffset, fb, W, H, segKey); } \
        else { ALVRL_RENDER_ATTR(MED, false, SM);                                                                                            \
            ALVRL_NAME(k_render)<MED, false, SM><<<numWork, ALVRL_CTA_SEGS, ALVRL_RENDER_SMEM, st>>>(P, pixSegs, slicePixels, work, repRecs, repOffset, fb, W, H, segKey); } \
    } while (0)
#ifdef ALVRL_FAST
    const int vis = P.scene.visMode;
    if (P.medium.type == 1) { if (vis == 2) ALVRL_LAUNCH_RENDER(1, 2); else if (vis == 1) ALVRL_LAUNCH_RENDER(1, 1); else ALVRL_LAUNCH_RENDER(1, 0); }
    else if (P.medium.grey) { if (vis == 2) ALVRL_LAUNCH_RENDER(2, 2); else if (vis == 1) ALVRL_LAUNCH_RENDER(2, 1); else ALVRL_LAUNCH_RENDER(2, 0); }
    else { if (vis == 2) ALVRL_LAUNCH_RENDER(0, 2); else if (vis == 1) ALVRL_LAUNCH_RENDER(0, 1); else ALVRL_LAUNCH_RENDER(0, 0); }
#else
    if (P.medium.type == 1) ALVRL_LAUNCH_RENDER(1, 0); else ALVRL_LAUNCH_RENDER(0, 0);
#endif
}

} // namespace alvrl
