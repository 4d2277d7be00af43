/* transport_launch.inl -- launchers shared by the two flavours (included after transport.cuh). */
namespace alvrl {

void ALVRL_NAME(launch_build_R)(const TransportParams &P, const SegRec *rowSegs, uint32_t numRows, const VrlRec *vrls, float2 *R,
                                uint32_t ldR, cudaStream_t st) {
    if (numRows == 0 || P.numVrls == 0) return;
    const uint32_t rowBlocks = (numRows + ALVRL_CTA_SEGS - 1) / ALVRL_CTA_SEGS;
    /* aim at >= 16 CTAs per SM (148 SMs) so that the tail is short, VRL chunks being multiples of the tile */
    const uint32_t target = 148u * 16u;
    uint32_t chunks = (target + rowBlocks - 1) / rowBlocks;
    const uint32_t maxChunks = (P.numVrls + ALVRL_TILE_VRLS - 1) / ALVRL_TILE_VRLS;
    if (chunks > maxChunks) chunks = maxChunks;
    if (chunks < 1) chunks = 1;
    uint32_t per = (P.numVrls + chunks - 1) / chunks;
    per = ((per + ALVRL_TILE_VRLS - 1) / ALVRL_TILE_VRLS) * ALVRL_TILE_VRLS;
    chunks = (P.numVrls + per - 1) / per;
    dim3 grid(rowBlocks, chunks);
    if (P.medium.type == 0) ALVRL_NAME(k_build_R)<0><<<grid, ALVRL_CTA_SEGS, 0, st>>>(P, rowSegs, numRows, vrls, R, ldR, per);
    else ALVRL_NAME(k_build_R)<1><<<grid, ALVRL_CTA_SEGS, 0, st>>>(P, rowSegs, numRows, vrls, R, ldR, per);
}

void ALVRL_NAME(launch_render)(const TransportParams &P, bool clustered, const SegRec *pixSegs, const uint32_t *slicePixels,
                               const uint4 *work, uint32_t numWork, const VrlRec *repRecs, const uint32_t *repOffset, float4 *fb,
                               uint32_t W, uint32_t H, cudaStream_t st) {
    if (numWork == 0) return;
    if (P.medium.type == 0) {
        if (clustered) ALVRL_NAME(k_render)<0, true><<<numWork, ALVRL_CTA_SEGS, 0, st>>>(P, pixSegs, slicePixels, work, repRecs, repOffset, fb, W, H);
        else ALVRL_NAME(k_render)<0, false><<<numWork, ALVRL_CTA_SEGS, 0, st>>>(P, pixSegs, slicePixels, work, repRecs, repOffset, fb, W, H);
    } else {
        if (clustered) ALVRL_NAME(k_render)<1, true><<<numWork, ALVRL_CTA_SEGS, 0, st>>>(P, pixSegs, slicePixels, work, repRecs, repOffset, fb, W, H);
        else ALVRL_NAME(k_render)<1, false><<<numWork, ALVRL_CTA_SEGS, 0, st>>>(P, pixSegs, slicePixels, work, repRecs, repOffset, fb, W, H);
    }
}

} // namespace alvrl
