/*
 * film.cu -- the step after the path: reconstruction-filter splat and pass accumulation of the film.
 *
 * ImageBlock::put (include/mitsuba/render/imageblock.h:124-202) adds a sample at fractional position pos to every pixel
 * within the filter radius, weighted by the product of two look-ups in the pre-rasterised filter
 * (ReconstructionFilter::configure / evalDiscretized, src/libcore/rfilter.cpp:37-55, include/mitsuba/core/rfilter.h:76-77),
 * into the channels {spectrum, alpha, weight}; invalid samples (NaN / negative) are rejected (147-151); the film divides by
 * the weight channel when it is developed (Bitmap::convertMultiSpectrumAlphaWeight, src/libcore/bitmap.cpp:1617-1624).
 * The path renders one sample per pixel centre, so the scatter is a gather here: thread = destination pixel, sources
 * visited in raster order (y outer, x inner) -- the order a sequential put() loop over the samples adds them in -- with
 * the reference's operations (weight = wX * wY; dest += weight * value), no contraction (-fmad=false).
 */
#include "dev_common.cuh"
#include "kernels.h"

namespace alvrl {

__global__ void k_film_splat(const float4 *__restrict__ fb, uint32_t W, uint32_t H, FilmFilterDev f, float *__restrict__ acc) {
    const uint32_t X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= W || Y >= H) return;
    float s[5];
    float *dst = acc + ((size_t) Y * W + X) * 5u;
#pragma unroll
    for (int k = 0; k < 5; k++) s[k] = dst[k];
    const int R = f.taps;                                                 /* floor(radius): |dx| <= radius for integer offsets */
    for (int sy = (int) Y - R; sy <= (int) Y + R; sy++) {
        if (sy < 0 || sy >= (int) H) continue;
        const float wy = f.table[min((int) fabsf(__fmul_rn((float) ((int) Y - sy), f.scaleFactor)), ALVRL_FILTER_RESOLUTION)];
        for (int sx = (int) X - R; sx <= (int) X + R; sx++) {
            if (sx < 0 || sx >= (int) W) continue;
            const float4 v = __ldg(&fb[(size_t) sy * W + sx]);
            /* put(): every channel finite and >= 0, else the sample is dropped (alpha = weight = 1 always pass) */
            if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && v.x >= 0.0f && v.y >= 0.0f && v.z >= 0.0f)) continue;
            const float wx = f.table[min((int) fabsf(__fmul_rn((float) ((int) X - sx), f.scaleFactor)), ALVRL_FILTER_RESOLUTION)];
            const float w = __fmul_rn(wx, wy);
            s[0] = __fadd_rn(s[0], __fmul_rn(w, v.x)); s[1] = __fadd_rn(s[1], __fmul_rn(w, v.y)); s[2] = __fadd_rn(s[2], __fmul_rn(w, v.z));
            s[3] = __fadd_rn(s[3], __fmul_rn(w, 1.0f)); s[4] = __fadd_rn(s[4], __fmul_rn(w, 1.0f));
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) dst[k] = s[k];
}

__global__ void k_film_develop(const float *__restrict__ acc, uint32_t n, float *__restrict__ rgb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *a = acc + (size_t) i * 5u;
    const float weight = a[4], inv = weight == 0.0f ? 0.0f : __fdiv_rn(1.0f, weight);
    rgb[3 * (size_t) i] = __fmul_rn(a[0], inv); rgb[3 * (size_t) i + 1] = __fmul_rn(a[1], inv); rgb[3 * (size_t) i + 2] = __fmul_rn(a[2], inv);
}

void launch_film_splat(const float4 *fb, uint32_t W, uint32_t H, const FilmFilterDev &f, float *acc, cudaStream_t st) {
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    k_film_splat<<<g, b, 0, st>>>(fb, W, H, f, acc);
}
void launch_film_develop(const float *acc, uint32_t n, float *rgb, cudaStream_t st) {
    if (n) k_film_develop<<<(n + 255) / 256, 256, 0, st>>>(acc, n, rgb);
}

} // namespace alvrl
