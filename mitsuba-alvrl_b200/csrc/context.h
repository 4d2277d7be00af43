/*
 * context.h -- the handle behind the C ABI: host mirror of `vrlIntegrator` + `Preprocessor` + `vrlClusterInfo`
 * state (vrlIntegrator.cpp:17-115,1085-1122; Preprocessor.cpp:1593-1619) and the device buffers that replace it.
 */
#pragma once
#include <string>
#include <vector>
#include <memory>
#include <stdexcept>
#include <map>
#include <mutex>
#include <cuda_runtime.h>
#include "types.h"
#include "bvh.h"
#include "slices.h"
#include "host_sampler.h"

namespace alvrl {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define ALVRL_CUDA(call)                                                                                   \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            throw alvrl::Error(ALVRL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

/* Device-memory pool behind DevBuf.  A frame allocates and frees the same multi-gigabyte workspaces (the compacted slice
 * matrices of the cluster refinement, per-round scratch) every time, and cudaMalloc / cudaFree of such blocks cost tens to
 * hundreds of milliseconds each: released blocks are kept and handed out again (best fit, at most twice the request).
 * put() synchronises the device like cudaFree does, so a recycled block is never still in use.  trim() returns everything
 * to the driver (called when the last handle of a device is destroyed, or when an allocation fails). */
class DevPool {
public:
    static DevPool &get() { static DevPool pool; return pool; }
    void *take(size_t bytes, size_t &cap) {
        int dev = 0; cudaGetDevice(&dev);
        bytes = (bytes + 511) & ~(size_t) 511;
        {
            std::lock_guard<std::mutex> lk(m_);
            auto &fl = free_[dev];
            auto it = fl.lower_bound(bytes);
            if (it != fl.end() && it->first <= 2 * bytes + (1u << 20)) { void *p = it->second; cap = it->first; fl.erase(it); return p; }
        }
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            trim(dev);
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) throw Error(ALVRL_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
        cap = bytes;
        return p;
    }
    void put(void *p, size_t cap) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceSynchronize();
        std::lock_guard<std::mutex> lk(m_);
        free_[dev].emplace(cap, p);
    }
    void trim(int dev) {
        std::lock_guard<std::mutex> lk(m_);
        for (auto &kv : free_[dev]) cudaFree(kv.second);
        free_[dev].clear();
    }
    void addHandle(int dev) { std::lock_guard<std::mutex> lk(m_); handles_[dev]++; }
    void dropHandle(int dev) {
        bool last;
        { std::lock_guard<std::mutex> lk(m_); last = --handles_[dev] <= 0; }
        if (last) trim(dev);
    }
private:
    std::mutex m_;
    std::map<int, std::multimap<size_t, void *>> free_;
    std::map<int, int> handles_;
};

/* owning device buffer */
template <typename T> struct DevBuf {
    T *p = nullptr; size_t n = 0; size_t capBytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) DevPool::get().put(p, capBytes); p = nullptr; n = 0; capBytes = 0; }
    void alloc(size_t count) {                     /* n is a capacity: buffers only grow */
        if (count <= n && p) return;
        release();
        if (count) p = static_cast<T *>(DevPool::get().take(count * sizeof(T), capBytes));
        n = count;
    }
    void upload(const T *src, size_t count, cudaStream_t st) {
        alloc(count);
        if (count) { ALVRL_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st)); ALVRL_CUDA(cudaStreamSynchronize(st)); }
    }
    void upload(const std::vector<T> &v, cudaStream_t st) { upload(v.data(), v.size(), st); }
    void download(T *dst, size_t count, cudaStream_t st, size_t offset = 0) const {
        if (count) { ALVRL_CUDA(cudaMemcpyAsync(dst, p + offset, count * sizeof(T), cudaMemcpyDeviceToHost, st)); ALVRL_CUDA(cudaStreamSynchronize(st)); }
    }
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    void init() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); a = b = nullptr; }
    void start(cudaStream_t st) { cudaEventRecord(a, st); }
    float stop(cudaStream_t st) { cudaEventRecord(b, st); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

} // namespace alvrl

struct alvrl_ctx {
    alvrl_params P;
    int device = 0;
    cudaStream_t stream = nullptr;
    alvrl::Timer timer;
    alvrl_stats stats;
    int mathMode = 0;                 /* 0 fast flavour, 1 strict flavour (ALVRL_MATH env / alvrl_set_math_mode) */

    /* scene (host) */
    std::vector<float> verts; std::vector<uint32_t> tris, triMat;
    std::vector<float> albedo; std::vector<uint32_t> matBits;
    std::vector<float> extraBounds;
    bool haveMesh = false, haveMat = false, haveMedium = false, haveCam = false, haveVrls = false;
    bool sceneDirty = true, segsDirty = true;
    float kdMin[3], kdMax[3], sceneMin[3], sceneMax[3];
    MediumDev medium; CameraDev cam;
    std::vector<float> gridHost;

    /* scene (device) */
    FilmFilterDev film; bool haveFilm = false; uint32_t filmPasses = 0; alvrl::DevBuf<float> dFilm;
    alvrl::DevBuf<BvhNode> dNodes, dLeafNodes; alvrl::DevBuf<Bvh4Node> dNodes4; alvrl::DevBuf<TriRec> dTris; alvrl::DevBuf<TriFast> dTrisFast; alvrl::DevBuf<float4> dTriVerts, dOcc;
    alvrl::DevBuf<uint32_t> dTriMat, dMatBits; alvrl::DevBuf<float4> dMatAlbedo; alvrl::DevBuf<float> dGrid;
    SceneDev sceneDev; OccDev occHost;

    /* VRLs */
    std::vector<VrlRec> vrlHost; uint64_t particleCount = 0;
    alvrl::DevBuf<VrlRec> dVrls;
    bool vrlSidesValid = false;       /* per-VRL side bits of the compiled occluder set are in dVrls (occ_query.h) */

    /* per pixel (index y + H*x) */
    alvrl::DevBuf<SegRec> dPixSegs; alvrl::DevBuf<uint32_t> dHitPrim; alvrl::DevBuf<float> dHitT;
    /* VRL tracer (tracer.cu): the area emitter */
    std::vector<uint32_t> emTris; std::vector<float> emCdf; float emPower[3] = {0, 0, 0}; bool haveEmitter = false;
    alvrl::DevBuf<uint32_t> dEmTris; alvrl::DevBuf<float> dEmCdf;
    /* ground truth (volpath.cu): the emitter's radiance, 1 / area, Shape::isEmitter per triangle; {rgb, weight} accumulator */
    float emRadiance[3] = {0, 0, 0}, emInvArea = 0; alvrl::DevBuf<uint8_t> dTriEmitter; alvrl::DevBuf<float4> dVolpathAcc; alvrl::DevBuf<float> dVolpathRgb;
    /* specular chains (chain.cu): the segments below the camera segments, grouped by pixel */
    std::vector<float> optics; alvrl::DevBuf<float4> dMatOptics; bool anyDelta = false, chainsValid = false;
    std::vector<uint32_t> chainOffset;                 /* P + 1 */
    std::vector<uint4> chainMeta;                      /* {pixel, ordinal, path code, (material << 1) | in-medium} */
    alvrl::DevBuf<SegRec> dChainSegs; alvrl::DevBuf<uint4> dChainMeta; alvrl::DevBuf<uint32_t> dChainKey;
    alvrl::DevBuf<SegRec> dXSegs; alvrl::DevBuf<uint32_t> dXIdx, dXKey, dXFirst; alvrl::DevBuf<float2> dX;     /* chain rows of R */
    alvrl::DevBuf<float4> dSubLi; alvrl::DevBuf<uint32_t> dXList, dXPix, dXPixFirst; alvrl::DevBuf<uint4> dXWork;   /* chain render pass */
    bool havePrimary = false;

    /* slices / rows */
    /* slices live on the device: dRecIdx = the pixel ids in partition order (slice s = positions [sliceLo[s], sliceLo[s] +
     * sliceSize[s])), dPixelToSlice = m_slices; the host keeps the ranges only */
    std::vector<uint32_t> sliceLo, sliceSize; bool haveSlices = false;
    std::vector<float> sliceCentroid;              /* 6 per slice: position / direction box midpoints (Preprocessor.cpp:1337-1338) */
    alvrl::DevBuf<uint32_t> dRecIdx, dPixelToSlice;
    uint32_t numSlices() const { return (uint32_t) sliceSize.size(); }
    std::vector<uint32_t> rowOffset, rowPixel; bool haveRows = false;
    std::vector<float> sliceUndersampling; float globalPixelUndersampling = -1;
    alvrl::DevBuf<uint32_t> dRowPixel; alvrl::DevBuf<SegRec> dRowSegs;
    uint32_t sliceBegin = 0, sliceEnd = 0xffffffffu;

    /* R (column-major, see types.h) */
    /* R: only the rows of the handle's slice range are stored (C5: 1 TB over all rows).  dR.p is the address row 0 WOULD have,
     * so that every consumer keeps indexing dR.p[vrl * ldR + global row]; rows outside [builtRow0, builtRow1) do not exist */
    alvrl::DevBuf<float2> dRstore; struct { float2 *p = nullptr; } dR; uint32_t rShift = 0;
    uint32_t ldR = 0; bool haveR = false;
    uint32_t builtRow0 = 0, builtRow1 = 0;          /* rows of R the last build_R / set_R wrote */
    alvrl::DevBuf<float> dTape; uint32_t tapeK = 0; std::vector<float> userTape;

    /* clusters (vrlClusterInfo, vrlIntegrator.cpp:106-112) */
    std::vector<std::vector<uint32_t>> selectedVrls; std::vector<std::vector<float>> clusterWeight;
    std::vector<uint32_t> gcVrls, fallBackVrls; std::vector<float> gcWeight, fallBackWeight;
    bool haveClusters = false, haveFallback = false;
    std::vector<std::vector<uint32_t>> globalVrlsPerCluster;
    uint32_t nearTieSplits = 0;
    std::vector<uint8_t> columnFlagsOverride;     /* all-reduced zero / non-zero column flags (multi-GPU) */

    /* render lists */
    alvrl::DevBuf<uint32_t> dSlicePixels, dRepOffset; alvrl::DevBuf<uint4> dWork; alvrl::DevBuf<VrlRec> dRepRecs;
    alvrl::DevBuf<float4> dFb; alvrl::DevBuf<float> dRgb;
    uint32_t numWork = 0; bool renderListsDirty = true, pixelListsDirty = true;
    std::vector<uint4> workHost;      /* host mirror of dWork (slice-range selection) */

    /* sample streams */
    std::unique_ptr<alvrl::HostSampler> mainSampler;
    std::unique_ptr<alvrl::HostSampler> globalStream;   /* counter mode: the (CLUSTER, GLOBAL_ID) stream, kept for the lazy fallback */

    uint32_t numPixels() const { return cam.W * cam.H; }
    uint32_t K() const { return 2 * P.volVolSamples + P.volSurfSamples; }
};

namespace alvrl {
/* clustering.cu: Preprocessor::buildClusters on the device (Preprocessor.cpp:133-283) */
void build_clusters_device(alvrl_ctx *c, bool needFallback);
void column_nonzero_device(alvrl_ctx *c, std::vector<uint8_t> &flags);
float measure_fp32_peak_tflops();
/* slices_dev.cu: Preprocessor::getSlices on the device (Preprocessor.cpp:1200-1227,1349-1418) and its consumers */
bool build_slices_device(alvrl_ctx *c, const float *dPos, const float *dDir, uint32_t P, uint32_t targetNumSlices);
void slice_gather_rows_device(alvrl_ctx *c, const std::vector<uint32_t> &positions, std::vector<uint32_t> &rowPixel);
void slice_bucket_pixels_device(alvrl_ctx *c, const std::vector<uint32_t> &sliceStart, uint32_t total);
}
