/*
 * group.cu -- the multi-GPU entry points of include/alvrl.h (alvrl_group_*): N handles (one per GPU), one NCCL
 * communicator, and one call that renders a frame with the slices sharded over the ranks.
 *
 * What is exchanged: Preprocessor::cluster() classifies VRL columns as zero / non-zero over ALL rows of R
 * (Preprocessor.cpp:846-855) -> an N-byte all-reduce (MAX) of the per-rank flags, on device buffers; and the framebuffer
 * (each rank renders the pixels of its slices into a zeroed W x H x 4 buffer) -> ncclReduce(sum) to rank 0.  Everything else
 * of a slice -- its R rows, its Clustering object, its pixels -- is local to the rank that owns it.
 *
 * Two ways to form a group:
 *   alvrl_group_create_local : one process drives several GPUs (what a Mitsuba plugin is: a single process) --
 *                              ncclCommInitAll, one host thread per member during a frame;
 *   alvrl_group_create_rank  : one process per GPU (torchrun / MPI style) -- ncclCommInitRank with a unique id that the
 *                              host distributes by its own means.
 * NCCL is bound at run time (dlopen libnccl.so.2), so a single-GPU user needs no NCCL at all and a host that already
 * loaded NCCL (e.g. through PyTorch) shares that copy.
 */
#include <dlfcn.h>
#include <nccl.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "context.h"
#include "kernels.h"
#include "sharding.h"
#include "../../include/alvrl.h"

using namespace alvrl;

namespace alvrl { void column_nonzero_into(alvrl_ctx *c, uint8_t *dFlags); }

namespace {

thread_local std::string g_gerr;
int gfail(int code, const std::string &m) { g_gerr = m; return code; }

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char *names[] = {getenv("ALVRL_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { if (n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!lib) { err = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return false; }
#define ALVRL_NCCL_SYM(field, sym) *(void **) (&field) = dlsym(lib, sym); if (!field) { err = std::string("NCCL symbol missing: ") + sym; return false; }
        ALVRL_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") ALVRL_NCCL_SYM(CommInitRank, "ncclCommInitRank") ALVRL_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        ALVRL_NCCL_SYM(CommDestroy, "ncclCommDestroy") ALVRL_NCCL_SYM(AllReduce, "ncclAllReduce") ALVRL_NCCL_SYM(Reduce, "ncclReduce")
        ALVRL_NCCL_SYM(CommCount, "ncclCommCount") ALVRL_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ALVRL_NCCL_SYM
        return true;
    }
};
Nccl &nccl() { static Nccl n; return n; }

struct Member {
    alvrl_ctx *c = nullptr; int rank = 0; ncclComm_t comm = nullptr; bool owned = false;
    DevBuf<float4> fb; DevBuf<uint8_t> flags; DevBuf<float> rgb;
    uint32_t sliceBegin = 0, sliceEnd = 0;
    /* load balance: cost estimate per slice, corrected after every frame by the ranks' measured times (identical on every rank) */
    std::vector<double> sliceCost; uint32_t corrections = 0; DevBuf<float> times; cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
};

#define G_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) throw Error(ALVRL_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r_)); } while (0)
#define G_API(call) do { int rc_ = (call); if (rc_ != ALVRL_OK) throw Error(rc_, alvrl_last_error()); } while (0)

} // namespace

struct alvrl_group {
    int world = 1;
    std::vector<Member *> members;       /* local members (1 in the process-per-GPU form) */
    float msFrame = 0, msExchange = 0;
    ~alvrl_group() { for (Member *m : members) delete m; }
};

namespace {

/* one frame on one member: every rank builds the (deterministic) slices, takes its range, and renders its pixels */
double gnow_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void member_frame(alvrl_group *g, Member *m, float *rgbHost) {
    alvrl_ctx *c = m->c;
    ALVRL_CUDA(cudaSetDevice(c->device));
    const bool prof = getenv("ALVRL_PROFILE") != nullptr;
    double t0 = gnow_ms(), tPrev = t0;
    std::string log;
    auto lap = [&](const char *what) { if (prof) { cudaStreamSynchronize(c->stream); const double n = gnow_ms(); char b[64]; snprintf(b, sizeof(b), " %s %.1f", what, n - tPrev); log += b; tPrev = n; } };
    G_API(alvrl_build_slices(c));                                            /* Preprocessor::buildSlices: replicated, identical on every rank */
    const uint32_t S = c->numSlices();
    /* cost estimate per slice and the weights the ranges are cut on (sharding.h, shared with the CPU tests) */
    uint64_t totalPix = 0;
    for (uint32_t i = 0; i < S; i++) totalPix += c->sliceSize[i];
    const bool adaptive = g->world > 1 && !getenv("ALVRL_GROUP_STATIC");
    if (m->sliceCost.size() != S) initial_slice_costs(c->sliceSize.data(), S, m->sliceCost);
    std::vector<uint32_t> sizes(S);
    if (!adaptive || m->ev[0] == nullptr) {                                      /* first frame: the pixel counts themselves */
        for (uint32_t i = 0; i < S; i++) sizes[i] = c->sliceSize[i] + (uint32_t) (totalPix / 500u);
    } else cut_weights(m->sliceCost, sizes);
    balanced_slice_range(sizes.data(), S, g->world, m->rank, m->sliceBegin, m->sliceEnd);
    if (adaptive && m->ev[0] == nullptr) for (int k = 0; k < 4; k++) ALVRL_CUDA(cudaEventCreate(&m->ev[k]));
    if (adaptive) ALVRL_CUDA(cudaEventRecord(m->ev[0], c->stream));
    G_API(alvrl_set_slice_range(c, m->sliceBegin, m->sliceEnd));
    lap("slices");
    G_API(alvrl_sample_slice_mapping(c));
    G_API(alvrl_build_R(c));
    lap("mapping+R");
    if (adaptive) ALVRL_CUDA(cudaEventRecord(m->ev[1], c->stream));
    const uint32_t N = (uint32_t) c->vrlHost.size(), P = c->numPixels();
    if (g->world > 1) {                                                      /* zero / non-zero columns over ALL rows: OR across ranks, on the device */
        m->flags.alloc(N);
        column_nonzero_into(c, m->flags.p);
        G_NCCL(nccl().AllReduce(m->flags.p, m->flags.p, N, ncclUint8, ncclMax, m->comm, c->stream));
        std::vector<uint8_t> f(N);
        m->flags.download(f.data(), N, c->stream);
        G_API(alvrl_set_column_nonzero(c, f.data()));
    } else G_API(alvrl_set_column_nonzero(c, nullptr));
    lap("flags");
    if (adaptive) ALVRL_CUDA(cudaEventRecord(m->ev[2], c->stream));
    G_API(alvrl_build_clusters(c));
    lap("clusters");
    m->fb.alloc(P);
    ALVRL_CUDA(cudaMemsetAsync(m->fb.p, 0, (size_t) P * sizeof(float4), c->stream));
    G_API(alvrl_render_device(c, m->fb.p, c->stream));
    lap("render");
    if (adaptive) ALVRL_CUDA(cudaEventRecord(m->ev[3], c->stream));
    if (g->world > 1) G_NCCL(nccl().Reduce(m->fb.p, m->fb.p, (size_t) P * 4, ncclFloat, ncclSum, 0, m->comm, c->stream));
    if (m->rank == 0 && rgbHost) {
        m->rgb.alloc(3 * (size_t) P);
        launch_fb_to_rgb(m->fb.p, m->rgb.p, P, c->stream);
        c->stats.kernelLaunches++;
        m->rgb.download(rgbHost, 3 * (size_t) P, c->stream);
    }
    ALVRL_CUDA(cudaStreamSynchronize(c->stream));
    lap("reduce+image");
    if (adaptive) {
        /* what this rank's range cost (R + clusters + render; the waits in the two exchanges excluded), known to every rank
         * after an all-reduce of world floats; the estimates of a rank's slices are scaled by measured / predicted, damped */
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, m->ev[0], m->ev[1]); cudaEventElapsedTime(&b, m->ev[2], m->ev[3]);
        std::vector<float> t((size_t) g->world, 0.0f);
        t[m->rank] = a + b;
        m->times.upload(t, c->stream);
        G_NCCL(nccl().AllReduce(m->times.p, m->times.p, (size_t) g->world, ncclFloat, ncclSum, m->comm, c->stream));
        m->times.download(t.data(), (size_t) g->world, c->stream);
        if (correct_slice_costs(m->sliceCost, sizes.data(), g->world, t.data(), m->corrections)) m->corrections++;
    }
    if (prof) fprintf(stderr, "[alvrl group] rank %d slices [%u, %u):%s | total %.1f ms\n", m->rank, m->sliceBegin, m->sliceEnd, log.c_str(), gnow_ms() - t0);
}

} // namespace

extern "C" {

const char *alvrl_group_last_error(void) { return g_gerr.c_str(); }

int alvrl_group_unique_id(uint8_t id[ALVRL_GROUP_ID_BYTES]) {
    if (!nccl().load()) return gfail(ALVRL_ERR_CUDA, nccl().err);
    static_assert(sizeof(ncclUniqueId) <= ALVRL_GROUP_ID_BYTES, "ncclUniqueId does not fit ALVRL_GROUP_ID_BYTES");
    ncclUniqueId u;
    ncclResult_t r = nccl().GetUniqueId(&u);
    if (r != ncclSuccess) return gfail(ALVRL_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r));
    memset(id, 0, ALVRL_GROUP_ID_BYTES); memcpy(id, &u, sizeof(u));
    return ALVRL_OK;
}

int alvrl_group_create_rank(alvrl_handle h, int rank, int nranks, const uint8_t id[ALVRL_GROUP_ID_BYTES], alvrl_group_handle *out) {
    if (!h || !out || nranks < 1 || rank < 0 || rank >= nranks) return gfail(ALVRL_ERR_ARG, "alvrl_group_create_rank: bad arguments");
    alvrl_group *g = new alvrl_group();
    try {
        g->world = nranks;
        Member *m = new Member(); m->c = h; m->rank = rank; g->members.push_back(m);
        if (nranks > 1) {
            if (!id) throw Error(ALVRL_ERR_ARG, "alvrl_group_create_rank: a unique id is needed for more than one rank");
            if (!nccl().load()) throw Error(ALVRL_ERR_CUDA, nccl().err);
            ncclUniqueId u; memcpy(&u, id, sizeof(u));
            ALVRL_CUDA(cudaSetDevice(h->device));
            G_NCCL(nccl().CommInitRank(&m->comm, nranks, u, rank));
        }
    } catch (const Error &e) { delete g; return gfail(e.code, e.what()); }
    *out = g;
    return ALVRL_OK;
}

int alvrl_group_create_local(int ndev, const int *devices, const alvrl_params *p, alvrl_group_handle *out) {
    if (ndev < 1 || !devices || !p || !out) return gfail(ALVRL_ERR_ARG, "alvrl_group_create_local: bad arguments");
    alvrl_group *g = new alvrl_group();
    try {
        g->world = ndev;
        for (int i = 0; i < ndev; i++) {
            Member *m = new Member(); m->rank = i; m->owned = true; g->members.push_back(m);
            G_API(alvrl_create(devices[i], p, &m->c));
        }
        if (ndev > 1) {
            if (!nccl().load()) throw Error(ALVRL_ERR_CUDA, nccl().err);
            std::vector<ncclComm_t> comms(ndev);
            G_NCCL(nccl().CommInitAll(comms.data(), ndev, devices));
            for (int i = 0; i < ndev; i++) g->members[i]->comm = comms[i];
        }
    } catch (const Error &e) {
        for (Member *m : g->members) if (m->owned && m->c) alvrl_destroy(m->c);
        delete g; return gfail(e.code, e.what());
    }
    *out = g;
    return ALVRL_OK;
}

int alvrl_group_size(alvrl_group_handle g, int *world, int *local) {
    if (!g) return gfail(ALVRL_ERR_ARG, "null group");
    if (world) *world = g->world;
    if (local) *local = (int) g->members.size();
    return ALVRL_OK;
}

int alvrl_group_member(alvrl_group_handle g, int i, alvrl_handle *h, int *rank) {
    if (!g || i < 0 || i >= (int) g->members.size()) return gfail(ALVRL_ERR_ARG, "group member index out of range");
    if (h) *h = g->members[i]->c;
    if (rank) *rank = g->members[i]->rank;
    return ALVRL_OK;
}

int alvrl_group_comm_size(alvrl_group_handle g, int *n) {
    if (!g || !n) return gfail(ALVRL_ERR_ARG, "null argument");
    *n = 1;
    if (g->world > 1 && g->members[0]->comm) { ncclResult_t r = nccl().CommCount(g->members[0]->comm, n); if (r != ncclSuccess) return gfail(ALVRL_ERR_CUDA, "ncclCommCount failed"); }
    return ALVRL_OK;
}

int alvrl_group_get_range(alvrl_group_handle g, int i, uint32_t *b, uint32_t *e) {
    if (!g || i < 0 || i >= (int) g->members.size()) return gfail(ALVRL_ERR_ARG, "group member index out of range");
    *b = g->members[i]->sliceBegin; *e = g->members[i]->sliceEnd;
    return ALVRL_OK;
}

/* vrlIntegrator::preprocess (slices) + prepass + render pass of one frame over all ranks.  rgb_host (W*H*3 floats, [y][x][c]) is
 * filled on the process that holds rank 0; pass NULL elsewhere (or to leave the image in rank 0's device buffer). */
int alvrl_group_frame(alvrl_group_handle g, float *rgb_host) {
    if (!g) return gfail(ALVRL_ERR_ARG, "null group");
    for (Member *m : g->members) m->err.clear();
    if (g->members.size() == 1) {
        Member *m = g->members[0];
        try { member_frame(g, m, rgb_host); } catch (const std::exception &e) { m->err = e.what(); }
    } else {
        std::vector<std::thread> th;
        for (Member *m : g->members) th.emplace_back([g, m, rgb_host]() {
            try { member_frame(g, m, rgb_host); } catch (const std::exception &e) { m->err = e.what(); }
        });
        for (auto &t : th) t.join();
    }
    for (Member *m : g->members) if (!m->err.empty()) return gfail(ALVRL_ERR_CUDA, "rank " + std::to_string(m->rank) + ": " + m->err);
    return ALVRL_OK;
}

/* device pointer of a member's framebuffer after alvrl_group_frame (W*H float4; the reduced image on rank 0) */
int alvrl_group_framebuffer(alvrl_group_handle g, int i, void **rgba_device) {
    if (!g || i < 0 || i >= (int) g->members.size()) return gfail(ALVRL_ERR_ARG, "group member index out of range");
    *rgba_device = g->members[i]->fb.p;
    return ALVRL_OK;
}

void alvrl_group_destroy(alvrl_group_handle g) {
    if (!g) return;
    for (Member *m : g->members) {
        if (m->c) cudaSetDevice(m->c->device);
        m->fb.release(); m->flags.release(); m->rgb.release(); m->times.release();
        for (int k = 0; k < 4; k++) if (m->ev[k]) cudaEventDestroy(m->ev[k]);
        if (m->comm) nccl().CommDestroy(m->comm);
        if (m->owned && m->c) alvrl_destroy(m->c);
    }
    delete g;
}

} // extern "C"
