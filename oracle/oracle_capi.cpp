/*
 * oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY (see oracle_core.hpp).
 *
 * extern "C" surface of the CPU oracle: the orc_* functions mirror the alvrl_* entry points of
 * include/alvrl.h one to one (same arguments, same meaning) so that parity tests drive both
 * sides with the same script.  The top-level flow restates vrlIntegrator::preprocess / prepass /
 * Li (src/integrators/vrl/vrlIntegrator.cpp:237-356,386-599) and Preprocessor::buildSlices /
 * sampleSliceMapping / buildClusters (Preprocessor.cpp:133-283,1130-1193,1502-1525).
 */
#include "oracle_prep.hpp"
#include "../mitsuba-alvrl_b200/csrc/shapes.h"   /* the hand-over triangles of analytic shapes (see orc_add_rectangle) */
#include <thread>
#include <set>
#include <memory>
#include <chrono>
#include <cstdio>

using namespace orc;

namespace {

thread_local std::string g_err;

struct Ctx {
    alvrl_params P;
    Scene scene; Medium medium; Camera cam;
    bool haveMesh = false, haveMat = false, haveMedium = false, haveCam = false, haveVrls = false;
    std::vector<VRL> vrls; uint64_t particleCount = 0;
    std::vector<Ray> rays; std::vector<Intersection> hits;       // per pixel index y + H*x
    /* area emitter of the VRL tracer: the triangles of its shape, TriMesh::prepareSamplingTable (trimesh.cpp:388-403) */
    std::vector<uint32_t> emTris; std::vector<Float> emCdf; Float emInvArea = 0; Spec emPower, emRadiance; bool haveEmitter = false;
    std::vector<uint8_t> triIsEmitter;                            // Shape::isEmitter of the triangle's shape
    bool havePrimary = false;
    struct ChainSeg { Ray ray; Intersection its; Spec weight; bool inMedium; uint32_t code; };
    std::vector<std::vector<ChainSeg>> chains; bool haveChains = false; bool anyDelta = false;   // vrlIntegrator.cpp:445-511
    std::vector<uint32_t> pixelToSlice; std::vector<SliceData> slices; bool haveSlices = false;
    std::vector<uint32_t> rowOffset, rowPixel; bool haveRows = false;
    std::vector<Float> sliceUndersampling; Float globalPixelUndersampling = -1;
    std::vector<std::set<std::pair<uint32_t, Float>>> localities;      // m_localities: neighbour slices (index, 6-D distance), Preprocessor.h
    std::vector<VrlContribution> R; bool haveR = false;
    std::vector<std::vector<uint32_t>> selectedVrls; std::vector<std::vector<Float>> clusterWeight;
    std::vector<uint32_t> gcVrls, fallBackVrls; std::vector<Float> gcWeight, fallBackWeight; bool haveClusters = false;
    std::unique_ptr<Sampler> mainSampler;
    const float *tape = nullptr; uint64_t tapeLen = 0; std::vector<float> tapeStore;
    int threads = 1;
    bool noVisibility = false;
    Float grazeTol = 0; std::vector<uint8_t> Rgraze;                   // test instrumentation: fragile visibility decisions per R entry
    uint32_t sliceBegin = 0, sliceEnd = 0xffffffffu;
    alvrl_stats stats;
    std::vector<uint8_t> columnFlagsOverride;      // all-reduced zero / non-zero column flags (multi-rank runs)
    uint32_t nearTieSplits = 0;
    uint64_t clusterVarianceSteps = 0, clusterSplits = 0;

    Ctx() { memset(&stats, 0, sizeof(stats)); }
    uint32_t P_() const { return cam.W * cam.H; }
    uint32_t K() const { return 2 * P.volVolSamples + P.volSurfSamples; }
    Sampler *newStream() {
        if (P.rngMode == ALVRL_RNG_MODE_SFMT) return new SfmtSampler(P.seed);
        return new CounterSampler(P.seed);
    }
    IntegratorCore core() {
        IntegratorCore c; c.scene = &scene; c.medium = &medium; c.volVolSamples = P.volVolSamples;
        c.volSurfSamples = P.volSurfSamples; c.shortVrls = P.shortVrls != 0; c.noVisibility = noVisibility;
        c.grazeTol = grazeTol;
        return c;
    }
};

int seterr(int code, const std::string &m) { g_err = m; return code; }
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception &e) { return seterr(ALVRL_ERR_ARG, e.what()); } return ALVRL_OK;

void tracePrimary(Ctx *c) {
    if (c->havePrimary) return;
    uint32_t W = c->cam.W, H = c->cam.H;
    c->rays.resize((size_t) W * H); c->hits.resize((size_t) W * H);
    auto work = [&](uint32_t x0, uint32_t x1) {
        for (uint32_t i = x0; i < x1; i++)
            for (uint32_t j = 0; j < H; j++) {
                size_t idx = (size_t) i * H + j;
                c->rays[idx] = c->cam.sampleRay(i + 0.5f, j + 0.5f);       // Preprocessor.cpp:1144-1146
                c->scene.rayIntersect(c->rays[idx], c->hits[idx]);
            }
    };
    int T = std::max(1, c->threads);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work, (uint32_t) ((uint64_t) W * t / T), (uint32_t) ((uint64_t) W * (t + 1) / T));
    for (auto &t : th) t.join();
    c->havePrimary = true;
}

/* The specular chains of LiInternal (vrlIntegrator.cpp:445-511) below every camera segment whose surface has delta components:
 * the segments LiInternal recurses into, in its own order (component 0 and everything below it before component 1), each with
 * the `weight` its VRL contributions are multiplied by.  The roulette draw of a branch (485: rRec.nextSample1D()) comes from the
 * counter stream of (pixel, path code), path code = 2 * parent's + component, the camera segment being 1. */
static const int kChainMaxDepth = 30;
static void chainBelow(Ctx *c, uint32_t pixel, const Ray &ray, const Intersection &its, bool inMedium, Spec throughputWithEtaSq,
                       Spec weight, int depth, uint32_t code, std::vector<Ctx::ChainSeg> &out) {
    const uint32_t bits = c->scene.matBits[its.material];
    if (!(bits & ALVRL_BSDF_DELTA)) return;                                  // 447-448: no specular chains
    if (c->scene.optics.size() != c->scene.matBits.size()) throw std::runtime_error("delta material without alvrl_set_material_optics");
    Spec transmittance(1.0f);                                                // 450-459
    if (inMedium) { MediumSamplingRecord mRec; c->medium.eval(Ray(ray.o, ray.d, 0, its.t), mRec); transmittance = mRec.transmittance; }
    if (transmittance.isZero()) return;
    if (depth > kChainMaxDepth) return;
    const HitFrame fr = c->scene.hitFrame(its);
    const V3 wi = fr.toLocal(-ray.d);                                        // skdtree.h:427
    const int compCount = (bits & ALVRL_BSDF_DIELECTRIC) ? 2 : 1;
    for (int i = 0; i < compCount; i++) {                                    // 467-504
        V3 wo; Float eta = 1;
        Spec bsdfWeight = sampleDelta(bits, c->scene.optics[its.material], wi, i, wo, eta);
        if (bsdfWeight.isZero()) continue;
        Spec throughputWithEtaSq2 = throughputWithEtaSq * transmittance * bsdfWeight * (eta * eta);
        Float maxRRprob = depth >= c->P.specularForcedRRdepth ? (Float) 0.98 : (Float) 1;
        Float rrProb = std::min(maxRRprob, throughputWithEtaSq2.max());
        const uint32_t childCode = code * 2u + (uint32_t) i;
        if (rrProb <= 0) continue;
        if (rrProb < 1) {
            const Float u = alvrl_rng_uniform(alvrl_rng_key(c->P.seed, ALVRL_RNG_CHAIN, pixel, childCode), 0);
            if (u > rrProb) continue;
        }
        throughputWithEtaSq2 /= rrProb;
        Ray ray2(its.p, fr.toWorld(wo), Epsilon, std::numeric_limits<Float>::infinity());   // 490: RayDifferential(p, d, time)
        bool inMedium2 = inMedium;
        if (bits & ALVRL_MAT_TRANSITION)                                                      // 491-493, records.inl:81-86
            inMedium2 = dot(ray2.d, its.n) > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
        Spec weightSpec = weight * transmittance * bsdfWeight / rrProb;                        // 500
        Intersection its2;
        if (!c->scene.rayIntersect(ray2, its2)) continue;                                      // 416-423: an infinite segment adds nothing
        out.push_back(Ctx::ChainSeg{ray2, its2, weightSpec, inMedium2, childCode});
        chainBelow(c, pixel, ray2, its2, inMedium2, throughputWithEtaSq2, weightSpec, depth + 1, childCode, out);
    }
}
void traceChains(Ctx *c) {
    tracePrimary(c);
    if (c->haveChains) return;
    const uint32_t P = c->P_();
    c->chains.assign(P, {});
    c->anyDelta = false;
    for (uint32_t b : c->scene.matBits) if (b & ALVRL_BSDF_DELTA) c->anyDelta = true;
    if (c->anyDelta)
        for (uint32_t pix = 0; pix < P; pix++)
            if (c->hits[pix].isValid())       // Li: rRec.newQuery(ESensorRay, sensor medium): depth 1, the camera sits in the medium
                chainBelow(c, pix, c->rays[pix], c->hits[pix], true, Spec(c->P.initialSpecularThroughput), Spec(1.0f), 1, 1u, c->chains[pix]);
    c->haveChains = true;
}

/* getVRLContributions for one row, vrlIntegrator.cpp:792-825 with vrlContributions != NULL */
void buildRow(Ctx *c, IntegratorCore &core, Sampler *sampler, uint32_t row, VrlContribution *out) {
    uint32_t N = c->vrls.size();
    uint32_t pixel = c->rowPixel[row];
    const Ray &ray = c->rays[pixel];
    const Intersection &its = c->hits[pixel];
    for (uint32_t v = 0; v < N; v++) out[v] = VrlContribution{0, 0};
    if (!its.isValid()) return;                                             // vrlIntegrator.cpp:418-423
    if (c->medium.sigmaS.isZero()) return;                                   // 795-797
    for (int s = 0; s < c->P.Rsamples; s++) {                                // LiInternal samples loop, 427-442 (quirk B3)
        for (uint32_t v = 0; v < N; v++) {
            Float normalization = 1.0 / c->particleCount;
            Float contribution, variance;
            if (s == 0) sampler->setContext(ALVRL_RNG_R, row, v);
            core.grazed = false;
            core.integrateVRL(ray, its, c->vrls[v], sampler, &contribution, &variance, Spec(1.0f));
            if (core.grazed && !c->Rgraze.empty()) c->Rgraze[(size_t) row * N + v] = 1;
            out[v].mean += contribution * normalization;
            out[v].var += variance * normalization * normalization;
        }
    }
    /* the chain below the camera segment (LiInternal 462-505 with vrlContributions != NULL): every segment inside the medium
     * adds its weighted contributions; its draws come from the stream of (row + (ordinal << 24), vrl) */
    if (c->anyDelta) {
        const std::vector<Ctx::ChainSeg> &ch = c->chains[pixel];
        for (size_t e = 0; e < ch.size(); e++) {
            if (!ch[e].inMedium) continue;
            for (uint32_t v = 0; v < N; v++) {
                Float normalization = 1.0 / c->particleCount;
                Float contribution, variance;
                sampler->setContext(ALVRL_RNG_R, row + ((uint32_t) (e + 1) << 24), v);
                core.integrateVRL(ch[e].ray, ch[e].its, c->vrls[v], sampler, &contribution, &variance, ch[e].weight);
                out[v].mean += contribution * normalization;
                out[v].var += variance * normalization * normalization;
            }
        }
    }
}

} // namespace

extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }

void orc_params_default(alvrl_params *p) {
    memset(p, 0, sizeof(*p));
    p->shortVrls = 1; p->vrlTargetNum = 500; p->maxParticleDepth = -1; p->specularForcedRRdepth = 100;
    p->initialSpecularThroughput = 20; p->volVolSamples = 2; p->volSurfSamples = 2; p->globalCluster = 0;
    p->globalUndersampling = -1; p->localRefinement = 1; p->localUndersampling = -1; p->fallBackUndersampling = 5;
    p->targetNumSlices = 100; p->targetPixelUndersampling = 64; p->sliceCurvatureFactor = 0.5f;
    p->neighbourCount = 0; p->neighbourWeight = 0; p->Rsamples = 1; p->depthCorrection = 1; p->maxPasses = 1;
    p->rngMode = ALVRL_RNG_MODE_COUNTER; p->seed = 0; p->anyHitShadowRays = 1; p->workerCount = 1; p->rrDepth = 5;
}

int orc_create(int, const alvrl_params *p, void **out) {
    /* vrlIntegrator ctor checks, vrlIntegrator.cpp:149-156; Preprocessor ctor, Preprocessor.cpp:36-38 */
    if (p->volVolSamples != 0 && p->volVolSamples < 2) return seterr(ALVRL_ERR_ARG, "Need at least 2 volVolSamples for variance estimate");
    if (p->volSurfSamples != 0 && p->volSurfSamples < 2) return seterr(ALVRL_ERR_ARG, "Need at least 2 volSurfSamples for variance estimate");
    if (p->targetNumSlices < 1) return seterr(ALVRL_ERR_ARG, "Invalid target number of slices!");
    Ctx *c = new Ctx();
    c->P = *p;
    c->mainSampler.reset(c->newStream());
    *out = c;
    return ALVRL_OK;
}
void orc_destroy(void *h) { delete (Ctx *) h; }
int orc_set_threads(void *h, int n) { ((Ctx *) h)->threads = std::max(1, n); return ALVRL_OK; }
int orc_set_no_visibility(void *h, int on) { ((Ctx *) h)->noVisibility = on != 0; return ALVRL_OK; }
/* test instrumentation: build_R also records which entries had a shadow ray whose occlusion decision is within tol of flipping */
int orc_set_graze_tolerance(void *h, float tol) { ((Ctx *) h)->grazeTol = tol; return ALVRL_OK; }
int orc_get_R_graze(void *h, uint32_t r0, uint32_t r1, uint8_t *out) {
    Ctx *c = (Ctx *) h;
    if (!c->haveR || c->Rgraze.empty()) return seterr(ALVRL_ERR_STATE, "set_graze_tolerance and build_R first");
    size_t N = c->vrls.size();
    memcpy(out, &c->Rgraze[(size_t) r0 * N], (size_t) (r1 - r0) * N);
    return ALVRL_OK;
}

int orc_set_mesh(void *h, const float *v, uint32_t nv, const uint32_t *tris, uint32_t nt, const uint32_t *mat) {
    Ctx *c = (Ctx *) h;
    c->scene.verts.resize(nv);
    for (uint32_t i = 0; i < nv; i++) c->scene.verts[i] = V3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    c->scene.tris.assign(tris, tris + 3 * (size_t) nt);
    c->scene.triMat.assign(mat, mat + nt);
    for (size_t i = 0; i < 3 * (size_t) nt; i++) if (tris[i] >= nv) return seterr(ALVRL_ERR_ARG, "triangle index out of range");
    c->scene.finalize();
    c->haveMesh = true; c->havePrimary = false;
    return ALVRL_OK;
}
/* alvrl_add_rectangle / alvrl_add_sphere: the triangles an analytic shape is handed over as.  NOT a restatement -- there is
 * nothing in the reference to restate (it intersects the analytic shapes): both sides append the triangles of csrc/shapes.h,
 * and tests/test_shapes_cpu.py checks those against the shapes' definitions (src/shapes/rectangle.cpp, src/shapes/sphere.cpp)
 * and, through this oracle's ray caster, against the analytic intersections. */
static int appendShape(Ctx *c, const std::vector<float> &v, const std::vector<uint32_t> &t, uint32_t material, uint32_t *first, uint32_t *count) {
    const uint32_t base = (uint32_t) c->scene.verts.size(), before = (uint32_t) c->scene.triMat.size();
    for (size_t i = 0; i < v.size(); i += 3) c->scene.verts.push_back(V3(v[i], v[i + 1], v[i + 2]));
    for (uint32_t i : t) c->scene.tris.push_back(base + i);
    c->scene.triMat.resize(c->scene.tris.size() / 3, material);
    if (first) *first = before;
    if (count) *count = (uint32_t) (t.size() / 3);
    c->scene.finalize();
    c->haveMesh = true; c->havePrimary = false;
    return ALVRL_OK;
}
int orc_add_rectangle(void *h, const float toWorld[16], int flipNormals, uint32_t material, uint32_t *firstTriangle) {
    ORC_TRY
    std::vector<float> v; std::vector<uint32_t> t;
    alvrl::tessellate_rectangle(toWorld, flipNormals != 0, v, t);
    return appendShape((Ctx *) h, v, t, material, firstTriangle, nullptr);
    ORC_CATCH
}
int orc_add_sphere(void *h, const float center[3], float radius, int flipNormals, uint32_t thetaSteps, uint32_t material,
                   uint32_t *firstTriangle, uint32_t *triangleCount) {
    ORC_TRY
    std::vector<float> v; std::vector<uint32_t> t;
    alvrl::tessellate_sphere(center, radius, flipNormals != 0, thetaSteps ? thetaSteps : 64u, v, t);
    return appendShape((Ctx *) h, v, t, material, firstTriangle, triangleCount);
    ORC_CATCH
}
int orc_set_materials(void *h, const float *albedo, const uint32_t *bits, uint32_t nm) {
    Ctx *c = (Ctx *) h;
    c->scene.albedo.resize(nm); c->scene.matBits.assign(bits, bits + nm);
    for (uint32_t i = 0; i < nm; i++) c->scene.albedo[i] = Spec(albedo[3 * i], albedo[3 * i + 1], albedo[3 * i + 2]);
    c->haveMat = true;
    return ALVRL_OK;
}
int orc_set_material_optics(void *h, const float *optics, uint32_t nm) {
    Ctx *c = (Ctx *) h;
    if (!c->haveMat || nm != c->scene.matBits.size()) return seterr(ALVRL_ERR_ARG, "set_material_optics: set_materials first, same count");
    c->scene.optics.resize(nm);
    for (uint32_t i = 0; i < nm; i++) memcpy(c->scene.optics[i].v, optics + 12 * (size_t) i, 12 * sizeof(float));
    c->havePrimary = false; c->haveChains = false;
    return ALVRL_OK;
}
int orc_set_extra_bounds(void *h, const float *p, uint32_t n) {
    Ctx *c = (Ctx *) h;
    c->scene.extraBounds.clear();
    for (uint32_t i = 0; i < n; i++) c->scene.extraBounds.push_back(V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    if (c->haveMesh) c->scene.finalize();
    return ALVRL_OK;
}
int orc_set_medium_homogeneous(void *h, const float a[3], const float s[3], float w, int32_t phase, float g) {
    Ctx *c = (Ctx *) h;
    c->medium.setHomogeneous(a, s, w); c->medium.phaseType = phase; c->medium.g = g;
    c->haveMedium = true;
    return ALVRL_OK;
}
int orc_set_medium_grid(void *h, const float *density, const int32_t res[3], const float mn[3], const float mx[3], float scale,
                        const float albedo[3], const float sBase[3], int32_t phase, float g) {
    Ctx *c = (Ctx *) h;
    int r[3] = {res[0], res[1], res[2]};
    c->medium.setGrid(density, r, V3(mn[0], mn[1], mn[2]), V3(mx[0], mx[1], mx[2]), scale, albedo, sBase);
    c->medium.phaseType = phase; c->medium.g = g;
    c->haveMedium = true;
    return ALVRL_OK;
}
/* GridDataSource::loadFromFile (src/volume/gridvolume.cpp:217-287) in front of orc_set_medium_grid: "VOL", version 3, type,
 * resolution, channels, the AABB (unless `min` / `max` were given, 112-117, 272-280), the voxels from float 12 on (285).  The
 * density is looked up with lookupFloat (337-389): EFloat32 voxels as they are, EUInt8 voxels through m_densityMap
 * (212-215) -- resolved here once per voxel instead of once per lookup, the eight corner values are the same floats. */
int orc_set_medium_grid_file(void *h, const char *path, const float *mn, const float *mx, float scale,
                             const float albedo[3], const float sBase[3], int32_t phase, float g) {
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return seterr(ALVRL_ERR_IO, "cannot open volume data file");
    std::vector<unsigned char> bytes;
    unsigned char buf[65536]; size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + got);
    fclose(f);
    if (bytes.size() < 48) return seterr(ALVRL_ERR_IO, "Encountered an invalid volume data file (truncated header)");
    if (bytes[0] != 'V' || bytes[1] != 'O' || bytes[2] != 'L') return seterr(ALVRL_ERR_ARG, "Encountered an invalid volume data file (incorrect header identifier)");
    if (bytes[3] != 3) return seterr(ALVRL_ERR_ARG, "Encountered an invalid volume data file (incorrect file version)");
    int32_t head[5]; memcpy(head, &bytes[4], 20);                       /* little-endian host assumed (test infrastructure) */
    float box[6]; memcpy(box, &bytes[24], 24);
    const int32_t type = head[0], res[3] = {head[1], head[2], head[3]}, channels = head[4];
    if (type == 2) return seterr(ALVRL_ERR_UNSUPPORTED, "Error: float16 volumes are not yet supported!");
    if (type != 1 && type != 3) return seterr(type == 4 ? ALVRL_ERR_UNSUPPORTED : ALVRL_ERR_ARG, "Encountered a volume data file of unknown type");
    if (channels != 1) return seterr(channels == 3 ? ALVRL_ERR_UNSUPPORTED : ALVRL_ERR_ARG, "only one-channel density volumes");
    if (res[0] < 1 || res[1] < 1 || res[2] < 1) return seterr(ALVRL_ERR_ARG, "Encountered an invalid volume data file (resolution out of range)");
    const size_t n = (size_t) res[0] * res[1] * res[2];
    if (bytes.size() < 48 + n * (type == 1 ? 4 : 1)) return seterr(ALVRL_ERR_IO, "Encountered an invalid volume data file (fewer voxels than the header announces)");
    std::vector<float> density(n);
    if (type == 1) memcpy(density.data(), &bytes[48], 4 * n);
    else {
        Float densityMap[256];
        for (int i = 0; i < 255; i++) densityMap[i] = i / 255.0f;
        densityMap[255] = 1.0f;
        for (size_t i = 0; i < n; i++) density[i] = densityMap[bytes[48 + i]];
    }
    if ((mn == nullptr) != (mx == nullptr)) return seterr(ALVRL_ERR_ARG, "give both bbox_min and bbox_max, or neither");
    return orc_set_medium_grid(h, density.data(), res, mn ? mn : box, mx ? mx : box + 3, scale, albedo, sBase, phase, g);
}
int orc_set_camera(void *h, const float s2c[16], const float c2w[16], uint32_t W, uint32_t H, float nearClip, float farClip) {
    Ctx *c = (Ctx *) h;
    memcpy(c->cam.s2c, s2c, 64); memcpy(c->cam.c2w, c2w, 64);
    c->cam.W = W; c->cam.H = H; c->cam.nearClip = nearClip; c->cam.farClip = farClip;
    c->haveCam = true; c->havePrimary = false;
    return ALVRL_OK;
}
int orc_set_vrls(void *h, const float *s, const float *e, const float *p, uint32_t n, uint64_t particleCount) {
    Ctx *c = (Ctx *) h;
    c->vrls.clear();
    for (uint32_t i = 0; i < n; i++) {                                     // vrlVector::put, VRL.h:148-158
        VRL v; v.start = V3(s[3 * i], s[3 * i + 1], s[3 * i + 2]); v.end = V3(e[3 * i], e[3 * i + 1], e[3 * i + 2]);
        v.power = Spec(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        if (!v.power.isValid()) return seterr(ALVRL_ERR_ARG, "invalid parsed VRL power");     // VRL.h:51-53
        if (c->haveMedium && c->medium.sigmaS.isZero()) continue;
        if (v.power.isZero()) continue;
        if (distance(v.start, v.end) == 0) continue;
        c->vrls.push_back(v);
    }
    c->particleCount = particleCount ? particleCount : c->vrls.size();
    c->haveVrls = true;
    return ALVRL_OK;
}
int orc_get_num_vrls(void *h, uint32_t *n) { *n = ((Ctx *) h)->vrls.size(); return ALVRL_OK; }
int orc_set_sample_tape(void *h, const float *tape, uint64_t n) {
    Ctx *c = (Ctx *) h;
    if (!tape) { c->tape = nullptr; c->tapeLen = 0; c->tapeStore.clear(); return ALVRL_OK; }
    c->tapeStore.assign(tape, tape + n); c->tape = c->tapeStore.data(); c->tapeLen = n;
    return ALVRL_OK;
}

/* Preprocessor::buildSlices, Preprocessor.cpp:1130-1193 */
int orc_build_slices(void *h) {
    Ctx *c = (Ctx *) h;
    if (!(c->haveMesh && c->haveCam && c->haveMat)) return seterr(ALVRL_ERR_STATE, "scene incomplete");
    ORC_TRY
    double t0 = now_ms();
    tracePrimary(c);
    uint32_t P = c->P_();
    std::vector<V3> gather(P), dirs(P);
    Float directionScale = distance(c->scene.sceneAABB.min, c->scene.sceneAABB.max) / 8 * c->P.sliceCurvatureFactor;
    const Float nan = std::numeric_limits<Float>::quiet_NaN();
    for (uint32_t i = 0; i < P; i++) {
        if (c->hits[i].isValid()) { gather[i] = c->hits[i].p; dirs[i] = directionScale * c->hits[i].n; }
        else { gather[i] = V3(nan); dirs[i] = V3(nan); }
    }
    c->pixelToSlice = SliceBuilder::getSlices(gather, dirs, c->P.targetNumSlices, c->slices);
    c->haveSlices = true; c->haveRows = false; c->haveR = false; c->haveClusters = false;
    c->stats.msSlices = (float) (now_ms() - t0);
    c->stats.numSlices = c->slices.size();
    ORC_CATCH
}
int orc_build_slices_from_gather(void *h, const float *pos, const float *dir) {
    Ctx *c = (Ctx *) h;
    if (!c->haveCam) return seterr(ALVRL_ERR_STATE, "camera not set");
    ORC_TRY
    uint32_t P = c->P_();
    std::vector<V3> gather(P), dirs(P);
    for (uint32_t i = 0; i < P; i++) { gather[i] = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]); dirs[i] = V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]); }
    c->pixelToSlice = SliceBuilder::getSlices(gather, dirs, c->P.targetNumSlices, c->slices);
    c->haveSlices = true; c->haveRows = false; c->haveR = false; c->haveClusters = false;
    c->stats.numSlices = c->slices.size();
    ORC_CATCH
}
int orc_get_gather_points(void *h, float *pos, float *dir) {
    Ctx *c = (Ctx *) h;
    tracePrimary(c);
    Float directionScale = distance(c->scene.sceneAABB.min, c->scene.sceneAABB.max) / 8 * c->P.sliceCurvatureFactor;
    const Float nan = std::numeric_limits<Float>::quiet_NaN();
    for (uint32_t i = 0; i < c->P_(); i++) {
        V3 g(nan), d(nan);
        if (c->hits[i].isValid()) { g = c->hits[i].p; d = directionScale * c->hits[i].n; }
        pos[3 * i] = g.x; pos[3 * i + 1] = g.y; pos[3 * i + 2] = g.z; dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
    }
    return ALVRL_OK;
}

/* Preprocessor::sampleSliceMapping, Preprocessor.cpp:1502-1525 */
/* Preprocessor::buildLocalities, Preprocessor.cpp:1241-1293: the neighbourCount slices closest in the 6-D (position centroid,
 * scaled direction centroid) distance, found by a replace-the-current-maximum scan.  Kept as written: `maxInd` is NOT reset
 * between slices (1263), and a slot that no distance ever beat (NaN centroids of one-pixel slices, SURVEY B-list) keeps the
 * entry of the previous slice.  The reference's two scratch arrays are uninitialised VLAs; they start zeroed / infinite here.
 * With neighbourCount = 0 the reference reads distances[0] of an empty array (undefined, result unused): nothing to restate. */
static void buildLocalities(Ctx *c) {
    const uint32_t S = (uint32_t) c->slices.size(), nc = (uint32_t) std::max(0, c->P.neighbourCount);
    c->localities.assign(S, {});
    auto dist = [&](uint32_t i, uint32_t j) {
        return SliceBuilder::sliceDistance(c->slices[i].positionCentroid, c->slices[i].directionCentroid,
                                           c->slices[j].positionCentroid, c->slices[j].directionCentroid);
    };
    if (S <= nc) {                                                          // 1245-1258: everybody is everybody's neighbour
        for (uint32_t i = 0; i < S; i++) for (uint32_t j = 0; j < S; j++) if (i != j) c->localities[i].insert(std::make_pair(j, dist(i, j)));
        return;
    }
    if (nc == 0) return;
    std::vector<Float> distances(nc, std::numeric_limits<Float>::infinity()); std::vector<uint32_t> indices(nc, 0);
    uint32_t maxInd = 0;
    for (uint32_t i = 0; i < S; i++) {
        for (uint32_t j = 0; j < nc; j++) distances[j] = std::numeric_limits<Float>::infinity();
        for (uint32_t j = 0; j < S; j++) {
            if (i == j) continue;
            const Float d = dist(i, j);
            if (d < distances[maxInd]) {
                distances[maxInd] = d; indices[maxInd] = j;
                for (uint32_t k = 0; k < nc; k++) if (distances[k] > distances[maxInd]) maxInd = k;
            }
        }
        for (uint32_t x = 0; x < nc; x++) c->localities[i].insert(std::make_pair(indices[x], distances[x]));
    }
}

int orc_sample_slice_mapping(void *h) {
    Ctx *c = (Ctx *) h;
    if (!c->haveSlices) return seterr(ALVRL_ERR_STATE, "build_slices first");
    ORC_TRY
    double t0 = now_ms();
    size_t S = c->slices.size();
    c->rowOffset.assign(S + 1, 0); c->rowPixel.clear(); c->sliceUndersampling.resize(S);
    size_t totalPix = 0, totalRep = 0;
    for (size_t i = 0; i < S; i++) {
        c->mainSampler->setContext(ALVRL_RNG_SLICEMAP, (uint32_t) i, 0);
        std::vector<uint32_t> px = SliceBuilder::sampleRepresentativePixels(c->slices[i], c->P.targetPixelUndersampling, c->mainSampler.get());
        c->rowPixel.insert(c->rowPixel.end(), px.begin(), px.end());
        c->rowOffset[i + 1] = c->rowPixel.size();
        c->sliceUndersampling[i] = ((Float) px.size()) / c->slices[i].gatherIdx.size();
        totalRep += px.size(); totalPix += c->slices[i].gatherIdx.size();
    }
    buildLocalities(c);                                                     // 1517
    c->globalPixelUndersampling = ((Float) totalRep) / totalPix;
    c->haveRows = true; c->haveR = false; c->haveClusters = false;
    c->stats.msSliceMapping = (float) (now_ms() - t0);
    c->stats.numRows = c->rowPixel.size();
    ORC_CATCH
}

static int buildR_impl(Ctx *c, std::vector<float> *recordTape) {
    if (!c->haveRows || !c->haveVrls || !c->haveMedium) return seterr(ALVRL_ERR_STATE, "sample_slice_mapping / set_vrls / set_medium first");
    ORC_TRY
    double t0 = now_ms();
    traceChains(c);
    uint32_t N = c->vrls.size(), G = c->rowPixel.size(), S = c->slices.size();
    c->R.assign((size_t) G * N, VrlContribution{0, 0});
    if (c->grazeTol > 0) c->Rgraze.assign((size_t) G * N, 0); else c->Rgraze.clear();
    uint32_t sb = std::min(c->sliceBegin, S), se = std::min(c->sliceEnd, S);
    uint64_t shadow = 0;
    if (c->tape) {
        int T = std::max(1, c->threads);
        std::vector<std::thread> th; std::vector<uint64_t> sh(T, 0);
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
            TapeSampler smp(c->tape, c->tapeLen, c->K() * c->P.Rsamples, N);
            IntegratorCore core = c->core();
            /* rows are independent in the addressed streams: split the row range evenly over the threads */
            const uint32_t rb = c->rowOffset[sb], re = c->rowOffset[se];
            const uint32_t r0 = rb + (uint64_t) (re - rb) * t / T, r1 = rb + (uint64_t) (re - rb) * (t + 1) / T;
            for (uint32_t row = r0; row < r1; row++) buildRow(c, core, &smp, row, &c->R[(size_t) row * N]);
            sh[t] = core.shadowRays;
        });
        for (auto &t : th) t.join();
        for (uint64_t s : sh) shadow += s;
    } else if (c->P.rngMode == ALVRL_RNG_MODE_SFMT || recordTape) {
        /* reference stream order: prepass loops (workerCount == 1, vrlIntegrator.cpp:322-333) or one
         * cloned sampler per Rbuilder over contiguous slice ranges (305-321, 1048-1051) */
        int w = std::max(1, c->P.workerCount);
        std::vector<std::unique_ptr<Sampler>> clones;
        if (w > 1) for (int i = 0; i < w; i++) clones.emplace_back(c->mainSampler->clone());
        std::vector<std::thread> th; std::vector<uint64_t> sh(w, 0);
        auto work = [&](int id) {
            Sampler *smp = w > 1 ? clones[id].get() : c->mainSampler.get();
            std::unique_ptr<RecordingSampler> rec;
            if (recordTape) { rec.reset(new RecordingSampler(smp, *recordTape, c->K() * c->P.Rsamples, N)); smp = rec.get(); }
            IntegratorCore core = c->core();
            uint32_t s0 = ((uint64_t) id * S) / w, s1 = ((uint64_t) (id + 1) * S) / w;
            for (uint32_t row = c->rowOffset[s0]; row < c->rowOffset[s1]; row++) buildRow(c, core, smp, row, &c->R[(size_t) row * N]);
            sh[id] = core.shadowRays;
        };
        if (w > 1) { for (int i = 0; i < w; i++) th.emplace_back(work, i); for (auto &t : th) t.join(); }
        else work(0);
        for (uint64_t s : sh) shadow += s;
    } else {
        int T = std::max(1, c->threads);
        std::vector<std::thread> th; std::vector<uint64_t> sh(T, 0);
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
            CounterSampler smp(c->P.seed);
            IntegratorCore core = c->core();
            /* rows are independent in the addressed streams: split the row range evenly over the threads */
            const uint32_t rb = c->rowOffset[sb], re = c->rowOffset[se];
            const uint32_t r0 = rb + (uint64_t) (re - rb) * t / T, r1 = rb + (uint64_t) (re - rb) * (t + 1) / T;
            for (uint32_t row = r0; row < r1; row++) buildRow(c, core, &smp, row, &c->R[(size_t) row * N]);
            sh[t] = core.shadowRays;
        });
        for (auto &t : th) t.join();
        for (uint64_t s : sh) shadow += s;
    }
    c->haveR = true; c->haveClusters = false;
    c->stats.msBuildR = (float) (now_ms() - t0);
    c->stats.pairsPreprocess += (uint64_t) (c->rowOffset[se] - c->rowOffset[sb]) * N * c->P.Rsamples;
    c->stats.shadowRays += shadow;
    c->stats.numVrls = N;
    ORC_CATCH
}
int orc_build_R(void *h) { return buildR_impl((Ctx *) h, nullptr); }
/* build R in this context's stream order and record every uniform into the fixed-slot tape layout */
int orc_build_R_record_tape(void *h, float *tapeOut, uint64_t n) {
    Ctx *c = (Ctx *) h;
    uint64_t need = (uint64_t) c->rowPixel.size() * c->vrls.size() * c->K() * c->P.Rsamples;
    if (n < need) return seterr(ALVRL_ERR_ARG, "tape buffer too small");
    if (c->P.workerCount > 1) return seterr(ALVRL_ERR_ARG, "record with workerCount == 1");
    std::vector<float> tape(need, 0.5f);
    int rc = buildR_impl(c, &tape);
    memcpy(tapeOut, tape.data(), need * sizeof(float));
    return rc;
}

/* Preprocessor::buildClusters, Preprocessor.cpp:133-283,838-898 */
int orc_build_clusters(void *h) {
    Ctx *c = (Ctx *) h;
    if (!c->haveR) return seterr(ALVRL_ERR_STATE, "build_R first");
    ORC_TRY
    double t0 = now_ms();
    uint32_t N = c->vrls.size(), G = c->rowPixel.size(), S = c->slices.size();
    if (c->globalPixelUndersampling < 0) fail("Invalid pixel undersampling. Did you forget to call buildSlices first?");
    MatView flat; flat.nVrls = N;
    for (uint32_t r = 0; r < G; r++) flat.rows.push_back(&c->R[(size_t) r * N]);
    Sampler *smp = c->mainSampler.get();
    smp->setContext(ALVRL_RNG_CLUSTER, ALVRL_RNG_GLOBAL_ID, 0);
    ClusterStats cstats;
    /* cluster(), 838-898 */
    std::vector<std::vector<uint32_t>> globalVrlsPerCluster;
    std::vector<uint32_t> nonZero, zero;
    for (uint32_t i = 0; i < N; i++) {
        Float sum = 0;                                                       // totalVrlContribution, 936-945
        for (uint32_t r = 0; r < G; r++) sum += flat.rows[r][i].mean;
        bool nz = sum != 0;
        if (c->columnFlagsOverride.size() == N) nz = c->columnFlagsOverride[i] != 0;
        if (nz) nonZero.push_back(i); else zero.push_back(i);
    }
    if (!nonZero.empty()) {
        if (c->P.globalCluster) {                                           // clusterRefinement, 899-912
            std::vector<std::vector<uint32_t>> init(1, nonZero);
            std::vector<double> lw(G, 1.0 / G);
            Clustering cl(init, flat, lw, c->globalPixelUndersampling, 1, &cstats);
            if (!cl.refine(c->P.globalUndersampling, smp)) fail("Couldn't refine global clustering!");
            globalVrlsPerCluster = cl.getVrlsPerCluster();
            c->nearTieSplits += cl.nearTieSplits;
        } else globalVrlsPerCluster.assign(1, nonZero);
    }
    if (!zero.empty()) globalVrlsPerCluster.push_back(zero);
    /* global clustering + fallback, 165-181 */
    {
        std::vector<double> lw(G, 1.0 / G);
        Clustering gcl(globalVrlsPerCluster, flat, lw, c->globalPixelUndersampling, 1, &cstats);
        gcl.sampleRepresentatives(c->gcVrls, c->gcWeight, smp);
        if (!gcl.refine(c->P.fallBackUndersampling, smp)) fail("couldn't refine global clustering! (but all VRLs should be non-zero!)");
        gcl.sampleRepresentatives(c->fallBackVrls, c->fallBackWeight, smp);
        c->nearTieSplits += gcl.nearTieSplits;
    }
    /* refinePerSlice, 199-283 */
    if (c->P.neighbourWeight > 0 && (std::min(c->sliceBegin, S) != 0 || std::min(c->sliceEnd, S) != S))
        fail("neighbourWeight > 0: the local matrices take rows of neighbour slices, R must hold all slices (no slice range)");
    c->selectedVrls.assign(S, {}); c->clusterWeight.assign(S, {});
    bool sfmt = c->P.rngMode == ALVRL_RNG_MODE_SFMT;
    int w = sfmt ? std::max(1, c->P.workerCount) : std::max(1, c->threads);
    std::vector<std::unique_ptr<Sampler>> clones;
    if (w > 1 || !sfmt) for (int i = 0; i < w; i++) clones.emplace_back(c->mainSampler->clone());   // ClusterRefiner ctor, 738
    std::vector<uint32_t> ties(w, 0); std::vector<ClusterStats> cs(w);
    std::vector<std::string> errs(w);
    auto work = [&](int id) {
        try {
        Sampler *s = (w > 1 || !sfmt) ? clones[id].get() : smp;
        uint32_t s0 = ((uint64_t) id * S) / w, s1 = ((uint64_t) (id + 1) * S) / w;
        for (uint32_t i = s0; i < s1; i++) {
            /* refineSlice, 254-283; getLocalMatrix, 779-827 */
            MatView L; L.nVrls = N;
            uint32_t nr = c->rowOffset[i + 1] - c->rowOffset[i];
            for (uint32_t r = c->rowOffset[i]; r < c->rowOffset[i + 1]; r++) L.rows.push_back(&c->R[(size_t) r * N]);
            std::vector<double> lw;
            if (c->P.neighbourWeight <= 0) {
                lw.assign(nr, 0.0);
                for (uint32_t k = 0; k < nr; k++) lw[k] = 1.0 / nr;
            } else {
                /* the rows of the neighbour slices follow (set order: by slice index), weighted by inverse distance; the slice's
                 * own rows share (1 - neighbourWeight) of the total.  Float arithmetic as written (796-820): 1.0 / dist is a
                 * double quotient stored to Float, everything else is Float, the final weights are widened to double. */
                const auto &loc = c->localities[i];
                std::vector<Float> neighbourWeights(loc.size());
                Float summedNeighbourWeight = 0;
                size_t j = 0;
                for (auto it = loc.begin(); it != loc.end(); ++it, ++j) {
                    for (uint32_t r = c->rowOffset[it->first]; r < c->rowOffset[it->first + 1]; r++) L.rows.push_back(&c->R[(size_t) r * N]);
                    neighbourWeights[j] = (Float) (1.0 / it->second);
                    summedNeighbourWeight += neighbourWeights[j];
                }
                const Float sliceWeight = summedNeighbourWeight * (1 - c->P.neighbourWeight) / c->P.neighbourWeight;
                const Float normalization = 1 / (sliceWeight + summedNeighbourWeight);
                for (uint32_t k = 0; k < nr; k++) lw.push_back(sliceWeight * normalization / nr);
                j = 0;
                for (auto it = loc.begin(); it != loc.end(); ++it, ++j) {
                    const uint32_t nrj = c->rowOffset[it->first + 1] - c->rowOffset[it->first];
                    for (uint32_t k = 0; k < nrj; k++) lw.push_back(neighbourWeights[j] * normalization / nrj);
                }
            }
            s->setContext(ALVRL_RNG_CLUSTER, i, 0);
            Clustering cl(globalVrlsPerCluster, L, lw, c->sliceUndersampling[i], c->P.depthCorrection, &cs[id]);
            if (!c->P.localRefinement) { cl.sampleRepresentatives(c->selectedVrls[i], c->clusterWeight[i], s); continue; }
            if (cl.refine(c->P.localUndersampling, s)) cl.sampleRepresentatives(c->selectedVrls[i], c->clusterWeight[i], s);
            else { c->selectedVrls[i] = c->fallBackVrls; c->clusterWeight[i] = c->fallBackWeight; }
            ties[id] += cl.nearTieSplits;
        }
        } catch (const std::exception &e) { errs[id] = e.what(); }
    };
    if (w > 1) { std::vector<std::thread> th; for (int i = 0; i < w; i++) th.emplace_back(work, i); for (auto &t : th) t.join(); }
    else work(0);
    for (auto &e : errs) if (!e.empty()) fail(e);
    for (int i = 0; i < w; i++) { c->nearTieSplits += ties[i]; cstats.splits += cs[i].splits; cstats.varianceSteps += cs[i].varianceSteps; }
    c->clusterSplits = cstats.splits; c->clusterVarianceSteps = cstats.varianceSteps;
    c->haveClusters = true;
    c->stats.msClusters = (float) (now_ms() - t0);
    ORC_CATCH
}
int orc_get_column_nonzero(void *h, uint8_t *flags) {
    Ctx *c = (Ctx *) h;
    if (!c->haveR) return seterr(ALVRL_ERR_STATE, "build_R first");
    uint32_t N = c->vrls.size(), G = c->rowPixel.size();
    for (uint32_t i = 0; i < N; i++) { Float sum = 0; for (uint32_t r = 0; r < G; r++) sum += c->R[(size_t) r * N + i].mean; flags[i] = sum != 0; }
    return ALVRL_OK;
}
int orc_set_column_nonzero(void *h, const uint8_t *flags) {
    Ctx *c = (Ctx *) h;
    if (!flags) c->columnFlagsOverride.clear(); else c->columnFlagsOverride.assign(flags, flags + c->vrls.size());
    return ALVRL_OK;
}
int orc_prepass(void *h) {
    int rc;
    if ((rc = orc_sample_slice_mapping(h))) return rc;
    if ((rc = orc_build_R(h))) return rc;
    return orc_build_clusters(h);
}

/* Li for every pixel centre, vrlIntegrator.cpp:386-393,542-599; image layout [y][x][c].
 * pixelSubset (optional): only these pixel indices are evaluated (bounded CPU baseline sample). */
static int render_impl(Ctx *c, float *rgb, bool clustered, const uint32_t *pixelSubset, uint32_t nSubset) {
    ORC_TRY
    double t0 = now_ms();
    traceChains(c);
    uint32_t W = c->cam.W, H = c->cam.H, P = W * H, N = c->vrls.size();
    uint32_t count = pixelSubset ? nSubset : P;
    if (!pixelSubset) for (size_t i = 0; i < (size_t) P * 3; i++) rgb[i] = 0;
    int T = std::max(1, c->threads);
    std::vector<std::thread> th; std::vector<uint64_t> pairs(T, 0), sh(T, 0);
    uint32_t S = c->slices.size();
    uint32_t sb = std::min(c->sliceBegin, S), se = std::min(c->sliceEnd, S);
    for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
        CounterSampler smp(c->P.seed);
        IntegratorCore core = c->core();
        for (uint32_t q = (uint64_t) count * t / T; q < (uint64_t) count * (t + 1) / T; q++) {
            uint32_t pix = pixelSubset ? pixelSubset[q] : q;
            uint32_t x = pix / H, y = pix % H;
            const Intersection &its = c->hits[pix];
            Spec Li(0.0f);
            if (its.isValid() && !c->medium.sigmaS.isZero()) {
                if (clustered) {
                    uint32_t slice = c->pixelToSlice[pix];
                    if (slice != ALVRL_NO_SLICE && (slice < sb || slice >= se)) continue;
                    const std::vector<uint32_t> &vr = slice == ALVRL_NO_SLICE ? c->fallBackVrls : c->selectedVrls[slice];
                    const std::vector<Float> &wt = slice == ALVRL_NO_SLICE ? c->fallBackWeight : c->clusterWeight[slice];
                    for (size_t i = 0; i < vr.size(); i++) {
                        smp.setContext(ALVRL_RNG_RENDER, pix, (uint32_t) i);
                        Li += wt.at(i) * core.integrateVRL(c->rays[pix], its, c->vrls[vr.at(i)], &smp, nullptr, nullptr);
                    }
                    Li /= (Float) c->particleCount;
                    pairs[t] += vr.size();
                    /* LiSpec (462-505): the chain's segments use the slice of the original camera ray (552-566) and carry their weight */
                    if (c->anyDelta) for (size_t e = 0; e < c->chains[pix].size(); e++) {
                        const Ctx::ChainSeg &cs = c->chains[pix][e];
                        if (!cs.inMedium) continue;
                        Spec LiS(0.0f);
                        for (size_t i = 0; i < vr.size(); i++) {
                            smp.setContext(ALVRL_RNG_RENDER, pix + ((uint32_t) (e + 1) << 24), (uint32_t) i);
                            LiS += wt.at(i) * core.integrateVRL(cs.ray, cs.its, c->vrls[vr.at(i)], &smp, nullptr, nullptr);
                        }
                        LiS /= (Float) c->particleCount;
                        Li += LiS * cs.weight;                                // 598: return Li * weight
                        pairs[t] += vr.size();
                    }
                } else {
                    for (uint32_t v = 0; v < N; v++) {                       // getVRLContributions, 803-816
                        Float normalization = 1.0 / c->particleCount;
                        smp.setContext(ALVRL_RNG_RENDER, pix, v);
                        Spec vc = core.integrateVRL(c->rays[pix], its, c->vrls[v], &smp, nullptr, nullptr);
                        vc *= normalization;
                        Li += vc;
                    }
                    pairs[t] += N;
                    if (c->anyDelta) for (size_t e = 0; e < c->chains[pix].size(); e++) {      // 462-505 around getVRLContributions
                        const Ctx::ChainSeg &cs = c->chains[pix][e];
                        if (!cs.inMedium) continue;
                        for (uint32_t v = 0; v < N; v++) {
                            Float normalization = 1.0 / c->particleCount;
                            smp.setContext(ALVRL_RNG_RENDER, pix + ((uint32_t) (e + 1) << 24), v);
                            Spec vc = core.integrateVRL(cs.ray, cs.its, c->vrls[v], &smp, nullptr, nullptr, cs.weight);
                            vc *= normalization;
                            Li += vc;
                        }
                        pairs[t] += N;
                    }
                }
            }
            float *o = pixelSubset ? rgb + 3 * (size_t) q : rgb + 3 * ((size_t) y * W + x);
            o[0] = Li[0]; o[1] = Li[1]; o[2] = Li[2];
        }
        sh[t] = core.shadowRays;
    });
    for (auto &t : th) t.join();
    for (int t = 0; t < T; t++) { c->stats.pairsRender += pairs[t]; c->stats.shadowRays += sh[t]; }
    c->stats.msRender = (float) (now_ms() - t0);
    ORC_CATCH
}
int orc_render(void *h, float *rgb) {
    Ctx *c = (Ctx *) h;
    if (!c->haveClusters) return seterr(ALVRL_ERR_STATE, "build_clusters first");
    return render_impl(c, rgb, true, nullptr, 0);
}
int orc_render_unclustered(void *h, float *rgb) {
    Ctx *c = (Ctx *) h;
    if (!c->haveVrls || !c->haveMedium) return seterr(ALVRL_ERR_STATE, "set_vrls / set_medium first");
    return render_impl(c, rgb, false, nullptr, 0);
}
int orc_render_pixels(void *h, const uint32_t *pixels, uint32_t n, float *rgb, int clustered) {
    Ctx *c = (Ctx *) h;
    if (clustered && !c->haveClusters) return seterr(ALVRL_ERR_STATE, "build_clusters first");
    return render_impl(c, rgb, clustered != 0, pixels, n);
}
int orc_set_slice_range(void *h, uint32_t b, uint32_t e) { Ctx *c = (Ctx *) h; c->sliceBegin = b; c->sliceEnd = e; return ALVRL_OK; }

/* ---- getters / setters ---------------------------------------------------------------------- */
int orc_get_stats(void *h, alvrl_stats *out) { *out = ((Ctx *) h)->stats; return ALVRL_OK; }
int orc_get_cluster_diag(void *h, uint32_t *nearTies, uint64_t *splits, uint64_t *varianceSteps) {
    Ctx *c = (Ctx *) h; *nearTies = c->nearTieSplits; *splits = c->clusterSplits; *varianceSteps = c->clusterVarianceSteps; return ALVRL_OK;
}
int orc_get_primary_hits(void *h, uint32_t *prim, float *t, float *p, float *n) {
    Ctx *c = (Ctx *) h;
    if (!(c->haveMesh && c->haveCam)) return seterr(ALVRL_ERR_STATE, "scene incomplete");
    tracePrimary(c);
    for (uint32_t i = 0; i < c->P_(); i++) {
        const Intersection &its = c->hits[i];
        if (prim) prim[i] = its.prim;
        if (t) t[i] = its.t;
        if (p) { p[3 * i] = its.p.x; p[3 * i + 1] = its.p.y; p[3 * i + 2] = its.p.z; }
        if (n) { n[3 * i] = its.n.x; n[3 * i + 1] = its.n.y; n[3 * i + 2] = its.n.z; }
    }
    return ALVRL_OK;
}
int orc_get_primary_ties(void *h, uint8_t *tie) {
    Ctx *c = (Ctx *) h; tracePrimary(c);
    for (uint32_t i = 0; i < c->P_(); i++) tie[i] = c->hits[i].tie;
    return ALVRL_OK;
}
int orc_get_pixel_to_slice(void *h, uint32_t *out) {
    Ctx *c = (Ctx *) h;
    if (!c->haveSlices) return seterr(ALVRL_ERR_STATE, "build_slices first");
    if (!c->pixelToSlice.empty()) memcpy(out, c->pixelToSlice.data(), c->pixelToSlice.size() * 4);
    return ALVRL_OK;
}
int orc_get_num_slices(void *h, uint32_t *ns, uint32_t *nr) {
    Ctx *c = (Ctx *) h; *ns = c->slices.size(); *nr = c->haveRows ? c->rowPixel.size() : 0; return ALVRL_OK;
}
int orc_get_rep_pixels(void *h, uint32_t *off, uint32_t *px) {
    Ctx *c = (Ctx *) h;
    if (!c->haveRows) return seterr(ALVRL_ERR_STATE, "sample_slice_mapping first");
    if (!c->rowOffset.empty()) memcpy(off, c->rowOffset.data(), c->rowOffset.size() * 4);
    if (!c->rowPixel.empty()) memcpy(px, c->rowPixel.data(), c->rowPixel.size() * 4);      /* (no rows: every pixel missed) */
    return ALVRL_OK;
}
int orc_set_rep_pixels(void *h, const uint32_t *off, const uint32_t *px, uint32_t ns) {
    Ctx *c = (Ctx *) h;
    if (!c->haveSlices || ns != c->slices.size()) return seterr(ALVRL_ERR_STATE, "slice count mismatch");
    c->rowOffset.assign(off, off + ns + 1); c->rowPixel.assign(px, px + off[ns]);
    c->sliceUndersampling.resize(ns);
    size_t totalPix = 0;
    for (uint32_t i = 0; i < ns; i++) {
        c->sliceUndersampling[i] = ((Float) (off[i + 1] - off[i])) / c->slices[i].gatherIdx.size();
        totalPix += c->slices[i].gatherIdx.size();
    }
    c->globalPixelUndersampling = ((Float) off[ns]) / totalPix;
    c->haveRows = true; c->haveR = false; c->haveClusters = false;
    return ALVRL_OK;
}
int orc_get_R(void *h, uint32_t r0, uint32_t r1, float *mv) {
    Ctx *c = (Ctx *) h;
    if (!c->haveR) return seterr(ALVRL_ERR_STATE, "build_R first");
    size_t N = c->vrls.size();
    memcpy(mv, &c->R[(size_t) r0 * N], (size_t) (r1 - r0) * N * sizeof(VrlContribution));
    return ALVRL_OK;
}
int orc_set_R(void *h, const float *mv) {
    Ctx *c = (Ctx *) h;
    if (!c->haveRows || !c->haveVrls) return seterr(ALVRL_ERR_STATE, "rows / vrls first");
    size_t n = (size_t) c->rowPixel.size() * c->vrls.size();
    c->R.resize(n); memcpy(c->R.data(), mv, n * sizeof(VrlContribution));
    c->haveR = true; c->haveClusters = false;
    return ALVRL_OK;
}
int orc_get_cluster_counts(void *h, uint32_t *off, uint32_t *ng, uint32_t *nf) {
    Ctx *c = (Ctx *) h;
    if (!c->haveClusters) return seterr(ALVRL_ERR_STATE, "build_clusters first");
    off[0] = 0;
    for (size_t i = 0; i < c->selectedVrls.size(); i++) off[i + 1] = off[i] + c->selectedVrls[i].size();
    *ng = c->gcVrls.size(); *nf = c->fallBackVrls.size();
    return ALVRL_OK;
}
int orc_get_clusters(void *h, uint32_t *vrls, float *weights, uint32_t *gv, float *gw, uint32_t *fv, float *fw) {
    Ctx *c = (Ctx *) h;
    if (!c->haveClusters) return seterr(ALVRL_ERR_STATE, "build_clusters first");
    size_t o = 0;
    for (size_t i = 0; i < c->selectedVrls.size(); i++)
        for (size_t j = 0; j < c->selectedVrls[i].size(); j++, o++) { vrls[o] = c->selectedVrls[i][j]; weights[o] = c->clusterWeight[i][j]; }
    for (size_t j = 0; j < c->gcVrls.size(); j++) { if (gv) gv[j] = c->gcVrls[j]; if (gw) gw[j] = c->gcWeight[j]; }
    for (size_t j = 0; j < c->fallBackVrls.size(); j++) { if (fv) fv[j] = c->fallBackVrls[j]; if (fw) fw[j] = c->fallBackWeight[j]; }
    return ALVRL_OK;
}
int orc_set_clusters(void *h, const uint32_t *off, uint32_t ns, const uint32_t *vrls, const float *weights,
                     const uint32_t *fv, const float *fw, uint32_t nf) {
    Ctx *c = (Ctx *) h;
    c->selectedVrls.assign(ns, {}); c->clusterWeight.assign(ns, {});
    for (uint32_t i = 0; i < ns; i++) {
        c->selectedVrls[i].assign(vrls + off[i], vrls + off[i + 1]);
        c->clusterWeight[i].assign(weights + off[i], weights + off[i + 1]);
    }
    c->fallBackVrls.assign(fv, fv + nf); c->fallBackWeight.assign(fw, fw + nf);
    c->haveClusters = true;
    return ALVRL_OK;
}
int orc_trace_rays(void *h, const float *o, const float *d, const float *mint, const float *maxt, uint32_t n,
                   uint32_t *prim, float *t, uint8_t *tie) {
    Ctx *c = (Ctx *) h;
    if (!c->haveMesh) return seterr(ALVRL_ERR_STATE, "set_mesh first");
    for (uint32_t i = 0; i < n; i++) {
        Ray ray(V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), mint[i], maxt[i]);
        Float u, v, tt; uint32_t pr; bool ti;
        c->scene.closestHit(ray, true, tt, pr, u, v, &ti);
        prim[i] = pr; if (t) t[i] = tt; if (tie) tie[i] = ti;
    }
    return ALVRL_OK;
}
int orc_eval_transmittance(void *h, const float *p1, const int32_t *onSurf, const float *p2, uint32_t n, float *T) {
    Ctx *c = (Ctx *) h;
    if (!c->haveMesh || !c->haveMedium) return seterr(ALVRL_ERR_STATE, "set_mesh / set_medium first");
    for (uint32_t i = 0; i < n; i++) {
        Spec s = evalTransmittance(c->scene, c->medium, V3(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]), onSurf && onSurf[i],
                                   V3(p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]), false);
        T[3 * i] = s[0]; T[3 * i + 1] = s[1]; T[3 * i + 2] = s[2];
    }
    return ALVRL_OK;
}
/* SFMT known-answer access for tests (src/tests/test_random.cpp:433-481) */
int orc_sfmt_ulongs(uint64_t seed, uint64_t *out, uint32_t n) { Random r(seed); for (uint32_t i = 0; i < n; i++) out[i] = r.nextULong(); return ALVRL_OK; }
int orc_sfmt_floats(uint64_t seed, float *out, uint32_t n) { Random r(seed); for (uint32_t i = 0; i < n; i++) out[i] = r.nextFloat(); return ALVRL_OK; }
int orc_sfmt_clone_ulongs(uint64_t seed, uint32_t skipParent, uint64_t *out, uint32_t n) {
    Random parent(seed); for (uint32_t i = 0; i < skipParent; i++) parent.nextULong();
    Random child(&parent); for (uint32_t i = 0; i < n; i++) out[i] = child.nextULong(); return ALVRL_OK;
}
/* single integrateVRL call: row of (mean, var, r, g, b) */
int orc_integrate_pair(void *h, uint32_t pixel, uint32_t vrl, const float *uniforms, uint32_t nu, float out[5]) {
    Ctx *c = (Ctx *) h;
    ORC_TRY
    tracePrimary(c);
    TapeSampler smp(uniforms, nu, nu, 1);
    smp.setContext(0, 0, 0);
    IntegratorCore core = c->core();
    Float m = 0, v = 0;
    Spec s(0.0f);
    if (c->hits[pixel].isValid()) s = core.integrateVRL(c->rays[pixel], c->hits[pixel], c->vrls[vrl], &smp, &m, &v);
    out[0] = m; out[1] = v; out[2] = s[0]; out[3] = s[1]; out[4] = s[2];
    ORC_CATCH
}

/* the chain segments below the camera segments, grouped by pixel (layout: include/alvrl.h::alvrl_get_chain_segments) */
int orc_get_chain_segments(void *h, uint32_t *offset, float *segs) {
    Ctx *c = (Ctx *) h;
    ORC_TRY
    traceChains(c);
    const uint32_t P = c->P_();
    uint32_t total = 0;
    for (uint32_t pix = 0; pix < P; pix++) {
        offset[pix] = total;
        for (const Ctx::ChainSeg &cs : c->chains[pix]) {
            if (segs) {
                float *o = segs + 16 * (size_t) total;
                o[0] = cs.ray.o.x; o[1] = cs.ray.o.y; o[2] = cs.ray.o.z; o[3] = cs.ray.d.x; o[4] = cs.ray.d.y; o[5] = cs.ray.d.z;
                o[6] = cs.its.p.x; o[7] = cs.its.p.y; o[8] = cs.its.p.z; o[9] = distance(cs.its.p, cs.ray.o);
                o[10] = cs.weight[0]; o[11] = cs.weight[1]; o[12] = cs.weight[2]; o[13] = cs.inMedium ? 1.0f : 0.0f;
                o[14] = (float) cs.code; o[15] = (float) cs.its.material;
            }
            total++;
        }
    }
    offset[P] = total;
    ORC_CATCH
}

int orc_set_seed(void *h, uint64_t seed) {                                 // the next progressive pass (integrator.cpp:398-434)
    Ctx *c = (Ctx *) h;
    if (c->P.rngMode == ALVRL_RNG_MODE_SFMT) return seterr(ALVRL_ERR_UNSUPPORTED, "set_seed: counter stream only");
    if (seed != c->P.seed) {
        c->P.seed = seed;
        c->mainSampler.reset(c->newStream());
        c->haveChains = false; c->haveRows = false; c->haveR = false; c->haveClusters = false;
    }
    return ALVRL_OK;
}

/* ---- VRL tracer: vrlTracer.h:14-58, 91-230 --------------------------------------------------------------------------- */
int orc_set_area_emitter(void *h, const uint32_t *tris, uint32_t n, const float radiance[3]) {
    Ctx *c = (Ctx *) h;
    if (!c->haveMesh || !n) return seterr(ALVRL_ERR_ARG, "set_area_emitter: set_mesh first, at least one triangle");
    c->emTris.assign(tris, tris + n);
    /* DiscreteDistribution::append / normalize (pmf.h:48-52,101-114) over Triangle::surfaceArea (triangle.cpp:61-67) */
    c->emCdf.assign(1, 0.0f);
    for (uint32_t i = 0; i < n; i++) {
        if (tris[i] >= c->scene.numTris()) return seterr(ALVRL_ERR_ARG, "set_area_emitter: triangle index out of range");
        const V3 &p0 = c->scene.verts[c->scene.tris[3 * tris[i]]], &p1 = c->scene.verts[c->scene.tris[3 * tris[i] + 1]], &p2 = c->scene.verts[c->scene.tris[3 * tris[i] + 2]];
        V3 sideA = p1 - p0, sideB = p2 - p0;
        c->emCdf.push_back(c->emCdf.back() + 0.5f * cross(sideA, sideB).length());
    }
    const Float sum = c->emCdf.back(), normalization = 1.0f / sum;
    for (size_t i = 1; i < c->emCdf.size(); ++i) c->emCdf[i] *= normalization;
    c->emCdf.back() = 1.0f;
    c->emInvArea = 1.0f / sum;
    c->emPower = Spec(radiance[0], radiance[1], radiance[2]) * (Float) M_PI * sum;       // area.cpp:198
    c->emRadiance = Spec(radiance[0], radiance[1], radiance[2]);
    c->triIsEmitter.assign(c->scene.numTris(), 0);
    for (uint32_t i = 0; i < n; i++) c->triIsEmitter[tris[i]] = 1;
    c->haveEmitter = true;
    return ALVRL_OK;
}
namespace {
inline void sincosF(Float theta, Float *s, Float *cs) { *s = (Float) ::sin((double) theta); *cs = (Float) ::cos((double) theta); }   // math::sincos, pinned through double
inline V3 squareToUniformSphere(Float sx, Float sy) {                       // warp.cpp:25-31
    Float z = 1.0f - 2.0f * sy;
    Float r = safe_sqrt(1.0f - z * z);
    Float sinPhi, cosPhi;
    sincosF((Float) (2.0f * M_PI * sx), &sinPhi, &cosPhi);
    return V3(r * cosPhi, r * sinPhi, z);
}
inline V3 squareToCosineHemisphere(Float sx, Float sy) {                    // warp.cpp:43-52, 81-102
    Float r1 = 2.0f * sx - 1.0f, r2 = 2.0f * sy - 1.0f;
    Float phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (Float) ((M_PI / 4.0f) * (r2 / r1)); }
    else { r = r2; phi = (Float) ((M_PI / 2.0f) - (r1 / r2) * (M_PI / 4.0f)); }
    Float cosPhi, sinPhi;
    sincosF(phi, &sinPhi, &cosPhi);
    Float px = r * cosPhi, py = r * sinPhi;
    Float z = safe_sqrt(1.0f - px * px - py * py);
    if (z == 0) z = 1e-10f;
    return V3(px, py, z);
}
inline void coordinateSystem(const V3 &a, V3 &b, V3 &c) {                    // util.cpp:592-601
    if (std::abs(a.x) > std::abs(a.y)) { Float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z); c = V3(a.z * invLen, 0.0f, -a.x * invLen); }
    else { Float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z); c = V3(0.0f, a.z * invLen, -a.y * invLen); }
    b = cross(c, a);
}
/* BSDF::sample(bRec, pdf, sample) of the three surface models of the path, in the local frame (wi, wo with z along the shading
 * normal): smooth dielectric (dielectric.cpp:281-364; mode ERadiance scales transmitted radiance by the squared relative index,
 * EImportance does not, 322-326), smooth conductor (conductor.cpp:254-283), diffuse (diffuse.cpp:129-148).  Returns the weight
 * f * cos / pdf (zero: no sample); pdf is the discrete probability for the delta components, the solid-angle density otherwise. */
inline Spec sampleSurfaceBsdf(uint32_t bits, const Spec &albedo, const Optics *opt, const V3 &wi, Float bsx, Float bsy, bool radianceMode,
                              V3 &woL, Float &bEta, Float &bsdfPdf, bool &delta) {
    Spec bsdfWeight(0.0f);
    bEta = 1.0f; bsdfPdf = 0; delta = false; woL = V3(0.0f);
    if (bits & ALVRL_BSDF_DIELECTRIC) {
        const Optics &o = *opt;
        const Float e = o.v[0], invE = 1 / e;
        Float cosThetaT;
        Float F = fresnelDielectricExt(wi.z, cosThetaT, e);
        delta = true;
        if (bsx <= F) { woL = V3(-wi.x, -wi.y, wi.z); bEta = 1.0f; bsdfPdf = F; bsdfWeight = Spec(o.v[6], o.v[7], o.v[8]); }
        else {
            Float scale = -(cosThetaT < 0 ? invE : e);
            woL = V3(scale * wi.x, scale * wi.y, cosThetaT);
            bEta = cosThetaT < 0 ? e : invE;
            bsdfPdf = 1 - F;
            Float factor = radianceMode ? (cosThetaT < 0 ? invE : e) : 1.0f;
            bsdfWeight = Spec(o.v[9], o.v[10], o.v[11]) * (factor * factor);
        }
    } else if (bits & ALVRL_BSDF_CONDUCTOR) {
        const Optics &o = *opt;
        if (wi.z > 0) {
            delta = true;
            woL = V3(-wi.x, -wi.y, wi.z); bsdfPdf = 1;
            bsdfWeight = Spec(o.v[6], o.v[7], o.v[8]) * Spec(fresnelConductorExact(wi.z, o.v[0], o.v[3]), fresnelConductorExact(wi.z, o.v[1], o.v[4]), fresnelConductorExact(wi.z, o.v[2], o.v[5]));
        }
    } else if ((bits & ALVRL_BSDF_SMOOTH) && wi.z > 0) {
        woL = squareToCosineHemisphere(bsx, bsy);
        bsdfPdf = INV_PI * woL.z;
        bsdfWeight = albedo;
    }
    return bsdfWeight;
}
inline V3 frameToWorld(const V3 &n, const V3 &v) { V3 s, t; coordinateSystem(n, s, t); return s * v.x + t * v.y + n * v.z; }   // Frame(n).toWorld
struct TracedVrl { Spec power; V3 start, end; };

/* vrlTracer::traceOneParticle, vrlTracer.h:91-230; the VRLs it stores (vrlVector::put filter applied) are appended to out */
void traceOneParticle(Ctx *c, Sampler *smp, std::vector<TracedVrl> &out) {
    const bool shortVrls = c->P.shortVrls != 0;
    const int maxDepth = c->P.maxParticleDepth, rrDepth = c->P.rrDepth;
    Medium &med = c->medium;
    /* scene->sampleEmitterPosition(pRec, next2D()): one emitter, m_emitterPDF = {0, 1}: the sample and the value pass unchanged
     * (scene.cpp:958-974); TriMesh::samplePosition (trimesh.cpp:412-423), Triangle::sample (triangle.cpp:24-45) */
    Float sx = smp->next1D(), sy = smp->next1D();
    {
        std::vector<Float>::const_iterator entry = std::lower_bound(c->emCdf.begin(), c->emCdf.end(), sy);        // pmf.h:123-135
        size_t index = std::min(c->emCdf.size() - 2, (size_t) std::max((ptrdiff_t) 0, entry - c->emCdf.begin() - 1));
        while ((c->emCdf[index + 1] - c->emCdf[index]) == 0 && index < c->emCdf.size() - 1) ++index;
        sy = (sy - c->emCdf[index]) / (c->emCdf[index + 1] - c->emCdf[index]);                                      // sampleReuse, 164-169
        const uint32_t tri = c->emTris[index];
        const V3 &p0 = c->scene.verts[c->scene.tris[3 * tri]], &p1 = c->scene.verts[c->scene.tris[3 * tri + 1]], &p2 = c->scene.verts[c->scene.tris[3 * tri + 2]];
        Float a = safe_sqrt(1.0f - sx);                                                                             // squareToUniformTriangle, warp.cpp:76-79
        Float bx = 1 - a, by = a * sy;
        V3 sideA = p1 - p0, sideB = p2 - p0;
        V3 p = p0 + (sideA * bx) + (sideB * by);
        V3 n = normalize(cross(sideA, sideB));
        Spec power = c->emPower;                                                                                    // area.cpp:94-98; / emPdf = 1
        /* emitter->sampleDirection(dRec, pRec, next2D()), area.cpp:115-123 */
        Float dx = smp->next1D(), dy = smp->next1D();
        V3 local = squareToCosineHemisphere(dx, dy);
        V3 d = frameToWorld(n, local);
        power *= Spec(1.0f);
        if (power.isZero()) return;
        bool inMedium = true;                                           // emitter->getMedium(): the emitter sits in the scene's medium
        TracedVrl cur; cur.power = power; cur.start = p;                // handleEmission (nextParticle is counted by the caller)
        bool curMedium = inMedium;
        auto endCurrent = [&](const V3 &q) {                            // endCurrentVrl + vrlVector::put, VRL.h:148-158
            if (distance(cur.start, q) == 0) return;
            cur.end = q;
            if (!curMedium || med.sigmaS.isZero()) return;
            if (cur.power.isZero()) return;
            if (distance(cur.start, cur.end) == 0) return;
            out.push_back(cur);
        };
        Ray ray(p, d, Epsilon, std::numeric_limits<Float>::infinity());
        int depth = 1;
        Spec throughput(1.0f);
        Float eta = 1.0f;
        while (!throughput.isZero() && (depth <= maxDepth || maxDepth < 0)) {
            Intersection its;
            c->scene.rayIntersect(ray, its);                            // rayIntersectAll: no special shapes here
            /* medium->sampleDistance(Ray(ray, 0, its.t), mRec, sampler), homogeneous.cpp:275-352 (EBalance) */
            bool scattered = false;
            Spec mTrans(1.0f), mSigmaS; Float pdfFailure = 1, pdfSuccess = 1; V3 mP;
            if (inMedium && med.type == 1) {
                /* HeterogeneousMedium::sampleDistance, heterogeneous.cpp:589-616 (simpson): one uniform, the optical depth it asks
                 * for, the march that finds where the ray reaches it */
                Float desiredDensity = -((Float) ::log((double) (1 - smp->next1D())));
                Float integratedDensity, tt, densityAtMinT, densityAtT;
                Ray seg(ray.o, ray.d, 0, its.t);
                bool success = false;
                if (med.invertDensityIntegral(seg, desiredDensity, integratedDensity, tt, densityAtMinT, densityAtT)) {
                    mP = seg(tt);
                    success = true;
                    mSigmaS = med.hetAlbedo * densityAtT;
                }
                Float expVal = fastexp(-integratedDensity);
                pdfFailure = expVal; pdfSuccess = expVal * densityAtT; mTrans = Spec(expVal);
                scattered = success && pdfSuccess > 0;
            } else if (inMedium) {
                Float rnd = smp->next1D(), sampledDistance;
                Float samplingDensity = 0;
                if (rnd < med.samplingWeight) {
                    rnd /= med.samplingWeight;
                    int channel = std::min((int) (smp->next1D() * 3), 2);
                    samplingDensity = med.sigmaT[channel];
                    sampledDistance = -((Float) ::log((double) (1 - rnd))) / samplingDensity;
                } else sampledDistance = std::numeric_limits<Float>::infinity();
                Float distSurf = its.t - 0;
                bool success = true;
                if (sampledDistance < distSurf) {
                    Float t = sampledDistance + 0;
                    mP = ray.o + t * ray.d;
                    mSigmaS = med.sigmaS;
                    if (mP.x == ray.o.x && mP.y == ray.o.y && mP.z == ray.o.z) success = false;
                } else { sampledDistance = distSurf; success = false; }
                pdfFailure = 0; pdfSuccess = 0;
                for (int i = 0; i < 3; ++i) { Float tmp = fastexp(-med.sigmaT[i] * sampledDistance); pdfFailure += tmp; pdfSuccess += med.sigmaT[i] * tmp; }
                pdfFailure /= 3; pdfSuccess /= 3;
                for (int i = 0; i < 3; ++i) mTrans[i] = fastexp(med.sigmaT[i] * (-sampledDistance));
                pdfSuccess = pdfSuccess * med.samplingWeight;
                pdfFailure = med.samplingWeight * pdfFailure + (1 - med.samplingWeight);
                if (mTrans.max() < 1e-20) mTrans = Spec(0.0f);
                scattered = success;
            }
            if (inMedium && scattered) {
                throughput *= mTrans * mSigmaS / pdfSuccess;
                /* medium->getPhaseFunction()->sample(pRec, sampler): returns 1 (isotropic.cpp:62-67, hg.cpp:74-98) */
                Float px = smp->next1D(), py = smp->next1D();
                V3 wo;
                if (med.phaseType == ALVRL_PHASE_ISOTROPIC) wo = squareToUniformSphere(px, py);
                else {
                    Float cosTheta;
                    if (std::abs(med.g) < Epsilon) cosTheta = 1 - 2 * px;
                    else { Float sqrTerm = (1 - med.g * med.g) / (1 - med.g + 2 * med.g * px); cosTheta = (1 + med.g * med.g - sqrTerm * sqrTerm) / (2 * med.g); }
                    Float sinTheta = safe_sqrt(1.0f - cosTheta * cosTheta), sinPhi, cosPhi;
                    sincosF((Float) (2 * M_PI * py), &sinPhi, &cosPhi);
                    wo = frameToWorld(ray.d, V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));      // Frame(-pRec.wi), wi = -ray.d
                }
                V3 endPoint;
                if (shortVrls) endPoint = mP;
                else { if (its.isValid()) endPoint = its.p; else break; }
                endCurrent(endPoint);                                                                   // handleMediumScattering
                cur.power = throughput * power; cur.start = mP; curMedium = inMedium;
                ray = Ray(mP, wo, 0, std::numeric_limits<Float>::infinity());
            } else if (its.isValid()) {
                if (inMedium) throughput *= mTrans / pdfFailure;
                const uint32_t bits = c->scene.matBits[its.material];
                const HitFrame fr = c->scene.hitFrame(its);
                const V3 wi = fr.toLocal(-ray.d);
                Float bsx = smp->next1D(), bsy = smp->next1D();
                V3 woL; Float bEta, bPdf; bool bDelta;                                                  // EImportance: particles
                const Spec bsdfWeight = sampleSurfaceBsdf(bits, c->scene.albedo[its.material], (bits & ALVRL_BSDF_DELTA) ? &c->scene.optics[its.material] : nullptr,
                                                          wi, bsx, bsy, false, woL, bEta, bPdf, bDelta);
                if (bsdfWeight.isZero()) { endCurrent(its.p); break; }
                V3 wiW = -ray.d, woW = fr.toWorld(woL);
                Float wiDotGeoN = dot(its.n, wiW), woDotGeoN = dot(its.n, woW);
                if (wiDotGeoN * wi.z <= 0 || woDotGeoN * woL.z <= 0) { endCurrent(its.p); break; }     // [Veach, p. 158]
                throughput *= bsdfWeight;
                eta *= bEta;
                if (bits & ALVRL_MAT_TRANSITION) inMedium = woDotGeoN > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
                endCurrent(its.p);                                                                      // handleSurfaceScattering
                cur.power = throughput * power; cur.start = its.p; curMedium = inMedium;
                ray = Ray(its.p, woW, Epsilon, std::numeric_limits<Float>::infinity());
            } else break;
            if (depth++ >= rrDepth) {
                Float q = std::min(throughput.max() * eta * eta, (Float) 0.95f);
                if (smp->next1D() >= q) break;
                throughput /= q;
            }
        }
    }
}
} // namespace
int orc_trace_vrls(void *h, uint32_t target) {
    Ctx *c = (Ctx *) h;
    if (!c->haveEmitter || !c->haveMedium || !c->haveMat) return seterr(ALVRL_ERR_STATE, "trace_vrls: set_area_emitter / set_medium / set_materials first");
    ORC_TRY
    if (!target) target = (uint32_t) c->P.vrlTargetNum;
    CounterSampler smp(c->P.seed);
    std::vector<TracedVrl> all;
    uint64_t particles = 0;
    while (all.size() < target) {                                          // randomWalk, vrlTracer.h:29-40
        smp.setContext(ALVRL_RNG_TRACER, (uint32_t) particles, 0);
        particles++;                                                       // m_vrls->nextParticle()
        traceOneParticle(c, &smp, all);
        if (particles > (1ull << 31)) throw std::runtime_error("trace_vrls: no VRLs are being generated");
    }
    c->vrls.clear();
    for (const TracedVrl &t : all) { VRL v; v.power = t.power; v.start = t.start; v.end = t.end; c->vrls.push_back(v); }
    c->particleCount = particles;
    c->haveVrls = true; c->haveR = false; c->haveClusters = false;
    ORC_CATCH
}
int orc_get_vrls(void *h, float *s, float *e, float *p, uint64_t *pc) {
    Ctx *c = (Ctx *) h;
    for (size_t i = 0; i < c->vrls.size(); i++) {
        const VRL &v = c->vrls[i];
        s[3 * i] = v.start.x; s[3 * i + 1] = v.start.y; s[3 * i + 2] = v.start.z; e[3 * i] = v.end.x; e[3 * i + 1] = v.end.y; e[3 * i + 2] = v.end.z;
        p[3 * i] = v.power[0]; p[3 * i + 1] = v.power[1]; p[3 * i + 2] = v.power[2];
    }
    if (pc) *pc = c->particleCount;
    return ALVRL_OK;
}

/* ---- ground truth: the volumetric path tracer restricted to the paths VRLs represent --------------------------------
 * VolumetricPathTracer::Li / Li_original with onlyVRLpaths (src/integrators/path/volpath.cpp:76-460) and
 * rayIntersectAndLookForEmitter (484-535), driven like SamplingIntegrator::renderBlock (src/librender/integrator.cpp:210-268:
 * one sample at the pixel centre when the sample count is 1, jittered otherwise; this fork's rule, 240-246), for the scenes of
 * this path: one homogeneous medium, one area emitter (Scene::sampleAttenuatedEmitterDirect, scene.cpp:854-899;
 * AreaLight::sampleDirect / pdfDirect / eval, src/emitters/area.cpp:103-108,158-183; Shape::sampleDirect / pdfDirect,
 * shape.cpp:102-126; TriMesh::samplePosition, trimesh.cpp:412-423), diffuse / smooth dielectric / smooth conductor surfaces
 * (diffuse.cpp:110-148, dielectric.cpp:281-332, conductor.cpp:268-283), no ENull surfaces (so the emitter search is one
 * intersection), no environment emitter, no subsurface integrator.  The film is the box filter: every sample lands in its
 * own pixel and the developed value is the sum times 1 / (sum of weights) (bitmap.cpp:1617-1624); invalid samples are
 * rejected like ImageBlock::put (imageblock.h:147-151).  Outer sample j of pixel p draws from the counter stream of
 * (ALVRL_RNG_VOLPATH, p, j) in the reference's order; its internalSamples walks continue that stream (volpath.cpp:111-120). */
namespace {
struct VolpathCfg { bool only, volToVol, volToSurf, singleScatter, strictNormals, hideEmitters; int internalSamples, maxDepth, rrDepth; };
struct DirectRec { V3 ref, refN, p, n, d; Float dist = 0, pdf = 0; };

inline Float miWeight(Float pdfA, Float pdfB) { pdfA *= pdfA; pdfB *= pdfB; return pdfA / (pdfA + pdfB); }             // volpath.cpp:537-540

/* HomogeneousMedium::sampleDistance(Ray(ray, 0, itsT), mRec, sampler), homogeneous.cpp:275-352 (strategy = balance) */
struct MRec { Spec transmittance, sigmaS; Float pdfFailure = 1, pdfSuccess = 1; V3 p; };
inline bool sampleDistanceH(Medium &med, const Ray &ray, Float itsT, Sampler *smp, MRec &m) {
    if (med.type == 1) {                                                    // heterogeneous.cpp:589-616 (simpson)
        Float desiredDensity = -((Float) ::log((double) (1 - smp->next1D())));
        Float integratedDensity, tt, densityAtMinT, densityAtT;
        Ray seg(ray.o, ray.d, 0, itsT);
        bool success = false;
        if (med.invertDensityIntegral(seg, desiredDensity, integratedDensity, tt, densityAtMinT, densityAtT)) {
            m.p = seg(tt);
            success = true;
            m.sigmaS = med.hetAlbedo * densityAtT;
        }
        Float expVal = fastexp(-integratedDensity);
        m.pdfFailure = expVal; m.pdfSuccess = expVal * densityAtT; m.transmittance = Spec(expVal);
        return success && m.pdfSuccess > 0;
    }
    Float rnd = smp->next1D(), sampledDistance;
    Float samplingDensity = 0;
    if (rnd < med.samplingWeight) {
        rnd /= med.samplingWeight;
        int channel = std::min((int) (smp->next1D() * 3), 2);
        samplingDensity = med.sigmaT[channel];
        sampledDistance = -((Float) ::log((double) (1 - rnd))) / samplingDensity;
    } else sampledDistance = std::numeric_limits<Float>::infinity();
    Float distSurf = itsT - 0;
    bool success = true;
    if (sampledDistance < distSurf) {
        Float t = sampledDistance + 0;
        m.p = ray.o + t * ray.d;
        m.sigmaS = med.sigmaS;
        if (m.p.x == ray.o.x && m.p.y == ray.o.y && m.p.z == ray.o.z) success = false;
    } else { sampledDistance = distSurf; success = false; }
    m.pdfFailure = 0; m.pdfSuccess = 0;
    for (int i = 0; i < 3; ++i) { Float tmp = fastexp(-med.sigmaT[i] * sampledDistance); m.pdfFailure += tmp; m.pdfSuccess += med.sigmaT[i] * tmp; }
    m.pdfFailure /= 3; m.pdfSuccess /= 3;
    for (int i = 0; i < 3; ++i) m.transmittance[i] = fastexp(med.sigmaT[i] * (-sampledDistance));
    m.pdfSuccess = m.pdfSuccess * med.samplingWeight;
    m.pdfFailure = med.samplingWeight * m.pdfFailure + (1 - med.samplingWeight);
    if (m.transmittance.max() < 1e-20) m.transmittance = Spec(0.0f);
    return success;
}

/* emitter->sampleDirect(dRec, sample): AreaLight::sampleDirect over Shape::sampleDirect over TriMesh::samplePosition
 * (area.cpp:158-174, shape.cpp:102-115, trimesh.cpp:412-423, triangle.cpp:24-45); returns radiance / pdf, fills dRec */
inline Spec emitterSampleDirect(Ctx *c, DirectRec &dRec, Float sx, Float sy) {
    /* one emitter: m_emitterPDF.sampleReuse(sample.x) leaves the sample unchanged, emPdf = 1 (scene.cpp:860-862) */
    std::vector<Float>::const_iterator entry = std::lower_bound(c->emCdf.begin(), c->emCdf.end(), sy);                  // pmf.h:123-135
    size_t index = std::min(c->emCdf.size() - 2, (size_t) std::max((ptrdiff_t) 0, entry - c->emCdf.begin() - 1));
    while ((c->emCdf[index + 1] - c->emCdf[index]) == 0 && index < c->emCdf.size() - 1) ++index;
    sy = (sy - c->emCdf[index]) / (c->emCdf[index + 1] - c->emCdf[index]);
    const uint32_t tri = c->emTris[index];
    const V3 &p0 = c->scene.verts[c->scene.tris[3 * tri]], &p1 = c->scene.verts[c->scene.tris[3 * tri + 1]], &p2 = c->scene.verts[c->scene.tris[3 * tri + 2]];
    Float a = safe_sqrt(1.0f - sx);
    Float bx = 1 - a, by = a * sy;
    V3 sideA = p1 - p0, sideB = p2 - p0;
    dRec.p = p0 + (sideA * bx) + (sideB * by);
    dRec.n = normalize(cross(sideA, sideB));
    dRec.pdf = c->emInvArea;
    dRec.d = dRec.p - dRec.ref;
    Float distSquared = dRec.d.lengthSquared();
    dRec.dist = std::sqrt(distSquared);
    dRec.d = dRec.d / dRec.dist;
    Float dp = std::abs(dot(dRec.d, dRec.n));
    dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) return c->emRadiance / dRec.pdf;
    dRec.pdf = 0.0f;
    return Spec(0.0f);
}
/* Scene::evalTransmittance(ref, refOnSurface, p, true, ...) in the medium or in vacuum, scene.cpp:619-679 without ENull surfaces */
inline Spec shadowTransmittance(Ctx *c, const V3 &ref, bool refOnSurface, const V3 &p, bool inMedium) {
    V3 d = p - ref;
    Float remaining = d.length();
    d = d / remaining;
    Ray ray(ref, d, refOnSurface ? Epsilon : 0, remaining * (1 - ShadowEpsilon));
    if (!(remaining > 0)) return Spec(1.0f);
    Float t, u, v; uint32_t prim;
    if (c->scene.closestHit(ray, true, t, prim, u, v, nullptr)) return Spec(0.0f);
    if (!inMedium) return Spec(1.0f);
    return c->medium.evalTransmittance(Ray(ray.o, ray.d, 0, std::min(t, remaining)));
}
inline Float pdfEmitterDirect(Ctx *c, const DirectRec &dRec) {                                                           // scene.cpp:949-952, area.cpp:176-183, shape.cpp:117-121
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) return c->emInvArea * (dRec.dist * dRec.dist) / std::abs(dot(dRec.d, dRec.n));
    return 0.0f;
}

/* PhaseFunction::sample(pRec, pdf, sampler) with pRec.wi = wi: isotropic.cpp:69-74, hg.cpp:74-104; returns wo, the value is 1 */
inline V3 samplePhase(const Medium &med, const V3 &wi, Float px, Float py, Float &pdf) {
    if (med.phaseType == ALVRL_PHASE_ISOTROPIC) { pdf = INV_FOURPI; return squareToUniformSphere(px, py); }
    Float cosTheta;
    if (std::abs(med.g) < Epsilon) cosTheta = 1 - 2 * px;
    else { Float sqrTerm = (1 - med.g * med.g) / (1 - med.g + 2 * med.g * px); cosTheta = (1 + med.g * med.g - sqrTerm * sqrTerm) / (2 * med.g); }
    Float sinTheta = safe_sqrt(1.0f - cosTheta * cosTheta), sinPhi, cosPhi;
    sincosF((Float) (2 * M_PI * py), &sinPhi, &cosPhi);
    V3 wo = frameToWorld(-wi, V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));                                       // Frame(-pRec.wi)
    pdf = med.phaseEval(wi, wo);
    return wo;
}

Spec volpathLiOriginal(Ctx *c, const VolpathCfg &cfg, Sampler *smp, const Ray &r, bool sensorInMedium) {
    Medium &med = c->medium;
    Ray ray(r);
    Spec Li(0.0f);
    Float eta = 1.0f;
    bool vrlFirstVertexOK = false, vrlSecondVertexOK = false, prevWasDiffuseSurface = false, prevWasVolume = false;
    bool inMedium = sensorInMedium;
    int depth = 1;
    bool emitted = true, indirectMedium = true;                          // rRec.type: ERadiance, later ERadianceNoEmission
    Intersection its;
    c->scene.rayIntersect(ray, its);
    Spec throughput(1.0f);
    bool scattered = false;
    auto vrlDirectOK = [&]() {
        /* (!rRec.depth == 2 || ...) of the reference is ((!depth) == 2 || ...): the first operand is never true */
        return !cfg.only || (depth != 1 && ((prevWasVolume || prevWasDiffuseSurface) && (!prevWasDiffuseSurface || cfg.volToSurf) && (!prevWasVolume || cfg.volToVol)));
    };
    /* rayIntersectAndLookForEmitter without ENull surfaces: one intersection; an emitter that is hit gives its unattenuated radiance */
    auto lookForEmitter = [&](const Ray &q, DirectRec &dRec, Spec &value) {
        bool surface = c->scene.rayIntersect(q, its);
        if (surface && c->triIsEmitter[its.prim]) {
            dRec.p = its.p; dRec.n = its.n; dRec.d = q.d; dRec.dist = its.t;                                            // setQuery, records.inl:170-178
            value = dot(its.n, -q.d) <= 0 ? Spec(0.0f) : c->emRadiance;                                                 // area.cpp:103-108
        }
    };
    while (depth <= cfg.maxDepth || cfg.maxDepth < 0) {
        if (cfg.only && depth > 2 && !(vrlFirstVertexOK && vrlSecondVertexOK)) break;
        MRec mRec;
        if (inMedium && sampleDistanceH(med, ray, its.t, smp, mRec)) {
            if (cfg.singleScatter) indirectMedium = false;
            if (depth == 1) { if (cfg.volToVol) vrlFirstVertexOK = true; }
            if (depth == 2) vrlSecondVertexOK = true;
            if (depth >= cfg.maxDepth && cfg.maxDepth != -1) break;
            throughput *= mRec.sigmaS * mRec.transmittance / mRec.pdfSuccess;
            DirectRec dRec; dRec.ref = mRec.p; dRec.refN = V3(0.0f);
            if (vrlDirectOK()) {
                Float sx = smp->next1D(), sy = smp->next1D();
                Spec value = emitterSampleDirect(c, dRec, sx, sy);                                                      // sampleAttenuatedEmitterDirect, scene.cpp:854-874
                if (dRec.pdf != 0) value *= shadowTransmittance(c, dRec.ref, false, dRec.p, inMedium);
                if (!value.isZero()) {
                    Float phaseVal = med.phaseEval(-ray.d, dRec.d);
                    if (phaseVal != 0) {
                        Float phasePdf = phaseVal;                                                                      // PhaseFunction::pdf = eval
                        const Float weight = miWeight(dRec.pdf, phasePdf);
                        Li += throughput * value * phaseVal * weight;
                    }
                }
            }
            /* phase function sampling: sample(pRec, pdf, sampler) returns 1 (isotropic.cpp:69-74, hg.cpp:99-104) */
            Float phasePdf;
            Float px = smp->next1D(), py = smp->next1D();
            V3 wo = samplePhase(med, -ray.d, px, py, phasePdf);
            ray = Ray(mRec.p, wo, 0, std::numeric_limits<Float>::infinity());
            Spec value(0.0f);
            lookForEmitter(ray, dRec, value);
            if (!value.isZero() && vrlDirectOK()) {
                const Float emitterPdf = pdfEmitterDirect(c, dRec);
                Li += throughput * value * miWeight(phasePdf, emitterPdf);
            }
            if (!indirectMedium) break;
            emitted = false;
            prevWasVolume = true; prevWasDiffuseSurface = false;
        } else {
            if (inMedium) throughput *= mRec.transmittance / mRec.pdfFailure;
            if (!its.isValid()) break;                                   // no environment emitter
            if (c->triIsEmitter[its.prim] && emitted && (!cfg.hideEmitters || scattered) && (!cfg.only || (vrlFirstVertexOK && vrlSecondVertexOK)))
                Li += throughput * (dot(its.n, -ray.d) <= 0 ? Spec(0.0f) : c->emRadiance);
            if (depth >= cfg.maxDepth && cfg.maxDepth != -1) break;
            const uint32_t bits = c->scene.matBits[its.material];
            const HitFrame fr = c->scene.hitFrame(its);
            const V3 wi = fr.toLocal(-ray.d);
            Float wiDotGeoN = -dot(its.n, ray.d), wiDotShN = wi.z;
            if (wiDotGeoN * wiDotShN < 0 && cfg.strictNormals) break;
            const bool smooth = !(bits & ALVRL_BSDF_DELTA) && (bits & ALVRL_BSDF_SMOOTH);
            DirectRec dRec; dRec.ref = its.p; dRec.refN = (bits & ALVRL_BSDF_DIELECTRIC) ? V3(0.0f) : its.n;            // records.inl:160-164
            const Intersection here = its;
            if (smooth && (!cfg.only || (vrlFirstVertexOK && vrlSecondVertexOK))) {
                Float sx = smp->next1D(), sy = smp->next1D();
                Spec value = emitterSampleDirect(c, dRec, sx, sy);                                                      // scene.cpp:876-899
                if (dRec.pdf != 0) {
                    bool shadowMedium = inMedium;                                                                       // its.getTargetMedium(dRec.d)
                    if (bits & ALVRL_MAT_TRANSITION) shadowMedium = dot(dRec.d, here.n) > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
                    value *= shadowTransmittance(c, here.p, true, dRec.p, shadowMedium);
                }
                if (!value.isZero()) {
                    const V3 woL = fr.toLocal(dRec.d);
                    Spec bsdfVal(0.0f);                                                                                 // diffuse.cpp:110-118
                    if (!(wi.z <= 0 || woL.z <= 0)) bsdfVal = c->scene.albedo[here.material] * (INV_PI * woL.z);
                    Float woDotGeoN = dot(here.n, dRec.d);
                    if (!bsdfVal.isZero() && (!cfg.strictNormals || woDotGeoN * woL.z > 0)) {
                        Float bsdfPdf = (wi.z <= 0 || woL.z <= 0) ? 0.0f : INV_PI * woL.z;                              // diffuse.cpp:120-127, warp.h
                        const Float weight = miWeight(dRec.pdf, bsdfPdf);
                        Li += throughput * value * bsdfVal * weight;
                    }
                }
            }
            /* BSDF sampling: sample(bRec, pdf, nextSample2D()), mode = ERadiance */
            Float bsx = smp->next1D(), bsy = smp->next1D();
            V3 woL; Float bEta, bsdfPdf; bool delta;
            const Spec bsdfWeight = sampleSurfaceBsdf(bits, c->scene.albedo[here.material], (bits & ALVRL_BSDF_DELTA) ? &c->scene.optics[here.material] : nullptr,
                                                      wi, bsx, bsy, true, woL, bEta, bsdfPdf, delta);
            if (bsdfWeight.isZero()) break;
            const V3 wo = fr.toWorld(woL);
            Float woDotGeoN = dot(here.n, wo);
            if (woDotGeoN * woL.z <= 0 && cfg.strictNormals) break;
            if (depth == 1 && delta) depth--;                            // 'undo' initial specular vertices
            if (cfg.volToSurf) { if (depth == 1 && inMedium && !delta) vrlFirstVertexOK = true; }
            prevWasVolume = false;
            prevWasDiffuseSurface = !delta;
            ray = Ray(here.p, wo, Epsilon, std::numeric_limits<Float>::infinity());
            throughput *= bsdfWeight;
            eta *= bEta;
            if (bits & ALVRL_MAT_TRANSITION) inMedium = dot(here.n, ray.d) > 0 ? (bits & ALVRL_MAT_EXTERIOR_MEDIUM) != 0 : (bits & ALVRL_MAT_INTERIOR_MEDIUM) != 0;
            Spec value(0.0f);
            lookForEmitter(ray, dRec, value);
            if (!value.isZero() && (!cfg.only || (vrlFirstVertexOK && vrlSecondVertexOK))) {
                const Float emitterPdf = !delta ? pdfEmitterDirect(c, dRec) : 0;
                Li += throughput * value * miWeight(bsdfPdf, emitterPdf);
            }
            emitted = false;
        }
        if (depth++ >= cfg.rrDepth) {
            Float q = std::min(throughput.max() * eta * eta, (Float) 0.95f);
            if (smp->next1D() >= q) break;
            throughput /= q;
        }
        scattered = true;
    }
    if (cfg.only && !(vrlFirstVertexOK && vrlSecondVertexOK)) Li *= 0;
    return Li;
}
} // namespace
int orc_volpath_render(void *h, uint32_t spp, uint32_t internalSamples, uint32_t flags, int32_t maxDepth, float *rgb) {
    Ctx *c = (Ctx *) h;
    if (!c->haveEmitter || !c->haveMedium || !c->haveMat || !c->haveCam) return seterr(ALVRL_ERR_STATE, "volpath_render: set_area_emitter / set_medium / set_materials / set_camera first");
    if (!spp || !internalSamples) return seterr(ALVRL_ERR_ARG, "volpath_render: spp and internalSamples must be positive");
    ORC_TRY
    VolpathCfg cfg;
    cfg.only = flags & ALVRL_VOLPATH_ONLY_VRL_PATHS; cfg.volToVol = flags & ALVRL_VOLPATH_VOL_TO_VOL; cfg.volToSurf = flags & ALVRL_VOLPATH_VOL_TO_SURF;
    cfg.singleScatter = flags & ALVRL_VOLPATH_SINGLE_SCATTER; cfg.strictNormals = flags & ALVRL_VOLPATH_STRICT_NORMALS; cfg.hideEmitters = flags & ALVRL_VOLPATH_HIDE_EMITTERS;
    cfg.internalSamples = (int) internalSamples; cfg.maxDepth = maxDepth; cfg.rrDepth = c->P.rrDepth;
    const bool centre = spp == 1 || (flags & ALVRL_VOLPATH_CENTRE_SAMPLES);
    const uint32_t W = c->cam.W, H = c->cam.H;
    auto work = [&](uint32_t x0, uint32_t x1) {
        CounterSampler smp(c->P.seed);
        for (uint32_t x = x0; x < x1; x++) for (uint32_t y = 0; y < H; y++) {
            const uint32_t pix = x * H + y;
            Spec sum(0.0f); Float weight = 0;
            for (uint32_t j = 0; j < spp; j++) {
                smp.setContext(ALVRL_RNG_VOLPATH, pix, j);
                Float ox = 0.5f, oy = 0.5f;
                if (!centre) { ox = smp.next1D(); oy = smp.next1D(); }                                                   // integrator.cpp:240-246
                const Ray ray = c->cam.sampleRay((Float) x + ox, (Float) y + oy);
                Spec Li(0.0f);                                                                                          // volpath.cpp:111-120
                for (int i = 0; i < cfg.internalSamples; i++) Li += volpathLiOriginal(c, cfg, &smp, ray, true);
                Li = Li / (Float) cfg.internalSamples;
                if (!Li.isValid()) continue;                                                                            // imageblock.h:147-151
                sum += Li; weight += 1.0f;
            }
            const Float invWeight = weight == 0 ? 0 : (Float) 1 / weight;
            float *o = rgb + 3 * ((size_t) y * W + x);
            for (int k = 0; k < 3; k++) o[k] = sum[k] * invWeight;
        }
    };
    int T = std::max(1, c->threads);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work, (uint32_t) ((uint64_t) W * t / T), (uint32_t) ((uint64_t) W * (t + 1) / T));
    for (auto &t : th) t.join();
    ORC_CATCH
}

/* ---- test hooks for the chi-square tests of the sampling routines (the reference's own strategy for them:
 * src/tests/test_chisquare.cpp:508-573 phase functions, 575-623 emitters) ------------------------------------------------- */
int orc_test_phase_sample(int32_t phaseType, float g, const float wi[3], const float *u, uint32_t n, float *wo, float *pdf) {
    Medium med; med.phaseType = phaseType; med.g = g;
    for (uint32_t i = 0; i < n; i++) {
        Float p; V3 w = samplePhase(med, V3(wi[0], wi[1], wi[2]), u[2 * i], u[2 * i + 1], p);
        wo[3 * i] = w.x; wo[3 * i + 1] = w.y; wo[3 * i + 2] = w.z; pdf[i] = p;
    }
    return ALVRL_OK;
}
int orc_test_phase_eval(int32_t phaseType, float g, const float wi[3], const float *wo, uint32_t n, float *val) {
    Medium med; med.phaseType = phaseType; med.g = g;
    for (uint32_t i = 0; i < n; i++) val[i] = med.phaseEval(V3(wi[0], wi[1], wi[2]), V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]));
    return ALVRL_OK;
}
/* Medium::sampleDistance(Ray(o, d, 0, itsT), mRec, sampler) of the handle's medium, n times on two uniforms each (the
 * homogeneous medium draws two, the grid medium one): distance of the sampled point from o (itsT when no interaction was
 * sampled), success flag, pdfSuccess, pdfFailure, transmittance */
int orc_test_sample_distance(void *h, const float o[3], const float d[3], float itsT, const float *u, uint32_t n,
                             float *t, uint8_t *success, float *pdfSuccess, float *pdfFailure, float *transmittance) {
    Ctx *c = (Ctx *) h;
    ORC_TRY
    if (!c->haveMedium) return seterr(ALVRL_ERR_STATE, "set a medium first");
    const Ray ray(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), 0, itsT);
    for (uint32_t i = 0; i < n; i++) {
        TapeSampler smp(u + 2 * (size_t) i, 2, 2, 1);
        smp.setContext(0, 0, 0);
        MRec m;
        const bool ok = sampleDistanceH(c->medium, ray, itsT, &smp, m);
        success[i] = ok ? 1 : 0;
        t[i] = ok ? (m.p - ray.o).length() : itsT;
        pdfSuccess[i] = m.pdfSuccess; pdfFailure[i] = m.pdfFailure;
        transmittance[3 * i] = m.transmittance[0]; transmittance[3 * i + 1] = m.transmittance[1]; transmittance[3 * i + 2] = m.transmittance[2];
    }
    ORC_CATCH
}
/* BSDF::sample of one surface model in the local frame (test_chisquare.cpp:398-506): bits = ALVRL_BSDF_* of the material,
 * optics12 as alvrl_set_material_optics; delta[i] = 1 for a discrete component (pdf = its probability) */
int orc_test_bsdf_sample(uint32_t bits, const float albedo[3], const float *optics12, const float wi[3], const float *u, uint32_t n,
                         int radianceMode, float *wo, float *pdf, float *weight, uint8_t *delta) {
    Optics o; if (optics12) memcpy(o.v, optics12, 12 * sizeof(float));
    for (uint32_t i = 0; i < n; i++) {
        V3 w; Float eta, p; bool d;
        Spec wt = sampleSurfaceBsdf(bits, Spec(albedo[0], albedo[1], albedo[2]), optics12 ? &o : nullptr, V3(wi[0], wi[1], wi[2]), u[2 * i], u[2 * i + 1],
                                    radianceMode != 0, w, eta, p, d);
        wo[3 * i] = w.x; wo[3 * i + 1] = w.y; wo[3 * i + 2] = w.z; pdf[i] = p; delta[i] = d ? 1 : 0;
        weight[3 * i] = wt[0]; weight[3 * i + 1] = wt[1]; weight[3 * i + 2] = wt[2];
    }
    return ALVRL_OK;
}
/* emitter->sampleDirect from a reference point in the volume (refN = 0): direction, solid-angle pdf, radiance / pdf */
int orc_test_emitter_sample_direct(void *h, const float ref[3], const float *u, uint32_t n, float *d, float *pdf, float *value) {
    Ctx *c = (Ctx *) h;
    if (!c->haveEmitter) return seterr(ALVRL_ERR_STATE, "set_area_emitter first");
    for (uint32_t i = 0; i < n; i++) {
        DirectRec dRec; dRec.ref = V3(ref[0], ref[1], ref[2]); dRec.refN = V3(0.0f);
        Spec v = emitterSampleDirect(c, dRec, u[2 * i], u[2 * i + 1]);
        d[3 * i] = dRec.d.x; d[3 * i + 1] = dRec.d.y; d[3 * i + 2] = dRec.d.z; pdf[i] = dRec.pdf;
        value[3 * i] = v[0]; value[3 * i + 1] = v[1]; value[3 * i + 2] = v[2];
    }
    return ALVRL_OK;
}
/* the density of having sampled direction d from ref: intersect, and if the emitter is what the ray meets, emitter->pdfDirect of the
 * query record (test_chisquare.cpp:364-378; volpath.cpp:521-524 + scene.cpp:949-952) */
int orc_test_emitter_pdf_direct(void *h, const float ref[3], const float *d, uint32_t n, float *pdf) {
    Ctx *c = (Ctx *) h;
    if (!c->haveEmitter) return seterr(ALVRL_ERR_STATE, "set_area_emitter first");
    for (uint32_t i = 0; i < n; i++) {
        V3 dir(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        Ray q(V3(ref[0], ref[1], ref[2]), dir, 0, std::numeric_limits<Float>::infinity());
        Intersection its;
        pdf[i] = 0;
        if (c->scene.rayIntersect(q, its) && c->triIsEmitter[its.prim]) {
            DirectRec dRec; dRec.ref = q.o; dRec.refN = V3(0.0f);
            dRec.p = its.p; dRec.n = its.n; dRec.d = dir; dRec.dist = its.t;
            pdf[i] = pdfEmitterDirect(c, dRec);
        }
    }
    return ALVRL_OK;
}

/* Film: ReconstructionFilter::configure / evalDiscretized (src/libcore/rfilter.cpp:37-55, include/mitsuba/core/rfilter.h:76-77),
 * ImageBlock::put for one sample per pixel centre (include/mitsuba/render/imageblock.h:144-185; the film is one block, samples
 * put in raster order), the division by the weight channel of Bitmap::convertMultiSpectrumAlphaWeight (bitmap.cpp:1617-1624).
 * frames: nPasses images [y][x][c]; out: the developed film. */
int orc_film(uint32_t W, uint32_t H, int filter, float param, const float *frames, uint32_t nPasses, float *out) {
    const int RES = 31;                                                    // MTS_FILTER_RESOLUTION
    Float radius, stddev = 0.5f;
    if (filter == ALVRL_FILTER_BOX) radius = (param > 0 ? param : 0.5f) + 1e-5f;        // box.cpp:38
    else if (filter == ALVRL_FILTER_TENT) radius = 1.0f;                               // tent.cpp:34
    else if (filter == ALVRL_FILTER_GAUSSIAN) { stddev = param > 0 ? param : 0.5f; radius = 4 * stddev; }   // gaussian.cpp:30-35
    else return seterr(ALVRL_ERR_ARG, "unknown filter");
    auto eval = [&](Float x) -> Float {
        if (filter == ALVRL_FILTER_BOX) return std::abs(x) <= radius ? 1.0f : 0.0f;    // box.cpp:46-48
        if (filter == ALVRL_FILTER_TENT) return std::max((Float) 0.0f, 1.0f - std::abs(x / radius));   // tent.cpp:42-44
        Float alpha = -1.0f / (2.0f * stddev * stddev);                                // gaussian.cpp:52-58
        return std::max((Float) 0.0f, fastexp(alpha * x * x) - fastexp(alpha * radius * radius));
    };
    Float values[RES + 1], sum = 0.0f;
    for (int i = 0; i < RES; ++i) { Float value = eval((radius * i) / RES); values[i] = value; sum += value; }
    values[RES] = 0.0f;
    const Float scaleFactor = RES / radius;
    sum *= 2 * radius / RES;
    const Float normalization = 1.0f / sum;
    for (int i = 0; i < RES; ++i) values[i] *= normalization;
    auto evalDiscretized = [&](Float x) { return values[std::min((int) std::abs(x * scaleFactor), RES)]; };

    const int channels = 5;
    std::vector<Float> bitmap((size_t) W * H * channels, 0.0f);
    std::vector<Float> weightsX(2 * (size_t) std::ceil(radius) + 2), weightsY(weightsX.size());
    for (uint32_t pass = 0; pass < nPasses; pass++) {
        const float *img = frames + (size_t) pass * W * H * 3;
        for (uint32_t sy = 0; sy < H; sy++) for (uint32_t sx = 0; sx < W; sx++) {
            const float *v = img + ((size_t) sy * W + sx) * 3;
            const Float value[5] = {v[0], v[1], v[2], 1.0f, 1.0f};
            bool bad = false;
            for (int i = 0; i < channels; ++i) if (!std::isfinite(value[i]) || value[i] < 0) bad = true;     // 147-151
            if (bad) continue;
            const Float posx = (sx + 0.5f) - 0.5f, posy = (sy + 0.5f) - 0.5f;         // offset 0, border cropped by the film
            const int minx = std::max((int) std::ceil(posx - radius), 0), miny = std::max((int) std::ceil(posy - radius), 0),
                      maxx = std::min((int) std::floor(posx + radius), (int) W - 1), maxy = std::min((int) std::floor(posy + radius), (int) H - 1);
            for (int x = minx, idx = 0; x <= maxx; ++x) weightsX[idx++] = evalDiscretized(x - posx);
            for (int y = miny, idx = 0; y <= maxy; ++y) weightsY[idx++] = evalDiscretized(y - posy);
            for (int y = miny, yr = 0; y <= maxy; ++y, ++yr) {
                const Float weightY = weightsY[yr];
                Float *dest = bitmap.data() + ((size_t) y * W + minx) * channels;
                for (int x = minx, xr = 0; x <= maxx; ++x, ++xr) {
                    const Float weight = weightsX[xr] * weightY;
                    for (int k = 0; k < channels; ++k) *dest++ += weight * value[k];
                }
            }
        }
    }
    for (size_t k = 0; k < (size_t) W * H; k++) {
        const Float weight = bitmap[k * channels + 4], invWeight = weight == 0 ? 0 : (Float) 1 / weight;
        for (int i = 0; i < 3; i++) out[3 * k + i] = bitmap[k * channels + i] * invWeight;
    }
    return ALVRL_OK;
}

} // extern "C"
