/*
 * vrl_plugin.cpp -- host side of the drop-in: `integrator type="vrl"` as a Mitsuba plugin whose hot path runs in
 * libalvrl.so (include/alvrl.h) instead of the reference's CPU loops.
 *
 * Mirrors src/integrators/vrl/vrlIntegrator.cpp: same XML parameter names, defaults and constructor errors (128-208),
 * `preprocess` loads the ASCII `vrlFile` and builds the slices (237-267), `prepass` runs sampleSliceMapping -> "Building
 * R" -> buildClusters (270-356).  Where the reference answers Li() per camera sample on worker threads (386-393), a GPU
 * replacement renders the whole frame at once: render() hands the framebuffer to the film ("Threading").
 * Exported plugin ABI: CreateInstance / GetDescription (include/mitsuba/core/cobject.h:99-107, vrlIntegrator.cpp:1127).
 */
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include "../../../include/alvrl.h"
#ifdef ALVRL_WITH_MITSUBA
#error "inside a Mitsuba tree: replace mts_mock.h by the Mitsuba headers as described in INTEGRATION.md"
#else
#include "mts_mock.h"
#endif

namespace {

class vrlIntegrator : public mts::Integrator {
public:
    explicit vrlIntegrator(const mts::Properties &props) {
        if (props.hasProperty("nc"))
            mts::LogError("Neighbourcount is now called 'neighbourCount' instead of 'nc'!");            /* 129-131 */
        alvrl_params_default(&m_p);
        m_p.shortVrls = props.getBoolean("shortVrls", true);
        m_p.vrlTargetNum = props.getInteger("vrlTargetNum", 500);
        m_p.maxParticleDepth = props.getInteger("maxParticleDepth", -1);
        m_p.specularForcedRRdepth = props.getInteger("specularForcedRRdepth", 100);
        m_p.initialSpecularThroughput = props.getFloat("initialSpecularThroughput", 20);
        m_p.volVolSamples = props.getInteger("volVolSamples", 2);
        if (m_p.volVolSamples != 0 && m_p.volVolSamples < 2)
            mts::LogError("Need at least 2 volVolSamples for variance estimate, but received: " + std::to_string(m_p.volVolSamples));
        m_p.volSurfSamples = props.getInteger("volSurfSamples", 2);
        if (m_p.volSurfSamples != 0 && m_p.volSurfSamples < 2)
            mts::LogError("Need at least 2 volSurfSamples for variance estimate, but received: " + std::to_string(m_p.volSurfSamples));
        m_p.globalCluster = props.getBoolean("globalCluster", false);
        m_p.globalUndersampling = props.getFloat("globalUndersampling", -1);
        m_p.localRefinement = props.getBoolean("localRefinement", true);
        m_p.localUndersampling = props.getFloat("localUndersampling", -1);
        m_p.fallBackUndersampling = props.getFloat("fallBackUndersampling", 5);
        m_p.targetNumSlices = props.getInteger("targetNumSlices", 100);
        m_p.targetPixelUndersampling = props.getFloat("targetPixelUndersampling", 64);
        m_p.sliceCurvatureFactor = props.getFloat("sliceCurvatureFactor", 0.5f);
        m_p.neighbourCount = props.getInteger("neighbourCount", 0);
        m_p.neighbourWeight = props.getFloat("neighbourWeight", 0.0f);
        m_p.Rsamples = props.getInteger("Rsamples", 1);
        m_p.depthCorrection = props.getFloat("depthCorrection", 1);
        m_p.numVrlFalseColor = props.getBoolean("numVrlFalseColor", false);
        m_p.slicesFalseColor = props.getBoolean("slicesFalseColor", false);
        m_p.convergenceFalseColor = props.getBoolean("convergenceFalseColor", false);
        m_vrlFile = props.getString("vrlFile", "");
        /* inherited (src/librender/integrator.cpp:53,272-298,348-349): queried so that the loader does not warn */
        m_numPasses = props.getInteger("numPasses", 1);                                                  /* SamplingIntegrator, 53 */
        m_p.rrDepth = props.getInteger("rrDepth", 5);                     /* MonteCarloIntegrator, 270-306; the VRL tracer's roulette (279) */
        m_maxDepth = props.getInteger("maxDepth", -1);
        m_strictNormals = props.getBoolean("strictNormals", false); m_hideEmitters = props.getBoolean("hideEmitters", false);
        if (m_p.rrDepth <= 0) mts::LogError("'rrDepth' must be set to a value greater than zero!");
        if (m_maxDepth <= 0 && m_maxDepth != -1) mts::LogError("'maxDepth' must be set to -1 (infinite) or a value greater than zero!");
        m_p.maxPasses = m_maxPasses = props.getInteger("maxPasses", 1);                                  /* ProgressiveMonteCarloIntegrator, 348-349 */
        m_dumpPasses = props.getBoolean("dumpPasses", false);
        /* device selection has no XML equivalent in the reference */
        m_device = props.getInteger("cudaDevice", 0);
        m_p.seed = (uint64_t) props.getInteger("seed", 0);
        /* polar steps of the triangles a `sphere` shape is handed over as (alvrl_add_sphere; 0 = the library's default) */
        m_sphereTessellation = props.getInteger("sphereTessellation", 0);
        if (m_sphereTessellation != 0 && m_sphereTessellation < 3) mts::LogError("'sphereTessellation' must be 0 (default) or at least 3");
    }
    /* the unserializing constructor and serialize(): what travels to a network node (vrlIntegrator.cpp:210-235 after the base
     * classes, src/librender/integrator.cpp:56-63, 307-321, 351-360) -- the reference's own field order; like the reference, the
     * clustering parameters do not travel */
    explicit vrlIntegrator(mts::Stream *stream) {
        alvrl_params_default(&m_p);
        m_numPasses = stream->readInt();
        m_p.rrDepth = stream->readInt(); m_maxDepth = stream->readInt(); m_strictNormals = stream->readBool(); m_hideEmitters = stream->readBool();
        m_p.maxPasses = m_maxPasses = stream->readInt(); m_dumpPasses = stream->readBool();
        m_p.volVolSamples = stream->readInt(); m_p.volSurfSamples = stream->readInt();
        m_p.globalCluster = stream->readBool(); m_p.localRefinement = stream->readBool();
        m_p.specularForcedRRdepth = stream->readInt(); m_p.initialSpecularThroughput = stream->readFloat();
        m_p.shortVrls = stream->readBool();
    }
    void serialize(mts::Stream *stream) const override {
        stream->writeInt(m_numPasses);
        stream->writeInt(m_p.rrDepth); stream->writeInt(m_maxDepth); stream->writeBool(m_strictNormals); stream->writeBool(m_hideEmitters);
        stream->writeInt(m_maxPasses); stream->writeBool(m_dumpPasses);
        stream->writeInt(m_p.volVolSamples); stream->writeInt(m_p.volSurfSamples);
        stream->writeBool(m_p.globalCluster != 0); stream->writeBool(m_p.localRefinement != 0);
        stream->writeInt(m_p.specularForcedRRdepth); stream->writeFloat(m_p.initialSpecularThroughput);
        stream->writeBool(m_p.shortVrls != 0);
    }
    ~vrlIntegrator() override { if (m_h) alvrl_destroy(m_h); }

    const alvrl_params &params() const { return m_p; }

    /* vrlIntegrator::preprocess, 237-267 */
    bool preprocess(const mts::Scene *scene) override {
        if (!m_vrlFile.empty()) {
            if (scene->media.size() != 1)                                                                /* 244-248 */
                mts::LogError("When loading VRLs from a file, the scene should (currently) contain exactly one medium, which will be the "
                              "medium where all VRLs will 'live'");
        } else {
            /* the VRLs are traced in every prepass (279-280): the device tracer walks one medium and starts on one area emitter */
            if (scene->media.size() != 1) mts::LogError("the device VRL tracer needs exactly one medium in the scene");
            if (scene->emitters.size() != 1) mts::LogError("the device VRL tracer needs exactly one area emitter (attached to a triangle mesh) in the scene");
            if (scene->emitters[0].meshIndex >= (scene->emitters[0].onAnalyticShape ? scene->shapes.size() : scene->meshes.size()))
                mts::LogError("the emitter's shape is not a mesh of the scene");
        }
        alvrl_params hp = m_p; hp.maxPasses = 1;                              /* the passes are driven from render() below */
        /* the false-colour debug outputs (vrlIntegrator.cpp:199-201) are produced here, from the library's slice map and cluster
         * counts (falseColorPass below); the library itself only ever renders radiance */
        hp.numVrlFalseColor = hp.slicesFalseColor = hp.convergenceFalseColor = 0;
        if (m_p.numVrlFalseColor || m_p.slicesFalseColor || m_p.convergenceFalseColor) {
            for (const mts::TriMeshView &tm : scene->meshes) if (tm.bsdf != 0) mts::LogError("the false-colour outputs are not supported with specular surfaces (the chain segments' weights stay on the device)");
            for (const mts::AnalyticShapeView &sh : scene->shapes) if (sh.bsdf != 0) mts::LogError("the false-colour outputs are not supported with specular surfaces (the chain segments' weights stay on the device)");
            if (m_p.slicesFalseColor && !m_p.numVrlFalseColor && !(m_p.globalCluster || m_p.localRefinement))
                mts::LogError("requested slices false color image without clustering!");                 /* 438-440 */
        }
        chk(alvrl_create(m_device, &hp, &m_h));
        /* triangle soup + one diffuse material per mesh */
        std::vector<float> verts, albedo, optics; std::vector<uint32_t> tris, mat, bits;
        bool anyDelta = false;
        std::vector<uint32_t> firstTri;
        /* BSDF type bits (bsdf.h:230-284) and the media on the two sides of the shape (shape.h:427-433): what the specular
         * chains of LiInternal read (vrlIntegrator.cpp:445-511).  One material per shape. */
        auto material = [&](const mts::SurfaceView &sv) {
            albedo.insert(albedo.end(), sv.reflectance, sv.reflectance + 3);
            uint32_t b = sv.smooth ? ALVRL_BSDF_SMOOTH : 0u;
            if (sv.bsdf == 1) b = ALVRL_BSDF_DIELECTRIC; else if (sv.bsdf == 2) b = ALVRL_BSDF_CONDUCTOR;
            if (sv.mediumTransition) b |= ALVRL_MAT_TRANSITION | (sv.interiorMedium ? ALVRL_MAT_INTERIOR_MEDIUM : 0u) | (sv.exteriorMedium ? ALVRL_MAT_EXTERIOR_MEDIUM : 0u);
            anyDelta = anyDelta || (b & ALVRL_BSDF_DELTA);
            bits.push_back(b);
            const float o[12] = {sv.eta[0], sv.eta[1], sv.eta[2], sv.k[0], sv.k[1], sv.k[2], sv.specularReflectance[0], sv.specularReflectance[1],
                                 sv.specularReflectance[2], sv.specularTransmittance[0], sv.specularTransmittance[1], sv.specularTransmittance[2]};
            optics.insert(optics.end(), o, o + 12);
            return (uint32_t) (bits.size() - 1);
        };
        for (size_t m = 0; m < scene->meshes.size(); m++) {
            const mts::TriMeshView &tm = scene->meshes[m];
            const uint32_t base = (uint32_t) (verts.size() / 3);
            firstTri.push_back((uint32_t) (tris.size() / 3));
            verts.insert(verts.end(), tm.positions, tm.positions + 3 * (size_t) tm.vertexCount);
            for (uint32_t i = 0; i < 3 * tm.triangleCount; i++) tris.push_back(base + tm.indices[i]);
            mat.insert(mat.end(), tm.triangleCount, material(tm));
        }
        if (!tris.empty()) chk(alvrl_set_mesh(m_h, verts.data(), (uint32_t) (verts.size() / 3), tris.data(), (uint32_t) (tris.size() / 3), mat.data()));
        /* analytic shapes go in as triangles (alvrl_add_rectangle: the same surface; alvrl_add_sphere: vertices on the sphere) */
        std::vector<uint32_t> shapeFirstTri, shapeTriCount;
        for (const mts::AnalyticShapeView &sh : scene->shapes) {
            uint32_t first = 0, count = 2;
            const uint32_t id = material(sh);
            if (sh.type == mts::AnalyticShapeView::ERectangle) chk(alvrl_add_rectangle(m_h, sh.toWorld, sh.flipNormals, id, &first));
            else chk(alvrl_add_sphere(m_h, sh.center, sh.radius, sh.flipNormals, (uint32_t) m_sphereTessellation, id, &first, &count));
            shapeFirstTri.push_back(first); shapeTriCount.push_back(count);
        }
        chk(alvrl_set_materials(m_h, albedo.data(), bits.data(), (uint32_t) bits.size()));
        if (anyDelta) chk(alvrl_set_material_optics(m_h, optics.data(), (uint32_t) bits.size()));
        chk(alvrl_set_extra_bounds(m_h, scene->sensor.position, 1));                                    /* scene.cpp:387-413 */
        const mts::MediumView &md = scene->media[0];
        if (md.homogeneous) chk(alvrl_set_medium_homogeneous(m_h, md.sigmaA, md.sigmaS, md.mediumSamplingWeight, md.phaseType, md.g));
        else chk(alvrl_set_medium_grid(m_h, md.grid, md.res, md.bboxMin, md.bboxMax, md.scale, md.albedo, md.sigmaS, md.phaseType, md.g));
        const mts::SensorView &s = scene->sensor;
        chk(alvrl_set_camera(m_h, s.sampleToCamera, s.cameraToWorld, s.width, s.height, s.nearClip, s.farClip));
        if (!m_vrlFile.empty()) chk(alvrl_load_vrl_file(m_h, m_vrlFile.c_str()));                       /* 249-251 */
        else {                                                                                          /* the emitter's shape: Scene::getEmitters */
            const mts::EmitterView &em = scene->emitters[0];
            const uint32_t emFirst = em.onAnalyticShape ? shapeFirstTri[em.meshIndex] : firstTri[em.meshIndex];
            std::vector<uint32_t> emTris(em.onAnalyticShape ? shapeTriCount[em.meshIndex] : scene->meshes[em.meshIndex].triangleCount);
            for (uint32_t i = 0; i < (uint32_t) emTris.size(); i++) emTris[i] = emFirst + i;
            chk(alvrl_set_area_emitter(m_h, emTris.data(), (uint32_t) emTris.size(), em.radiance));
        }
        if (m_p.globalCluster || m_p.localRefinement) chk(alvrl_build_slices(m_h));                     /* 254-265 */
        m_pass = 0;
        return true;
    }
    /* vrlIntegrator::prepass, 270-356: one call per progressive pass */
    bool prepass(const mts::Scene *) override {
        if (!m_h) mts::LogError("VRL filename given, but vrls were not loaded!");                       /* 285-286 */
        /* the reference's sampler keeps advancing from pass to pass (integrator.cpp:432-433); a pass of the counter stream is
         * addressed by its seed */
        if (m_pass > 0) chk(alvrl_set_seed(m_h, m_p.seed + (uint64_t) m_pass));
        m_pass++;
        if (m_vrlFile.empty()) chk(alvrl_trace_vrls(m_h, 0));                                           /* 276-280: vrlTracer::randomWalk */
        if (m_p.globalCluster || m_p.localRefinement) chk(alvrl_prepass(m_h));
        return true;
    }
    /* one render pass: every pixel centre through Li (386-393; MonteCarloIntegrator::render) */
    /* numVrlFalseColor / slicesFalseColor for scenes without delta surfaces: LiInternal returns LiDirect (447-448), which is the
     * value getClusteredVrlContributions / getVRLContributions compute instead of radiance (574-584, 806-807) for every camera ray
     * that hits something, and zero for the rays that leave the scene (418-423).  convergenceFalseColor only takes effect behind
     * delta surfaces (quirk B4: the early return at 447-448 comes first), so without them it is the radiance image. */
    void falseColorPass(mts::Scene *scene, std::vector<float> &rgb) {
        const uint32_t W = scene->sensor.width, H = scene->sensor.height, P = W * H;
        std::vector<uint32_t> prim(P), toSlice(P, ALVRL_NO_SLICE), offset;
        chk(alvrl_get_primary_hits(m_h, prim.data(), nullptr, nullptr, nullptr));
        const bool clustered = m_p.globalCluster || m_p.localRefinement;
        uint32_t nVrls = 0, nGlobal = 0, nFallback = 0, nSlices = 0, nRows = 0;
        if (clustered) {
            chk(alvrl_get_pixel_to_slice(m_h, toSlice.data()));
            chk(alvrl_get_num_slices(m_h, &nSlices, &nRows));
            if (m_p.numVrlFalseColor) {
                offset.resize(nSlices + 1);
                chk(alvrl_get_cluster_counts(m_h, offset.data(), &nGlobal, &nFallback));
                chk(alvrl_get_num_vrls(m_h, &nVrls));
            }
        }
        for (uint32_t y = 0; y < H; y++)
            for (uint32_t x = 0; x < W; x++) {
                const uint32_t i = y + H * x;                                   /* m_ci->m_slices[y + sizeY * x], 558 */
                float c[3] = {0, 0, 0};
                if (prim[i] != ALVRL_NO_HIT) {
                    const uint32_t sl = toSlice[i];
                    if (m_p.numVrlFalseColor) {
                        if (!clustered) c[0] = c[1] = c[2] = 1.0f;              /* Li = weight, 806-807 */
                        else {
                            uint32_t count = sl == ALVRL_NO_SLICE ? nFallback : offset[sl + 1] - offset[sl];
                            if (count == 0) count = nFallback;                  /* a slice without a list of its own renders with the fallback list */
                            c[0] = c[1] = c[2] = (float) count / (float) nVrls;  /* 574-575 */
                        }
                    } else if (sl == ALVRL_NO_SLICE) c[0] = c[1] = c[2] = 0.5f;   /* "fallback clustering in gray", 577-578 */
                    else {                                                      /* 580-584, in the reference's unsigned arithmetic */
                        c[0] = (float) (((sl + sl * sl) % 43u) / 43.0);
                        c[1] = (float) (((7u * sl + 2u * sl * sl + 7u) % 41u) / 41.0);
                        c[2] = (float) (((23u * sl + 5u * sl * sl + sl * sl * sl + 17u) % 53u) / 53.0);
                    }
                }
                float *o = &rgb[3 * ((size_t) y * W + x)];
                o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
            }
    }
    bool renderPass(mts::Scene *scene, std::vector<float> &rgb) {
        rgb.resize((size_t) scene->sensor.width * scene->sensor.height * 3);
        if (m_p.numVrlFalseColor || m_p.slicesFalseColor) { falseColorPass(scene, rgb); return true; }
        if (m_p.globalCluster || m_p.localRefinement) chk(alvrl_render(m_h, rgb.data()));
        else chk(alvrl_render_unclustered(m_h, rgb.data()));
        return true;
    }
    /* ProgressiveMonteCarloIntegrator::render (src/librender/integrator.cpp:380-440): prepass + render pass, maxPasses times,
     * into a film that is not cleared in between; the film's reconstruction filter and the division by the accumulated weights
     * run on the device (alvrl_film_*).  One pass through a box filter is the render pass's image itself. */
    bool render(mts::Scene *scene) override {
        std::vector<float> rgb;
        const int filter = scene->film ? scene->film->rfilter : ALVRL_FILTER_BOX;
        const bool endless = m_maxPasses < 0;                                /* "while (m_maxPasses < 0 || pass <= m_maxPasses)", 398 */
        const bool useFilm = endless || m_maxPasses > 1 || filter != ALVRL_FILTER_BOX;
        if (useFilm) chk(alvrl_film_configure(m_h, filter, scene->film ? scene->film->rfilterParam : 0.0f));
        m_cancelled = false;
        double prepassCpu = 0, prepassWall = 0, renderCpu = 0, renderWall = 0;     /* cumulative, as integrator.cpp:394-430 */
        int pass = 1;
        for (; (endless || pass <= m_maxPasses) && !m_cancelled; pass++) {  /* cancel() takes effect between passes */
            const double w0 = wallSeconds(), c0 = cpuSeconds();
            if (!prepass(scene)) return false;
            const double w1 = wallSeconds(), c1 = cpuSeconds();
            if (!renderPass(scene, rgb)) return false;
            /* the frame alvrl_render left on the device; a false-colour pass was made here, on the host */
            if (useFilm) chk(alvrl_film_put(m_h, (m_p.numVrlFalseColor || m_p.slicesFalseColor) ? rgb.data() : nullptr));
            prepassWall += w1 - w0; prepassCpu += c1 - c0; renderWall += wallSeconds() - w1; renderCpu += cpuSeconds() - c1;
            if (endless && m_dumpPasses && scene->film) {                     /* without a pass count every pass is dumped, 428-430 */
                std::vector<float> sofar(rgb.size());
                chk(alvrl_film_develop(m_h, sofar.data()));
                scene->film->dumpPass(passFileName(scene->destinationFile, pass, prepassCpu, prepassWall, renderCpu, renderWall), sofar.data(),
                                      scene->sensor.width, scene->sensor.height);
            }
        }
        pass--;
        if (useFilm) chk(alvrl_film_develop(m_h, rgb.data()));
        if (scene->film) scene->film->setImage(rgb.data(), scene->sensor.width, scene->sensor.height);
        /* dumpPasses: with a pass count only the last pass is dumped (integrator.cpp:436-438), under a name that carries the
         * cumulative prepass / render times and, from passFileSuffix (vrlIntegrator.cpp:357-364), the two StatsCounters */
        if (!endless && m_dumpPasses && scene->film && pass >= 1)
            scene->film->dumpPass(passFileName(scene->destinationFile, pass, prepassCpu, prepassWall, renderCpu, renderWall), rgb.data(),
                                  scene->sensor.width, scene->sensor.height);
        return true;
    }
    /* ProgressiveMonteCarloIntegrator::dumpPass's file name (integrator.cpp:365-374) */
    std::string passFileName(const std::string &origFile, int pass, double prepassCpu, double prepassWall, double renderCpu, double renderWall) const {
        alvrl_stats st; memset(&st, 0, sizeof(st));
        chk(alvrl_get_stats(m_h, &st));
        char buf[512];
        snprintf(buf, sizeof(buf), "_pass%03d_precpu%.4e_prewall%.4e_rencpu%.4e_renwall%.4e_prevrl%.4e_renvrl%.4e.blahExtensionTODO", pass,
                 prepassCpu, prepassWall, renderCpu, renderWall, (double) (float) st.pairsPreprocess, (double) (float) st.pairsRender);
        return origFile + buf;
    }
    static double wallSeconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    static double cpuSeconds() { return (double) std::clock() / CLOCKS_PER_SEC; }        /* user + system of the process, as cpu_timer */
    /* SamplingIntegrator::cancel (integrator.cpp:68-71) stops the block scheduler; here: no further pass is started */
    void cancel() override { m_cancelled = true; }
    alvrl_handle handle() const { return m_h; }
private:
    static void chk(int rc) { if (rc != ALVRL_OK) mts::LogError(alvrl_last_error()); }
    alvrl_params m_p; std::string m_vrlFile; int m_device = 0, m_maxPasses = 1, m_pass = 0; alvrl_handle m_h = nullptr;
    int m_sphereTessellation = 0;
    int m_numPasses = 1, m_maxDepth = -1; bool m_strictNormals = false, m_hideEmitters = false, m_dumpPasses = false;
    volatile bool m_cancelled = false;
};

} // namespace

extern "C" {
void *CreateInstance(const mts::Properties &props) { return new vrlIntegrator(props); }
const char *GetDescription() { return "An implementation of Adaptive Lightslice for Virtual Ray Lights (B200 device path)"; }
/* test hook: the parsed parameter block of an instance */
void alvrl_plugin_get_params(void *inst, alvrl_params *out) { *out = static_cast<vrlIntegrator *>(inst)->params(); }
void alvrl_plugin_destroy(void *inst) { delete static_cast<vrlIntegrator *>(inst); }
/* test hooks: serialize an instance into a byte buffer (returns the size; buf may be NULL) and build one from such a buffer */
int alvrl_plugin_serialize(void *inst, uint8_t *buf, int cap) {
    mts::Stream st; static_cast<vrlIntegrator *>(inst)->serialize(&st);
    const int n = (int) st.bytes().size();
    if (buf && cap >= n) memcpy(buf, st.bytes().data(), (size_t) n);
    return n;
}
int alvrl_plugin_unserialize(const uint8_t *buf, int n, void **inst, char *err, int errLen) {
    try {
        mts::Stream st;
        st.write(buf, (size_t) n);
        st.seek(0);
        *inst = new vrlIntegrator(&st);
        return 0;
    } catch (const std::exception &e) { strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; return -1; }
}
/* test hooks for building a Properties object from C */
void *alvrl_plugin_props_new() { return new mts::Properties(); }
void alvrl_plugin_props_free(void *p) { delete static_cast<mts::Properties *>(p); }
void alvrl_plugin_props_set_int(void *p, const char *k, int v) { static_cast<mts::Properties *>(p)->setInteger(k, v); }
void alvrl_plugin_props_set_float(void *p, const char *k, float v) { static_cast<mts::Properties *>(p)->setFloat(k, v); }
void alvrl_plugin_props_set_bool(void *p, const char *k, int v) { static_cast<mts::Properties *>(p)->setBoolean(k, v != 0); }
void alvrl_plugin_props_set_string(void *p, const char *k, const char *v) { static_cast<mts::Properties *>(p)->setString(k, v); }
/* CreateInstance with the reference's error behaviour (Log(EError) throws) turned into a status + message */
int alvrl_plugin_create(void *props, void **inst, char *err, int errLen) {
    try { *inst = CreateInstance(*static_cast<mts::Properties *>(props)); return 0; }
    catch (const std::exception &e) { strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; return -1; }
}
/* test hooks: a mock Scene assembled from flat arrays (the arrays stay owned by the caller), and one whole frame through
 * the integrator's virtuals in the order the host calls them -- preprocess, prepass, render (scene.cpp:416-449,
 * integrator.cpp:380-440) -- with a Film that hands the image back */
void *alvrl_plugin_scene_new() { return new mts::Scene(); }
void alvrl_plugin_scene_free(void *s) { delete static_cast<mts::Scene *>(s); }
/* the BSDF and media of the mesh added last: bsdf 1 = dielectric (eta[0] = intIOR / extIOR), 2 = conductor (eta, k rgb) */
void alvrl_plugin_scene_set_mesh_bsdf(void *s, int bsdf, const float *eta, const float *k, int mediumTransition, int interiorMedium, int exteriorMedium) {
    mts::TriMeshView &m = static_cast<mts::Scene *>(s)->meshes.back();
    m.bsdf = bsdf; m.smooth = bsdf == 0;
    for (int i = 0; i < 3; i++) { m.eta[i] = eta ? eta[i] : 1.0f; m.k[i] = k ? k[i] : 0.0f; }
    m.mediumTransition = mediumTransition != 0; m.interiorMedium = interiorMedium != 0; m.exteriorMedium = exteriorMedium != 0;
}
void alvrl_plugin_scene_add_mesh(void *s, const float *positions, uint32_t nv, const uint32_t *indices, uint32_t nt, const float *reflectance, int smooth) {
    mts::TriMeshView m; m.positions = positions; m.vertexCount = nv; m.indices = indices; m.triangleCount = nt;
    m.reflectance[0] = reflectance[0]; m.reflectance[1] = reflectance[1]; m.reflectance[2] = reflectance[2]; m.smooth = smooth != 0;
    static_cast<mts::Scene *>(s)->meshes.push_back(m);
}
/* analytic shapes (mts::AnalyticShapeView): a `rectangle` under toWorld (row-major 4x4), a `sphere` by centre and radius */
void alvrl_plugin_scene_add_rectangle(void *s, const float *toWorld, int flipNormals, const float *reflectance) {
    mts::AnalyticShapeView v; v.type = mts::AnalyticShapeView::ERectangle; memcpy(v.toWorld, toWorld, 16 * sizeof(float)); v.flipNormals = flipNormals != 0;
    memcpy(v.reflectance, reflectance, 3 * sizeof(float));
    static_cast<mts::Scene *>(s)->shapes.push_back(v);
}
void alvrl_plugin_scene_add_sphere(void *s, const float *center, float radius, int flipNormals, const float *reflectance) {
    mts::AnalyticShapeView v; v.type = mts::AnalyticShapeView::ESphere; memcpy(v.center, center, 3 * sizeof(float)); v.radius = radius; v.flipNormals = flipNormals != 0;
    memcpy(v.reflectance, reflectance, 3 * sizeof(float));
    static_cast<mts::Scene *>(s)->shapes.push_back(v);
}
/* the BSDF and media of the analytic shape added last (as alvrl_plugin_scene_set_mesh_bsdf) */
void alvrl_plugin_scene_set_shape_bsdf(void *s, int bsdf, const float *eta, const float *k, int mediumTransition, int interiorMedium, int exteriorMedium) {
    mts::AnalyticShapeView &m = static_cast<mts::Scene *>(s)->shapes.back();
    m.bsdf = bsdf; m.smooth = bsdf == 0;
    for (int i = 0; i < 3; i++) { m.eta[i] = eta ? eta[i] : 1.0f; m.k[i] = k ? k[i] : 0.0f; }
    m.mediumTransition = mediumTransition != 0; m.interiorMedium = interiorMedium != 0; m.exteriorMedium = exteriorMedium != 0;
}
/* an area emitter on the analytic shape with this index (in the order they were added) */
void alvrl_plugin_scene_add_area_emitter_on_shape(void *s, uint32_t shapeIndex, const float *radiance) {
    mts::EmitterView e; e.meshIndex = shapeIndex; e.onAnalyticShape = true; e.radiance[0] = radiance[0]; e.radiance[1] = radiance[1]; e.radiance[2] = radiance[2];
    static_cast<mts::Scene *>(s)->emitters.push_back(e);
}
void alvrl_plugin_scene_add_medium_homogeneous(void *s, const float *sigmaA, const float *sigmaS, float weight, int phaseType, float g) {
    mts::MediumView m; memset(&m, 0, sizeof(m));
    m.homogeneous = true; m.mediumSamplingWeight = weight; m.phaseType = phaseType; m.g = g;
    for (int k = 0; k < 3; k++) { m.sigmaA[k] = sigmaA[k]; m.sigmaS[k] = sigmaS[k]; }
    static_cast<mts::Scene *>(s)->media.push_back(m);
}
void alvrl_plugin_scene_set_sensor(void *s, const float *sampleToCamera, const float *cameraToWorld, uint32_t w, uint32_t h, float nearClip,
                                   float farClip, const float *position) {
    mts::SensorView &v = static_cast<mts::Scene *>(s)->sensor;
    memcpy(v.sampleToCamera, sampleToCamera, 16 * sizeof(float)); memcpy(v.cameraToWorld, cameraToWorld, 16 * sizeof(float));
    v.width = w; v.height = h; v.nearClip = nearClip; v.farClip = farClip;
    memcpy(v.position, position, 3 * sizeof(float));
}
/* an area emitter on the mesh with this index (in the order of alvrl_plugin_scene_add_mesh) */
void alvrl_plugin_scene_add_area_emitter(void *s, uint32_t meshIndex, const float *radiance) {
    mts::EmitterView e; e.meshIndex = meshIndex; e.radiance[0] = radiance[0]; e.radiance[1] = radiance[1]; e.radiance[2] = radiance[2];
    static_cast<mts::Scene *>(s)->emitters.push_back(e);
}
int alvrl_plugin_render_frame_filtered(void *inst, void *scene, int rfilter, float rfilterParam, float *rgbOut, char *err, int errLen);
int alvrl_plugin_render_frame(void *inst, void *scene, float *rgbOut, char *err, int errLen) {
    return alvrl_plugin_render_frame_filtered(inst, scene, ALVRL_FILTER_BOX, 0.0f, rgbOut, err, errLen);
}
int alvrl_plugin_render_frame_filtered(void *inst, void *scene, int rfilter, float rfilterParam, float *rgbOut, char *err, int errLen) {
    struct CopyFilm : mts::Film {
        float *dst; explicit CopyFilm(float *d) : dst(d) {}
        void setImage(const float *rgb, uint32_t w, uint32_t h) override { memcpy(dst, rgb, (size_t) w * h * 3 * sizeof(float)); }
    } film(rgbOut);
    film.rfilter = rfilter; film.rfilterParam = rfilterParam;
    mts::Scene *sc = static_cast<mts::Scene *>(scene);
    mts::Integrator *it = static_cast<vrlIntegrator *>(inst);
    try {
        sc->film = &film;
        it->preprocess(sc); it->render(sc);           /* render() = ProgressiveMonteCarloIntegrator::render: prepass + render pass per pass */
        sc->film = nullptr;
        return 0;
    } catch (const std::exception &e) { sc->film = nullptr; strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; return -1; }
}
/* one frame with a destination file: returns the pass file name the film was handed (empty: dumpPasses is off) */
int alvrl_plugin_render_frame_dump(void *inst, void *scene, const char *destinationFile, float *rgbOut, char *passFile, int passFileLen, char *err, int errLen) {
    struct DumpFilm : mts::Film {
        float *dst; std::string name; std::vector<float> dumped;
        explicit DumpFilm(float *d) : dst(d) {}
        void setImage(const float *rgb, uint32_t w, uint32_t h) override { memcpy(dst, rgb, (size_t) w * h * 3 * sizeof(float)); }
        void dumpPass(const std::string &f, const float *rgb, uint32_t w, uint32_t h) override { name = f; dumped.assign(rgb, rgb + (size_t) w * h * 3); }
    } film(rgbOut);
    mts::Scene *sc = static_cast<mts::Scene *>(scene);
    mts::Integrator *it = static_cast<vrlIntegrator *>(inst);
    try {
        sc->film = &film; sc->destinationFile = destinationFile;
        it->preprocess(sc); it->render(sc);
        sc->film = nullptr;
        strncpy(passFile, film.name.c_str(), passFileLen - 1); passFile[passFileLen - 1] = 0;
        if (!film.name.empty() && memcmp(film.dumped.data(), rgbOut, film.dumped.size() * sizeof(float)) != 0) { strncpy(err, "the dumped pass is not the developed film", errLen - 1); return -2; }
        return 0;
    } catch (const std::exception &e) { sc->film = nullptr; strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; return -1; }
}
/* maxPasses < 0: render until cancelled.  The film of this hook cancels the integrator from inside the cancelAfter-th pass dump
 * (as a GUI thread would between passes); names receives the dumped pass files, one per line */
int alvrl_plugin_render_until_cancelled(void *inst, void *scene, const char *destinationFile, int cancelAfter, float *rgbOut, char *names, int namesLen,
                                        char *err, int errLen) {
    struct CancellingFilm : mts::Film {
        float *dst; mts::Integrator *it; int after, dumps = 0; std::string names; std::vector<float> last;
        CancellingFilm(float *d, mts::Integrator *i, int a) : dst(d), it(i), after(a) {}
        void setImage(const float *rgb, uint32_t w, uint32_t h) override { memcpy(dst, rgb, (size_t) w * h * 3 * sizeof(float)); }
        void dumpPass(const std::string &f, const float *rgb, uint32_t w, uint32_t h) override {
            names += f + "\n"; last.assign(rgb, rgb + (size_t) w * h * 3);
            if (++dumps >= after) it->cancel();
        }
    };
    mts::Scene *sc = static_cast<mts::Scene *>(scene);
    mts::Integrator *it = static_cast<vrlIntegrator *>(inst);
    CancellingFilm film(rgbOut, it, cancelAfter);
    try {
        sc->film = &film; sc->destinationFile = destinationFile;
        it->preprocess(sc); it->render(sc);
        sc->film = nullptr;
        strncpy(names, film.names.c_str(), namesLen - 1); names[namesLen - 1] = 0;
        if (film.last.empty() || memcmp(film.last.data(), rgbOut, film.last.size() * sizeof(float)) != 0) { strncpy(err, "the last dumped pass is not the final film", errLen - 1); return -2; }
        return film.dumps;
    } catch (const std::exception &e) { sc->film = nullptr; strncpy(err, e.what(), errLen - 1); err[errLen - 1] = 0; return -1; }
}
int alvrl_plugin_unqueried(void *props) { return (int) static_cast<mts::Properties *>(props)->getUnqueried().size(); }
}
