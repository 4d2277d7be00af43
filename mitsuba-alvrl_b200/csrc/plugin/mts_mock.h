/*
 * mts_mock.h -- the slice of Mitsuba 0.6's plugin interface that the `vrl` integrator touches, mocked so that the plugin
 * shim (vrl_plugin.cpp) compiles and is testable in an image without Mitsuba (Boost, Xerces-C, OpenEXR and SCons are not
 * installed here, so the real headers do not compile).  A maintainer building inside a Mitsuba tree defines
 * ALVRL_WITH_MITSUBA and these declarations are replaced by <mitsuba/render/scene.h>.
 *
 * Mirrors: Properties (include/mitsuba/core/properties.h), the CreateInstance/GetDescription plugin ABI
 * (include/mitsuba/core/cobject.h:99-107), Integrator::preprocess/render hooks (include/mitsuba/render/integrator.h:49-130,
 * 483-511) and the few Scene/Sensor/Film/Medium/TriMesh accessors preprocess() marshals from.
 */
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace mts {

/* Log(EError, ...) throws std::runtime_error (src/libcore/logger.cpp:100,147) */
[[noreturn]] inline void LogError(const std::string &msg) { throw std::runtime_error(msg); }

class Properties {
public:
    void setBoolean(const std::string &k, bool v) { m_b[k] = v; }
    void setInteger(const std::string &k, int v) { m_i[k] = v; }
    void setFloat(const std::string &k, float v) { m_f[k] = v; }
    void setString(const std::string &k, const std::string &v) { m_s[k] = v; }
    bool hasProperty(const std::string &k) const { return m_b.count(k) || m_i.count(k) || m_f.count(k) || m_s.count(k); }
    bool getBoolean(const std::string &k, bool d) const { m_q.insert(k); auto it = m_b.find(k); return it == m_b.end() ? d : it->second; }
    int getInteger(const std::string &k, int d) const { m_q.insert(k); auto it = m_i.find(k); return it == m_i.end() ? d : it->second; }
    float getFloat(const std::string &k, float d) const {
        m_q.insert(k);
        auto it = m_f.find(k); if (it != m_f.end()) return it->second;
        auto ii = m_i.find(k); return ii == m_i.end() ? d : (float) ii->second;
    }
    std::string getString(const std::string &k, const std::string &d) const { m_q.insert(k); auto it = m_s.find(k); return it == m_s.end() ? d : it->second; }
    /* the loader warns about attributes that were never queried (src/librender/scenehandler.cpp:792-795) */
    std::vector<std::string> getUnqueried() const {
        std::vector<std::string> out;
        for (auto &kv : m_b) if (!m_q.count(kv.first)) out.push_back(kv.first);
        for (auto &kv : m_i) if (!m_q.count(kv.first)) out.push_back(kv.first);
        for (auto &kv : m_f) if (!m_q.count(kv.first)) out.push_back(kv.first);
        for (auto &kv : m_s) if (!m_q.count(kv.first)) out.push_back(kv.first);
        return out;
    }
private:
    std::map<std::string, bool> m_b; std::map<std::string, int> m_i; std::map<std::string, float> m_f; std::map<std::string, std::string> m_s;
    mutable std::set<std::string> m_q;
};

/* Stream (include/mitsuba/core/stream.h): what serialize() / the unserializing constructor use -- typed values in sequence */
class Stream {
public:
    void write(const void *p, size_t n) { put(p, n); }                       /* Stream::write, stream.h */
    void writeInt(int v) { put(&v, sizeof(v)); }
    void writeFloat(float v) { put(&v, sizeof(v)); }
    void writeBool(bool v) { uint8_t b = v ? 1 : 0; put(&b, 1); }           /* stream.h: writeBool = writeUChar */
    int readInt() { int v; get(&v, sizeof(v)); return v; }
    float readFloat() { float v; get(&v, sizeof(v)); return v; }
    bool readBool() { uint8_t b; get(&b, 1); return b != 0; }
    const std::vector<uint8_t> &bytes() const { return m_data; }
    void seek(size_t pos) { m_pos = pos; }
private:
    void put(const void *p, size_t n) { const uint8_t *b = static_cast<const uint8_t *>(p); m_data.insert(m_data.end(), b, b + n); }
    void get(void *p, size_t n) {
        if (m_pos + n > m_data.size()) LogError("Stream: read beyond the end of the stream");
        std::copy(m_data.begin() + m_pos, m_data.begin() + m_pos + n, static_cast<uint8_t *>(p)); m_pos += n;
    }
    std::vector<uint8_t> m_data; size_t m_pos = 0;
};

/* what preprocess() reads from `const Scene *` (triangle meshes with diffuse BSDFs, one medium, a perspective sensor) */
/* bsdf: 0 diffuse (reflectance), 1 smooth dielectric (eta = intIOR / extIOR), 2 smooth conductor (eta, k rgb); the shape's
 * interior / exterior medium is the scene's medium or none (Shape::getInteriorMedium / getExteriorMedium) */
struct SurfaceView {            /* Shape::getBSDF / getInteriorMedium / getExteriorMedium of any shape */
    float reflectance[3] = {0.5f, 0.5f, 0.5f}; bool smooth = true;
    int bsdf = 0; float eta[3] = {1, 1, 1}, k[3] = {0, 0, 0}, specularReflectance[3] = {1, 1, 1}, specularTransmittance[3] = {1, 1, 1};
    bool mediumTransition = false, interiorMedium = false, exteriorMedium = false;
};
struct TriMeshView : SurfaceView { const float *positions; uint32_t vertexCount; const uint32_t *indices; uint32_t triangleCount; };
/* the analytic shapes of the reference's scenes: `rectangle` (src/shapes/rectangle.cpp: toWorld, flipNormals) and `sphere`
 * (src/shapes/sphere.cpp: centre and radius after the constructor folded toWorld into them, flipNormals) */
struct AnalyticShapeView : SurfaceView {
    enum Type { ERectangle = 0, ESphere = 1 } type = ERectangle;
    float toWorld[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; float center[3] = {0, 0, 0}, radius = 1; bool flipNormals = false;
};
struct MediumView {
    bool homogeneous; float sigmaA[3], sigmaS[3]; float mediumSamplingWeight; int phaseType; float g;
    const float *grid; int res[3]; float bboxMin[3], bboxMax[3]; float scale; float albedo[3];
};
struct SensorView { float sampleToCamera[16], cameraToWorld[16]; uint32_t width, height; float nearClip, farClip; float position[3]; };
/* Film: its reconstruction filter (Film::getReconstructionFilter: 0 box, 1 tent, 2 gaussian = ALVRL_FILTER_*; param <= 0 = the
 * filter's default radius / stddev) and where the developed image goes */
class Film {
public:
    virtual ~Film() {}
    int rfilter = 0; float rfilterParam = 0.0f;
    virtual void setImage(const float *rgb, uint32_t width, uint32_t height) = 0;
    /* ProgressiveMonteCarloIntegrator::dumpPass (src/librender/integrator.cpp:361-378): setDestinationFile(passFile), develop(),
     * setDestinationFile(original) -- here the developed image arrives with the file name */
    virtual void dumpPass(const std::string &passFile, const float *rgb, uint32_t width, uint32_t height) { (void) passFile; (void) rgb; (void) width; (void) height; }
};
/* an area emitter attached to a triangle mesh (Scene::getEmitters, src/emitters/area.cpp): mesh index + radiance */
struct EmitterView { uint32_t meshIndex; float radiance[3]; bool onAnalyticShape = false; /* meshIndex then counts Scene::shapes */ };
class Scene {
public:
    std::vector<TriMeshView> meshes; std::vector<AnalyticShapeView> shapes; std::vector<MediumView> media; std::vector<EmitterView> emitters; SensorView sensor; Film *film = nullptr;
    std::string destinationFile;                          /* Scene::getDestinationFile */
};

class Integrator {
public:
    virtual ~Integrator() {}
    virtual bool preprocess(const Scene *scene) = 0;
    virtual bool prepass(const Scene *scene) = 0;
    virtual bool render(Scene *scene) = 0;
    virtual void serialize(Stream *stream) const = 0;      /* network rendering: the object travels to the nodes */
    virtual void cancel() {}
};

} // namespace mts

extern "C" {
void *CreateInstance(const mts::Properties &props);     /* include/mitsuba/core/cobject.h:99-107 */
const char *GetDescription();
}
