/*
 * plugin_on_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  The plugin shim (csrc/plugin/vrl_plugin.cpp, unchanged) compiled
 * against the CPU oracle's mirror of the C ABI (orc_* in oracle/liborc.so) instead of libalvrl.so, so that the shim's host
 * logic -- marshalling of meshes, analytic shapes, materials, the medium, the sensor and the emitter out of the scene, the
 * order of the calls in preprocess / prepass / render, the progressive passes into the film -- runs in the CPU suite on a
 * machine without a GPU (tests/test_plugin_oracle_cpu.py).  Built by that test into a temporary directory; never shipped,
 * never loaded by the product.
 *
 * Every alvrl_* call of the shim is renamed to its orc_* mirror.  What the oracle does not mirror as a handle method is
 * adapted here: alvrl_load_vrl_file (the 9-floats-per-line reader, VRL.h:43-54,120-128, in front of orc_set_vrls) and the
 * film (alvrl_film_configure / put / develop over orc_film, which takes all passes at once).
 */
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define alvrl_create orc_create
#define alvrl_destroy orc_destroy
#define alvrl_last_error orc_last_error
#define alvrl_params_default orc_params_default
#define alvrl_set_mesh orc_set_mesh
#define alvrl_add_rectangle orc_add_rectangle
#define alvrl_add_sphere orc_add_sphere
#define alvrl_set_materials orc_set_materials
#define alvrl_set_material_optics orc_set_material_optics
#define alvrl_set_extra_bounds orc_set_extra_bounds
#define alvrl_set_medium_homogeneous orc_set_medium_homogeneous
#define alvrl_set_medium_grid orc_set_medium_grid
#define alvrl_set_vrls orc_set_vrls
#define alvrl_set_area_emitter orc_set_area_emitter
#define alvrl_set_seed orc_set_seed
#define alvrl_build_slices orc_build_slices
#define alvrl_trace_vrls orc_trace_vrls
#define alvrl_prepass orc_prepass
#define alvrl_get_stats orc_get_stats
#define alvrl_get_primary_hits orc_get_primary_hits
#define alvrl_get_pixel_to_slice orc_get_pixel_to_slice
#define alvrl_get_num_slices orc_get_num_slices
#define alvrl_get_cluster_counts orc_get_cluster_counts
#define alvrl_get_num_vrls orc_get_num_vrls
#define alvrl_set_camera t_set_camera
#define alvrl_render t_render
#define alvrl_render_unclustered t_render_unclustered
#define alvrl_load_vrl_file t_load_vrl_file
#define alvrl_film_configure t_film_configure
#define alvrl_film_put t_film_put
#define alvrl_film_develop t_film_develop
#include "../../include/alvrl.h"

extern "C" {
int orc_set_camera(alvrl_handle h, const float s2c[16], const float c2w[16], uint32_t W, uint32_t H, float nearClip, float farClip);
int orc_render(alvrl_handle h, float *rgb);
int orc_render_unclustered(alvrl_handle h, float *rgb);
int orc_film(uint32_t W, uint32_t H, int filter, float param, const float *frames, uint32_t n, float *out);
}

namespace {
struct HandleState { uint32_t W = 0, H = 0; std::vector<float> last, passes; int filter = 0; float param = 0; uint32_t nPasses = 0; bool haveFilm = false; };
std::map<void *, HandleState> g_state;
}

extern "C" {
int t_set_camera(alvrl_handle h, const float s2c[16], const float c2w[16], uint32_t W, uint32_t H, float nearClip, float farClip) {
    HandleState &s = g_state[(void *) h]; s.W = W; s.H = H;
    return orc_set_camera(h, s2c, c2w, W, H, nearClip, farClip);
}
int t_render(alvrl_handle h, float *rgb) {
    const int rc = orc_render(h, rgb);
    HandleState &s = g_state[(void *) h];
    if (rc == 0) s.last.assign(rgb, rgb + 3 * (size_t) s.W * s.H);
    return rc;
}
int t_render_unclustered(alvrl_handle h, float *rgb) {
    const int rc = orc_render_unclustered(h, rgb);
    HandleState &s = g_state[(void *) h];
    if (rc == 0) s.last.assign(rgb, rgb + 3 * (size_t) s.W * s.H);
    return rc;
}
int t_load_vrl_file(alvrl_handle h, const char *path) {
    std::ifstream f(path);
    if (!f) return ALVRL_ERR_IO;
    std::vector<float> s, e, p;
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss(line);
        float v[9]; int k = 0;
        while (k < 9 && (ss >> v[k])) k++;
        if (k < 9) break;
        s.insert(s.end(), v, v + 3); e.insert(e.end(), v + 3, v + 6); p.insert(p.end(), v + 6, v + 9);
    }
    return alvrl_set_vrls(h, s.data(), e.data(), p.data(), (uint32_t) (s.size() / 3), 0);
}
int t_film_configure(alvrl_handle h, int filter, float param) {
    HandleState &s = g_state[(void *) h];
    s.filter = filter; s.param = param; s.passes.clear(); s.nPasses = 0; s.haveFilm = true;
    return 0;
}
int t_film_put(alvrl_handle h, const float *rgb) {
    HandleState &s = g_state[(void *) h];
    if (!s.haveFilm) return ALVRL_ERR_STATE;
    const size_t n = 3 * (size_t) s.W * s.H;
    if (rgb) s.passes.insert(s.passes.end(), rgb, rgb + n);
    else { if (s.last.size() != n) return ALVRL_ERR_STATE; s.passes.insert(s.passes.end(), s.last.begin(), s.last.end()); }
    s.nPasses++;
    return 0;
}
int t_film_develop(alvrl_handle h, float *rgb) {
    HandleState &s = g_state[(void *) h];
    if (!s.haveFilm || !s.nPasses) return ALVRL_ERR_STATE;
    return orc_film(s.W, s.H, s.filter, s.param, s.passes.data(), s.nPasses, rgb);
}
}

#include "../../mitsuba-alvrl_b200/csrc/plugin/vrl_plugin.cpp"
