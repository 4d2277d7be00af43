/*
 * shapes.h -- the reference's analytic shapes as triangles (host only): the path's queries run on triangles, so a host that
 * holds a `rectangle` or a `sphere` hands it over through alvrl_add_rectangle / alvrl_add_sphere.
 *
 *   rectangle  src/shapes/rectangle.cpp: the square [-1, 1]^2 of the xy plane under toWorld, normal = toWorld(Normal(0, 0, 1))
 *              (101-105), flipNormals = toWorld * scale(1, 1, -1) (82-83).  Two triangles over the corners in the order of
 *              its own createTriMesh (170-196) cover exactly the same surface; only the hit ids differ (two triangles for
 *              one primitive).  The winding is chosen so that the triangles' geometric normal is the shape's normal, which
 *              for a mirroring transform means reversing it (the normal goes with the inverse transpose).
 *   sphere     src/shapes/sphere.cpp: centre, radius and flipNormals describe the surface (the constructor folds toWorld into
 *              them, 108-127); the normal is normalize(p - centre), negated by flipNormals (245-251).  The triangles are an
 *              approximation: rings of constant polar angle with shared poles and a closed seam (2 P (T - 2) triangles for T
 *              polar steps and P = 2 T azimuthal steps; no degenerate triangles, unlike the reference's preview mesh of
 *              389-458, which duplicates the seam and collapses a triangle per pole quad).  The vertices lie on the sphere;
 *              the surface deviates from it by at most radius * (1 - cos(rho)), rho = half the diagonal of a (dTheta, dPhi) cell =
 *              0.5 sqrt(2) pi / (T - 1) -- 6.2e-4 radius for T = 64.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace alvrl {

/* appends 4 vertices and 2 triangles; toWorld row-major 4x4 (affine) */
inline void tessellate_rectangle(const float toWorld[16], bool flipNormals, std::vector<float> &verts, std::vector<uint32_t> &tris) {
    const float *m = toWorld;
    /* determinant of the linear part, with the third column negated by flipNormals (rectangle.cpp:82-83) */
    const double sz = flipNormals ? -1.0 : 1.0;
    const double a = m[0], b = m[1], c = sz * m[2], d = m[4], e = m[5], f = sz * m[6], g = m[8], h = m[9], i = sz * m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (!(det != 0)) throw std::runtime_error("rectangle: 'toWorld' is singular");          /* before anything is appended */
    const bool reverse = det < 0;
    const float corners[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};                  /* rectangle.cpp:179-182 */
    const uint32_t base = (uint32_t) (verts.size() / 3);
    for (int k = 0; k < 4; k++)
        for (int r = 0; r < 3; r++) verts.push_back(m[4 * r + 0] * corners[k][0] + m[4 * r + 1] * corners[k][1] + m[4 * r + 3]);
    const uint32_t t[2][3] = {{0, 1, 2}, {2, 3, 0}};                                   /* rectangle.cpp:190-196 */
    for (int k = 0; k < 2; k++) {
        tris.push_back(base + t[k][0]);
        tris.push_back(base + (reverse ? t[k][2] : t[k][1]));
        tris.push_back(base + (reverse ? t[k][1] : t[k][2]));
    }
}

/* appends 2 + 2 T (T - 2) vertices and 4 T (T - 2) triangles (T = thetaSteps >= 3) */
inline void tessellate_sphere(const float center[3], float radius, bool flipNormals, uint32_t thetaSteps,
                              std::vector<float> &verts, std::vector<uint32_t> &tris) {
    if (!(radius > 0)) throw std::runtime_error("Cannot create spheres of radius <= 0");     /* sphere.cpp:130-131 */
    if (thetaSteps < 3) throw std::runtime_error("sphere: at least 3 polar steps");
    const uint32_t T = thetaSteps, P = 2 * T;
    const uint32_t base = (uint32_t) (verts.size() / 3);
    auto put = [&](double x, double y, double z) {
        verts.push_back((float) (center[0] + radius * x)); verts.push_back((float) (center[1] + radius * y)); verts.push_back((float) (center[2] + radius * z));
    };
    const double dTheta = M_PI / (T - 1), dPhi = 2 * M_PI / P;
    put(0, 0, 1);                                                  /* north pole: vertex 0 */
    for (uint32_t i = 1; i + 1 < T; i++) {                         /* rings 1 .. T-2: vertex 1 + (i - 1) P + j */
        const double st = std::sin(i * dTheta), ct = std::cos(i * dTheta);
        for (uint32_t j = 0; j < P; j++) put(st * std::cos(j * dPhi), st * std::sin(j * dPhi), ct);
    }
    put(0, 0, -1);                                                 /* south pole: the last vertex */
    const uint32_t south = base + 1 + (T - 2) * P;
    auto ring = [&](uint32_t i, uint32_t j) { return base + 1 + (i - 1) * P + (j % P); };
    auto tri = [&](uint32_t a, uint32_t b, uint32_t c) {           /* (a, b, c) counter-clockwise seen from outside */
        tris.push_back(a); tris.push_back(flipNormals ? c : b); tris.push_back(flipNormals ? b : c);
    };
    for (uint32_t j = 0; j < P; j++) tri(base, ring(1, j), ring(1, j + 1));
    for (uint32_t i = 1; i + 2 < T; i++)
        for (uint32_t j = 0; j < P; j++) {
            tri(ring(i, j), ring(i + 1, j), ring(i + 1, j + 1));
            tri(ring(i, j), ring(i + 1, j + 1), ring(i, j + 1));
        }
    for (uint32_t j = 0; j < P; j++) tri(south, ring(T - 2, j + 1), ring(T - 2, j));
}

}   // namespace alvrl
