// CPU experiment behind the 4-wide tree of the fast flavour (bvh.h::collapse4): node visits per shadow ray on the C4 occluder
// scene, binary skip-pointer tree vs the collapsed 4-wide tree, and the slowest ray of a warp of 32 semi-coherent rays
// (neighbouring origins, targets on one VRL).  Also checks that both trees give the same occlusion decisions.
//   python -c "import alvrl_loader,numpy as np; p=alvrl_loader.load(); v,t,m=p.scenes.occluder_mesh(); v.astype(np.float32).tofile('/tmp/verts.bin'); t.astype(np.uint32).tofile('/tmp/tris.bin')"
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include tools/micro/bvh_visits.cpp -o /tmp/bvh_visits && /tmp/bvh_visits /tmp/verts.bin /tmp/tris.bin
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#include <algorithm>
#include <cmath>
#include <cstring>
#include "../../mitsuba-alvrl_b200/csrc/bvh.h"
using namespace alvrl;
static uint32_t asU(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
struct Ray { float o[3], d[3], len; };
static HostBvh bvh; static std::vector<TriFast> tf; static std::vector<Bvh4Node> n4;
static bool triHit(const TriFast &t, const Ray &r) {
    float den = t.p.x * r.d[0] + t.p.y * r.d[1] + t.p.z * r.d[2], num = t.p.w - (t.p.x * r.o[0] + t.p.y * r.o[1] + t.p.z * r.o[2]);
    float tt = num / den; if (!(tt >= 0 && tt <= r.len)) return false;
    float P[3] = {r.o[0] + tt * r.d[0], r.o[1] + tt * r.d[1], r.o[2] + tt * r.d[2]};
    float u = t.q.x * P[0] + t.q.y * P[1] + t.q.z * P[2] + t.q.w, v = t.r.x * P[0] + t.r.y * P[1] + t.r.z * P[2] + t.r.w;
    return u >= 0 && v >= 0 && u + v <= 1;
}
static bool boxHit(const float *lo, const float *hi, const Ray &r, const float *inv) {
    float t0 = 0, t1 = r.len;
    for (int k = 0; k < 3; k++) { float a = (lo[k] - r.o[k]) * inv[k], b = (hi[k] - r.o[k]) * inv[k]; t0 = std::max(t0, std::min(a, b)); t1 = std::min(t1, std::max(a, b)); }
    return t0 <= t1;
}
static uint32_t travSkip(const Ray &r, bool &hit) {
    float inv[3] = {1 / r.d[0], 1 / r.d[1], 1 / r.d[2]}; uint32_t node = 0, nv = 0; hit = false; const uint32_t N = (uint32_t) bvh.nodes.size();
    while (node < N && !hit) { const BvhNode &nd = bvh.nodes[node]; nv++;
        float lo[3] = {nd.lo.x, nd.lo.y, nd.lo.z}, hi[3] = {nd.hi.x, nd.hi.y, nd.hi.z};
        if (boxHit(lo, hi, r, inv)) { uint32_t lf = asU(nd.hi.w); if (lf) { for (uint32_t i = 0; i < (lf & 15) && !hit; i++) hit = triHit(tf[(lf >> 4) + i], r); node = asU(nd.lo.w); } else node++; }
        else node = asU(nd.lo.w); }
    return nv;
}
static uint32_t trav4(const Ray &r, bool &hit, uint32_t &maxSp) {
    float inv[3] = {1 / r.d[0], 1 / r.d[1], 1 / r.d[2]}; int stack[256]; uint32_t sp = 0; uint32_t nv = 0; hit = false; const int DONE = (int) 0x80000000u; int cur = 0;
    while (cur != DONE && !hit) {
        if (cur >= 0) { const Bvh4Node &n = n4[cur]; nv++;
            const float *lx = &n.lox.x, *ly = &n.loy.x, *lz = &n.loz.x, *hx = &n.hix.x, *hy = &n.hiy.x, *hz = &n.hiz.x; const int *ch = &n.child.x;
            for (int k = 0; k < 4; k++) { float lo[3] = {lx[k], ly[k], lz[k]}, hi[3] = {hx[k], hy[k], hz[k]}; if (ch[k] != DONE && boxHit(lo, hi, r, inv)) stack[sp++] = ch[k]; }
            maxSp = std::max(maxSp, sp); cur = sp ? stack[--sp] : DONE;
        } else { uint32_t lf = (uint32_t) ~cur; for (uint32_t i = 0; i < (lf & 15) && !hit; i++) hit = triHit(tf[(lf >> 4) + i], r); cur = sp ? stack[--sp] : DONE; } }
    return nv;
}
int main(int argc, char **argv) {
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t nv = ftell(f) / 12; fseek(f, 0, SEEK_SET);
    std::vector<float> V(nv * 3); if (!fread(V.data(), 4, nv * 3, f)) return 1; fclose(f);
    f = fopen(argv[2], "rb"); fseek(f, 0, SEEK_END); size_t nt = ftell(f) / 12; fseek(f, 0, SEEK_SET);
    std::vector<uint32_t> T(nt * 3); if (!fread(T.data(), 4, nt * 3, f)) return 1; fclose(f);
    BvhBuilder b(V.data(), T.data(), (uint32_t) nt); b.build(bvh);
    tf.resize(nt); for (size_t i = 0; i < nt; i++) { uint32_t t = bvh.triOrder[i]; tf[i] = makeTriFast(&V[3 * T[3 * t]], &V[3 * T[3 * t + 1]], &V[3 * T[3 * t + 2]]); }
    const uint32_t depth = collapse4(bvh, n4);
    printf("triangles %zu, binary nodes %zu, 4-wide nodes %zu (depth %u)\n", nt, bvh.nodes.size(), n4.size(), depth);
    std::mt19937 rng(1); std::uniform_real_distribution<float> U(0.02f, 0.98f), J(-0.04f, 0.04f), U01(0, 1);
    const int NW = 6000; double sSkip = 0, mSkip = 0, s4 = 0, m4 = 0; int nocc = 0, mism = 0; uint32_t maxSp = 0;
    for (int w = 0; w < NW; w++) {
        float c[3] = {U(rng), U(rng), U(rng)}, s[3] = {U(rng), U(rng), U(rng)}, e[3] = {U(rng), U(rng), U(rng)}; uint32_t mx1 = 0, mx2 = 0;
        for (int l = 0; l < 32; l++) { Ray r; float tt = U01(rng);
            for (int k = 0; k < 3; k++) { r.o[k] = std::min(0.98f, std::max(0.02f, c[k] + J(rng))); r.d[k] = s[k] + tt * (e[k] - s[k]) - r.o[k]; }
            r.len = std::sqrt(r.d[0] * r.d[0] + r.d[1] * r.d[1] + r.d[2] * r.d[2]); for (int k = 0; k < 3; k++) r.d[k] /= r.len;
            bool h1, h2; uint32_t a = travSkip(r, h1), b4 = trav4(r, h2, maxSp); mism += h1 != h2;
            nocc += h1; sSkip += a; s4 += b4; mx1 = std::max(mx1, a); mx2 = std::max(mx2, b4); }
        mSkip += mx1; m4 += mx2; }
    const double n = NW * 32.0;
    printf("occluded %.3f, decisions that differ %d | binary skip-pointer tree: %.1f node visits per ray, %.1f for the slowest of 32 | 4-wide: %.1f, %.1f (deepest stack %u)\n",
           nocc / n, mism, sSkip / n, mSkip / NW, s4 / n, m4 / NW, maxSp);
    return mism != 0;
}
