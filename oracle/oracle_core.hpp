/*
 * oracle_core.hpp -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (IEEE fp32, no FMA contraction, no fast-math) of the reference's VRL hot path:
 * value types, SFMT sampler, triangle scene with TriAccel closest hit, perspective sensor,
 * homogeneous / heterogeneous media, phase functions, diffuse BSDF and the integrateVRL estimator.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this code; the product (libalvrl.so) never does.
 *
 * The reference itself cannot be compiled in this image (no Boost/Xerces/OpenEXR headers, see
 * DESIGN.md).  What the reference's own test suite pins of this path is reproduced: the SFMT stream
 * (known-answer vector, src/tests/test_random.cpp) and the chi-square test of the phase-function,
 * BSDF and emitter sampling routines (src/tests/test_chisquare.cpp; tests/test_chisquare_cpu.py).
 * Everything else -- integrateVRL, R, slices, clusters, images -- is "parity unpinned" by the
 * reference's own tests: this file *defines* the reference result, line by line from the cited
 * sources (paths relative to the reference tree).  Where no vector exists the restatement is
 * checked against independent computations instead (quadrature of the estimator's integrand and of
 * the grid medium's optical depth, tests/test_estimator_quadrature_cpu.py; the ground-truth path
 * tracer and a brute-force estimator, tests/test_volpath.py).
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <limits>
#include <algorithm>
#include <string>
#include <stdexcept>
#include "../include/alvrl.h"
#include "../include/alvrl_rng.h"

namespace orc {

typedef float Float;
static const Float Epsilon = 1e-4f;        // include/mitsuba/core/constants.h:32
static const Float ShadowEpsilon = 1e-3f;  // constants.h:33
static const Float INV_PI = 0.31830988618379067154f;
static const Float INV_FOURPI = 0.07957747154594766788f;

[[noreturn]] inline void fail(const std::string &msg) { throw std::runtime_error(msg); } // Log(EError) throws, logger.cpp:147

/* ---- include/mitsuba/core/vector.h, point.h -------------------------------------------- */
struct V3 {
    Float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(Float a) : x(a), y(a), z(a) {}
    V3(Float a, Float b, Float c) : x(a), y(b), z(c) {}
    Float operator[](int i) const { return (&x)[i]; }
    Float &operator[](int i) { return (&x)[i]; }
    V3 operator+(const V3 &o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3 &o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(Float f) const { return V3(x * f, y * f, z * f); }
    V3 operator/(Float f) const { Float r = (Float) 1 / f; return V3(x * r, y * r, z * r); } // vector.h:~560: recip multiply
    Float lengthSquared() const { return x * x + y * y + z * z; }
    Float length() const { return std::sqrt(lengthSquared()); }
    bool isFinite() const { return std::isfinite(x) && std::isfinite(y) && std::isfinite(z); }
};
inline V3 operator*(Float f, const V3 &v) { return V3(f * v.x, f * v.y, f * v.z); }
inline Float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3 &a, const V3 &b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline V3 normalize(const V3 &v) { return v / v.length(); }
inline Float distance(const V3 &a, const V3 &b) { return (a - b).length(); }
inline Float distanceSquared(const V3 &a, const V3 &b) { return (a - b).lengthSquared(); }
inline Float safe_sqrt(Float v) { return std::sqrt(std::max((Float) 0, v)); }          // math.h:260-267
inline double safe_sqrt(double v) { return std::sqrt(std::max(0.0, v)); }
inline Float fastexp(Float v) { return (Float) ::exp((double) v); }                        // math.h:185-187
/* asinh / sinh / atan / tan: the reference calls the libm overloads of whatever glibc the build host has (version
 * unpinned, results differ by an ulp between versions).  The oracle pins them to the correctly rounded value. */
inline Float o_asinh(Float v) { return (Float) ::asinh((double) v); }
inline Float o_sinh(Float v) { return (Float) ::sinh((double) v); }
inline Float o_atan(Float v) { return (Float) ::atan((double) v); }
inline Float o_tan(Float v) { return (Float) ::tan((double) v); }

/* ---- include/mitsuba/core/spectrum.h (RGB, SPECTRUM_SAMPLES = 3) ------------------------ */
struct Spec {
    Float s[3];
    Spec() { s[0] = s[1] = s[2] = 0; }
    Spec(Float v) { s[0] = s[1] = s[2] = v; }
    Spec(Float r, Float g, Float b) { s[0] = r; s[1] = g; s[2] = b; }
    Float operator[](int i) const { return s[i]; }
    Float &operator[](int i) { return s[i]; }
    Spec operator*(const Spec &o) const { return Spec(s[0] * o.s[0], s[1] * o.s[1], s[2] * o.s[2]); }
    Spec operator*(Float f) const { return Spec(s[0] * f, s[1] * f, s[2] * f); }
    Spec operator+(const Spec &o) const { return Spec(s[0] + o.s[0], s[1] + o.s[1], s[2] + o.s[2]); }
    Spec operator/(Float f) const { Float r = 1.0f / f; return Spec(s[0] * r, s[1] * r, s[2] * r); } // spectrum.h:415-425
    Spec &operator*=(const Spec &o) { for (int i = 0; i < 3; i++) s[i] *= o.s[i]; return *this; }
    Spec &operator*=(Float f) { for (int i = 0; i < 3; i++) s[i] *= f; return *this; }
    Spec &operator/=(Float f) { Float r = 1.0f / f; for (int i = 0; i < 3; i++) s[i] *= r; return *this; }
    Spec &operator+=(const Spec &o) { for (int i = 0; i < 3; i++) s[i] += o.s[i]; return *this; }
    bool isZero() const { return s[0] == 0 && s[1] == 0 && s[2] == 0; }
    bool isValid() const { for (int i = 0; i < 3; i++) if (!std::isfinite(s[i]) || s[i] < 0.0f) return false; return true; } // spectrum.h:467-472
    Float max() const { return std::max(std::max(s[0], s[1]), s[2]); }
    Float getLuminance() const { return s[0] * 0.212671f + s[1] * 0.715160f + s[2] * 0.072169f; } // spectrum.h:725-727
};
inline Spec operator*(Float f, const Spec &a) { return a * f; }

/* ---- SFMT-19937, src/libcore/random.cpp ----------------------------------------------- */
class Random {
public:
    static const int N = 19937 / 128 + 1, N32 = N * 4, N64 = N * 2;       // random.cpp:72-82
    explicit Random(uint64_t seed) { init_gen_rand(seed); }
    explicit Random(Random *parent) { seedFrom(parent); }
    void seedFrom(Random *parent) {                                         // random.cpp:528-533
        uint64_t buf[N64];
        for (int i = 0; i < N64; ++i) buf[i] = parent->nextULong();
        init_by_array(reinterpret_cast<uint32_t *>(buf), N64 * 2);          // random.cpp:540-549
    }
    uint64_t nextULong() {                                                  // random.cpp:288-297
        if (idx >= N32) { gen_rand_all(); idx = 0; }
        uint64_t r = st64()[idx / 2];
        idx += 2;
        return r;
    }
    Float nextFloat() {                                                     // random.cpp:630-639
        uint32_t u = (uint32_t) ((nextULong() & 0xFFFFFFFFull) >> 9) | 0x3f800000u;
        float f; memcpy(&f, &u, 4);
        return f - 1.0f;
    }
private:
    uint32_t st[N32];
    int idx;
    uint64_t *st64() { return reinterpret_cast<uint64_t *>(st); }
    static void lshift128(uint32_t *out, const uint32_t *in, int shift) {  // random.cpp:160-172
        uint64_t th = ((uint64_t) in[3] << 32) | in[2], tl = ((uint64_t) in[1] << 32) | in[0];
        uint64_t oh = th << (shift * 8), ol = tl << (shift * 8);
        oh |= tl >> (64 - shift * 8);
        out[0] = (uint32_t) ol; out[1] = (uint32_t) (ol >> 32); out[2] = (uint32_t) oh; out[3] = (uint32_t) (oh >> 32);
    }
    static void rshift128(uint32_t *out, const uint32_t *in, int shift) {  // random.cpp:138-150
        uint64_t th = ((uint64_t) in[3] << 32) | in[2], tl = ((uint64_t) in[1] << 32) | in[0];
        uint64_t oh = th >> (shift * 8), ol = tl >> (shift * 8);
        ol |= th << (64 - shift * 8);
        out[0] = (uint32_t) ol; out[1] = (uint32_t) (ol >> 32); out[2] = (uint32_t) oh; out[3] = (uint32_t) (oh >> 32);
    }
    static void do_recursion(uint32_t *r, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d) {
        static const uint32_t MSK[4] = {0xdfffffefU, 0xddfecb7fU, 0xbffaffffU, 0xbffffff6U};   // random.cpp:89-92
        uint32_t x[4], y[4];
        lshift128(x, a, 1);      // SL2
        rshift128(y, c, 1);      // SR2
        for (int i = 0; i < 4; i++)                                          // random.cpp:204-219
            r[i] = a[i] ^ x[i] ^ ((b[i] >> 11) & MSK[i]) ^ y[i] ^ (d[i] << 18);
    }
    void gen_rand_all() {                                                   // random.cpp:376-390
        const int POS1 = 122;
        uint32_t *r1 = &st[(N - 2) * 4], *r2 = &st[(N - 1) * 4];
        int i;
        for (i = 0; i < N - POS1; ++i) {
            uint32_t out[4];
            do_recursion(out, &st[i * 4], &st[(i + POS1) * 4], r1, r2);
            memcpy(&st[i * 4], out, 16);
            r1 = r2; r2 = &st[i * 4];
        }
        for (; i < N; ++i) {
            uint32_t out[4];
            do_recursion(out, &st[i * 4], &st[(i + POS1 - N) * 4], r1, r2);
            memcpy(&st[i * 4], out, 16);
            r1 = r2; r2 = &st[i * 4];
        }
    }
    void period_certification() {                                           // random.cpp:318-345
        static const uint32_t parity[4] = {0x00000001U, 0, 0, 0x13c9e684U};
        int inner = 0;
        for (int i = 0; i < 4; ++i) inner ^= st[i] & parity[i];
        for (int i = 16; i > 0; i >>= 1) inner ^= inner >> i;
        inner &= 1;
        if (inner == 1) return;
        for (int i = 0; i < 4; ++i) {
            uint32_t work = 1;
            for (int j = 0; j < 32; ++j) {
                if ((work & parity[i]) != 0) { st[i] ^= work; return; }
                work = work << 1;
            }
        }
    }
    void init_gen_rand(uint64_t seed) {                                     // random.cpp:397-406 (Mitsuba-specific)
        uint64_t *p = st64();
        p[0] = seed;
        for (int i = 1; i < N64; ++i)
            p[i] = 6364136223846793005ULL * (p[i - 1] ^ (p[i - 1] >> 62)) + i;
        idx = N32;
        period_certification();
    }
    static uint32_t func1(uint32_t x) { return (x ^ (x >> 27)) * (uint32_t) 1664525UL; }
    static uint32_t func2(uint32_t x) { return (x ^ (x >> 27)) * (uint32_t) 1566083941UL; }
    void init_by_array(const uint32_t *init_key, int key_length) {          // random.cpp:408-469
        int i, j, count, lag = 11, size = N32;
        int mid = (size - lag) / 2;
        uint32_t r;
        memset(st, 0x8b, sizeof(st));
        count = (key_length + 1 > N32) ? key_length + 1 : N32;
        r = func1(st[0] ^ st[mid] ^ st[N32 - 1]);
        st[mid] += r; r += key_length; st[mid + lag] += r; st[0] = r;
        count--;
        for (i = 1, j = 0; (j < count) && (j < key_length); j++) {
            r = func1(st[i] ^ st[(i + mid) % N32] ^ st[(i + N32 - 1) % N32]);
            st[(i + mid) % N32] += r; r += init_key[j] + i;
            st[(i + mid + lag) % N32] += r; st[i] = r; i = (i + 1) % N32;
        }
        for (; j < count; j++) {
            r = func1(st[i] ^ st[(i + mid) % N32] ^ st[(i + N32 - 1) % N32]);
            st[(i + mid) % N32] += r; r += i;
            st[(i + mid + lag) % N32] += r; st[i] = r; i = (i + 1) % N32;
        }
        for (j = 0; j < N32; j++) {
            r = func2(st[i] + st[(i + mid) % N32] + st[(i + N32 - 1) % N32]);
            st[(i + mid) % N32] ^= r; r -= i;
            st[(i + mid + lag) % N32] ^= r; st[i] = r; i = (i + 1) % N32;
        }
        idx = N32;
        period_certification();
    }
};

/* ---- Sampler: src/samplers/independent.cpp:71-103 + the counter stream of alvrl_rng.h ---- */
struct Sampler {
    virtual ~Sampler() {}
    virtual Float next1D() = 0;
    virtual Sampler *clone() = 0;
    /* (domain, a, b) addressing for the counter stream; no-op for SFMT */
    virtual void setContext(uint32_t, uint32_t, uint32_t) {}
    /* counter stream: the draws of Clustering::split come from the sub-stream of the cluster [begin, end) (alvrl_rng.h);
     * a sequential stream (SFMT, tapes) just keeps drawing */
    virtual void enterNode(uint32_t, uint32_t) {}
    virtual void leaveNode() {}
    /* the draws of a cluster's split do not depend on what was drawn before (counter stream) */
    virtual bool replayable() const { return false; }
    uint64_t draws = 0;
};
struct SfmtSampler : Sampler {
    Random rnd;
    explicit SfmtSampler(uint64_t seed) : rnd(seed) {}
    explicit SfmtSampler(Random *parent) : rnd(parent) {}
    Float next1D() override { draws++; return rnd.nextFloat(); }
    Sampler *clone() override { return new SfmtSampler(&rnd); }          // independent.cpp:71-80
};
struct CounterSampler : Sampler {
    uint64_t seed; uint32_t key = 0, k = 0, outerKey = 0, outerK = 0; bool inNode = false;
    explicit CounterSampler(uint64_t s) : seed(s) {}
    void setContext(uint32_t domain, uint32_t a, uint32_t b) override { key = alvrl_rng_key(seed, domain, a, b); k = 0; inNode = false; }
    void enterNode(uint32_t begin, uint32_t end) override { if (inNode) fail("CounterSampler::enterNode nested"); outerKey = key; outerK = k; key = alvrl_rng_node_key(key, begin, end); k = 0; inNode = true; }
    void leaveNode() override { if (inNode) { key = outerKey; k = outerK; inNode = false; } }
    Float next1D() override { draws++; return alvrl_rng_uniform(key, k++); }
    bool replayable() const override { return true; }
    Sampler *clone() override { return new CounterSampler(seed); }
};
/* Replays a tape laid out like alvrl_set_sample_tape: K slots per (row, vrl) */
struct TapeSampler : Sampler {
    const float *tape; uint64_t n; uint32_t K, N; uint64_t base = 0; uint32_t k = 0;
    TapeSampler(const float *t, uint64_t n_, uint32_t K_, uint32_t N_) : tape(t), n(n_), K(K_), N(N_) {}
    void setContext(uint32_t, uint32_t a, uint32_t b) override { base = ((uint64_t) a * N + b) * K; k = 0; }
    Float next1D() override { draws++; uint64_t i = base + (k++); if (i >= n) fail("sample tape exhausted"); return tape[i]; }
    Sampler *clone() override { return new TapeSampler(tape, n, K, N); }
};
/* Records what an inner sampler produced into the fixed-slot tape layout */
struct RecordingSampler : Sampler {
    Sampler *inner; std::vector<float> &tape; uint32_t K, N; uint64_t base = 0; uint32_t k = 0;
    RecordingSampler(Sampler *in, std::vector<float> &t, uint32_t K_, uint32_t N_) : inner(in), tape(t), K(K_), N(N_) {}
    void setContext(uint32_t d, uint32_t a, uint32_t b) override { inner->setContext(d, a, b); base = ((uint64_t) a * N + b) * K; k = 0; }
    Float next1D() override { draws++; Float u = inner->next1D(); if (k < K) tape[base + k] = u; k++; return u; }
    void enterNode(uint32_t b, uint32_t e) override { inner->enterNode(b, e); }
    void leaveNode() override { inner->leaveNode(); }
    Sampler *clone() override { fail("RecordingSampler::clone"); }
};

/* ---- geometry: triaccel.h, aabb.h, skdtree.cpp ------------------------------------------ */
struct Ray {
    V3 o, d, dRcp; Float mint, maxt;
    Ray() : mint(Epsilon), maxt(std::numeric_limits<Float>::infinity()) {}
    Ray(const V3 &o_, const V3 &d_, Float mint_, Float maxt_) : o(o_), d(d_), mint(mint_), maxt(maxt_) {
        dRcp = V3((Float) 1 / d.x, (Float) 1 / d.y, (Float) 1 / d.z);      // ray.h setDirection
    }
    V3 operator()(Float t) const { return o + t * d; }
};

struct AABB {
    V3 min, max;
    AABB() : min(std::numeric_limits<Float>::infinity()), max(-std::numeric_limits<Float>::infinity()) {}
    void expandBy(const V3 &p) {
        for (int i = 0; i < 3; i++) { min[i] = std::min(min[i], p[i]); max[i] = std::max(max[i], p[i]); }
    }
    /* aabb.h:308-338 */
    bool rayIntersect(const Ray &ray, Float &nearT, Float &farT) const {
        nearT = -std::numeric_limits<Float>::infinity();
        farT = std::numeric_limits<Float>::infinity();
        for (int i = 0; i < 3; i++) {
            const Float origin = ray.o[i], minVal = min[i], maxVal = max[i];
            if (ray.d[i] == 0) {
                if (origin < minVal || origin > maxVal) return false;
            } else {
                Float t1 = (minVal - origin) * ray.dRcp[i];
                Float t2 = (maxVal - origin) * ray.dRcp[i];
                if (t1 > t2) std::swap(t1, t2);
                nearT = std::max(t1, nearT);
                farT = std::min(t2, farT);
                if (!(nearT <= farT)) return false;
            }
        }
        return true;
    }
};

struct TriAccel {                                                          // triaccel.h:37-59
    uint32_t k; Float n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv;
    int load(const V3 &A, const V3 &B, const V3 &C) {                      // triaccel.h:61-95
        static const int waldModulo[4] = {1, 2, 0, 1};
        V3 b = C - A, c = B - A, N = cross(c, b);
        k = 0;
        for (int j = 0; j < 3; j++) if (std::abs(N[j]) > std::abs(N[k])) k = j;
        uint32_t u = waldModulo[k], v = waldModulo[k + 1];
        const Float n_k = N[k], denom = b[u] * c[v] - b[v] * c[u];
        if (denom == 0) { k = 3; return 1; }
        n_u = N[u] / n_k; n_v = N[v] / n_k; n_d = dot(A, N) / n_k;
        b_nu = b[u] / denom; b_nv = -b[v] / denom;
        a_u = A[u]; a_v = A[v];
        c_nu = c[v] / denom; c_nv = -c[u] / denom;
        return 0;
    }
    /* test instrumentation, not part of the reference: plane parameter and barycentrics without the accept tests */
    bool planeHit(const Ray &ray, Float &u, Float &v, Float &t) const {
        Float o_u, o_v, o_k, d_u, d_v, d_k;
        switch (k) {
            case 0: o_u = ray.o[1]; o_v = ray.o[2]; o_k = ray.o[0]; d_u = ray.d[1]; d_v = ray.d[2]; d_k = ray.d[0]; break;
            case 1: o_u = ray.o[2]; o_v = ray.o[0]; o_k = ray.o[1]; d_u = ray.d[2]; d_v = ray.d[0]; d_k = ray.d[1]; break;
            case 2: o_u = ray.o[0]; o_v = ray.o[1]; o_k = ray.o[2]; d_u = ray.d[0]; d_v = ray.d[1]; d_k = ray.d[2]; break;
            default: return false;
        }
        t = (n_d - o_u * n_u - o_v * n_v - o_k) / (d_u * n_u + d_v * n_v + d_k);
        if (!std::isfinite(t)) return false;
        const Float hu = o_u + t * d_u - a_u, hv = o_v + t * d_v - a_v;
        u = hv * b_nu + hu * b_nv;
        v = hu * c_nu + hv * c_nv;
        return true;
    }
    bool rayIntersect(const Ray &ray, Float mint, Float maxt, Float &u, Float &v, Float &t) const { // triaccel.h:97-158
        Float o_u, o_v, o_k, d_u, d_v, d_k;
        switch (k) {
            case 0: o_u = ray.o[1]; o_v = ray.o[2]; o_k = ray.o[0]; d_u = ray.d[1]; d_v = ray.d[2]; d_k = ray.d[0]; break;
            case 1: o_u = ray.o[2]; o_v = ray.o[0]; o_k = ray.o[1]; d_u = ray.d[2]; d_v = ray.d[0]; d_k = ray.d[1]; break;
            case 2: o_u = ray.o[0]; o_v = ray.o[1]; o_k = ray.o[2]; d_u = ray.d[0]; d_v = ray.d[1]; d_k = ray.d[2]; break;
            default: return false;
        }
        t = (n_d - o_u * n_u - o_v * n_v - o_k) / (d_u * n_u + d_v * n_v + d_k);
        if (t < mint || t > maxt) return false;
        const Float hu = o_u + t * d_u - a_u;
        const Float hv = o_v + t * d_v - a_v;
        u = hv * b_nu + hu * b_nv;
        v = hu * c_nu + hv * c_nv;
        return u >= 0 && v >= 0 && u + v <= 1.0f;
    }
};

struct Intersection {
    Float t = std::numeric_limits<Float>::infinity();
    V3 p, n;              // barycentric position; shading normal (= face normal: no vertex normals)
    uint32_t prim = ALVRL_NO_HIT, material = 0;
    Float wiz = 0;        // Frame::cosTheta(its.wi) = dot(-ray.d, shFrame.n), skdtree.h:427
    bool tie = false;     // >= 2 triangles at the minimal t (quirk B13)
    bool isValid() const { return t != std::numeric_limits<Float>::infinity(); }
};

/* per-material optics (include/alvrl.h::alvrl_set_material_optics): dielectric eta = intIOR / extIOR in [0]; conductor
 * eta rgb in [0..2], k rgb in [3..5]; specularReflectance [6..8], specularTransmittance [9..11] */
struct Optics { Float v[12]; };

struct Scene {
    std::vector<V3> verts;
    std::vector<uint32_t> tris, triMat;
    std::vector<Spec> albedo;
    std::vector<uint32_t> matBits;
    std::vector<Optics> optics;
    std::vector<TriAccel> accel;
    std::vector<V3> extraBounds;
    AABB kdAABB, sceneAABB;

    size_t numTris() const { return triMat.size(); }
    void finalize() {
        accel.resize(numTris());
        AABB box;
        for (size_t i = 0; i < numTris(); i++) {
            const V3 &A = verts[tris[3 * i]], &B = verts[tris[3 * i + 1]], &C = verts[tris[3 * i + 2]];
            accel[i].load(A, B, C);
            box.expandBy(A); box.expandBy(B); box.expandBy(C);
        }
        /* gkdtree.h:1213-1220: enlarge (max uses the already enlarged min) */
        const Float eps = 1e-3f;
        box.min = box.min - ((box.max - box.min) * eps + V3(eps));
        box.max = box.max + ((box.max - box.min) * eps + V3(eps));
        kdAABB = box;
        sceneAABB = box;                                                    // scene.cpp:387-413
        for (const V3 &p : extraBounds) sceneAABB.expandBy(p);
    }
    /* ShapeKDTree::rayIntersect core (skdtree.cpp:112-204) over a brute-force loop instead of the
     * kd-tree.  shadowOverload selects the adaptive-epsilon variant without the Epsilon floor
     * (skdtree.cpp:154-157).  Tie rule (quirk B13): lowest triangle index among the minimal t. */
    bool closestHit(const Ray &ray, bool shadowOverload, Float &tOut, uint32_t &prim, Float &uOut, Float &vOut, bool *tie = nullptr) {
        Float mint, maxt;
        tOut = std::numeric_limits<Float>::infinity();
        prim = ALVRL_NO_HIT;
        if (tie) *tie = false;
        if (!kdAABB.rayIntersect(ray, mint, maxt)) return false;
        Float rayMinT = ray.mint;
        if (rayMinT == Epsilon) {
            Float m = std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
            if (!shadowOverload) m = std::max(m, Epsilon);
            rayMinT *= m;
        }
        if (rayMinT > mint) mint = rayMinT;
        if (ray.maxt < maxt) maxt = ray.maxt;
        if (!(maxt > mint)) return false;
        bool found = false;
        for (size_t i = 0; i < accel.size(); i++) {
            Float u, v, t;
            if (accel[i].rayIntersect(ray, mint, maxt, u, v, t)) {
                if (!found || t < tOut) { tOut = t; prim = (uint32_t) i; uOut = u; vOut = v; found = true; if (tie) *tie = false; }
                else if (t == tOut && tie) *tie = true;
            }
        }
        return found;
    }
    /* Test instrumentation, not part of the reference: is the occlusion decision of this shadow ray within `tol` of
     * flipping (the "documented grazing-hit ties" of north_star)?  A triangle is hit LOOSELY when the accept tests pass
     * with every bound relaxed by tol (barycentric units; tol x segment length on t) and ROBUSTLY when they pass with
     * every bound tightened; the decision is fragile when some triangle is hit loosely and none robustly. */
    bool shadowDecisionFragile(const Ray &ray, Float tol) const {
        Float mint = ray.mint, maxt = ray.maxt;
        if (mint == Epsilon) mint *= std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
        const Float dt = tol * std::max(maxt, (Float) 1e-3f);
        const Float dlo = mint > 0 ? std::min(dt, 0.25f * mint) : dt;        /* a surface origin sits at t = 0 << mint by design */
        bool loose = false, robust = false;
        for (size_t i = 0; i < accel.size(); i++) {
            Float u, v, t;
            if (!accel[i].planeHit(ray, u, v, t)) continue;
            if (t >= mint - dlo && t <= maxt + dt && u >= -tol && v >= -tol && u + v <= 1 + tol) loose = true;
            if (t >= mint + dlo && t <= maxt - dt && u >= tol && v >= tol && u + v <= 1 - tol) robust = true;
        }
        return loose && !robust;
    }
    /* ShapeKDTree::rayIntersect(ray, its) + fillIntersectionRecord<true> (skdtree.cpp:112-142, skdtree.h:343-428) */
    /* shading frame of a triangle hit: n = face normal, s = normalize(dpdu - n dot(n, dpdu)) with dpdu = p1 - p0, t = cross(n, s)
     * (skdtree.h:367-378,395-396,426; util.cpp:603-608) */
    struct HitFrame hitFrame(const Intersection &its) const;
    bool rayIntersect(const Ray &ray, Intersection &its) {
        Float u = 0, v = 0;
        its = Intersection();
        if (!closestHit(ray, false, its.t, its.prim, u, v, &its.tie)) { its.t = std::numeric_limits<Float>::infinity(); return false; }
        const V3 &p0 = verts[tris[3 * its.prim]], &p1 = verts[tris[3 * its.prim + 1]], &p2 = verts[tris[3 * its.prim + 2]];
        const V3 b(1 - u - v, u, v);
        its.p = p0 * b.x + p1 * b.y + p2 * b.z;
        V3 side1 = p1 - p0, side2 = p2 - p0;
        V3 faceNormal = cross(side1, side2);
        Float length = faceNormal.length();
        if (!(faceNormal.x == 0 && faceNormal.y == 0 && faceNormal.z == 0)) faceNormal = faceNormal / length;
        its.n = faceNormal;
        its.material = triMat[its.prim];
        its.wiz = dot(-ray.d, its.n);
        return true;
    }
};

/* ---- delta BSDFs of the specular chains (vrlIntegrator.cpp:445-511) --------------------------- */
/* fresnelDielectricExt, src/libcore/util.cpp:651-681 */
inline Float fresnelDielectricExt(Float cosThetaI_, Float &cosThetaT_, Float eta) {
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    Float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    Float cosThetaI = std::abs(cosThetaI_);
    Float cosThetaT = std::sqrt(cosThetaTSqr);
    Float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    Float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
/* fresnelConductorExact, util.cpp:739-761, one channel */
inline Float fresnelConductorExact(Float cosThetaI, Float eta, Float k) {
    Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    Float temp1 = eta * eta - k * k - sinThetaI2, a2pb2 = safe_sqrt(temp1 * temp1 + k * k * eta * eta * 4), a = safe_sqrt((a2pb2 + temp1) * 0.5f);
    Float term1 = a2pb2 + cosThetaI2, term2 = a * (2 * cosThetaI);
    Float Rs2 = (term1 - term2) / (term1 + term2);
    Float term3 = a2pb2 * cosThetaI2 + sinThetaI4, term4 = term2 * sinThetaI2;
    Float Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}
/* Frame of a triangle hit: shFrame.n = face normal, s and t from computeShadingFrame(n, dpdu = p1 - p0) (skdtree.h:367-426,
 * util.cpp:603-608) */
struct HitFrame {
    V3 s, t, n;
    V3 toLocal(const V3 &v) const { return V3(dot(v, s), dot(v, t), dot(v, n)); }
    V3 toWorld(const V3 &v) const { return s * v.x + t * v.y + n * v.z; }
};
inline HitFrame Scene::hitFrame(const Intersection &its) const {
    const V3 &p0 = verts[tris[3 * its.prim]], &p1 = verts[tris[3 * its.prim + 1]];
    HitFrame f; f.n = its.n;
    const V3 dpdu = p1 - p0;
    f.s = normalize(dpdu - f.n * dot(f.n, dpdu));
    f.t = cross(f.n, f.s);
    return f;
}
/* BSDF::sample(bRec, Point2(0.5)) with bRec.component = comp, mode = ERadiance, of the smooth dielectric (dielectric.cpp:335-387,
 * 218-226) and the smooth conductor (conductor.cpp:254-268).  Returns the weight, fills wo (local) and eta. */
inline Spec sampleDelta(uint32_t bits, const Optics &o, const V3 &wi, int comp, V3 &wo, Float &etaOut) {
    const Spec specR(o.v[6], o.v[7], o.v[8]), specT(o.v[9], o.v[10], o.v[11]);
    if (bits & ALVRL_BSDF_DIELECTRIC) {
        const Float eta = o.v[0], invEta = 1 / eta;
        Float cosThetaT;
        Float F = fresnelDielectricExt(wi.z, cosThetaT, eta);
        if (comp == 0) { wo = V3(-wi.x, -wi.y, wi.z); etaOut = 1.0f; return specR * F; }
        Float scale = -(cosThetaT < 0 ? invEta : eta);
        wo = V3(scale * wi.x, scale * wi.y, cosThetaT);
        etaOut = cosThetaT < 0 ? eta : invEta;
        Float factor = cosThetaT < 0 ? invEta : eta;                        // ERadiance
        return specT * (factor * factor * (1 - F));
    }
    /* conductor: one component */
    if (comp != 0 || wi.z <= 0) return Spec(0.0f);
    wo = V3(-wi.x, -wi.y, wi.z); etaOut = 1.0f;
    return specR * Spec(fresnelConductorExact(wi.z, o.v[0], o.v[3]), fresnelConductorExact(wi.z, o.v[1], o.v[4]), fresnelConductorExact(wi.z, o.v[2], o.v[5]));
}

/* ---- sensor: src/sensors/perspective.cpp:247-269 ---------------------------------------- */
struct Camera {
    Float s2c[16], c2w[16];
    uint32_t W = 0, H = 0; Float nearClip = 0, farClip = 0;
    V3 invRes() const { return V3(1.0f / W, 1.0f / H, 0); }
    static V3 xformPoint(const Float *m, const V3 &p) {                   // transform.h:108-125
        Float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
        Float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
        Float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
        Float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
        if (w == 1.0f) return V3(x, y, z);
        return V3(x, y, z) / w;
    }
    static V3 xformAffine(const Float *m, const V3 &p) {                  // transform.h:128-136
        return V3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                  m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
    }
    static V3 xformVector(const Float *m, const V3 &v) {                  // transform.h:175-183
        return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
                  m[8] * v.x + m[9] * v.y + m[10] * v.z);
    }
    Ray sampleRay(Float px, Float py) const {
        V3 ir = invRes();
        V3 nearP = xformPoint(s2c, V3(px * ir.x, py * ir.y, 0.0f));
        V3 d = normalize(nearP);
        Float invZ = 1.0f / d.z;
        return Ray(xformAffine(c2w, V3(0.0f)), xformVector(c2w, d), nearClip * invZ, farClip * invZ);
    }
    V3 position() const { return xformAffine(c2w, V3(0.0f)); }
};

/* ---- media ---------------------------------------------------------------------------------- */
struct MediumSamplingRecord { Spec transmittance, sigmaS; Float pdfFailure = 0; };

struct Medium {
    int type = 0;                 // 0 homogeneous, 1 heterogeneous grid (simpson)
    Spec sigmaA, sigmaS, sigmaT;  // Medium base class (medium.cpp:27-37); for type 1: sigmaS = sigmaS_base (quirk B2)
    Float samplingWeight = 0;
    int phaseType = ALVRL_PHASE_ISOTROPIC; Float g = 0;
    /* heterogeneous */
    std::vector<float> density; int res[3] = {0, 0, 0}; V3 bmin, bmax; Float scale = 1, stepSize = 0; Spec hetAlbedo;

    void setHomogeneous(const Float a[3], const Float s[3], Float w) {    // homogeneous.cpp:156-184
        type = 0;
        for (int i = 0; i < 3; i++) { sigmaA[i] = a[i]; sigmaS[i] = s[i]; sigmaT[i] = a[i] + s[i]; }
        samplingWeight = w;
        if (samplingWeight == -1 || samplingWeight < 0) {
            samplingWeight = -1;
            for (int i = 0; i < 3; ++i) {
                Float alb = sigmaS[i] / sigmaT[i];
                if (alb > samplingWeight && sigmaT[i] != 0) samplingWeight = alb;
            }
            if (samplingWeight > 0) samplingWeight = std::max(samplingWeight, (Float) 0.5f);
        }
    }
    void setGrid(const float *dens, const int r[3], const V3 &mn, const V3 &mx, Float sc, const Float alb[3], const Float sBase[3]) {
        type = 1;
        density.assign(dens, dens + (size_t) r[0] * r[1] * r[2]);
        for (int i = 0; i < 3; i++) { res[i] = r[i]; hetAlbedo[i] = alb[i]; sigmaS[i] = sBase[i]; }
        bmin = mn; bmax = mx; scale = sc;
        /* gridvolume.cpp:188-215: stepSize = min_i 0.5 * extent_i / (res_i - 1) */
        stepSize = std::numeric_limits<Float>::infinity();
        V3 extents = bmax - bmin;
        for (int i = 0; i < 3; ++i) stepSize = std::min(stepSize, 0.5f * extents[i] / (Float) (res[i] - 1));
    }
    /* GridDataSource::lookupFloat, gridvolume.cpp:337-388, float32 data, identity toWorld.
     * worldToGrid = scale((res-1)/extents) * translate(-min)  (gridvolume.cpp:188-196), applied with
     * transformAffine (transform.h:128-136): g = s*p + (s * -min). */
    Float lookupDensity(const V3 &_p) {
        V3 extents = bmax - bmin;
        V3 sc((Float) (res[0] - 1) / extents.x, (Float) (res[1] - 1) / extents.y, (Float) (res[2] - 1) / extents.z);
        V3 tr(sc.x * (-bmin.x), sc.y * (-bmin.y), sc.z * (-bmin.z));
        V3 p(sc.x * _p.x + tr.x, sc.y * _p.y + tr.y, sc.z * _p.z + tr.z);
        const int x1 = (int) std::floor(p.x), y1 = (int) std::floor(p.y), z1 = (int) std::floor(p.z);
        const int x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
        if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= res[0] || y2 >= res[1] || z2 >= res[2]) return 0;
        const Float fx = p.x - x1, fy = p.y - y1, fz = p.z - z1, _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
        const float *d = density.data();
        const size_t rx = res[0], ry = res[1];
        const Float d000 = d[(z1 * ry + y1) * rx + x1], d001 = d[(z1 * ry + y1) * rx + x2],
                    d010 = d[(z1 * ry + y2) * rx + x1], d011 = d[(z1 * ry + y2) * rx + x2],
                    d100 = d[(z2 * ry + y1) * rx + x1], d101 = d[(z2 * ry + y1) * rx + x2],
                    d110 = d[(z2 * ry + y2) * rx + x1], d111 = d[(z2 * ry + y2) * rx + x2];
        return ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz
             + ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
    }
    /* HeterogeneousMedium::integrateDensity, heterogeneous.cpp:301-376 (composite Simpson; the
     * returned optical depth already contains m_scale) */
    Float integrateDensity(const Ray &ray) {
        AABB box; box.min = bmin; box.max = bmax;
        Float mint, maxt;
        if (!box.rayIntersect(ray, mint, maxt)) return 0.0f;
        mint = std::max(mint, ray.mint);
        maxt = std::min(maxt, ray.maxt);
        Float length = maxt - mint, maxComp = 0;
        V3 p = ray(mint), pLast = ray(maxt);
        for (int i = 0; i < 3; ++i)
            maxComp = std::max(std::max(maxComp, std::abs(p[i])), std::abs(pLast[i]));
        if (length < 1e-6f * maxComp) return 0.0f;
        uint32_t nSteps = (uint32_t) std::ceil(length / stepSize);
        nSteps += nSteps % 2;
        const Float stepSz = length / nSteps;
        const V3 increment = ray.d * stepSz;
        Float integratedDensity = lookupDensity(p) + lookupDensity(pLast);
        /* HETVOL_EARLY_EXIT is defined (heterogeneous.cpp:31, 336-340, 353-360): once the running Simpson sum passes
         * -log(Epsilon) of optical depth the march stops and the optical depth is +infinity (transmittance exactly 0) */
        const Float stopAfterDensity = -(Float) ::log((double) Epsilon);      // math::fastlog, math.h:193-195
        const Float stopValue = stopAfterDensity * 3.0f / (stepSz * scale);
        p = p + increment;
        Float m = 4;
        for (uint32_t i = 1; i < nSteps; ++i) {
            integratedDensity += m * lookupDensity(p);
            m = 6 - m;
            if (integratedDensity > stopValue) return std::numeric_limits<Float>::infinity();
            V3 next = p + increment;
            if (p.x == next.x && p.y == next.y && p.z == next.z) break;
            p = next;
        }
        return integratedDensity * scale * stepSz * (1.0f / 3.0f);
    }
    /* HeterogeneousMedium::invertDensityIntegral, heterogeneous.cpp:422-545: composite Simpson march until the optical depth
     * reaches desiredDensity, then Newton-bisection on the quadratic through the last three lookups */
    bool invertDensityIntegral(const Ray &ray, Float desiredDensity, Float &integratedDensity, Float &t, Float &densityAtMinT, Float &densityAtT) {
        integratedDensity = densityAtMinT = densityAtT = 0.0f;
        AABB box; box.min = bmin; box.max = bmax;
        Float mint, maxt;
        if (!box.rayIntersect(ray, mint, maxt)) return false;
        mint = std::max(mint, ray.mint);
        maxt = std::min(maxt, ray.maxt);
        Float length = maxt - mint, maxComp = 0;
        V3 p = ray(mint), pLast = ray(maxt);
        for (int i = 0; i < 3; ++i) maxComp = std::max(std::max(maxComp, std::abs(p[i])), std::abs(pLast[i]));
        if (length < 1e-6f * maxComp) return false;
        uint32_t nSteps = (uint32_t) std::ceil(length / (2 * stepSize));
        Float stepSz = length / nSteps, multiplier = (1.0f / 6.0f) * stepSz * scale;
        V3 fullStep = ray.d * stepSz, halfStep = fullStep * .5f;
        Float node1 = lookupDensity(p);
        if (ray.mint == mint) densityAtMinT = node1 * scale;
        else densityAtMinT = 0.0f;
        for (uint32_t i = 0; i < nSteps; ++i) {
            Float node2 = lookupDensity(p + halfStep), node3 = lookupDensity(p + fullStep),
                  newDensity = integratedDensity + multiplier * (node1 + node2 * 4 + node3);
            if (newDensity >= desiredDensity) {
                Float a = 0, b = stepSz, x = a, fx = integratedDensity - desiredDensity, stepSizeSqr = stepSz * stepSz, temp = scale / stepSizeSqr;
                int it = 1;
                while (true) {
                    Float dfx = temp * (node1 * stepSizeSqr - (3 * node1 - 4 * node2 + node3) * stepSz * x + 2 * (node1 - 2 * node2 + node3) * x * x);
                    x -= fx / dfx;
                    if (x <= a || x >= b || dfx == 0) x = 0.5f * (b + a);
                    Float intval = integratedDensity + temp * (1.0f / 6.0f) * (x * (6 * node1 * stepSizeSqr - 3 * (3 * node1 - 4 * node2 + node3) * stepSz * x
                                                                                  + 4 * (node1 - 2 * node2 + node3) * x * x));
                    fx = intval - desiredDensity;
                    if (std::abs(fx) < 1e-6f) {
                        t = mint + stepSz * i + x;
                        integratedDensity = intval;
                        densityAtT = temp * (node1 * stepSizeSqr - (3 * node1 - 4 * node2 + node3) * stepSz * x + 2 * (node1 - 2 * node2 + node3) * x * x);
                        return true;
                    } else if (++it > 30) return false;
                    if (fx > 0) b = x;
                    else a = x;
                }
            }
            V3 next = p + fullStep;
            if (p.x == next.x && p.y == next.y && p.z == next.z) break;
            integratedDensity = newDensity;
            node1 = node3;
            p = next;
        }
        return false;
    }
    /* Medium::evalTransmittance */
    Spec evalTransmittance(const Ray &ray) {
        if (type == 0) {                                                   // homogeneous.cpp:266-273
            Float negLength = ray.mint - ray.maxt;
            Spec T;
            for (int i = 0; i < 3; ++i) T[i] = sigmaT[i] != 0 ? fastexp(sigmaT[i] * negLength) : (Float) 1.0f;
            return T;
        }
        return Spec(fastexp(-integrateDensity(ray)));                     // heterogeneous.cpp:546-548 (simpson)
    }
    /* Medium::eval */
    void eval(const Ray &ray, MediumSamplingRecord &mRec) {
        if (type == 0) {                                                   // homogeneous.cpp:354-396 (EBalance)
            Float distance = ray.maxt - ray.mint;
            Float pdfFailure = 0;
            for (int i = 0; i < 3; ++i) { Float temp = fastexp(-sigmaT[i] * distance); pdfFailure += temp; }
            pdfFailure /= 3;
            for (int i = 0; i < 3; ++i) mRec.transmittance[i] = fastexp(sigmaT[i] * (-distance));
            mRec.pdfFailure = pdfFailure * samplingWeight + (1 - samplingWeight);
            mRec.sigmaS = sigmaS;
            if (mRec.transmittance.max() < 1e-20) mRec.transmittance = Spec(0.0f);
            return;
        }
        /* heterogeneous.cpp:665-691 (simpson) */
        Float expVal = fastexp(-integrateDensity(ray));
        V3 p = ray(ray.maxt);
        Float maxtDensity = lookupDensity(p) * scale;
        mRec.sigmaS = hetAlbedo * maxtDensity;
        mRec.transmittance = Spec(expVal);
        mRec.pdfFailure = expVal;
    }
    Float phaseEval(const V3 &wi, const V3 &wo) const {
        if (phaseType == ALVRL_PHASE_ISOTROPIC) return INV_FOURPI;        // isotropic.cpp:76-78
        Float temp = 1.0f + g * g + 2.0f * g * dot(wi, wo);               // hg.cpp:107-110
        return INV_FOURPI * (1 - g * g) / (temp * std::sqrt(temp));
    }
};

struct VRL { Spec power; V3 start, end; };                               // VRL.h:17-100

/* Scene::evalTransmittance, scene.cpp:619-679, for scenes without ENull surfaces (appendix A10):
 * any surface hit inside the segment is an occluder. */
inline Spec evalTransmittance(Scene &scene, Medium &medium, const V3 &p1, bool p1OnSurface, const V3 &p2, bool p2OnSurface,
                              uint32_t *hitPrim = nullptr, bool *tie = nullptr) {
    V3 d = p2 - p1;
    Float remaining = d.length();
    d = d / remaining;
    Float lengthFactor = p2OnSurface ? (1 - ShadowEpsilon) : 1;
    Ray ray(p1, d, p1OnSurface ? Epsilon : 0, remaining * lengthFactor);
    if (hitPrim) *hitPrim = ALVRL_NO_HIT;
    if (remaining > 0) {
        Float t, u, v; uint32_t prim;
        bool surface = scene.closestHit(ray, true, t, prim, u, v, tie);
        if (hitPrim) *hitPrim = prim;
        if (surface) return Spec(0.0f);
        return medium.evalTransmittance(Ray(ray.o, ray.d, 0, std::min(t, remaining)));
    }
    return Spec(1.0f);
}

/* ---- vrlIntegrator.cpp: sampling + estimator ------------------------------------------- */
struct IntegratorCore {
    Scene *scene; Medium *medium; int volVolSamples, volSurfSamples; bool shortVrls;
    uint64_t shadowRays = 0;
    bool noVisibility = false;   // test hook: skip the occlusion query (T = medium only)
    Float grazeTol = 0;          // test hook: > 0 -> `grazed` collects "a shadow ray's decision was within grazeTol of flipping"
    bool grazed = false;

    Spec transUV(const V3 &a, bool aSurf, const V3 &b) {
        shadowRays++;
        if (noVisibility) {
            V3 d = b - a; Float remaining = d.length(); d = d / remaining;
            return medium->evalTransmittance(Ray(a, d, 0, remaining));
        }
        if (grazeTol > 0 && !grazed) {
            V3 d = b - a; Float remaining = d.length();
            if (remaining > 0) { d = d / remaining; grazed = scene->shadowDecisionFragile(Ray(a, d, aSurf ? Epsilon : 0, remaining), grazeTol); }
        }
        return evalTransmittance(*scene, *medium, a, aSurf, b, false);
    }

    /* vrlIntegrator.cpp:962-1032 */
    static Float getClosestPoints(V3 S1P0, V3 S1P1, V3 S2P0, V3 S2P1, V3 &S1h, V3 &S2h) {
        V3 u = S1P1 - S1P0, v = S2P1 - S2P0, w = S1P0 - S2P0;
        Float a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
        Float D = a * c - b * b;
        Float sc, sN, sD = D, tc, tN, tD = D;
        if (D < Epsilon * u.lengthSquared() * v.lengthSquared()) {
            sN = 0.0; sD = 1.0; tN = e; tD = c;
        } else {
            sN = (b * e - c * d);
            tN = (a * e - b * d);
            if (sN < 0.0) { sN = 0.0; tN = e; tD = c; }
            else if (sN > sD) { sN = sD; tN = e + b; tD = c; }
        }
        if (tN < 0.0) {
            tN = 0.0;
            if (-d < 0.0) sN = 0.0;
            else if (-d > a) sN = sD;
            else { sN = -d; sD = a; }
        } else if (tN > tD) {
            tN = tD;
            if ((-d + b) < 0.0) sN = 0;
            else if ((-d + b) > a) sN = sD;
            else { sN = (-d + b); sD = a; }
        }
        sc = sN / sD;
        tc = tN / tD;
        V3 dP = w + (sc * u) - (tc * v);
        S1h = S1P0 + sc * (S1P1 - S1P0);
        S2h = S2P0 + tc * (S2P1 - S2P0);
        return dP.length();
    }
    static Float A(Float x, Float h, Float sinTheta) { return o_asinh((x / h) * sinTheta); }   // 955-957
    /* vrlIntegrator.cpp:916-953 */
    static Float sampleVtoDistance(const Ray &eyeRay, const V3 &itsP, const VRL &vrl, V3 &V, Float uniform) {
        if (distance(vrl.start, vrl.end) == 0) { V = vrl.start; return 1; }
        Float cosTheta = dot(normalize(eyeRay.d), normalize(vrl.end - vrl.start));
        Float sinTheta = safe_sqrt(1 - cosTheta * cosTheta);
        if (sinTheta < Epsilon) {
            V = vrl.start + uniform * (vrl.end - vrl.start);
            return 1 / distance(vrl.end, vrl.start);
        }
        V3 Uh, Vh;
        Float h = getClosestPoints(eyeRay.o, itsP, vrl.start, vrl.end, Uh, Vh);
        Float V0c = -1 * distance(Vh, vrl.start);
        Float V1c = distance(Vh, vrl.end);
        Float newV = h * o_sinh(A(V0c, h, sinTheta) + (uniform * (A(V1c, h, sinTheta) - A(V0c, h, sinTheta))));
        newV = newV / sinTheta;
        Float result = 1.0f / std::sqrt(h * h + newV * newV * sinTheta * sinTheta);
        Float denom = (A(V1c, h, sinTheta) - A(V0c, h, sinTheta)) / sinTheta;
        newV += distance(Vh, vrl.start);
        V = vrl.start + newV * (normalize(vrl.end - vrl.start));
        return result / denom;
    }
    /* vrlIntegrator.cpp:889-914 */
    static Float KullaSampling(V3 Apt, V3 B, V3 D, V3 &result, Sampler *sampler) {
        V3 dir = normalize(B - Apt);
        Float dotPr = dot(dir, D - Apt);
        V3 I = Apt + (dotPr * dir);
        Float Dis = distance(D, I);
        Float angle_a = o_atan(distance(Apt, I) / Dis);
        Float angle_b = o_atan(distance(I, B) / Dis);
        if (dotPr > 0) {
            angle_a *= -1;
            if (distance(Apt, I) > distance(Apt, B)) angle_b *= -1;
        }
        Float uniform = sampler->next1D();
        Float t = Dis * o_tan(((1.0f - uniform) * angle_a) + (uniform * angle_b));
        Float pdf = Dis / ((angle_b - angle_a) * (Dis * Dis + t * t));
        result = I + (t * dir);
        return pdf;
    }
    /* vrlIntegrator.cpp:860-876 (finite eye segment; infinite rays are dropped earlier, quirk B5) */
    static Float sampleUVKulla(const Ray &eyeRay, const V3 &itsP, const VRL &vrl, V3 &U, V3 &V, Sampler *sampler) {
        Float result = sampleVtoDistance(eyeRay, itsP, vrl, V, sampler->next1D());
        Float dist = distance(itsP, eyeRay.o);
        V3 Apt = eyeRay.o, B = eyeRay.o + (dist * eyeRay.d);
        result *= KullaSampling(Apt, B, V, U, sampler);
        return result;
    }

    /* integrateVRL, vrlIntegrator.cpp:603-785 (single medium: eyeMedium == vrlMedium, appendix A10) */
    Spec integrateVRL(const Ray &ray, const Intersection &its, const VRL &vrl, Sampler *sampler,
                      Float *contrib, Float *variance, Spec weight = Spec(1.0f)) {
        if (contrib) *contrib = 0;
        if (variance) *variance = 0;
        if (medium->sigmaS.isZero()) return Spec(0.0f);
        V3 U, V, S = vrl.start, E = ray.o, Usurf = its.p;
        V3 SV = normalize(vrl.end - vrl.start), VU, EU = ray.d;
        MediumSamplingRecord eyeMRec, vrlMRec;
        Float samplingPDF;
        Spec totalContribution(0.0f), transmittanceUV;
        std::vector<Float> volVolSampleLum(volVolSamples, 0.0f);
        for (int sample = 0; sample < volVolSamples; sample++) {
            samplingPDF = sampleUVKulla(ray, its.p, vrl, U, V, sampler);
            if (distance(U, V) == 0) continue;
            VU = normalize(U - V);
            transmittanceUV = transUV(U, false, V);
            if (transmittanceUV.isZero()) continue;
            medium->eval(Ray(E, EU, 0, distance(E, U)), eyeMRec);
            medium->eval(Ray(S, SV, 0, distance(S, V)), vrlMRec);
            Spec contribution = weight;
            contribution *= vrl.power;
            contribution *= vrlMRec.sigmaS * eyeMRec.sigmaS / samplingPDF;
            contribution *= 1 / distanceSquared(U, V);
            contribution *= vrlMRec.transmittance;
            contribution *= transmittanceUV;
            contribution *= eyeMRec.transmittance;
            if (shortVrls) contribution /= vrlMRec.pdfFailure;
            contribution *= medium->phaseEval(-VU, -EU);
            contribution *= medium->phaseEval(-SV, VU);
            if (contribution.isValid()) {
                totalContribution += contribution / (Float) volVolSamples;
                volVolSampleLum[sample] = contribution.getLuminance();
            }
        }
        Float mean = 0, M2 = 0;
        for (int i = 0; i < volVolSamples; i++) {
            Float delta = volVolSampleLum[i] - mean;
            mean += delta / (i + 1);
            M2 += delta * (volVolSampleLum[i] - mean);
        }
        if (contrib && volVolSamples > 0) *contrib += mean;
        if (variance && volVolSamples > 0) *variance += M2 / ((volVolSamples - 1) * volVolSamples);

        U = Usurf;
        Spec transmittanceEUsurf(0.0f);
        if (its.isValid()) {
            if (distance(Usurf, E) != 0) {
                medium->eval(Ray(ray.o, ray.d, 0, distance(Usurf, E)), eyeMRec);    // quirk B1: vrlMedium
                transmittanceEUsurf = eyeMRec.transmittance;
            }
        }
        std::vector<Float> volSurfSampleLum(volSurfSamples, 0.0f);
        bool smooth = (scene->matBits[its.material] & ALVRL_BSDF_SMOOTH) != 0;
        if (!transmittanceEUsurf.isZero() && smooth) {
            for (int sample = 0; sample < volSurfSamples; sample++) {
                samplingPDF = KullaSampling(vrl.start, vrl.end, its.p, V, sampler);   // sampleV, 838-842
                if (distance(U, V) == 0) continue;
                VU = normalize(U - V);
                Spec tUV = transUV(U, true, V);
                medium->eval(Ray(S, SV, 0, distance(S, V)), vrlMRec);
                Spec contribution = weight;
                contribution *= vrl.power;
                contribution *= medium->sigmaS / samplingPDF;                           // base-class getSigmaS(), quirk B2
                contribution *= 1 / distanceSquared(U, V);
                contribution *= vrlMRec.transmittance;
                contribution *= tUV;
                contribution *= transmittanceEUsurf;
                if (shortVrls) contribution /= vrlMRec.pdfFailure;
                contribution *= medium->phaseEval(-SV, VU);
                /* SmoothDiffuse::eval, diffuse.cpp:110-118; wo = its.toLocal(-VU) */
                Float cosWo = dot(-VU, its.n);
                Spec bsdf(0.0f);
                if (!(its.wiz <= 0 || cosWo <= 0)) bsdf = scene->albedo[its.material] * (INV_PI * cosWo);
                contribution *= bsdf;
                if (contribution.isValid()) {
                    totalContribution += contribution / (Float) volSurfSamples;
                    volSurfSampleLum[sample] = contribution.getLuminance();
                }
            }
        }
        mean = 0; M2 = 0;
        for (int i = 0; i < volSurfSamples; i++) {
            Float delta = volSurfSampleLum[i] - mean;
            mean += delta / (i + 1);
            M2 += delta * (volSurfSampleLum[i] - mean);
        }
        if (contrib && volSurfSamples > 0) *contrib += mean;
        if (variance && volSurfSamples > 0) *variance += M2 / ((volSurfSamples - 1) * volSurfSamples);
        return totalContribution;
    }
};

} // namespace orc
