/*
 * host_sampler.h -- host-side sample streams of the plugin.
 *
 *  SfmtStream    the reference's IndependentSampler over SFMT-19937 (src/samplers/independent.cpp:52-103,
 *                src/libcore/random.cpp): 64-bit seeding is Mitsuba's own recurrence (random.cpp:397-406), clone()
 *                draws 312 words from the parent and runs init_by_array (random.cpp:408-469,528-549), next1D() keeps
 *                the low 32 bits >> 9 (random.cpp:630-639).  Used when rngMode = ALVRL_RNG_MODE_SFMT so that the
 *                prepass consumes exactly the reference's stream (workerCount decides the clone layout).
 *  CounterStream the addressed stream of include/alvrl_rng.h.
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include "../../include/alvrl_rng.h"

namespace alvrl {

class HostSampler {
public:
    virtual ~HostSampler() {}
    virtual float next1D() = 0;
    virtual HostSampler *clone() = 0;
    virtual void setContext(uint32_t /*domain*/, uint32_t /*a*/, uint32_t /*b*/) {}
    /* addressed streams expose (key, position) so that device code can continue them (refine.cuh) */
    virtual bool counterState(uint32_t & /*key*/, uint32_t & /*pos*/) const { return false; }
    virtual void setCounterPos(uint32_t /*pos*/) {}
    /* counter stream: the draws of Clustering::split come from the sub-stream of the cluster [begin, end) (alvrl_rng.h) */
    virtual void enterNode(uint32_t /*begin*/, uint32_t /*end*/) {}
    virtual void leaveNode() {}
};

class Sfmt19937 {
    enum { kWords128 = 156, kWords32 = 624, kWords64 = 312, kPos1 = 122 };
    uint32_t s[kWords32];
    int idx;
    void certify() {
        static const uint32_t parity[4] = {0x00000001u, 0u, 0u, 0x13c9e684u};
        uint32_t inner = 0;
        for (int i = 0; i < 4; i++) inner ^= s[i] & parity[i];
        inner ^= inner >> 16; inner ^= inner >> 8; inner ^= inner >> 4; inner ^= inner >> 2; inner ^= inner >> 1;
        if (inner & 1) return;
        for (int i = 0; i < 4; i++)
            for (uint32_t bit = 1; bit; bit <<= 1)
                if (parity[i] & bit) { s[i] ^= bit; return; }
    }
    /* one step of the recursion on 128-bit lanes: r = a ^ (a << 8) ^ ((b >> 11) & mask) ^ (c >> 8) ^ (d << 18) */
    static void step(uint32_t *r, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d) {
        static const uint32_t mask[4] = {0xdfffffefu, 0xddfecb7fu, 0xbffaffffu, 0xbffffff6u};
        const uint32_t x0 = a[0] << 8, x1 = (a[1] << 8) | (a[0] >> 24), x2 = (a[2] << 8) | (a[1] >> 24), x3 = (a[3] << 8) | (a[2] >> 24);
        const uint32_t y0 = (c[0] >> 8) | (c[1] << 24), y1 = (c[1] >> 8) | (c[2] << 24), y2 = (c[2] >> 8) | (c[3] << 24), y3 = c[3] >> 8;
        const uint32_t o0 = a[0] ^ x0 ^ ((b[0] >> 11) & mask[0]) ^ y0 ^ (d[0] << 18);
        const uint32_t o1 = a[1] ^ x1 ^ ((b[1] >> 11) & mask[1]) ^ y1 ^ (d[1] << 18);
        const uint32_t o2 = a[2] ^ x2 ^ ((b[2] >> 11) & mask[2]) ^ y2 ^ (d[2] << 18);
        const uint32_t o3 = a[3] ^ x3 ^ ((b[3] >> 11) & mask[3]) ^ y3 ^ (d[3] << 18);
        r[0] = o0; r[1] = o1; r[2] = o2; r[3] = o3;
    }
    void refill() {
        const uint32_t *r1 = &s[4 * (kWords128 - 2)], *r2 = &s[4 * (kWords128 - 1)];
        for (int i = 0; i < kWords128; i++) {
            const int j = (i + kPos1 < kWords128) ? i + kPos1 : i + kPos1 - kWords128;
            step(&s[4 * i], &s[4 * i], &s[4 * j], r1, r2);
            r1 = r2; r2 = &s[4 * i];
        }
    }
public:
    explicit Sfmt19937(uint64_t seed) {
        uint64_t prev = seed;
        memcpy(&s[0], &prev, 8);
        for (int i = 1; i < kWords64; i++) {
            prev = 6364136223846793005ull * (prev ^ (prev >> 62)) + (uint64_t) i;
            memcpy(&s[2 * i], &prev, 8);
        }
        idx = kWords32;
        certify();
    }
    explicit Sfmt19937(Sfmt19937 &parent) {
        uint32_t key[kWords32];
        for (int i = 0; i < kWords64; i++) { uint64_t w = parent.next64(); memcpy(&key[2 * i], &w, 8); }
        seedArray(key, kWords32);
    }
    void seedArray(const uint32_t *key, int len) {
        auto f1 = [](uint32_t x) { return (x ^ (x >> 27)) * 1664525u; };
        auto f2 = [](uint32_t x) { return (x ^ (x >> 27)) * 1566083941u; };
        const int n = kWords32, lag = 11, mid = (n - lag) / 2;
        memset(s, 0x8b, sizeof(s));
        int count = (len + 1 > n) ? len + 1 : n;
        uint32_t r = f1(s[0] ^ s[mid] ^ s[n - 1]);
        s[mid] += r; r += (uint32_t) len; s[mid + lag] += r; s[0] = r;
        count--;
        int i = 1, j = 0;
        for (; j < count && j < len; j++) {
            r = f1(s[i] ^ s[(i + mid) % n] ^ s[(i + n - 1) % n]);
            s[(i + mid) % n] += r; r += key[j] + (uint32_t) i;
            s[(i + mid + lag) % n] += r; s[i] = r; i = (i + 1) % n;
        }
        for (; j < count; j++) {
            r = f1(s[i] ^ s[(i + mid) % n] ^ s[(i + n - 1) % n]);
            s[(i + mid) % n] += r; r += (uint32_t) i;
            s[(i + mid + lag) % n] += r; s[i] = r; i = (i + 1) % n;
        }
        for (j = 0; j < n; j++) {
            r = f2(s[i] + s[(i + mid) % n] + s[(i + n - 1) % n]);
            s[(i + mid) % n] ^= r; r -= (uint32_t) i;
            s[(i + mid + lag) % n] ^= r; s[i] = r; i = (i + 1) % n;
        }
        idx = n;
        certify();
    }
    uint64_t next64() {
        if (idx >= kWords32) { refill(); idx = 0; }
        uint64_t w; memcpy(&w, &s[idx], 8);
        idx += 2;
        return w;
    }
    float nextFloat() { return alvrl_bits_to_float((uint32_t) (next64() & 0xffffffffull)); }
};

class SfmtStream : public HostSampler {
    Sfmt19937 g;
public:
    explicit SfmtStream(uint64_t seed) : g(seed) {}
    explicit SfmtStream(Sfmt19937 &parent) : g(parent) {}
    float next1D() override { return g.nextFloat(); }
    HostSampler *clone() override { return new SfmtStream(g); }
    uint64_t next64() { return g.next64(); }
};

class CounterStream : public HostSampler {
    uint64_t seed; uint32_t key = 0, k = 0, outerKey = 0, outerK = 0; bool inNode = false;
public:
    explicit CounterStream(uint64_t s) : seed(s) {}
    void setContext(uint32_t domain, uint32_t a, uint32_t b) override { key = alvrl_rng_key(seed, domain, a, b); k = 0; inNode = false; }
    void enterNode(uint32_t begin, uint32_t end) override { leaveNode(); outerKey = key; outerK = k; key = alvrl_rng_node_key(key, begin, end); k = 0; inNode = true; }
    void leaveNode() override { if (inNode) { key = outerKey; k = outerK; inNode = false; } }
    float next1D() override { return alvrl_rng_uniform(key, k++); }
    HostSampler *clone() override { return new CounterStream(seed); }
    bool counterState(uint32_t &key_, uint32_t &pos) const override { key_ = inNode ? outerKey : key; pos = inNode ? outerK : k; return true; }
    void setCounterPos(uint32_t pos) override { if (inNode) outerK = pos; else k = pos; }
};

} // namespace alvrl
