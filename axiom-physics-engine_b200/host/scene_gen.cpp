// Deterministic synthetic-scene generator (see include/axcd_scene.h).  Host only.
#include "axcd_scene.h"

#include <cmath>
#include <cstring>

namespace {

// Bit-compatible with axiom::math::DeterministicRNG (include/axiom/math/random.hpp:25-56).
class Pcg32 {
public:
    explicit Pcg32(uint64_t seed) : s_(seed | 1ULL) {
        for (int i = 0; i < 10; ++i) next();
    }
    uint32_t next() {
        const uint64_t o = s_;
        s_ = o * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint32_t x = static_cast<uint32_t>(((o >> 18U) ^ o) >> 27U);
        const uint32_t r = static_cast<uint32_t>(o >> 59U);
        return (x >> r) | (x << ((~r + 1U) & 31));
    }
    // May return exactly 1.0f (random.hpp:55 quirk); the scene tolerates it.
    float unit() { return static_cast<float>(next()) / 4294967296.0f; }
    float range(float lo, float hi) { return lo + unit() * (hi - lo); }

private:
    uint64_t s_;
};

void unitVector(Pcg32& g, float out[3]) {
    float x, y, z, l2;
    do {
        x = g.range(-1.0f, 1.0f);
        y = g.range(-1.0f, 1.0f);
        z = g.range(-1.0f, 1.0f);
        l2 = x * x + y * y + z * z;
    } while (l2 > 1.0f || l2 < 1e-6f);
    const float inv = 1.0f / std::sqrt(l2);
    out[0] = x * inv;
    out[1] = y * inv;
    out[2] = z * inv;
}

uint32_t asBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
float asFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

int32_t generateInto(const AxcdSceneSpec& sp, Pcg32& g, float* xf, AxcdShape* shapes,
                     float* hull, uint32_t hullCap, uint32_t& hullUsed) {
    const float twoPi = 6.28318530717958647692f;
    for (uint32_t i = 0; i < sp.numBodies; ++i) {
        const float t = g.unit();
        float* T = xf + 10ull * i;
        T[0] = g.range(0.0f, sp.domain);
        T[1] = g.range(0.0f, sp.domain);
        T[2] = g.range(0.0f, sp.domain);
        float axis[3];
        unitVector(g, axis);
        const float angle = g.range(0.0f, twoPi);
        const float s = std::sin(angle * 0.5f), c = std::cos(angle * 0.5f);
        T[3] = axis[0] * s;
        T[4] = axis[1] * s;
        T[5] = axis[2] * s;
        T[6] = c;
        T[7] = T[8] = T[9] = 1.0f;
        AxcdShape& sh = shapes[i];
        if (t < sp.fracBox) {
            sh.type = AXCD_SHAPE_BOX;
            sh.p0 = g.range(sp.sizeMin, sp.sizeMax);
            sh.p1 = g.range(sp.sizeMin, sp.sizeMax);
            sh.p2 = g.range(sp.sizeMin, sp.sizeMax);
        } else if (t < sp.fracBox + sp.fracSphere) {
            sh.type = AXCD_SHAPE_SPHERE;
            sh.p0 = g.range(sp.sizeMin, sp.sizeMax);
            sh.p1 = sh.p2 = 0.0f;
        } else {
            const float ax = g.range(sp.sizeMin, sp.sizeMax);
            const float ay = g.range(sp.sizeMin, sp.sizeMax);
            const float az = g.range(sp.sizeMin, sp.sizeMax);
            if (hullUsed + sp.hullVerts > hullCap) return AXCD_ERR_OUT_OF_RANGE;
            sh.type = AXCD_SHAPE_CONVEX;
            sh.p0 = asFloat(hullUsed);
            sh.p1 = asFloat(sp.hullVerts);
            sh.p2 = 0.0f;
            for (uint32_t k = 0; k < sp.hullVerts; ++k) {
                float d[3];
                unitVector(g, d);
                float* v = hull + 3ull * (hullUsed + k);
                v[0] = d[0] * ax;
                v[1] = d[1] * ay;
                v[2] = d[2] * az;
            }
            hullUsed += sp.hullVerts;
        }
    }
    (void)asBits;
    return AXCD_OK;
}

}  // namespace

extern "C" {

int32_t axcd_scene_generate(const AxcdSceneSpec* spec, float* xf, AxcdShape* shapes,
                            float* hullXYZ, uint32_t hullCap, uint32_t firstHullVert,
                            uint32_t* outHullVerts) {
    if (!spec || !xf || !shapes) return AXCD_ERR_NULL_POINTER;
    Pcg32 g(spec->seed);
    uint32_t used = firstHullVert;
    const int32_t rc = generateInto(*spec, g, xf, shapes, hullXYZ, hullCap, used);
    if (outHullVerts) *outHullVerts = used - firstHullVert;
    return rc;
}

int32_t axcd_scene_generate_worlds(const AxcdSceneSpec* spec, uint32_t numWorlds, float* xf,
                                   AxcdShape* shapes, uint32_t* worldId, float* hullXYZ,
                                   uint32_t hullCap, uint32_t* outHullVerts) {
    if (!spec || !xf || !shapes || !worldId) return AXCD_ERR_NULL_POINTER;
    uint32_t used = 0;
    for (uint32_t w = 0; w < numWorlds; ++w) {
        Pcg32 g(spec->seed + w);
        const uint64_t base = static_cast<uint64_t>(w) * spec->numBodies;
        const int32_t rc =
            generateInto(*spec, g, xf + 10ull * base, shapes + base, hullXYZ, hullCap, used);
        if (rc) return rc;
        for (uint32_t i = 0; i < spec->numBodies; ++i) worldId[base + i] = w;
    }
    if (outHullVerts) *outHullVerts = used;
    return AXCD_OK;
}

void axcd_scene_rng_u32(uint64_t seed, uint32_t n, uint32_t* out) {
    Pcg32 g(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = g.next();
}

}  // extern "C"
