/*
 * axcd_scene.h — deterministic synthetic-scene generator (host only, no CUDA).
 *
 * Draws from a generator bit-compatible with axiom::math::DeterministicRNG
 * (reference: include/axiom/math/random.hpp:19-68, PCG-XSH-RR; goldens in SURVEY.md Appendix C) in a
 * fixed order per body: type, px,py,pz, axis (Marsaglia rejection as random.hpp:102-116),
 * angle in [0,2*pi), size parameters; scale = (1,1,1); rotation = Quat::fromAxisAngle
 * (src/math/quat.cpp:68-72 -> glm::angleAxis: w = cos(a/2), xyz = axis*sin(a/2)).
 * The same blobs feed the CUDA path and the CPU oracle (BASELINE.md "Inputs").
 */
#ifndef AXCD_SCENE_H
#define AXCD_SCENE_H
#include <stdint.h>
#include "axcd.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct AxcdSceneSpec {
    uint32_t numBodies;
    float fracBox;      /* P(box)                                   */
    float fracSphere;   /* P(sphere); hull probability is the rest  */
    float domain;       /* positions uniform in [0, domain)^3       */
    float sizeMin;      /* radius / half-extent / semi-axis range   */
    float sizeMax;
    uint32_t hullVerts; /* vertices per hull (on an ellipsoid)      */
    uint64_t seed;
} AxcdSceneSpec;

/* Writes numBodies transforms (10 floats each, axiom::math::Transform layout) and shapes; hull
 * vertices are appended to hullXYZ starting at vertex index firstHullVert (capacity hullCap
 * vertices); *outHullVerts = vertices appended.  Returns 0 or 601 when hullCap is too small.    */
AXCD_API int32_t axcd_scene_generate(const AxcdSceneSpec* spec, float* xf, AxcdShape* shapes,
                                     float* hullXYZ, uint32_t hullCap, uint32_t firstHullVert,
                                     uint32_t* outHullVerts);

/* numWorlds independent worlds of spec->numBodies bodies each, world w seeded spec->seed + w
 * (config C3: seed 1000 + worldId); bodies of world w occupy [w*numBodies, (w+1)*numBodies).    */
AXCD_API int32_t axcd_scene_generate_worlds(const AxcdSceneSpec* spec, uint32_t numWorlds,
                                            float* xf, AxcdShape* shapes, uint32_t* worldId,
                                            float* hullXYZ, uint32_t hullCap,
                                            uint32_t* outHullVerts);

/* First n outputs of the generator's RNG for a seed (known-answer tests).                       */
AXCD_API void axcd_scene_rng_u32(uint64_t seed, uint32_t n, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
