/*
 * axref.h — C interface of the CPU ORACLE for the collision hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (axiom-physics-engine_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" for broadphase pair sets, GJK distances and EPA contacts — the
 * reference snapshot has no collision code (src/collision/.gitkeep is empty), so no
 * reference test or golden vector exists for them.  The math building blocks this oracle is
 * composed of ARE pinned against the reference's own tests (tests/test_oracle_math.py cites them),
 * and the authored algorithms are validated independently: brute-force O(N^2) overlap,
 * closed-form sphere/box answers and a scipy QP cross-check (tests/test_oracle_*.py).
 */
#ifndef AXREF_H
#define AXREF_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct AxrefShape { uint32_t type; float p0, p1, p2; } AxrefShape;   /* == AxcdShape   */
typedef struct AxrefContact {                                                 /* == AxcdContact */
    uint32_t a, b; float px, py, pz, nx, ny, nz, depth; uint32_t status;
} AxrefContact;
typedef struct AxrefNarrowCfg {
    uint32_t gjkMaxIters, epaMaxIters, epaMaxFaces;
    float gjkTol, epaTol;
    uint32_t wantDistances;   /* 1: run GJK to convergence on every pair and report distances */
    uint32_t flags;           /* AXREF_BOXBOX_GJK_EPA: box-box pairs through GJK/EPA instead of the SAT */
} AxrefNarrowCfg;
enum { AXREF_BOXBOX_GJK_EPA = 1u };
typedef struct AxrefNarrowStats {
    uint64_t numContacts, numPenetrating, gjkFailures, epaFailures, gjkIterations;
} AxrefNarrowStats;

/* math KATs (pin the GLM-free restatement against the reference's own test expectations) */
void axref_quat_rotate(const float q[4], const float v[3], float out[3]);
void axref_quat_mul(const float p[4], const float q[4], float out[4]);
void axref_quat_to_mat3(const float q[4], float outColMajor[9]);
void axref_transform_point(const float xf[10], const float p[3], float out[3]);
void axref_rng_u32(uint64_t seed, uint32_t n, uint32_t* out);
void axref_rng_float(uint64_t seed, uint32_t n, float* out);
int  axref_aabb_intersects(const float a[6], const float b[6]);
float axref_vec3_dot(const float a[3], const float b[3]);
void axref_vec3_cross(const float a[3], const float b[3], float o[3]);
void axref_quat_conjugate(const float q[4], float o[4]);                                  /* quat.hpp:96 */
void axref_transform_direction(const float xf[10], const float d[3], float o[3]);         /* transform.cpp:95-100 */
void axref_inverse_transform_point(const float xf[10], const float p[3], float o[3]);     /* transform.cpp:113-122 */
void axref_inverse_transform_direction(const float xf[10], const float d[3], float o[3]); /* transform.cpp:124-131 */
void axref_aabb_expand_point(const float a[6], const float p[3], float o[6]);
void axref_aabb_merge(const float a[6], const float b[6], float o[6]);
void axref_aabb_center(const float a[6], float o[3]);
void axref_aabb_expand_margin(const float a[6], float margin, float o[6]);
void axref_aabb_from_center_extents(const float c[3], const float h[3], float o[6]);

/* stage 1: refit.  xf = n x 10 floats (axiom::math::Transform), out = n x 6 floats (AABB).    */
int32_t axref_refit(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                    uint32_t nHullVerts, float margin, float* outAabb, int nthreads);

/* the same with the box route selectable: mat4Route != 0 refits boxes through AABB::transform(Transform::toMatrix())
 * (src/math/aabb.cpp:8-35) instead of 8 x Transform::transformPoint — SURVEY.md 8(a) row a15.            */
int32_t axref_refit_route(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                          uint32_t nHullVerts, float margin, float* outAabb, int nthreads, int mat4Route);

/* stage 2: candidate pairs, canonical (a<b, sorted).  *outCount is the true count even when it
 * exceeds cap (only cap pairs are written).  worldId may be NULL.                             */
int32_t axref_broadphase_brute(const float* aabb, uint32_t n, const uint32_t* worldId,
                               uint32_t* outPairs, uint64_t cap, uint64_t* outCount);
int32_t axref_broadphase_grid(const float* aabb, uint32_t n, const uint32_t* worldId,
                              uint32_t* outPairs, uint64_t cap, uint64_t* outCount, int nthreads);

/* the same with collision filters: filt = n x 3 words (categoryBits, maskBits, groupIndex as int32),
 * gui::FilterInfo semantics (include/axiom/gui/body_inspector.hpp:38-42); NULL = no filtering.  */
int32_t axref_broadphase_brute_f(const float* aabb, uint32_t n, const uint32_t* worldId, const uint32_t* filt,
                                 uint32_t* outPairs, uint64_t cap, uint64_t* outCount);
int32_t axref_broadphase_grid_f(const float* aabb, uint32_t n, const uint32_t* worldId, const uint32_t* filt,
                                uint32_t* outPairs, uint64_t cap, uint64_t* outCount, int nthreads);

/* stage 3: narrowphase over given pairs (in the given order).  outDist may be NULL.           */
int32_t axref_narrowphase(const float* xf, const AxrefShape* shapes, uint32_t n,
                          const float* hullXYZ, uint32_t nHullVerts, const uint32_t* pairs,
                          uint64_t npairs, const AxrefNarrowCfg* cfg, AxrefContact* outContacts,
                          uint64_t cap, uint64_t* outCount, float* outDist,
                          AxrefNarrowStats* stats, int nthreads);

/* stage 4: contact manifolds (1..4 points per contact; box-box face clipping, single point otherwise),
 * one per contact in the given order.  == AxcdManifold.                                        */
typedef struct AxrefManifold {
    uint32_t a, b; float nx, ny, nz; uint32_t count; float px[4], py[4], pz[4], depth[4];
} AxrefManifold;
int32_t axref_manifolds(const float* xf, const AxrefShape* shapes, uint32_t n, const AxrefContact* contacts,
                        uint64_t ncontacts, AxrefManifold* out, uint64_t* outPointCount, int nthreads);

/* scene queries (brute force over all bodies).  AABB query: (query, body) hits in (query, body) order.
 * Ray cast: closest hit per ray; body == 0xffffffff when nothing is hit.  == AxcdRay / AxcdRayHit. */
typedef struct AxrefRay { float ox, oy, oz, dx, dy, dz, tMax; uint32_t world; } AxrefRay;
typedef struct AxrefRayHit { uint32_t body; float t, nx, ny, nz; uint32_t flags; } AxrefRayHit;
int32_t axref_query_aabbs(const float* aabb, uint32_t n, const uint32_t* worldId, const float* qboxes,
                          const uint32_t* qworld, uint32_t nq, uint32_t* outHits, uint64_t cap, uint64_t* outCount);
int32_t axref_raycast(const float* xf, const AxrefShape* shapes, const float* hullXYZ, const float* aabb, uint32_t n,
                      const uint32_t* worldId, const AxrefRay* rays, uint32_t nq, AxrefRayHit* out, int nthreads);

/* GJK-based CCD: time of impact of the given pairs under linear motion (disp = n x 3 displacements over
 * the step), conservative advancement on the exact GJK distance.  == AxcdSweep.                  */
typedef struct AxrefSweep { uint32_t hit; float toi, nx, ny, nz; uint32_t iterations; } AxrefSweep;
int32_t axref_ccd_pairs(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                        const uint32_t* pairs, uint64_t npairs, const float* disp, const AxrefNarrowCfg* cfg,
                        AxrefSweep* out, int nthreads);
/* CCD with rotation: rot = one rotation vector (angular velocity * dt) per body; q(t) = normalize(q + t/2 (w,0)(x)q). */
int32_t axref_ccd_pairs_angular(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                                const uint32_t* pairs, uint64_t npairs, const float* disp, const float* rot,
                                const AxrefNarrowCfg* cfg, AxrefSweep* out, int nthreads);
void axref_ccd_pose_at(const float* xf10, const float* disp3, const float* rot3, float t, float* out10);


/* one pair, for closed-form checks: returns 1 if contact. dist = core GJK distance minus radii
 * (<= 0 for contacts; exact only when cfg->wantDistances).                                    */
int32_t axref_collide_pair(const float xfA[10], const AxrefShape* sa, const float xfB[10],
                           const AxrefShape* sb, const float* hullXYZ, const AxrefNarrowCfg* cfg,
                           AxrefContact* out, float* outDist, uint32_t* outUsedEpa);

#ifdef __cplusplus
}
#endif
#endif
