/*
 * axcd.h — C ABI of the B200-native collision-detection path for the Axiom physics engine.
 *
 * The path: per-body world-space AABB refit -> broadphase candidate-pair finding (Morton radix
 * sort + LBVH) -> GJK (+EPA for penetrating pairs) narrowphase.  It fills the reference's empty
 * `axiom::collision` slot (reference: src/collision/.gitkeep, src/CMakeLists.txt:23-27) and is
 * driven by the call shape the reference documents for it (reference: CLAUDE.md:162-178,
 * include/axiom/core/profiler.hpp:13-23):
 *
 *      broadphase_.update();           -> axcd_refit + axcd_broadphase
 *      broadphase_.getPairCount();     -> AxcdStats.numPairs
 *      narrowphase_.detectCollisions();-> axcd_narrowphase
 *      narrowphase_.getContactCount(); -> AxcdStats.numContacts
 *
 * Every entry point returns an `axiom::core::ErrorCode` value (reference:
 * include/axiom/core/error_code.hpp:21-57); 0 == Success.  Nothing here throws or aborts on bad
 * input.  All pointers are plain host pointers owned by the caller; the context owns all device
 * memory.  One context == one CUDA device + one stream; a context is NOT thread-safe (the
 * reference's device-facing objects have the same contract, include/axiom/gpu/vk_command.hpp:15).
 *
 * There is no CPU fallback behind this ABI: without a CUDA device axcd_create fails with
 * AXCD_ERR_GPU_INIT (500) and no other entry point does any work.
 */
#ifndef AXCD_H
#define AXCD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define AXCD_API __declspec(dllexport)
#else
#define AXCD_API __attribute__((visibility("default")))
#endif

/* ---- status codes == axiom::core::ErrorCode (include/axiom/core/error_code.hpp:21-57) ------ */
enum {
    AXCD_OK = 0,                     /* ErrorCode::Success                                     */
    AXCD_ERR_OUT_OF_MEMORY = 200,    /* ErrorCode::OutOfMemory (host allocation)               */
    AXCD_ERR_NULL_POINTER = 202,     /* ErrorCode::NullPointer                                 */
    AXCD_ERR_INVALID_SHAPE = 300,    /* ErrorCode::InvalidShape                                */
    AXCD_ERR_GJK_NO_CONVERGE = 301,  /* ErrorCode::GJKFailedToConverge   (per-contact status)  */
    AXCD_ERR_EPA_NO_CONVERGE = 302,  /* ErrorCode::EPAFailedToConverge   (per-contact status)  */
    AXCD_ERR_GPU_INIT = 500,         /* ErrorCode::VulkanInitializationFailed slot: device init */
    AXCD_ERR_GPU_ALLOC = 502,        /* ErrorCode::BufferAllocationFailed                      */
    AXCD_ERR_GPU_INVALID_OP = 503,   /* ErrorCode::GPU_INVALID_OPERATION (wrong call order)    */
    AXCD_ERR_GPU_FAILED = 505,       /* ErrorCode::GPU_OPERATION_FAILED (CUDA / NCCL error)    */
    AXCD_ERR_INVALID_PARAM = 600,    /* ErrorCode::InvalidParameter                            */
    AXCD_ERR_OUT_OF_RANGE = 601      /* ErrorCode::OutOfRange (capacity exceeded)              */
};

/* ---- shape types: numeric order of debug::ShapeType (include/axiom/debug/physics_debug_draw.hpp:87-94)
 * Sphere, Box, Capsule, Convex and Cylinder are in scope; Plane/Mesh -> AXCD_ERR_INVALID_SHAPE.  */
enum { AXCD_SHAPE_SPHERE = 0, AXCD_SHAPE_BOX = 1, AXCD_SHAPE_CAPSULE = 2, AXCD_SHAPE_PLANE = 3,
       AXCD_SHAPE_CONVEX = 4, AXCD_SHAPE_MESH = 5,
       /* gui::ShapeType::Cylinder (include/axiom/gui/body_inspector.hpp:24).  debug::ShapeType has no cylinder, so
        * the value lies past that enum's range.                                                          */
       AXCD_SHAPE_CYLINDER = 6 };

/* Flattened debug::DebugShape (physics_debug_draw.hpp:97-112): 16-byte POD.
 *   Sphere : p0 = radius                                  (DebugShape::radius)
 *   Box    : p0,p1,p2 = halfExtents.x/y/z                  (DebugShape::halfExtents)
 *   Capsule: p0 = radius, p1 = height of the segment between the cap centres, local Y axis
 *            (DebugShape::radius / height; src/debug/physics_debug_draw.cpp:254-266)
 *   Cylinder: p0 = radius, p1 = height, local Y axis like the capsule; placed by Transform::transformPoint, so a
 *            non-uniform scale gives an elliptic cylinder.  The rim is resolved to 32768 directions (5e-9 of
 *            the radius, below float resolution).
 *   Convex : p0 = bit pattern of uint32 firstVertex, p1 = bit pattern of uint32 vertexCount,
 *            indexing the xyz-packed hull vertex pool (DebugShape::vertices / vertexCount,
 *            src/debug/physics_debug_draw.cpp:285-288).                                         */
typedef struct AxcdShape {
    uint32_t type;
    float p0, p1, p2;
} AxcdShape;

/* Narrowphase output record, derived from debug::DebugContactPoint
 * (physics_debug_draw.hpp:128-132: position, normal, penetrationDepth) plus the pair ids.
 * a < b are body indices; normal is unit length and points from body a to body b;
 * position is the midpoint of the two witness (surface) points; depth >= 0.
 * status: 0 ok, 301 GJK hit its iteration cap, 302 EPA hit its iteration/face cap.            */
typedef struct AxcdContact {
    uint32_t a, b;
    float px, py, pz;
    float nx, ny, nz;
    float depth;
    uint32_t status;
} AxcdContact;

/* POD config with in-header defaults (axcd_default_config), the reference's config idiom
 * (include/axiom/gui/physics_panel.hpp:34-43).                                                 */
typedef struct AxcdConfig {
    uint32_t maxBodies;     /* capacity: bodies                                                 */
    uint32_t maxPairs;      /* capacity: candidate pairs                                        */
    uint32_t maxContacts;   /* capacity: contacts                                               */
    uint32_t maxHullVerts;  /* capacity: hull vertex pool                                       */
    uint32_t numWorlds;     /* 1 = single scene; >1 = batched independent worlds (worldId req.) */
    float aabbMargin;       /* AABB::expand(float) applied after the tight fit; default 0       */
    uint32_t gjkMaxIters;   /* default 32                                                       */
    uint32_t epaMaxIters;   /* default 32                                                       */
    uint32_t epaMaxFaces;   /* default 64 (hard cap 64)                                         */
    float gjkTol;           /* relative GJK convergence tolerance, default 1e-6                 */
    float epaTol;           /* EPA termination tolerance, default 1e-4                          */
    uint32_t flags;         /* AXCD_FLAG_*                                                      */
    int32_t deviceOrdinal;  /* CUDA device                                                      */
    void* stream;           /* cudaStream_t to run on, or NULL for a context-owned stream       */
} AxcdConfig;

enum {
    /* Also write the GJK distance of every candidate pair (0 for contacts) so that
     * axcd_get_pair_distances works.  Off by default: separated pairs then stop at the first
     * separating axis.                                                                          */
    AXCD_FLAG_PAIR_DISTANCES = 1u,
    /* 2u is reserved (round 1 kept an alternative, slower EPA kernel behind it; removed).              */
    /* Temporal coherence: the AABB buffer holds persistent FAT boxes — a body's
     * box is rebuilt (tight box expanded by aabbMargin, AABB::expand(float), aabb.hpp:156-160) only when
     * its tight box leaves it — and axcd_broadphase reuses the previous candidate list when no body
     * moved out of its fat box (AxcdStats.movedBodies == 0 -> broadphaseSkipped = 1).  The candidate set
     * is then the overlap set of the fat boxes (a superset of the tight one); the contact set is
     * unchanged, because the narrowphase works on the exact shapes.  Not available in x-slab mode.  */
    AXCD_FLAG_TEMPORAL_COHERENCE = 4u,
    /* Box-box pairs are decided in closed form by default (15-axis separating-axis test: contact iff no
     * axis separates, depth = the smallest overlap, normal = that axis — what EPA converges to).  This
     * flag sends them through the generic GJK + EPA path instead, like hulls and capsules (A/B
     * measurements, and tests that validate one against the other).                               */
    AXCD_FLAG_BOXBOX_GJK_EPA = 8u,
    /* axcd_step / axcd_step_async normally replay a CUDA graph of the whole step once the launch
     * configuration has been stable for a step.  This flag (or AXCD_NO_GRAPH=1 in the environment) keeps
     * them on direct launches.                                                                       */
    AXCD_FLAG_NO_GRAPH = 16u,
    /* Boxes are refit through the reference's OTHER route, AABB::transform(Transform::toMatrix()) (8 corners of
     * the local box through the TRS matrix; src/math/aabb.cpp:8-35, src/math/transform.cpp:13-24), instead of
     * 8 x Transform::transformPoint.  Same box up to rounding; kept for hosts that refit that way.      */
    AXCD_FLAG_REFIT_MAT4_ROUTE = 32u
};

typedef struct AxcdStats {
    uint32_t numBodies, numPairs, numContacts, numPenetrating; /* numPenetrating = EPA runs     */
    uint32_t gjkFailures, epaFailures;  /* contacts whose status is 301 / 302                   */
    uint32_t requiredPairs, requiredContacts; /* sizes that would have been needed (for 601)    */
    float refitMs, sortMs, buildMs, pairMs, pairSortMs, gjkMs, epaMs, totalMs;
    /* gui::PhysicsWorldStats sinks (include/axiom/gui/physics_panel.hpp:26-30)                 */
    float broadphaseTime, narrowphaseTime;
    uint64_t bytesMoved;    /* algorithmic HBM bytes of the step              */
    uint32_t kernelLaunches; /* kernels launched by the last refit+broadphase+narrowphase       */
    uint32_t contactPointCount; /* gui::PhysicsWorldStats::contactPointCount (physics_panel.hpp:21):
                                   manifold points of the last axcd_build_manifolds, else 0      */
    uint32_t movedBodies;       /* AXCD_FLAG_TEMPORAL_COHERENCE: bodies whose fat box was rebuilt by
                                   the last refit (numBodies without the flag)                   */
    uint32_t broadphaseSkipped; /* 1 if the last axcd_broadphase reused the cached candidate list */
    uint32_t graphLaunched;     /* 1 if the last fused step was one CUDA graph launch: totalMs is then the
                                   only timing (per-stage times need the staged calls)               */
    uint32_t ghostBodies;       /* axcd_slab_step: ghost bodies received from the other ranks this step  */
    uint32_t sortFallback;      /* 1 if the Morton bucket sort gave up (a bucket over its capacity: clustered scene)
                                   and the LSD radix kernels sorted this step                             */
    uint32_t sortMaxBucket;     /* largest Morton bucket of the step (capacity 1024)                        */
    float exchangeMs;           /* axcd_slab_step: owned refit + ghost selection + NCCL handshake and exchange
                                   + unpack, CUDA events on the context stream (includes the one host read) */
} AxcdStats;

typedef struct AxcdContext AxcdContext; /* opaque */

AXCD_API void axcd_default_config(AxcdConfig* cfg);
/* Number of CUDA devices visible to the process (0 without a driver / device): valid deviceOrdinal range. */
AXCD_API int32_t axcd_device_count(void);
AXCD_API int32_t axcd_create(const AxcdConfig* cfg, AxcdContext** out);
AXCD_API void axcd_destroy(AxcdContext* ctx);

/* Static scene description.  hullXYZ is xyz-packed (12 B per vertex); worldId may be NULL when
 * numWorlds == 1.  Validates shape types/ranges -> 300 / 600 / 601.                            */
AXCD_API int32_t axcd_set_shapes(AxcdContext* ctx, const AxcdShape* shapes, uint32_t n,
                                 const float* hullXYZ, uint32_t nHullVerts,
                                 const uint32_t* worldId);

/* Per-step poses.  `transforms` is an array of axiom::math::Transform (40 B: position 0,
 * rotation (x,y,z,w) 12, scale 28; include/axiom/math/transform.hpp:18-22) with the given byte
 * stride (>= 40).  n must equal the n of axcd_set_shapes.  Host -> device copy.                */
AXCD_API int32_t axcd_set_transforms(AxcdContext* ctx, const void* transforms, uint32_t n,
                                     uint32_t strideBytes);
/* Buffer lifetime: the copy is enqueued on the context stream and the call returns.  From pageable memory
 * the driver stages the data before returning; from page-locked memory (axcd_pin_host_buffer, recommended
 * for the per-step buffer) the DMA is truly asynchronous, so the caller must leave the buffer untouched
 * until the next blocking call on this context (axcd_step, axcd_get_stats, any axcd_get_*) has returned.
 * The same rule holds for axcd_set_ghosts and axcd_set_body_keys.                                    */

/* Per-step poses without the scales: `poses` is an array of 28-byte records (position 0, rotation (x,y,z,w) 12
 * — the first 28 bytes of a Transform) with the given byte stride (>= 28; 40 reads them out of a Transform
 * array).  A rigid body's scale does not change from step to step, so after one axcd_set_transforms a step
 * only has to upload 28 of the 40 bytes per body; the scales stay as the last axcd_set_transforms left them.
 * Returns 503 before the first axcd_set_transforms.  Same asynchrony and buffer-lifetime rule as above.      */
AXCD_API int32_t axcd_set_poses(AxcdContext* ctx, const void* poses, uint32_t n, uint32_t strideBytes);

/* The three stages (asynchronous on the context stream) and the fused step (synchronises and
 * fills stats).  Each stage runs once per refit, in order: axcd_broadphase needs an axcd_refit
 * since the last set_transforms, axcd_narrowphase needs a fresh axcd_broadphase; calling a stage
 * out of order or twice returns 503 (GPU invalid operation).                                   */
AXCD_API int32_t axcd_refit(AxcdContext* ctx);
AXCD_API int32_t axcd_broadphase(AxcdContext* ctx);
AXCD_API int32_t axcd_narrowphase(AxcdContext* ctx);
AXCD_API int32_t axcd_step(AxcdContext* ctx, AxcdStats* outStats);
/* The fused step without the final synchronisation: refit + broadphase + narrowphase are enqueued on the
 * context stream (as ONE CUDA graph launch once the launch configuration — body count, filters, slab rule
 * — has been stable for a step, otherwise as the individual kernels) and the call returns.  Counts and
 * results are fetched with axcd_get_stats / axcd_get_* as after axcd_step.  A graph-launched step reports
 * totalMs only; broadphaseTime / narrowphaseTime and the per-stage times are filled by the staged calls
 * axcd_refit / axcd_broadphase / axcd_narrowphase, which always launch directly.                    */
AXCD_API int32_t axcd_step_async(AxcdContext* ctx);
/* Waits for the stream, then reports counts/timings of the last stages run.  Returns 601 if a
 * capacity was exceeded (requiredPairs / requiredContacts say by how much).                    */
AXCD_API int32_t axcd_get_stats(AxcdContext* ctx, AxcdStats* outStats);

/* Blocking getters (device -> host copies).                                                    */
AXCD_API int32_t axcd_get_aabbs(AxcdContext* ctx, void* outAabb24, uint32_t cap);
AXCD_API int32_t axcd_get_pairs(AxcdContext* ctx, uint32_t* outPairs2, uint32_t cap,
                                uint32_t* outCount); /* canonical: a<b, sorted by (a,b)         */
AXCD_API int32_t axcd_get_pair_distances(AxcdContext* ctx, float* outDist, uint32_t cap,
                                         uint32_t* outCount); /* needs AXCD_FLAG_PAIR_DISTANCES */
AXCD_API int32_t axcd_get_contacts(AxcdContext* ctx, AxcdContact* out, uint32_t cap,
                                   uint32_t* outCount); /* same (a,b) order                     */

/* ---- contact manifolds -------------------------------------------------
 * One manifold per contact, in contact order: 1..4 points that share the contact's normal.  Each
 * point is one debug::DebugContactPoint (position, normal, penetrationDepth;
 * include/axiom/debug/physics_debug_draw.hpp:128-132) and the total is what
 * gui::PhysicsWorldStats::contactPointCount reports (include/axiom/gui/physics_panel.hpp:21).
 * Box-box contacts are clipped feature against feature (reference face / incident face, at most
 * four points kept); capsule-box contacts clip the capsule's segment against the facing box face (one
 * or two points); every other shape pair keeps the narrowphase point.  Unused slots are zero.       */
typedef struct AxcdManifold {
    uint32_t a, b;          /* == the contact's                                                  */
    float nx, ny, nz;       /* == the contact's normal (a -> b)                                  */
    uint32_t count;         /* 1..4                                                              */
    float px[4], py[4], pz[4];
    float depth[4];
} AxcdManifold;
/* Asynchronous on the context stream; needs axcd_narrowphase (or axcd_step) first; once per
 * narrowphase (503 otherwise).  The device buffer (88 B x maxContacts) is allocated on first use. */
AXCD_API int32_t axcd_build_manifolds(AxcdContext* ctx);
AXCD_API int32_t axcd_get_manifolds(AxcdContext* ctx, AxcdManifold* out, uint32_t cap,
                                    uint32_t* outCount, uint32_t* outPointCount /* or NULL */);

/* ---- scene queries on the LBVH of the last broadphase ("scene queries",
 * reference: CLAUDE.md:88).  Both need axcd_broadphase (or axcd_step) first and block until done.
 *
 * AABB overlap query: for each query box (axiom::math::AABB, 24 B) every body whose world AABB meets
 * it under AABB::intersects (closed intervals, include/axiom/math/aabb.hpp:132-135).  outHits2
 * receives (query, body) index pairs sorted by (query, body).  queryWorld (one world id per query)
 * restricts each query to one world in batched mode; NULL = all worlds.  Returns 601 when cap is too
 * small; *outCount then says how many pairs there are.                                             */
AXCD_API int32_t axcd_query_aabbs(AxcdContext* ctx, const float* boxes6, const uint32_t* queryWorld,
                                  uint32_t nq, uint32_t* outHits2, uint32_t cap, uint32_t* outCount);

/* Closest-hit ray cast.  A body is hit iff the ray passes the slab test of its AABB within [0, tMax]
 * and its shape test does: spheres, oriented boxes and capsules in closed form; convex hulls by
 * conservative advancement on the GJK distance (t = first parameter at which the ray point is within
 * 1e-4 of the hull).  t is the parameter along (dx,dy,dz), which need not be unit length; the normal is
 * the unit surface normal at the hit (zero when the origin is inside the shape, t = 0).  No hit:
 * body = 0xffffffff, t = tMax.  Ties go to the lower body index.  `world` selects the world in batched
 * mode (ignored when numWorlds == 1).  flags is reserved (0).                                      */
typedef struct AxcdRay {
    float ox, oy, oz;
    float dx, dy, dz;
    float tMax;
    uint32_t world;
} AxcdRay;
typedef struct AxcdRayHit {
    uint32_t body;
    float t;
    float nx, ny, nz;
    uint32_t flags;
} AxcdRayHit;
AXCD_API int32_t axcd_raycast(AxcdContext* ctx, const AxcdRay* rays, uint32_t nq, AxcdRayHit* outHits);

/* GJK-based continuous collision detection ("GJK-based CCD", reference: CLAUDE.md:136): time of impact of
 * the given body pairs under LINEAR motion over one step.  displacement3 holds one (dx,dy,dz) per body
 * (n x 3 floats: velocity * dt); rotations stay fixed during the sweep.  Conservative advancement on the
 * exact GJK distance: toi in [0,1] is the first parameter at which the surfaces come within 1e-4 of each
 * other (hit = 1), normal points from a to b along the closest direction (zero if the pair already overlaps at t = 0);
 * hit = 0, toi = 1 when they do not meet during the step.  Needs shapes and transforms only.  Blocking. */
typedef struct AxcdSweep {
    uint32_t hit;
    float toi;
    float nx, ny, nz;
    uint32_t iterations;   /* conservative-advancement steps taken */
} AxcdSweep;
AXCD_API int32_t axcd_ccd_pairs(AxcdContext* ctx, const uint32_t* pairs2, uint32_t npairs,
                                const float* displacement3, AxcdSweep* out);

/* The same with rotation: rotation3 holds one rotation vector per body (n x 3 floats: angular velocity * dt,
 * radians, world frame).  Motion model: position p + displacement * t, orientation
 *     q(t) = normalize(q + t * 0.5 * (w, 0) (x) q)
 * — the first-order quaternion integration engines step with, so t = 1 is exactly the pose an integrator using it
 * reaches (and no sin / cos: the CUDA path and the CPU oracle agree bit for bit).  Conservative advancement: the
 * turning rate of that path never exceeds |w|, so gap / (linear approach + |wA| reachA + |wB| reachB), with reach =
 * the largest distance of a core point from its body's position, never oversteps the first contact; at most 64
 * steps, then the time reached is reported as a (conservative) hit.  Same result record and error behaviour as
 * axcd_ccd_pairs. */
AXCD_API int32_t axcd_ccd_pairs_angular(AxcdContext* ctx, const uint32_t* pairs2, uint32_t npairs,
                                        const float* displacement3, const float* rotation3, AxcdSweep* out);

/* Contact sink: a page-locked host buffer (axcd_pin_host_buffer, cudaHostAlloc, ...) of at least maxContacts
 * records that every narrowphase from now on fills with the step's contacts — the records axcd_get_contacts
 * returns, in the same order — so that the host copy travels while the narrowphase is still running instead of
 * in a separate copy after it.  Scenes decided in closed form: the narrowphase kernel reports, tile by tile, how
 * far the device array is complete, and the next blocking call on the context follows that and hands every
 * finished stretch to the copy engine on a second stream (DMA bursts behind the kernel); other scenes: one copy
 * kernel at the end of the narrowphase.  The sink is valid once a blocking call (axcd_step, axcd_get_stats,
 * axcd_get_contacts, ...) has returned; the count is AxcdStats::numContacts.  A new step on the same context
 * first completes the previous step's sink.  NULL detaches it.  600 if the buffer is not page-locked or too
 * small.  The device array stays complete: axcd_get_contacts, manifolds etc. work as before.                   */
AXCD_API int32_t axcd_set_contact_sink(AxcdContext* ctx, AxcdContact* hostBuffer, uint32_t capacity);

/* Collision filtering, gui::FilterInfo semantics (include/axiom/gui/body_inspector.hpp:38-42):
 * two bodies with the same non-zero groupIndex collide iff it is positive; otherwise both
 * (maskBits & other.categoryBits) must be non-zero.  Applied when candidate pairs are emitted.
 * AxcdFilter has the memory layout of gui::FilterInfo (uint32, uint32, int16 + 2 bytes of padding = 12
 * bytes), so an engine-side FilterInfo array passes through unchanged.  filters = one record per body
 * (n == the body count of axcd_set_shapes, 600 otherwise), or NULL to switch filtering off (the default;
 * axcd_set_shapes also switches it off).  Not available in x-slab mode (ghost records carry no filter).  */
typedef struct AxcdFilter {
    uint32_t categoryBits;
    uint32_t maskBits;
    int16_t groupIndex;
    uint16_t reserved_;      /* padding of gui::FilterInfo; ignored */
} AxcdFilter;
AXCD_API int32_t axcd_set_filters(AxcdContext* ctx, const AxcdFilter* filters, uint32_t n);

/* Sleeping bodies (debug::DebugRigidBody::isAwake, include/axiom/debug/physics_debug_draw.hpp:123;
 * gui::SleepInfo, include/axiom/gui/body_inspector.hpp:45-51): awake[i] == 0 marks body i asleep; a
 * pair of two sleeping bodies is not a candidate (applied where candidate pairs are emitted).  NULL
 * switches the rule off (all awake, the default).                                                 */
AXCD_API int32_t axcd_set_awake(AxcdContext* ctx, const uint8_t* awake, uint32_t n);   /* n == owned body count */

/* ---- one huge scene across several GPUs: x-slab mode ---------
 * Bodies [0, nOwned) are the ones this rank owns (set with axcd_set_shapes / axcd_set_transforms as
 * usual, keys = their global ids).  Every step the caller appends the ghost bodies received from the
 * other ranks with axcd_set_ghosts, after which the context holds nOwned + nGhosts bodies.  With the
 * slab rule enabled a candidate pair is kept only if max(min_a.x, min_b.x) lies in [xLo, xHi) — exactly
 * one rank keeps each pair — and pairs are oriented by key (key[a] < key[b]) instead of by local
 * index, so each contact is computed exactly as a single-GPU run would.  With these building-block calls hull
 * shapes are not accepted as ghosts (-> 300); axcd_slab_step ships a hull ghost's vertices too (capacity:
 * maxHullVerts beyond the owned vertices).                                                          */
AXCD_API int32_t axcd_set_slab(AxcdContext* ctx, float xLo, float xHi, uint32_t enable);
AXCD_API int32_t axcd_set_body_keys(AxcdContext* ctx, const uint32_t* keys, uint32_t first,
                                    uint32_t count);
/* Global ids (keys) of the bodies currently held: the owned ones, then this step's ghosts.  Blocking.    */
AXCD_API int32_t axcd_get_body_keys(AxcdContext* ctx, uint32_t* outKeys, uint32_t cap);
AXCD_API int32_t axcd_set_ghosts(AxcdContext* ctx, uint32_t nOwned, uint32_t nGhosts,
                                 const void* transforms40, const AxcdShape* shapes,
                                 const uint32_t* keys);

/* Device-side variant of the ghost hand-off (no host copies).  axcd_pack_ghosts (after axcd_refit):
 * for every destination rank r != myRank selects the owned bodies whose x-interval meets slab
 * [edges[r], edges[r+1]) and packs them as 64-byte records into a device buffer; outDevPtrs[r] /
 * outCounts[r] (host arrays of numRanks entries) receive the DEVICE pointer and the record count
 * (outDevPtrs[myRank] = NULL).  The caller moves the records rank to rank (NCCL) and appends what it
 * received — one contiguous device buffer of nGhosts records — with axcd_set_ghosts_device.
 * Returns 601 if a send buffer was too small (capacity: maxBodies - nOwned records each).       */
AXCD_API int32_t axcd_pack_ghosts(AxcdContext* ctx, const float* edges, uint32_t numRanks,
                                  uint32_t myRank, void** outDevPtrs, uint32_t* outCounts);
AXCD_API int32_t axcd_set_ghosts_device(AxcdContext* ctx, uint32_t nOwned, uint32_t nGhosts,
                                        const void* devRecords);

/* The whole exchange inside the library, over NCCL (ncclAllGather of the per-destination counts, then one
 * grouped ncclSend / ncclRecv of the ghost records, GPU to GPU): no host staging, no framework in between.
 * NCCL is bound at run time (dlopen libnccl.so.2): without it these calls return 500, everything else works.
 *   axcd_nccl_unique_id : rank 0 creates the 128-byte ncclUniqueId; the host broadcasts it by any means
 *   axcd_slab_init      : every rank, after axcd_set_shapes / axcd_set_transforms / axcd_set_body_keys of its
 *                         OWNED bodies: creates the communicator on the context's device (collective call),
 *                         allocates the send / receive buffers (capacity maxBodies - nOwned ghost records) and
 *                         switches the slab rule on for [edges[rank], edges[rank+1]); edges has numRanks+1 entries
 *   axcd_slab_init_comm : the same with a communicator the host already owns (ncclComm_t passed as void*)
 *   axcd_slab_step      : one step of the sharded scene — refit of the owned bodies, device-side ghost
 *                         selection, size handshake, exchange, append, then the fused step on owned + ghosts.
 *                         Collective: every rank of the communicator must call it.  _async returns after the
 *                         handshake's single host read, with the step enqueued.                          */
AXCD_API int32_t axcd_nccl_unique_id(void* out128);
AXCD_API int32_t axcd_slab_init(AxcdContext* ctx, const void* uniqueId128, uint32_t rank, uint32_t numRanks,
                                const float* edges);
AXCD_API int32_t axcd_slab_init_comm(AxcdContext* ctx, void* ncclComm, uint32_t rank, uint32_t numRanks,
                                     const float* edges);
AXCD_API int32_t axcd_slab_step_async(AxcdContext* ctx);
AXCD_API int32_t axcd_slab_step(AxcdContext* ctx, AxcdStats* outStats);

/* Page-locks (and later releases) a caller-owned host buffer so that the per-step copies — transforms in,
 * contacts / manifolds out — run at full PCIe rate; the engine does not need the CUDA headers for it.
 * Optional: every entry point also accepts pageable memory (about half the copy bandwidth).       */
AXCD_API int32_t axcd_pin_host_buffer(void* hostPtr, uint64_t bytes);
AXCD_API int32_t axcd_unpin_host_buffer(void* hostPtr);

/* == axiom::core::errorCodeToString (src/core/error_code.cpp:5-62); static storage.            */
AXCD_API const char* axcd_error_string(int32_t code);
/* CUDA error text of the last failure on this context (static storage), "" if none.           */
AXCD_API const char* axcd_last_device_error(AxcdContext* ctx);

/* ---- self-test hooks for the device primitives (tests/ only; not part of the plug-in path) -- */
/* Sorts n (key,value) pairs by the low `keyBits` bits of key with the hand-written radix sort. */
AXCD_API int32_t axcd_test_sort_pairs32(AxcdContext* ctx, uint32_t* keys, uint32_t* vals,
                                        uint32_t n, uint32_t keyBits);
AXCD_API int32_t axcd_test_sort_keys64(AxcdContext* ctx, uint64_t* keys, uint32_t n,
                                       uint32_t keyBits);
/* Device-resident timing of the (key,value) radix sort on n pseudo-random keys: average ms of
 * `iters` sorts (CUDA events on the context stream, inputs regenerated on the device each time). */
AXCD_API int32_t axcd_test_sort_bench(AxcdContext* ctx, uint32_t n, uint32_t keyBits, uint32_t iters,
                                      float* outMsPerSort);
/* The step's own Morton sort (identity payload) on caller keys: mode 1 = bucket sort with the LSD fallback armed,
 * mode 0 = LSD only.  *outFallback = 1 if the LSD kernels produced the result.  _bench_: device-resident timing
 * on n pseudo-random keys, average ms per sort.                                                      */
AXCD_API int32_t axcd_test_sort_morton(AxcdContext* ctx, const uint32_t* keys, uint32_t n, uint32_t keyBits,
                                       uint32_t mode, uint32_t* outKeys, uint32_t* outVals, uint32_t* outFallback);
AXCD_API int32_t axcd_test_sort_bench_morton(AxcdContext* ctx, uint32_t n, uint32_t keyBits, uint32_t iters,
                                             uint32_t mode, float* outMsPerSort);
/* Measured FP32 peak of the device: an FMA-chain kernel (8 independent chains per thread, `iters` x 128
 * FMAs per thread, 2048 threads per SM), best of three, in TFLOP/s.  The roof the GJK / EPA kernels are
 * reported against (BASELINE.md asks for a peak measured on the box, not the nominal one).        */
AXCD_API int32_t axcd_test_fp32_peak(AxcdContext* ctx, uint32_t iters, float* outTflops);

#ifdef __cplusplus
}
#endif
#endif /* AXCD_H */
