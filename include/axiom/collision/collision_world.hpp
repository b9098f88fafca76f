// C++20 host façade over the C ABI (include/axcd.h) for Axiom's empty `axiom::collision` slot.
//
// Mirrors the only usage the reference documents for this path (CLAUDE.md:162-178,
// include/axiom/core/profiler.hpp:13-23):
//
//     { AXIOM_PROFILE_SCOPE("Broadphase");  broadphase_.update();
//       AXIOM_PROFILE_VALUE("BroadphasePairs", broadphase_.getPairCount()); }
//     { AXIOM_PROFILE_SCOPE("Narrowphase"); narrowphase_.detectCollisions();
//       AXIOM_PROFILE_VALUE("ContactCount", narrowphase_.getContactCount()); }
//
// and the reference's conventions: factories return Result<std::unique_ptr<T>> with private
// constructors (include/axiom/gpu/vk_instance.hpp:34,95), owners are non-copyable / non-movable
// (vk_instance.hpp:40-43), fallible calls return Result<void> with static-lifetime messages
// (include/axiom/core/result.hpp:32), no exceptions (CLAUDE.md:139), camelCase methods and
// member_ suffix (CLAUDE.md:113-121).  Inside the Axiom tree the engine's own headers are used;
// standalone, minimal layout-identical stand-ins are defined so this header compiles on its own.
#pragma once

#include "axcd.h"

#include <cstddef>
#include <cstdint>
#include <memory>
#include <utility>

#if __has_include("axiom/core/result.hpp") && __has_include("axiom/math/transform.hpp")
#include "axiom/core/error_code.hpp"
#include "axiom/core/result.hpp"
#include "axiom/math/aabb.hpp"
#include "axiom/math/transform.hpp"
#define AXIOM_COLLISION_HAS_ENGINE_TYPES 1
#if __has_include("axiom/gui/body_inspector.hpp") && defined(AXIOM_COLLISION_WITH_GUI_TYPES)
#include "axiom/gui/body_inspector.hpp"   // gui::FilterInfo (pulls in the GUI headers: opt-in)
#define AXIOM_COLLISION_HAS_GUI_TYPES 1
#endif
#else
namespace axiom::core {
enum class ErrorCode : int {   // numeric values of include/axiom/core/error_code.hpp:21-57
    Success = 0, OutOfMemory = 200, NullPointer = 202, InvalidShape = 300, GJKFailedToConverge = 301,
    EPAFailedToConverge = 302, VulkanInitializationFailed = 500, BufferAllocationFailed = 502,
    GPU_INVALID_OPERATION = 503, GPU_OPERATION_FAILED = 505, InvalidParameter = 600, OutOfRange = 601
};
template <typename T>
class Result {   // the subset of include/axiom/core/result.hpp:15-230 this façade needs
public:
    static Result success(T v) { Result r; r.ok_ = true; r.value_ = std::move(v); return r; }
    static Result failure(ErrorCode c, const char* m) { Result r; r.code_ = c; r.message_ = m; return r; }
    bool isSuccess() const noexcept { return ok_; }
    bool isFailure() const noexcept { return !ok_; }
    T& value() noexcept { return value_; }
    ErrorCode errorCode() const noexcept { return code_; }
    const char* errorMessage() const noexcept { return message_; }
private:
    bool ok_ = false;
    T value_{};
    ErrorCode code_ = ErrorCode::Success;
    const char* message_ = "";
};
template <>
class Result<void> {
public:
    static Result success() { Result r; r.ok_ = true; return r; }
    static Result failure(ErrorCode c, const char* m) { Result r; r.code_ = c; r.message_ = m; return r; }
    bool isSuccess() const noexcept { return ok_; }
    bool isFailure() const noexcept { return !ok_; }
    ErrorCode errorCode() const noexcept { return code_; }
    const char* errorMessage() const noexcept { return message_; }
private:
    bool ok_ = false;
    ErrorCode code_ = ErrorCode::Success;
    const char* message_ = "";
};
}  // namespace axiom::core
namespace axiom::math {
struct Vec3 { float x, y, z; };
struct Quat { float x, y, z, w; };
struct Transform { Vec3 position; Quat rotation; Vec3 scale; };   // 40 B: transform.hpp:18-22
struct AABB { Vec3 min, max; };                                   // 24 B: aabb.hpp:18-21
}  // namespace axiom::math
#endif

namespace axiom::collision {

static_assert(sizeof(math::Transform) == 40, "Transform must be the reference's 40-byte record");
static_assert(sizeof(math::AABB) == 24, "AABB must be the reference's 24-byte record");

using Shape = AxcdShape;            // flattened debug::DebugShape (physics_debug_draw.hpp:97-112)
using ContactPoint = AxcdContact;   // debug::DebugContactPoint + pair ids (physics_debug_draw.hpp:128-132)
using ContactManifold = AxcdManifold;   // 1..4 DebugContactPoints sharing one normal
using Ray = AxcdRay;
using RayHit = AxcdRayHit;
using Sweep = AxcdSweep;
/// gui::FilterInfo's layout (include/axiom/gui/body_inspector.hpp:38-42): uint32 categoryBits, uint32 maskBits,
/// int16 groupIndex (+ 2 bytes of padding).  A FilterInfo array can be passed as-is.
using Filter = AxcdFilter;
static_assert(sizeof(Filter) == 12, "Filter must have gui::FilterInfo's 12-byte layout");
#ifdef AXIOM_COLLISION_HAS_GUI_TYPES
static_assert(sizeof(gui::FilterInfo) == sizeof(Filter) && offsetof(gui::FilterInfo, groupIndex) == offsetof(Filter, groupIndex),
              "gui::FilterInfo layout changed");
#endif
struct BodyPair { std::uint32_t a, b; };
struct QueryHit { std::uint32_t query, body; };

struct CollisionConfig : AxcdConfig {
    CollisionConfig() { axcd_default_config(this); }
};

/// Page-locks a caller-owned buffer for the lifetime of this object (full-rate host <-> device copies).
class PinnedRegion {
public:
    PinnedRegion(void* ptr, std::size_t bytes) : ptr_(axcd_pin_host_buffer(ptr, bytes) == AXCD_OK ? ptr : nullptr) {}
    ~PinnedRegion() { if (ptr_) axcd_unpin_host_buffer(ptr_); }
    PinnedRegion(const PinnedRegion&) = delete;
    PinnedRegion& operator=(const PinnedRegion&) = delete;
    bool pinned() const noexcept { return ptr_ != nullptr; }
private:
    void* ptr_;
};

/// Owns one device context (one GPU, one stream).  Not thread-safe, like the reference's
/// device-facing objects (include/axiom/gpu/vk_command.hpp:15).
class CollisionWorld {
public:
    static core::Result<std::unique_ptr<CollisionWorld>> create(const CollisionConfig& config) {
        AxcdContext* ctx = nullptr;
        const std::int32_t rc = axcd_create(&config, &ctx);
        if (rc != AXCD_OK) return fail<std::unique_ptr<CollisionWorld>>(rc);
        return core::Result<std::unique_ptr<CollisionWorld>>::success(
            std::unique_ptr<CollisionWorld>(new CollisionWorld(ctx)));
    }
    ~CollisionWorld() { axcd_destroy(ctx_); }
    CollisionWorld(const CollisionWorld&) = delete;
    CollisionWorld& operator=(const CollisionWorld&) = delete;
    CollisionWorld(CollisionWorld&&) = delete;
    CollisionWorld& operator=(CollisionWorld&&) = delete;

    /// shapes/hull/worldId must stay valid for the duration of the call only (copied to the device)
    core::Result<void> setShapes(const Shape* shapes, std::uint32_t count, const float* hullXYZ = nullptr,
                                 std::uint32_t hullVertexCount = 0, const std::uint32_t* worldId = nullptr) {
        bodyCount_ = count;
        return wrap(axcd_set_shapes(ctx_, shapes, count, hullXYZ, hullVertexCount, worldId));
    }
    core::Result<void> setTransforms(const math::Transform* transforms, std::uint32_t count) {
        return wrap(axcd_set_transforms(ctx_, transforms, count, sizeof(math::Transform)));
    }
    /// Position + rotation only (28-byte records at the given stride; the scales of the last setTransforms stay):
    /// what a step has to upload once the bodies' scales are on the device.  setPoses(transforms, n) reads the
    /// first 28 bytes of each Transform.
    core::Result<void> setPoses(const void* poses, std::uint32_t count, std::uint32_t strideBytes = 28) {
        return wrap(axcd_set_poses(ctx_, poses, count, strideBytes));
    }
    core::Result<void> setPoses(const math::Transform* transforms, std::uint32_t count) {
        return wrap(axcd_set_poses(ctx_, transforms, count, sizeof(math::Transform)));
    }
    /// A page-locked host buffer (>= maxContacts records) that every narrowphase from now on also fills with the
    /// step's contacts, so the host copy travels while the narrowphase runs; nullptr detaches.  Count: stats().
    core::Result<void> setContactSink(ContactPoint* pinnedHostBuffer, std::uint32_t capacity) {
        return wrap(axcd_set_contact_sink(ctx_, pinnedHostBuffer, capacity));
    }
    core::Result<void> refit() { return wrap(axcd_refit(ctx_)); }
    core::Result<void> broadphase() { return wrap(axcd_broadphase(ctx_)); }
    core::Result<void> narrowphase() { return wrap(axcd_narrowphase(ctx_)); }
    core::Result<AxcdStats> step() {
        AxcdStats s{};
        const std::int32_t rc = axcd_step(ctx_, &s);
        if (rc != AXCD_OK) return fail<AxcdStats>(rc);
        return core::Result<AxcdStats>::success(s);
    }
    core::Result<AxcdStats> stats() {
        AxcdStats s{};
        const std::int32_t rc = axcd_get_stats(ctx_, &s);
        if (rc != AXCD_OK) return fail<AxcdStats>(rc);
        return core::Result<AxcdStats>::success(s);
    }
    core::Result<void> getAABBs(math::AABB* out, std::uint32_t capacity) {
        return wrap(axcd_get_aabbs(ctx_, out, capacity));
    }
    core::Result<std::uint32_t> getPairs(BodyPair* out, std::uint32_t capacity) {
        std::uint32_t n = 0;
        const std::int32_t rc = axcd_get_pairs(ctx_, reinterpret_cast<std::uint32_t*>(out), capacity, &n);
        if (rc != AXCD_OK) return fail<std::uint32_t>(rc);
        return core::Result<std::uint32_t>::success(n);
    }
    core::Result<std::uint32_t> getContacts(ContactPoint* out, std::uint32_t capacity) {
        std::uint32_t n = 0;
        const std::int32_t rc = axcd_get_contacts(ctx_, out, capacity, &n);
        if (rc != AXCD_OK) return fail<std::uint32_t>(rc);
        return core::Result<std::uint32_t>::success(n);
    }
    /// Contact manifolds of the last narrowphase (box-box feature clipping, 1..4 points each);
    /// stats().contactPointCount then reports gui::PhysicsWorldStats::contactPointCount.
    core::Result<void> buildManifolds() { return wrap(axcd_build_manifolds(ctx_)); }
    core::Result<std::uint32_t> getManifolds(ContactManifold* out, std::uint32_t capacity,
                                             std::uint32_t* outPointCount = nullptr) {
        std::uint32_t n = 0;
        const std::int32_t rc = axcd_get_manifolds(ctx_, out, capacity, &n, outPointCount);
        if (rc != AXCD_OK) return fail<std::uint32_t>(rc);
        return core::Result<std::uint32_t>::success(n);
    }
    /// Scene queries on the LBVH of the last broadphase.  queryAABBs: (query, body) hits sorted by
    /// (query, body); on OutOfRange the value needed is written to *required.
    core::Result<std::uint32_t> queryAABBs(const math::AABB* boxes, std::uint32_t count, QueryHit* out,
                                           std::uint32_t capacity, const std::uint32_t* queryWorld = nullptr,
                                           std::uint32_t* required = nullptr) {
        std::uint32_t n = 0;
        const std::int32_t rc = axcd_query_aabbs(ctx_, reinterpret_cast<const float*>(boxes), queryWorld, count,
                                                 reinterpret_cast<std::uint32_t*>(out), capacity, &n);
        if (required) *required = n;
        if (rc != AXCD_OK) return fail<std::uint32_t>(rc);
        return core::Result<std::uint32_t>::success(n);
    }
    core::Result<void> rayCast(const Ray* rays, std::uint32_t count, RayHit* outHits) {
        return wrap(axcd_raycast(ctx_, rays, count, outHits));
    }
    /// GJK-based CCD: time of impact of body pairs under linear motion (displacements: bodyCount() x 3 floats).
    core::Result<void> sweepPairs(const BodyPair* pairs, std::uint32_t count, const math::Vec3* displacements,
                                  Sweep* out) {
        return wrap(axcd_ccd_pairs(ctx_, reinterpret_cast<const std::uint32_t*>(pairs), count,
                                   reinterpret_cast<const float*>(displacements), out));
    }
    /// The same with rotation: one rotation vector (angular velocity * dt) per body; see axcd_ccd_pairs_angular.
    core::Result<void> sweepPairs(const BodyPair* pairs, std::uint32_t count, const math::Vec3* displacements,
                                  const math::Vec3* rotations, Sweep* out) {
        return wrap(axcd_ccd_pairs_angular(ctx_, reinterpret_cast<const std::uint32_t*>(pairs), count,
                                           reinterpret_cast<const float*>(displacements),
                                           reinterpret_cast<const float*>(rotations), out));
    }
    /// The fused step without the final synchronisation (one CUDA graph launch once the launch
    /// configuration is stable); fetch counts with stats().
    core::Result<void> stepAsync() { return wrap(axcd_step_async(ctx_)); }

    /// Collision filtering with gui::FilterInfo semantics, one record per body; nullptr switches it off.
    core::Result<void> setFilters(const Filter* filters, std::uint32_t count) {
        return wrap(axcd_set_filters(ctx_, filters, count));
    }
#ifdef AXIOM_COLLISION_HAS_GUI_TYPES
    core::Result<void> setFilters(const gui::FilterInfo* filters, std::uint32_t count) {
        return wrap(axcd_set_filters(ctx_, reinterpret_cast<const Filter*>(filters), count));   // same layout (asserted above)
    }
#endif
    /// Sleeping bodies (debug::DebugRigidBody::isAwake): awake[i] == 0 marks body i asleep; pairs of two
    /// sleeping bodies are dropped.  nullptr switches the rule off.
    core::Result<void> setAwake(const std::uint8_t* awake, std::uint32_t count) {
        return wrap(axcd_set_awake(ctx_, awake, count));
    }

    /// One huge scene across several GPUs (x-slabs).  Call after setShapes / setTransforms of the OWNED bodies.
    /// globalIds[i] = id of owned body i in the whole scene; edges has numRanks + 1 entries.  initSlab with
    /// a unique id creates the NCCL communicator inside the library (collective over all ranks);
    /// initSlabWithComm takes an ncclComm_t the host already owns.  slabStep is collective as well.
    static core::Result<void> ncclUniqueId(void* out128) { return wrap(axcd_nccl_unique_id(out128)); }
    core::Result<void> setBodyKeys(const std::uint32_t* globalIds, std::uint32_t count) {
        return wrap(axcd_set_body_keys(ctx_, globalIds, 0, count));
    }
    core::Result<void> initSlab(const void* ncclUniqueId128, std::uint32_t rank, std::uint32_t numRanks, const float* edges) {
        return wrap(axcd_slab_init(ctx_, ncclUniqueId128, rank, numRanks, edges));
    }
    core::Result<void> initSlabWithComm(void* ncclComm, std::uint32_t rank, std::uint32_t numRanks, const float* edges) {
        return wrap(axcd_slab_init_comm(ctx_, ncclComm, rank, numRanks, edges));
    }
    core::Result<AxcdStats> slabStep() {
        AxcdStats s{};
        const std::int32_t rc = axcd_slab_step(ctx_, &s);
        if (rc != AXCD_OK) return fail<AxcdStats>(rc);
        return core::Result<AxcdStats>::success(s);
    }
    /// Global ids of the bodies currently held (owned, then this step's ghosts): pair / contact indices are
    /// local; map them through this to report in scene ids.
    core::Result<void> getBodyKeys(std::uint32_t* out, std::uint32_t capacity) {
        return wrap(axcd_get_body_keys(ctx_, out, capacity));
    }
    std::uint32_t bodyCount() const noexcept { return bodyCount_; }

private:
    explicit CollisionWorld(AxcdContext* ctx) : ctx_(ctx) {}
    template <typename T>
    static core::Result<T> fail(std::int32_t rc) {
        return core::Result<T>::failure(static_cast<core::ErrorCode>(rc), axcd_error_string(rc));
    }
    static core::Result<void> wrap(std::int32_t rc) {
        if (rc != AXCD_OK) return fail<void>(rc);
        return core::Result<void>::success();
    }
    AxcdContext* ctx_;
    std::uint32_t bodyCount_ = 0;
};

/// `broadphase_.update(); broadphase_.getPairCount();`  — the world must outlive this object.
class Broadphase {
public:
    explicit Broadphase(CollisionWorld& world) : world_(&world) {}
    core::Result<void> update() {
        auto r = world_->refit();
        if (r.isFailure()) return r;
        return world_->broadphase();
    }
    std::uint32_t getPairCount() {
        auto s = world_->stats();
        return s.isSuccess() ? s.value().numPairs : 0u;
    }
private:
    CollisionWorld* world_;
};

/// `narrowphase_.detectCollisions(); narrowphase_.getContactCount();`
class Narrowphase {
public:
    explicit Narrowphase(CollisionWorld& world) : world_(&world) {}
    core::Result<void> detectCollisions() { return world_->narrowphase(); }
    std::uint32_t getContactCount() {
        auto s = world_->stats();
        return s.isSuccess() ? s.value().numContacts : 0u;
    }
private:
    CollisionWorld* world_;
};

}  // namespace axiom::collision
