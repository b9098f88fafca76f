// Drives the C++ façade (include/axiom/collision/collision_world.hpp) the way the reference's
// PhysicsWorld::step would (CLAUDE.md:162-178).  Prints "pairs contacts"; exit code 0 on success,
// 77 when no CUDA device is present (so the CPU-only test can tell "linked fine" from "failed").
#include "axiom/collision/collision_world.hpp"
#include "axcd_scene.h"

#include <cstdio>
#include <vector>

using namespace axiom;

#ifdef AXIOM_EXPECT_ENGINE_TYPES   // set by the test that compiles this file against the Axiom tree's own headers
#ifndef AXIOM_COLLISION_HAS_ENGINE_TYPES
#error "the facade did not pick up axiom/core/result.hpp and axiom/math/transform.hpp"
#endif
static_assert(sizeof(math::Transform) == 40 && sizeof(math::AABB) == 24, "engine math types at the boundary");
#endif

int main() {
    const std::uint32_t n = 1000;   // config C0: seed 1, L = 10, 50/50 boxes and spheres
    AxcdSceneSpec spec{n, 0.5f, 0.5f, 10.0f, 0.25f, 0.5f, 16, 1};
    std::vector<math::Transform> xf(n);
    std::vector<collision::Shape> shapes(n);
    std::uint32_t hullUsed = 0;
    if (axcd_scene_generate(&spec, reinterpret_cast<float*>(xf.data()), shapes.data(), nullptr, 0, 0, &hullUsed) != 0)
        return 2;

    collision::CollisionConfig cfg;
    cfg.maxBodies = n;
    cfg.maxPairs = 16 * n;
    cfg.maxContacts = 16 * n;
    auto created = collision::CollisionWorld::create(cfg);
    if (created.isFailure()) {
        std::printf("create failed: %d %s\n", static_cast<int>(created.errorCode()), created.errorMessage());
        return created.errorCode() == core::ErrorCode::VulkanInitializationFailed ? 77 : 3;
    }
    auto& world = *created.value();
    if (world.setShapes(shapes.data(), n).isFailure()) return 4;
    if (world.setTransforms(xf.data(), n).isFailure()) return 5;
    collision::Broadphase broadphase(world);
    collision::Narrowphase narrowphase(world);
    if (broadphase.update().isFailure()) return 6;
    const std::uint32_t pairs = broadphase.getPairCount();
    if (narrowphase.detectCollisions().isFailure()) return 7;
    const std::uint32_t contacts = narrowphase.getContactCount();
    std::vector<collision::ContactPoint> out(contacts ? contacts : 1);
    auto got = world.getContacts(out.data(), static_cast<std::uint32_t>(out.size()));
    if (got.isFailure() || got.value() != contacts) return 8;
    for (std::uint32_t k = 0; k < contacts; ++k)
        if (!(out[k].a < out[k].b) || !(out[k].depth >= 0.0f)) return 9;
    // the "next" rows through the same façade: manifolds and scene queries
    if (world.buildManifolds().isFailure()) return 10;
    std::vector<collision::ContactManifold> man(contacts ? contacts : 1);
    std::uint32_t points = 0;
    auto gm = world.getManifolds(man.data(), static_cast<std::uint32_t>(man.size()), &points);
    if (gm.isFailure() || gm.value() != contacts || points < contacts) return 11;
    auto st = world.stats();
    if (st.isFailure() || st.value().contactPointCount != points) return 12;
    math::AABB everything{{-1e9f, -1e9f, -1e9f}, {1e9f, 1e9f, 1e9f}};
    std::vector<collision::QueryHit> hits(n);
    auto qh = world.queryAABBs(&everything, 1, hits.data(), n);
    if (qh.isFailure() || qh.value() != n) return 13;
    collision::Ray ray{xf[0].position.x, xf[0].position.y, xf[0].position.z - 50.0f, 0.0f, 0.0f, 1.0f, 100.0f, 0};
    collision::RayHit rh{};
    if (world.rayCast(&ray, 1, &rh).isFailure() || rh.body == 0xffffffffu) return 14;
    // filters (gui::FilterInfo layout), sleeping flags and the asynchronous fused step
    std::vector<collision::Filter> filt(n);
    for (std::uint32_t i = 0; i < n; ++i) filt[i] = collision::Filter{1u, 0xFFFFu, static_cast<std::int16_t>(-1), 0};
    if (world.setFilters(filt.data(), n).isFailure()) return 15;    // everybody in one negative group: nothing collides
    if (world.stepAsync().isFailure()) return 16;
    auto fs = world.stats();
    if (fs.isFailure() || fs.value().numPairs != 0) return 17;
    if (world.setFilters(nullptr, 0).isFailure()) return 18;
    std::vector<std::uint8_t> awake(n, 0);
    if (world.setAwake(awake.data(), n).isFailure()) return 19;     // everybody asleep: no pair has an awake body
    auto as = world.step();
    if (as.isFailure() || as.value().numPairs != 0) return 20;
    if (world.setAwake(nullptr, 0).isFailure()) return 21;
    auto again = world.step();
    if (again.isFailure() || again.value().numPairs != pairs || again.value().numContacts != contacts) return 22;
    // position + rotation only (28 of the 40 bytes of each Transform): the same scene again
    if (world.setPoses(xf.data(), n).isFailure()) return 23;
    auto posed = world.step();
    if (posed.isFailure() || posed.value().numPairs != pairs || posed.value().numContacts != contacts) return 24;
    std::printf("%u %u %u\n", pairs, contacts, points);
    return 0;
}
