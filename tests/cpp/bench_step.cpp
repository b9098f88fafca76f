// C++20 host driving the collision path through the façade at benchmark size, the way an Axiom
// PhysicsWorld::step would (CLAUDE.md:162-178): setTransforms (host -> device), Broadphase::update,
// Narrowphase::detectCollisions, getContacts (device -> host).  Prints one line:
//     bodies pairs contacts ms_per_step_device ms_per_step_end_to_end
// usage: bench_step [bodies=1000000] [domain=100] [seed=3] [steps=20] [pin]
// exit code 77 when no CUDA device is present.
#include "axiom/collision/collision_world.hpp"
#include "axcd_scene.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

using namespace axiom;

int main(int argc, char** argv) {
    const std::uint32_t n = argc > 1 ? static_cast<std::uint32_t>(std::atoll(argv[1])) : 1000000u;
    const float domain = argc > 2 ? static_cast<float>(std::atof(argv[2])) : 100.0f;
    const std::uint64_t seed = argc > 3 ? static_cast<std::uint64_t>(std::atoll(argv[3])) : 3u;
    const int steps = argc > 4 ? std::atoi(argv[4]) : 20;
    AxcdSceneSpec spec{n, 0.5f, 0.5f, domain, 0.25f, 0.5f, 16, seed};
    std::vector<math::Transform> xf(n);
    std::vector<collision::Shape> shapes(n);
    std::uint32_t hullUsed = 0;
    if (axcd_scene_generate(&spec, reinterpret_cast<float*>(xf.data()), shapes.data(), nullptr, 0, 0, &hullUsed) != 0) return 2;

    collision::CollisionConfig cfg;
    cfg.maxBodies = n;
    cfg.maxPairs = 8 * n;
    cfg.maxContacts = 8 * n;
    auto created = collision::CollisionWorld::create(cfg);
    if (created.isFailure()) {
        std::printf("create failed: %d %s\n", static_cast<int>(created.errorCode()), created.errorMessage());
        return created.errorCode() == core::ErrorCode::VulkanInitializationFailed ? 77 : 3;
    }
    auto& world = *created.value();
    if (world.setShapes(shapes.data(), n).isFailure()) return 4;
    collision::Broadphase broadphase(world);
    collision::Narrowphase narrowphase(world);
    std::vector<collision::ContactPoint> contacts(cfg.maxContacts);
    // optional 5th argument "pin": page-lock the per-step buffers (transforms in, contacts out)
    const bool pin = argc > 5 && std::string(argv[5]) == "pin";
    std::unique_ptr<collision::PinnedRegion> pinXf, pinCon;
    if (pin) {
        pinXf = std::make_unique<collision::PinnedRegion>(xf.data(), xf.size() * sizeof(math::Transform));
        pinCon = std::make_unique<collision::PinnedRegion>(contacts.data(), contacts.size() * sizeof(collision::ContactPoint));
        if (!pinXf->pinned() || !pinCon->pinned()) return 10;
    }

    double deviceMs = 0.0;
    std::uint32_t pairs = 0, ncon = 0;
    std::chrono::steady_clock::time_point t0;
    for (int s = -3; s < steps; ++s) {   // three warm-up steps
        if (s == 0) t0 = std::chrono::steady_clock::now();
        if (world.setTransforms(xf.data(), n).isFailure()) return 5;
        if (broadphase.update().isFailure()) return 6;
        if (narrowphase.detectCollisions().isFailure()) return 7;
        auto st = world.stats();
        if (st.isFailure()) return 8;
        pairs = st.value().numPairs;
        auto got = world.getContacts(contacts.data(), static_cast<std::uint32_t>(contacts.size()));
        if (got.isFailure()) return 9;
        ncon = got.value();
        if (s >= 0) deviceMs += st.value().totalMs;
    }
    const double wallMs =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%u %u %u %.4f %.4f\n", n, pairs, ncon, deviceMs / steps, wallMs / steps);
    return 0;
}
