// A C++20 host for the sharded scene (SURVEY.md 8(e), config C4 in small): R ranks as R host threads, one
// GPU and one CollisionWorld each, the ghost exchange over NCCL inside libaxcd.so (initSlab / slabStep).
// The union of the ranks' pair and contact sets (in global ids) must equal the single-GPU run exactly.
//   slab_host [bodies=200000] [domain=58.5] [seed=7] [ranks=2] [steps=5]
// Prints "ranks bodies pairs contacts ghosts ms_per_step_wall graph"; exit 77 without enough GPUs.
#include "axiom/collision/collision_world.hpp"
#include "axcd_scene.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

using namespace axiom;

namespace {
struct RankOut {
    std::vector<collision::BodyPair> pairs;
    std::vector<collision::ContactPoint> contacts;
    std::uint32_t ghosts = 0, graph = 0;
    double wallMs = 0.0;
    int rc = 0;
};

bool pairLess(const collision::BodyPair& x, const collision::BodyPair& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; }
bool contactLess(const collision::ContactPoint& x, const collision::ContactPoint& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; }
}  // namespace

int main(int argc, char** argv) {
    const std::uint32_t n = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 200000u;
    const float domain = argc > 2 ? std::strtof(argv[2], nullptr) : 58.5f;
    const std::uint64_t seed = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 7u;
    const std::uint32_t R = argc > 4 ? std::strtoul(argv[4], nullptr, 10) : 2u;
    const int steps = argc > 5 ? std::atoi(argv[5]) : 5;

    if ((std::uint32_t)axcd_device_count() < R) return 77;   // one GPU per rank

    AxcdSceneSpec spec{n, 0.5f, 0.5f, domain, 0.25f, 0.5f, 16, seed};
    std::vector<math::Transform> xf(n);
    std::vector<collision::Shape> shapes(n);
    std::uint32_t hullUsed = 0;
    if (axcd_scene_generate(&spec, reinterpret_cast<float*>(xf.data()), shapes.data(), nullptr, 0, 0, &hullUsed) != 0) return 2;

    // single-GPU answer on device 0
    std::vector<collision::BodyPair> refPairs;
    std::vector<collision::ContactPoint> refContacts;
    {
        collision::CollisionConfig cfg;
        cfg.maxBodies = n;
        cfg.maxPairs = 8 * n;
        cfg.maxContacts = 8 * n;
        auto created = collision::CollisionWorld::create(cfg);
        if (created.isFailure()) return created.errorCode() == core::ErrorCode::VulkanInitializationFailed ? 77 : 3;
        auto& w = *created.value();
        if (w.setShapes(shapes.data(), n).isFailure() || w.setTransforms(xf.data(), n).isFailure()) return 4;
        auto st = w.step();
        if (st.isFailure()) return 5;
        refPairs.resize(st.value().numPairs);
        refContacts.resize(st.value().numContacts);
        if (w.getPairs(refPairs.data(), (std::uint32_t)refPairs.size()).isFailure()) return 6;
        if (w.getContacts(refContacts.data(), (std::uint32_t)refContacts.size()).isFailure()) return 7;
    }

    // equal-count slab edges on the position x
    std::vector<float> xs(n);
    for (std::uint32_t i = 0; i < n; ++i) xs[i] = reinterpret_cast<const float*>(&xf[i])[0];
    std::vector<float> sorted = xs;
    std::sort(sorted.begin(), sorted.end());
    std::vector<float> edges(R + 1);
    edges[0] = -std::numeric_limits<float>::infinity();
    edges[R] = std::numeric_limits<float>::infinity();
    for (std::uint32_t r = 1; r < R; ++r) edges[r] = sorted[(std::size_t)r * n / R];

    char uid[128];
    {
        auto u = collision::CollisionWorld::ncclUniqueId(uid);
        if (u.isFailure()) return 77;   // no NCCL on this machine
    }
    std::vector<RankOut> out(R);
    std::vector<std::thread> th;
    for (std::uint32_t r = 0; r < R; ++r) {
        th.emplace_back([&, r] {
            RankOut& o = out[r];
            std::vector<std::uint32_t> gid;
            for (std::uint32_t i = 0; i < n; ++i)
                if (xs[i] >= edges[r] && xs[i] < edges[r + 1]) gid.push_back(i);
            const std::uint32_t m = (std::uint32_t)gid.size();
            std::vector<math::Transform> oxf(m);
            std::vector<collision::Shape> osh(m);
            for (std::uint32_t k = 0; k < m; ++k) {
                oxf[k] = xf[gid[k]];
                osh[k] = shapes[gid[k]];
            }
            collision::CollisionConfig cfg;
            cfg.deviceOrdinal = (std::int32_t)r;
            cfg.maxBodies = m + m / 2 + 4096;
            cfg.maxPairs = 8 * cfg.maxBodies;
            cfg.maxContacts = cfg.maxPairs;
            auto created = collision::CollisionWorld::create(cfg);
            if (created.isFailure()) { o.rc = created.errorCode() == core::ErrorCode::InvalidParameter ? 77 : 10; return; }
            auto& w = *created.value();
            if (w.setShapes(osh.data(), m).isFailure() || w.setTransforms(oxf.data(), m).isFailure() ||
                w.setBodyKeys(gid.data(), m).isFailure()) { o.rc = 11; return; }
            if (w.initSlab(uid, r, R, edges.data()).isFailure()) { o.rc = 12; return; }
            AxcdStats st{};
            const auto t0 = std::chrono::steady_clock::now();
            for (int s = 0; s < steps; ++s) {
                auto rs = w.slabStep();
                if (rs.isFailure()) { o.rc = 13; return; }
                st = rs.value();
            }
            o.wallMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
            o.ghosts = st.ghostBodies;
            o.graph = st.graphLaunched;
            std::vector<std::uint32_t> keys(st.numBodies);
            if (w.getBodyKeys(keys.data(), st.numBodies).isFailure()) { o.rc = 14; return; }
            o.pairs.resize(st.numPairs);
            o.contacts.resize(st.numContacts);
            if (st.numPairs && w.getPairs(o.pairs.data(), st.numPairs).isFailure()) { o.rc = 15; return; }
            if (st.numContacts && w.getContacts(o.contacts.data(), st.numContacts).isFailure()) { o.rc = 16; return; }
            for (auto& p : o.pairs) { p.a = keys[p.a]; p.b = keys[p.b]; }
            for (auto& c : o.contacts) { c.a = keys[c.a]; c.b = keys[c.b]; }
        });
    }
    for (auto& t : th) t.join();
    std::vector<collision::BodyPair> allPairs;
    std::vector<collision::ContactPoint> allContacts;
    std::uint32_t ghosts = 0, graph = 1;
    double wall = 0.0;
    for (auto& o : out) {
        if (o.rc) return o.rc;
        allPairs.insert(allPairs.end(), o.pairs.begin(), o.pairs.end());
        allContacts.insert(allContacts.end(), o.contacts.begin(), o.contacts.end());
        ghosts += o.ghosts;
        graph &= o.graph;
        wall = std::max(wall, o.wallMs);
    }
    std::sort(allPairs.begin(), allPairs.end(), pairLess);
    std::sort(allContacts.begin(), allContacts.end(), contactLess);
    if (allPairs.size() != refPairs.size() || std::memcmp(allPairs.data(), refPairs.data(), refPairs.size() * sizeof(collision::BodyPair)) != 0)
        return 20;   // the union of the ranks' candidate pairs is not the single-GPU set
    if (allContacts.size() != refContacts.size() ||
        std::memcmp(allContacts.data(), refContacts.data(), refContacts.size() * sizeof(collision::ContactPoint)) != 0)
        return 21;   // contacts differ (they must match bit for bit)
    std::printf("%u %u %zu %zu %u %.3f %u\n", R, n, allPairs.size(), allContacts.size(), ghosts, wall, graph);
    return 0;
}
