// GJK-based continuous collision detection ("GJK-based CCD", reference:
// CLAUDE.md:136): time of impact of body pairs under LINEAR motion by conservative advancement.  B moves
// by D = dispB - dispA relative to A over t in [0,1]; each step runs the exact GJK distance (cores minus
// radii); gap / (approach speed along the closest direction) is a lower bound of the time to contact for
// convex shapes, so t advances by it until the gap is <= tol (hit), the shapes move apart, or t > 1.
// One thread per pair; same expression trees as the CPU oracle.
#pragma once

#include "axcd_narrow.cuh"

namespace axcd {

constexpr float kCcdTol = 1e-4f;
constexpr int kCcdMaxIters = 48;
constexpr int kCcdThreads = 64;

struct SweepResult {
    uint32_t hit;
    float toi;
    V3 n;
    uint32_t iterations;
};

// Conservative advancement of core B (moved by D * t, t in [0,1]) against the fixed core A.
template <bool CYL>
__device__ __forceinline__ SweepResult sweepCores(const CoreT<CYL>& A, CoreT<CYL> B, V3 D, NarrowParams cfg) {
    cfg.wantDistances = 1u;
    const V3 c0 = B.c;
    const float rs = A.r + B.r;
    SweepResult out{0u, 1.0f, mk3(0.f, 0.f, 0.f), 0u};
    V3 nLast = mk3(0.f, 0.f, 0.f);
    float t = 0.0f;
    int it = 0;
    for (; it < kCcdMaxIters; ++it) {
        B.c = c0 + D * t;
        Simplex s;
        const GjkResult g = gjk(A, B, cfg, rs, s);
        if (g.state == GJK_OVERLAP) {   // cores touch at t: the normal is the last closest direction (zero at t = 0)
            out.hit = 1u;
            out.toi = t;
            out.n = nLast;
            break;
        }
        const float dist = sqrtf(g.vv);
        const float gap = dist - rs;
        const V3 n = -(g.v * (1.0f / dist));   // from a to b
        nLast = n;
        if (gap <= kCcdTol) {
            out.hit = 1u;
            out.toi = t;
            out.n = n;
            break;
        }
        const float approach = dot3(D, g.v) / dist;   // speed at which B closes in along the closest direction
        if (!(approach > 0.0f)) break;                // moving apart or sliding past
        t = t + gap / approach;
        if (!(t <= 1.0f)) break;
    }
    if (it == kCcdMaxIters) {
        // the advancement ran out of iterations (grazing / very slow approach) while still closing in within
        // the step: report the conservative answer — contact at the time reached — rather than a miss
        out.hit = 1u;
        out.toi = t;
        out.n = nLast;
    }
    out.iterations = (uint32_t)it;
    return out;
}

// ---- CCD with rotation ---------------------------------------------------------------------------------
// Each body moves by disp * t and turns by its rotation vector w (angular velocity * dt, world frame) under
// first-order integration, q(t) = normalize(q + t * 0.5 * (w, 0) (x) q) — algebraic, bit-identical to the oracle.
// The turning rate of that path never exceeds |w|, so with rho = the largest distance of a core point from its
// body's position, gap / (approach of the origins along the closest direction + |wA| rhoA + |wB| rhoB) is a lower
// bound of the time to contact; both cores are rebuilt from the poses at t every step.
constexpr int kCcdAngMaxIters = 64;

__device__ __forceinline__ BodyPose poseAt(const BodyPose& t0, V3 d, float4 dq, float t) {
    BodyPose o = t0;
    o.p = t0.p + d * t;
    const float4 u = make_float4(t0.q.x + dq.x * t, t0.q.y + dq.y * t, t0.q.z + dq.z * t, t0.q.w + dq.w * t);
    const float len2 = (u.x * u.x + u.y * u.y) + (u.z * u.z + u.w * u.w);
    const float inv = 1.0f / sqrtf(len2);
    o.q = make_float4(u.x * inv, u.y * inv, u.z * inv, u.w * inv);
    return o;
}
__device__ __forceinline__ float4 halfSpin(V3 w, float4 q) {   // 0.5 * (w, 0) (x) q
    const V3 qv = mk3(q.x, q.y, q.z);
    const V3 v = cross3(w, qv) + w * q.w;
    return make_float4(v.x * 0.5f, v.y * 0.5f, v.z * 0.5f, -dot3(w, qv) * 0.5f);
}
template <bool CYL>
__device__ __forceinline__ float coreReach(const CoreT<CYL>& k) {
    if (k.kind == CORE_POINT) return 0.0f;
    if (k.kind == CORE_SEGMENT) return sqrtf(dot3(k.e0, k.e0));
    if (k.kind == CORE_BOX || k.kind == CORE_CYLINDER) return sqrtf((dot3(k.e0, k.e0) + dot3(k.e1, k.e1)) + dot3(k.e2, k.e2));
    float best = 0.0f;
    for (uint32_t i = 0; i < k.nv; ++i) {
        const float4 v = __ldg(k.verts + i);
        const V3 lv = mk3(v.x * k.s.x, v.y * k.s.y, v.z * k.s.z);
        const float d2 = dot3(lv, lv);
        if (d2 > best) best = d2;
    }
    return sqrtf(best);
}

template <bool CYL>
__global__ void __launch_bounds__(kCcdThreads)
ccdAngularKernel(const uint2* __restrict__ pairs, uint32_t npairs, const float* __restrict__ xf, const uint4* __restrict__ shapes,
                 const float4* __restrict__ hull, const float* __restrict__ disp, const float* __restrict__ rot, NarrowParams cfg,
                 uint32_t* __restrict__ out) {
    const uint32_t k = blockIdx.x * kCcdThreads + threadIdx.x;
    if (k >= npairs) return;
    cfg.wantDistances = 1u;
    const uint2 pr = __ldg(pairs + k);
    const BodyPose ta = loadPose(xf, pr.x), tb = loadPose(xf, pr.y);
    const uint4 sa = __ldg(shapes + pr.x), sb = __ldg(shapes + pr.y);
    const V3 origin = ta.p;
    auto vec = [](const float* __restrict__ p, uint32_t i) { return mk3(__ldg(p + 3 * (size_t)i), __ldg(p + 3 * (size_t)i + 1), __ldg(p + 3 * (size_t)i + 2)); };
    const V3 D = vec(disp, pr.y) - vec(disp, pr.x);
    const V3 rotA = vec(rot, pr.x), rotB = vec(rot, pr.y);
    const float4 dqA = halfSpin(rotA, ta.q), dqB = halfSpin(rotB, tb.q);
    const V3 zero = mk3(0.f, 0.f, 0.f);
    const float spin = sqrtf(dot3(rotA, rotA)) * coreReach(makeCore<CYL>(ta, sa, hull, origin)) +
                       sqrtf(dot3(rotB, rotB)) * coreReach(makeCore<CYL>(tb, sb, hull, origin));
    SweepResult o{0u, 1.0f, zero, 0u};
    V3 nLast = zero;
    float t = 0.0f;
    int it = 0;
    for (; it < kCcdAngMaxIters; ++it) {
        const CoreT<CYL> A = makeCore<CYL>(poseAt(ta, zero, dqA, t), sa, hull, origin);
        const CoreT<CYL> B = makeCore<CYL>(poseAt(tb, D, dqB, t), sb, hull, origin);
        const float rs = A.r + B.r;
        Simplex s;
        const GjkResult g = gjk(A, B, cfg, rs, s);
        if (g.state == GJK_OVERLAP) {
            o.hit = 1u;
            o.toi = t;
            o.n = nLast;
            break;
        }
        const float dist = sqrtf(g.vv);
        const float gap = dist - rs;
        const V3 n = -(g.v * (1.0f / dist));
        nLast = n;
        if (gap <= kCcdTol) {
            o.hit = 1u;
            o.toi = t;
            o.n = n;
            break;
        }
        const float approach = dot3(D, g.v) / dist + spin;
        if (!(approach > 0.0f)) break;
        t = t + gap / approach;
        if (!(t <= 1.0f)) break;
    }
    if (it == kCcdAngMaxIters) {   // conservative: contact at the time reached
        o.hit = 1u;
        o.toi = t;
        o.n = nLast;
    }
    uint32_t* w = out + (size_t)k * 6;
    w[0] = o.hit;
    w[1] = __float_as_uint(o.toi);
    w[2] = __float_as_uint(o.n.x);
    w[3] = __float_as_uint(o.n.y);
    w[4] = __float_as_uint(o.n.z);
    w[5] = (uint32_t)it;
}

template <bool CYL>
__global__ void __launch_bounds__(kCcdThreads)
ccdKernel(const uint2* __restrict__ pairs, uint32_t npairs, const float* __restrict__ xf, const uint4* __restrict__ shapes,
          const float4* __restrict__ hull, const float* __restrict__ disp, NarrowParams cfg, uint32_t* __restrict__ out) {
    const uint32_t k = blockIdx.x * kCcdThreads + threadIdx.x;
    if (k >= npairs) return;
    const uint2 pr = __ldg(pairs + k);
    const BodyPose ta = loadPose(xf, pr.x), tb = loadPose(xf, pr.y);
    const V3 origin = ta.p;
    const V3 dA = mk3(__ldg(disp + 3 * (size_t)pr.x), __ldg(disp + 3 * (size_t)pr.x + 1), __ldg(disp + 3 * (size_t)pr.x + 2));
    const V3 dB = mk3(__ldg(disp + 3 * (size_t)pr.y), __ldg(disp + 3 * (size_t)pr.y + 1), __ldg(disp + 3 * (size_t)pr.y + 2));
    const SweepResult r = sweepCores(makeCore<CYL>(ta, __ldg(shapes + pr.x), hull, origin), makeCore<CYL>(tb, __ldg(shapes + pr.y), hull, origin),
                                     dB - dA, cfg);
    uint32_t* o = out + (size_t)k * 6;
    o[0] = r.hit;
    o[1] = __float_as_uint(r.toi);
    o[2] = __float_as_uint(r.n.x);
    o[3] = __float_as_uint(r.n.y);
    o[4] = __float_as_uint(r.n.z);
    o[5] = r.iterations;
}

}  // namespace axcd
