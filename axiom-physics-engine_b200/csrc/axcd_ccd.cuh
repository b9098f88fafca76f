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
