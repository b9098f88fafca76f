// Scene queries on the LBVH of the last broadphase ("scene queries",
// reference: CLAUDE.md:88): AABB overlap queries and closest-hit ray casts.  One thread per query,
// the same 64-byte nodes and register-carried descent as findPairsKernel.
//
// Semantics (shared with the CPU oracle, which answers by brute force over all bodies):
//   AABB query: every body whose AABB meets the query box under AABB::intersects (closed intervals,
//               include/axiom/math/aabb.hpp:132-135); hits come back sorted by (query, body).
//   Ray cast  : a body is hit iff the slab test against its AABB passes within [0, tMax] AND its shape
//               test reports t in [0, tMax]: sphere / oriented box / capsule in closed form, convex hulls by
//               conservative advancement on the GJK distance (first t with the gap <= 1e-4).  Reported t = max(shape t, AABB entry t), closest =
//               minimum (t, body index), so the answer does not depend on the traversal order.  The
//               slab test is monotone under box union (round-to-nearest is monotone), so pruning a node
//               whose entry t exceeds the best t so far can never drop the winner.
#pragma once

#include "axcd_lbvh.cuh"
#include "axcd_narrow.cuh"
#include "axcd_ccd.cuh"

namespace axcd {

constexpr int kQueryThreads = 64;
constexpr uint32_t kNoHit = 0xffffffffu;

struct QueryTree {
    const float4* leafLo;        // sorted leaves: (min.xyz, body index)
    const BvhNode* nodes;
    const uint32_t* sortedKeys;  // Morton keys in sorted order (world id in the high bits)
    const float* aabb;           // per body (used when n < 2: no tree)
    const uint32_t* worldId;     // per body or NULL
    uint32_t n;
    int worldShift;
    int hasWorlds;
};

// Sorted-leaf range [lo, hi] of world w (empty: lo > hi).
__device__ __forceinline__ void worldRange(const QueryTree& T, uint32_t w, uint32_t& lo, uint32_t& hi) {
    if (!T.hasWorlds) {
        lo = 0;
        hi = T.n - 1;
        return;
    }
    uint32_t a = 0, b = T.n;   // first k with world(k) >= w
    while (a < b) {
        const uint32_t m = (a + b) >> 1;
        if ((__ldg(T.sortedKeys + m) >> T.worldShift) < w) a = m + 1; else b = m;
    }
    lo = a;
    b = T.n;                   // first k with world(k) > w
    while (a < b) {
        const uint32_t m = (a + b) >> 1;
        if ((__ldg(T.sortedKeys + m) >> T.worldShift) <= w) a = m + 1; else b = m;
    }
    hi = a - 1;                // a == lo -> hi = lo - 1 (empty; lo == 0 wraps to 0xffffffff, handled by caller)
}

// ---- AABB overlap query ---------------------------------------------------------------------------
// FILL = false: counts[q] = number of hits.  FILL = true: writes the hit bodies into the query's segment
// [starts[q], starts[q] + counts[q]) of segB (unordered; sortSegmentsCoopKernel orders them).
template <bool FILL>
__global__ void __launch_bounds__(kQueryThreads)
queryAabbKernel(QueryTree T, const float* __restrict__ qboxes, const uint32_t* __restrict__ qworld, uint32_t nq,
                uint32_t* __restrict__ counts, const uint32_t* __restrict__ starts, uint32_t* __restrict__ segB) {
    const uint32_t q = blockIdx.x * kQueryThreads + threadIdx.x;
    if (q >= nq) return;
    const float* b = qboxes + (size_t)q * 6;
    const float l0 = __ldg(b), l1 = __ldg(b + 1), l2 = __ldg(b + 2), h0 = __ldg(b + 3), h1 = __ldg(b + 4), h2 = __ldg(b + 5);
    const uint32_t w = (T.hasWorlds && qworld) ? __ldg(qworld + q) : 0u;
    uint32_t c = 0;
    const uint32_t start = FILL ? starts[q] : 0u;
    if (T.n == 1) {
        const float* a = T.aabb;
        const bool worldOk = !(T.hasWorlds && qworld) || T.worldId[0] == w;
        if (worldOk && boxesIntersect(l0, l1, l2, h0, h1, h2, a[0], a[1], a[2], a[3], a[4], a[5])) {
            if (FILL) segB[start] = 0u;
            c = 1;
        }
    } else if (T.n >= 2) {
        uint32_t rlo = 0, rhi = T.n - 1;
        if (T.hasWorlds && qworld) worldRange(T, w, rlo, rhi);
        if (rlo <= rhi && rhi != 0xffffffffu) {
            uint32_t stack[kTravStack];
            int sp = 0;
            uint32_t ni = 0;
            while (true) {
                const float4* np = reinterpret_cast<const float4*>(T.nodes + ni);
                const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
                const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(np) + 3);
                const uint32_t first = q3.x, split = q3.y, last = q3.z;
                const bool hitL = first <= rhi && split >= rlo &&
                                  boxesIntersect(l0, l1, l2, h0, h1, h2, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
                const bool hitR = split + 1 <= rhi && last >= rlo &&
                                  boxesIntersect(l0, l1, l2, h0, h1, h2, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
                uint32_t next = kNoHit;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (!(s ? hitR : hitL)) continue;
                    const bool leaf = s ? (split + 1 == last) : (first == split);
                    const uint32_t child = s ? split + 1 : split;
                    if (!leaf) {
                        if (next == kNoHit) next = child;
                        else if (sp < kTravStack) stack[sp++] = child;
                        continue;
                    }
                    if (child < rlo || child > rhi) continue;
                    if (FILL) segB[start + c] = __float_as_uint(__ldg(&T.leafLo[child].w));
                    ++c;
                }
                if (next != kNoHit) ni = next;
                else if (sp > 0) ni = stack[--sp];
                else break;
            }
        }
    }
    if (!FILL) counts[q] = c;
}

// ---- ray cast -----------------------------------------------------------------------------------------
// Slab test of the ray against a box, clipped to [0, tMax]; *tNear = entry parameter (>= 0).
__device__ __forceinline__ bool raySlab(V3 o, V3 d, float lx, float ly, float lz, float hx, float hy, float hz,
                                        float tMax, float& tNear) {
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    const float lo[3] = {lx, ly, lz}, hi[3] = {hx, hy, hz};
    float tn = 0.0f, tf = tMax;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!(lo[k] <= hi[k])) return false;   // empty or NaN box
        if (dd[k] == 0.0f) {
            if (!(lo[k] <= oo[k] && oo[k] <= hi[k])) return false;
            continue;
        }
        const float inv = 1.0f / dd[k];
        const float t1 = (lo[k] - oo[k]) * inv, t2 = (hi[k] - oo[k]) * inv;
        const float a = (t1 < t2) ? t1 : t2, b = (t1 < t2) ? t2 : t1;
        if (a > tn) tn = a;
        if (b < tf) tf = b;
    }
    if (!(tn <= tf)) return false;
    tNear = tn;
    return true;
}

struct ShapeHit {
    bool hit;
    float t;
    V3 n;
};

__device__ __forceinline__ ShapeHit raySphere(V3 m /* origin - centre */, V3 d, float r, float tMax) {
    ShapeHit h{false, 0.0f, mk3(0.f, 0.f, 0.f)};
    const float dd = dot3(d, d), b = dot3(m, d), cc = dot3(m, m) - r * r;
    if (cc <= 0.0f) {
        h.hit = true;
        return h;
    }
    const float disc = b * b - dd * cc;
    if (!(disc >= 0.0f)) return h;
    const float t = (-b - sqrtf(disc)) / dd;
    if (!(t >= 0.0f && t <= tMax)) return h;
    h.hit = true;
    h.t = t;
    h.n = (m + d * t) * (1.0f / r);
    return h;
}

__device__ __noinline__ ShapeHit rayShape(V3 o, V3 d, float tMax, const BodyPose& t, uint4 sh) {
    ShapeHit h{false, 0.0f, mk3(0.f, 0.f, 0.f)};
    const V3 m = o - t.p;
    const float p0 = __uint_as_float(sh.y), p1 = __uint_as_float(sh.z), p2 = __uint_as_float(sh.w);
    if (sh.x == AXCD_SHAPE_SPHERE) return raySphere(m, d, p0, tMax);
    if (sh.x == AXCD_SHAPE_BOX) {
        V3 ax[3];
        quatToColumns(t.q, ax[0], ax[1], ax[2]);
        const float half[3] = {fabsf(p0 * t.s.x), fabsf(p1 * t.s.y), fabsf(p2 * t.s.z)};
        float ol[3], dl[3];
        bool inside = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ol[k] = dot3(m, ax[k]);
            dl[k] = dot3(d, ax[k]);
            if (!(fabsf(ol[k]) <= half[k])) inside = false;
        }
        if (inside) {
            h.hit = true;
            return h;
        }
        float tn = -3.0e38f, tf = 3.0e38f;
        int axis = -1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (dl[k] == 0.0f) {
                if (!(fabsf(ol[k]) <= half[k])) return h;
                continue;
            }
            const float inv = 1.0f / dl[k];
            const float t1 = (-half[k] - ol[k]) * inv, t2 = (half[k] - ol[k]) * inv;
            const float a = (t1 < t2) ? t1 : t2, b = (t1 < t2) ? t2 : t1;
            if (a > tn) { tn = a; axis = k; }
            if (b < tf) tf = b;
        }
        if (!(tn <= tf) || !(tf >= 0.0f) || axis < 0) return h;
        const float tt = (tn < 0.0f) ? 0.0f : tn;
        if (!(tt <= tMax)) return h;
        const V3 axn = (axis == 0) ? ax[0] : ((axis == 1) ? ax[1] : ax[2]);
        const float dla = (axis == 0) ? dl[0] : ((axis == 1) ? dl[1] : dl[2]);
        h.hit = true;
        h.t = tt;
        h.n = (dla > 0.0f) ? -axn : axn;
        return h;
    }
    if (sh.x == AXCD_SHAPE_CAPSULE) {
        // segment pa..pb = centre -+ column1 * (height/2 * scale.y), as the narrowphase core; radius unscaled
        V3 c0, c1, c2;
        quatToColumns(t.q, c0, c1, c2);
        const V3 e = c1 * ((p1 * 0.5f) * t.s.y);
        const float r = p0;
        const V3 oa = m + e;   // origin - pa
        const V3 ob = m - e;   // origin - pb
        const V3 ba = e * 2.0f;
        const float baba = dot3(ba, ba), baoa = dot3(ba, oa);
        {
            float sgm = (baba > 0.0f) ? baoa / baba : 0.0f;
            sgm = (sgm < 0.0f) ? 0.0f : ((sgm > 1.0f) ? 1.0f : sgm);
            const V3 q = oa - ba * sgm;
            if (dot3(q, q) <= r * r) {
                h.hit = true;
                return h;
            }
        }
        bool any = false;
        float best = 0.0f;
        V3 bn = mk3(0.f, 0.f, 0.f);
        const float dd = dot3(d, d), bard = dot3(ba, d), rdoa = dot3(d, oa), oaoa = dot3(oa, oa);
        const float a = baba * dd - bard * bard;
        if (a > 0.0f) {
            const float b = baba * rdoa - baoa * bard;
            const float c = (baba * oaoa - baoa * baoa) - (r * r) * baba;
            const float disc = b * b - a * c;
            if (disc >= 0.0f) {
                const float tc = (-b - sqrtf(disc)) / a;
                const float y = baoa + tc * bard;
                if (tc >= 0.0f && tc <= tMax && y > 0.0f && y < baba) {
                    any = true;
                    best = tc;
                    bn = ((oa + d * tc) - ba * (y / baba)) * (1.0f / r);
                }
            }
        }
        const ShapeHit ha = raySphere(oa, d, r, tMax), hb = raySphere(ob, d, r, tMax);
        if (ha.hit && (!any || ha.t < best)) { any = true; best = ha.t; bn = ha.n; }
        if (hb.hit && (!any || hb.t < best)) { any = true; best = hb.t; bn = hb.n; }
        h.hit = any;
        h.t = best;
        h.n = bn;
        return h;
    }
    return h;   // hulls: rayHull
}

// Ray against a convex hull: the ray origin is a point core at rest, the hull moves by -d * tMax; the time
// of impact of that sweep (conservative advancement on the GJK distance) is t / tMax.  Surface normal =
// from the hull towards the ray origin.
__device__ __noinline__ ShapeHit rayHull(V3 o, V3 d, float tMax, const BodyPose& t, uint4 sh, const float4* __restrict__ hull,
                                         const NarrowParams& cfg) {
    ShapeHit h{false, 0.0f, mk3(0.f, 0.f, 0.f)};
    CoreT<true> P;   // the swept path serves hulls and cylinders: always the cylinder-capable instantiation
    P.kind = CORE_POINT;
    P.c = mk3(0.f, 0.f, 0.f);
    P.e0 = P.e1 = P.e2 = P.c;
    P.s = P.c;
    P.verts = nullptr;
    P.nv = 0;
    P.r = 0.0f;
    const SweepResult sw = sweepCores(P, makeCore<true>(t, sh, hull, o), -(d * tMax), cfg);
    if (!sw.hit) return h;
    h.hit = true;
    h.t = sw.toi * tMax;
    h.n = -sw.n;
    return h;
}

struct RayBest {
    uint32_t body;
    float t;
    V3 n;
    uint32_t flags;
    bool have;
};

// Exact test of one body whose AABB the ray enters at tNear; updates the best hit.
template <bool HULLS>
__device__ __forceinline__ void rayTestBody(uint32_t body, float tNear, V3 o, V3 d, float tMax,
                                            const float* __restrict__ xf, const uint4* __restrict__ shapes,
                                            const float4* __restrict__ hull, const NarrowParams& cfg, RayBest& best) {
    const uint4 sh = __ldg(shapes + body);
    const uint32_t flags = 0u;
    ShapeHit h;
    if (!HULLS || (sh.x != AXCD_SHAPE_CONVEX && sh.x != AXCD_SHAPE_CYLINDER)) h = rayShape(o, d, tMax, loadPose(xf, body), sh);
    else h = rayHull(o, d, tMax, loadPose(xf, body), sh, hull, cfg);
    if (!h.hit) return;
    const float t = (h.t > tNear) ? h.t : tNear;
    if (!best.have || t < best.t || (t == best.t && body < best.body)) {
        best.have = true;
        best.body = body;
        best.t = t;
        best.n = h.n;
        best.flags = flags;
    }
}

// rays: 32-byte records (origin, direction, tMax, world); hits: 24-byte records.  HULLS = false is the lean
// instantiation for scenes without convex hulls (the GJK machinery of rayHull costs 56 more registers).
template <bool HULLS>
__global__ void __launch_bounds__(kQueryThreads)
raycastKernel(QueryTree T, const float4* __restrict__ rays, uint32_t nq, const float* __restrict__ xf,
              const uint4* __restrict__ shapes, const float4* __restrict__ hull, NarrowParams cfg,
              uint32_t* __restrict__ hits) {
    const uint32_t q = blockIdx.x * kQueryThreads + threadIdx.x;
    if (q >= nq) return;
    const float4 r0 = __ldg(rays + 2 * (size_t)q), r1 = __ldg(rays + 2 * (size_t)q + 1);
    const V3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r0.w, r1.x, r1.y);
    const float tMax = r1.z;
    const uint32_t w = __float_as_uint(r1.w);
    RayBest best{kNoHit, tMax, mk3(0.f, 0.f, 0.f), 0u, false};
    if (T.n == 1) {
        const float* a = T.aabb;
        float tNear;
        const bool worldOk = !T.hasWorlds || T.worldId[0] == w;
        if (worldOk && raySlab(o, d, a[0], a[1], a[2], a[3], a[4], a[5], tMax, tNear))
            rayTestBody<HULLS>(0u, tNear, o, d, tMax, xf, shapes, hull, cfg, best);
    } else if (T.n >= 2) {
        uint32_t rlo = 0, rhi = T.n - 1;
        if (T.hasWorlds) worldRange(T, w, rlo, rhi);
        if (rlo <= rhi && rhi != 0xffffffffu) {
            uint32_t stack[kTravStack];
            int sp = 0;
            uint32_t ni = 0;
            while (true) {
                const float4* np = reinterpret_cast<const float4*>(T.nodes + ni);
                const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
                const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(np) + 3);
                const uint32_t first = q3.x, split = q3.y, last = q3.z;
                const float tCap = best.have ? best.t : tMax;   // ties (t == best) must still be visited
                float tL = 0.0f, tR = 0.0f;
                bool hitL = first <= rhi && split >= rlo && raySlab(o, d, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tCap, tL);
                bool hitR = split + 1 <= rhi && last >= rlo && raySlab(o, d, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tCap, tR);
                // leaves are tested at once; internal children are visited nearer first
                if (hitL && first == split) {
                    if (split >= rlo && split <= rhi)
                        rayTestBody<HULLS>(__float_as_uint(__ldg(&T.leafLo[split].w)), tL, o, d, tMax, xf, shapes, hull, cfg, best);
                    hitL = false;
                }
                if (hitR && split + 1 == last) {
                    if (last >= rlo && last <= rhi)
                        rayTestBody<HULLS>(__float_as_uint(__ldg(&T.leafLo[last].w)), tR, o, d, tMax, xf, shapes, hull, cfg, best);
                    hitR = false;
                }
                uint32_t next = kNoHit;
                if (hitL && hitR) {
                    const bool leftFirst = tL <= tR;
                    next = leftFirst ? split : split + 1;
                    if (sp < kTravStack) stack[sp++] = leftFirst ? split + 1 : split;
                } else if (hitL) {
                    next = split;
                } else if (hitR) {
                    next = split + 1;
                }
                if (next != kNoHit) ni = next;
                else if (sp > 0) ni = stack[--sp];
                else break;
            }
        }
    }
    uint32_t* out = hits + (size_t)q * 6;
    out[0] = best.body;
    out[1] = __float_as_uint(best.t);
    out[2] = __float_as_uint(best.n.x);
    out[3] = __float_as_uint(best.n.y);
    out[4] = __float_as_uint(best.n.z);
    out[5] = best.flags;
}

}  // namespace axcd
