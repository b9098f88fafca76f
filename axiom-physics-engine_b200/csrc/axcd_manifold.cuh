// Stage 4: contact manifolds.  One manifold per contact, 1..4 points sharing
// the contact's normal.  Each point expands to a debug::DebugContactPoint (position, normal,
// penetrationDepth; reference: include/axiom/debug/physics_debug_draw.hpp:128-132); the sum of the
// point counts is gui::PhysicsWorldStats::contactPointCount (include/axiom/gui/physics_panel.hpp:21).
//
// Box-box contacts get feature clipping: reference face = the face of either box most aligned with
// the contact normal (ties: body a), incident face = the other box's face most anti-parallel to it,
// Sutherland-Hodgman against the reference face's four side planes, vertices on or below the
// reference face kept (position = midpoint between the vertex and its projection onto the face,
// depth = distance below the face), reduced to at most four (deepest, farthest from it, largest
// triangle, farthest outside that triangle).  Capsule-box contacts clip the capsule's segment against the
// facing box face (1..2 points); capsule-capsule contacts with (nearly) parallel axes get the two ends of the
// overlapping stretch.  Every other pair class, and a clip that comes out empty, keeps the narrowphase point.  Same expression trees as the CPU oracle.
//
// Tiles of 256 contacts per 128-thread block; the 88-byte records are staged in shared memory and
// written with coalesced 128-bit stores; box-box contacts are compacted and clipped densely.  Algorithmic bytes: 40 B contact + 2 x 56 B pose/shape
// gathers in, 88 B out per contact.
#pragma once

#include "axcd_narrow.cuh"

namespace axcd {

constexpr int kManThreads = 128;
constexpr int kManWords = sizeof(AxcdManifold) / 4;   // 22
static_assert(sizeof(AxcdManifold) == 88, "AxcdManifold layout");
static_assert((2 * kManThreads * sizeof(AxcdManifold)) % 16 == 0, "a tile's records are whole float4s");

// BoxFrame / makeBoxFrame: axcd_narrow.cuh (shared with the box-box SAT)

__device__ __forceinline__ int argmaxAbs3(float d0, float d1, float d2) {   // lowest index on ties
    int k = 0;
    float best = fabsf(d0);
    if (fabsf(d1) > best) { best = fabsf(d1); k = 1; }
    if (fabsf(d2) > best) { k = 2; }
    return k;
}
__device__ __forceinline__ V3 pickAxis(const BoxFrame& f, int i) { return (i == 0) ? f.ax[0] : ((i == 1) ? f.ax[1] : f.ax[2]); }
__device__ __forceinline__ float pickHalf(const BoxFrame& f, int i) { return (i == 0) ? f.h[0] : ((i == 1) ? f.h[1] : f.h[2]); }
__device__ __forceinline__ float pick3(float d0, float d1, float d2, int i) { return (i == 0) ? d0 : ((i == 1) ? d1 : d2); }

// Per-thread polygon storage in shared memory: column `tid` of a [words][kManThreads] array, so the
// same word of all threads is contiguous (conflict-free for uniform indices).
struct PolyCol {
    float* base;
    __device__ __forceinline__ V3 get(int k) const {
        return mk3(base[(3 * k) * kManThreads], base[(3 * k + 1) * kManThreads], base[(3 * k + 2) * kManThreads]);
    }
    __device__ __forceinline__ void set(int k, V3 v) const {
        base[(3 * k) * kManThreads] = v.x;
        base[(3 * k + 1) * kManThreads] = v.y;
        base[(3 * k + 2) * kManThreads] = v.z;
    }
    __device__ __forceinline__ float& f(int k) const { return base[k * kManThreads]; }
};

// Clips, keeps and reduces (A-centred frame).  On return poly.get(0..nk) are the kept positions in
// polygon order, dep.f(0..nk) their depths, and the returned mask says which of them survive the
// reduction to four; *outNk = nk (0: the narrowphase point stands).
__device__ __forceinline__ uint32_t boxBoxManifold(const BoxFrame& A, const BoxFrame& B, V3 n, PolyCol poly, PolyCol tmp,
                                                   int* outNk) {
    const float da0 = dot3(n, A.ax[0]), da1 = dot3(n, A.ax[1]), da2 = dot3(n, A.ax[2]);
    const float db0 = dot3(n, B.ax[0]), db1 = dot3(n, B.ax[1]), db2 = dot3(n, B.ax[2]);
    const int ia = argmaxAbs3(da0, da1, da2), ib = argmaxAbs3(db0, db1, db2);
    const float daI = pick3(da0, da1, da2, ia), dbI = pick3(db0, db1, db2, ib);
    const bool refIsA = fabsf(daI) >= fabsf(dbI);
    const BoxFrame& R = refIsA ? A : B;
    const BoxFrame& I = refIsA ? B : A;
    const int i = refIsA ? ia : ib;
    // outward normal of the reference face, facing the other box (n points from a to b)
    const float sgn = refIsA ? ((daI >= 0.0f) ? 1.0f : -1.0f) : ((dbI >= 0.0f) ? -1.0f : 1.0f);
    const V3 nr = pickAxis(R, i) * sgn;
    const float di0 = dot3(nr, I.ax[0]), di1 = dot3(nr, I.ax[1]), di2 = dot3(nr, I.ax[2]);
    const int j = argmaxAbs3(di0, di1, di2);
    const float sI = (pick3(di0, di1, di2, j) >= 0.0f) ? -1.0f : 1.0f;
    const int ju = (j + 1) % 3, jv = (j + 2) % 3;
    const V3 fc = I.c + pickAxis(I, j) * (sI * pickHalf(I, j));
    const V3 eu = pickAxis(I, ju) * pickHalf(I, ju), ev = pickAxis(I, jv) * pickHalf(I, jv);
    int np = 4;
    poly.set(0, (fc + eu) + ev);
    poly.set(1, (fc - eu) + ev);
    poly.set(2, (fc - eu) - ev);
    poly.set(3, (fc + eu) - ev);
    // clip against the reference face's side planes: s * dot(p - cR, ax_w) <= h_w  (ping-pong poly <-> tmp)
    PolyCol src = poly, dst = tmp;
    bool over = false;
    for (int side = 0; side < 4 && np > 0; ++side) {
        const int w = (i + 1 + (side >> 1)) % 3;
        const float s = (side & 1) ? -1.0f : 1.0f;
        const V3 pn = pickAxis(R, w) * s;
        const float hw = pickHalf(R, w);
        int nt = 0;
        V3 prev = src.get(np - 1);
        float dprev = dot3(prev - R.c, pn) - hw;
        for (int k = 0; k < np; ++k) {
            const V3 cur = src.get(k);
            const float dcur = dot3(cur - R.c, pn) - hw;
            const bool inPrev = dprev <= 0.0f, inCur = dcur <= 0.0f;
            // a convex quad clipped by four planes has at most 8 vertices; sign noise on a degenerate
            // (zero-extent) reference face could produce more crossings: never write past the 8 slots
            if (inPrev != inCur) {
                const float t = dprev / (dprev - dcur);
                if (nt < 8) dst.set(nt++, prev + (cur - prev) * t);
                else over = true;
            }
            if (inCur) {
                if (nt < 8) dst.set(nt++, cur);
                else over = true;
            }
            prev = cur;
            dprev = dcur;
        }
        np = nt;
        const PolyCol sw = src; src = dst; dst = sw;
    }
    if (over) np = 0;   // the narrowphase point stands
    // four passes: the clipped polygon is back in `poly`, `tmp` is free and takes the depths.
    // keep the vertices on or below the reference face (in place: nk <= k)
    int nk = 0;
    const float hi = pickHalf(R, i);
    for (int k = 0; k < np; ++k) {
        const V3 pk = poly.get(k);
        const float sep = dot3(pk - R.c, nr) - hi;
        if (sep <= 0.0f) {
            poly.set(nk, pk - nr * (sep * 0.5f));
            tmp.f(nk) = -sep;
            ++nk;
        }
    }
    *outNk = nk;
    if (nk == 0) return 0u;
    uint32_t keep = (1u << nk) - 1u;
    if (nk > 4) {
        // reduction: deepest, farthest from it, largest triangle, then the vertex farthest outside it
        int p0 = 0;
        float d0 = tmp.f(0);
        for (int k = 1; k < nk; ++k) {
            const float dk = tmp.f(k);
            if (dk > d0) { d0 = dk; p0 = k; }
        }
        const V3 q0 = poly.get(p0);
        int p1 = -1;
        float best = -1.0f;
        for (int k = 0; k < nk; ++k) {
            if (k == p0) continue;
            const V3 d = poly.get(k) - q0;
            const float dd = dot3(d, d);
            if (dd > best) { best = dd; p1 = k; }
        }
        const V3 q1 = poly.get(p1);
        const V3 e = q1 - q0;
        // signed areas against the edge p0->p1 go to the upper half of the depth column
        int p2 = -1;
        best = 0.0f;
        for (int k = 0; k < nk; ++k) {
            const float ar = dot3(cross3(e, poly.get(k) - q0), nr);
            tmp.f(8 + k) = ar;
            if (k == p0 || k == p1) continue;
            if (fabsf(ar) > best) { best = fabsf(ar); p2 = k; }
        }
        keep = (1u << p0) | (1u << p1);
        if (p2 >= 0) {
            keep |= 1u << p2;
            const V3 q2 = poly.get(p2);
            const float flip = (tmp.f(8 + p2) >= 0.0f) ? -1.0f : 1.0f;
            const V3 e12 = q2 - q1, e20 = q0 - q2;
            int p3 = -1;
            best = 0.0f;
            for (int k = 0; k < nk; ++k) {
                if (k == p0 || k == p1 || k == p2) continue;
                const V3 pk = poly.get(k);
                const float o01 = tmp.f(8 + k) * flip;
                const float o12 = dot3(cross3(e12, pk - q1), nr) * flip;
                const float o20 = dot3(cross3(e20, pk - q2), nr) * flip;
                float v = o01;
                if (o12 > v) v = o12;
                if (o20 > v) v = o20;
                if (v > best) { best = v; p3 = k; }
            }
            if (p3 >= 0) keep |= 1u << p3;
        }
    }
    return keep;
}

// Capsule against box: the capsule's core segment clipped against the side planes of the box face that
// faces the capsule; the ends of the clipped segment within the radius of that face are the contact points
// (1..2).  Returns the number of points written (0: the narrowphase point stands).  Positions are in the
// frame centred on body a.
__device__ __forceinline__ int capsuleBoxManifold(const BoxFrame& R, V3 cc, V3 e, float r, V3 toCap, V3* __restrict__ outPos,
                                                  float* __restrict__ outDep) {
    const V3 p0 = cc - e, seg = e * 2.0f;
    const float d0a = dot3(toCap, R.ax[0]), d1a = dot3(toCap, R.ax[1]), d2a = dot3(toCap, R.ax[2]);
    const int i = argmaxAbs3(d0a, d1a, d2a);
    const V3 nr = pickAxis(R, i) * ((pick3(d0a, d1a, d2a, i) >= 0.0f) ? 1.0f : -1.0f);
    float t0 = 0.0f, t1 = 1.0f;
    for (int side = 0; side < 4; ++side) {
        const int w = (i + 1 + (side >> 1)) % 3;
        const V3 pn = pickAxis(R, w) * ((side & 1) ? -1.0f : 1.0f);
        const float d0 = dot3(p0 - R.c, pn) - pickHalf(R, w);   // <= 0 inside
        const float dd = dot3(seg, pn);
        if (dd > 0.0f) {
            const float t = -d0 / dd;
            if (t < t1) t1 = t;
        } else if (dd < 0.0f) {
            const float t = -d0 / dd;
            if (t > t0) t0 = t;
        } else if (d0 > 0.0f) {
            return 0;   // parallel to the plane and outside it
        }
    }
    if (!(t0 <= t1)) return 0;
    const int cand = (t0 == t1) ? 1 : 2;
    const float hi = pickHalf(R, i);
    int cnt = 0;
    for (int k = 0; k < cand; ++k) {
        const V3 q = p0 + seg * (k ? t1 : t0);
        const float sep = (dot3(q - R.c, nr) - hi) - r;
        if (sep <= 0.0f) {
            outPos[cnt] = q - nr * (r + sep * 0.5f);
            outDep[cnt] = -sep;
            ++cnt;
        }
    }
    return cnt;
}

// Capsule against capsule with (nearly) parallel axes (|eA x eB|^2 <= 1e-2 |eA|^2 |eB|^2): the two ends of the
// stretch of A's segment that B's segment overlaps; an end is kept when the surfaces overlap there along the
// contact normal.  Frame centred on body a (A's segment is -eA..eA).  Returns the number of points (0: the
// narrowphase point stands).
__device__ __forceinline__ int capsuleCapsuleManifold(V3 eA, float rA, V3 cB, V3 eB, float rB, V3 n, V3* __restrict__ outPos,
                                                      float* __restrict__ outDep) {
    const float aa = dot3(eA, eA), bb = dot3(eB, eB);
    if (!(aa > 0.0f && bb > 0.0f)) return 0;
    const V3 cr = cross3(eA, eB);
    if (!(dot3(cr, cr) <= 1e-2f * (aa * bb))) return 0;
    const float tc = dot3(cB, eA) / aa, te = dot3(eB, eA) / aa;
    float lo = tc - te, hi = tc + te;
    if (lo > hi) { const float t = lo; lo = hi; hi = t; }
    if (lo < -1.0f) lo = -1.0f;
    if (hi > 1.0f) hi = 1.0f;
    if (!(lo < hi)) return 0;
    int cnt = 0;
    for (int k = 0; k < 2; ++k) {
        const V3 pA = eA * (k ? hi : lo);
        const float sB = dot3(pA - cB, eB) / bb;
        const V3 pB = cB + eB * sB;
        const float sep = dot3(pB - pA, n) - (rA + rB);
        if (sep <= 0.0f) {
            outPos[cnt] = ((pA + n * rA) + (pB - n * rB)) * 0.5f;
            outDep[cnt] = -sep;
            ++cnt;
        }
    }
    return cnt;
}

// One block handles tiles of kManTile contacts: every thread writes the single-point record of its
// contacts into the shared staging area, the tile's box-box contacts are compacted into a list, and the
// first threads of the block clip them with all lanes busy (polygons in shared-memory columns, nothing
// in local memory); the finished tile leaves with coalesced 128-bit stores.
constexpr int kManTile = 2 * kManThreads;

__global__ void __launch_bounds__(kManThreads)
manifoldKernel(const AxcdContact* __restrict__ contacts, const uint32_t* __restrict__ contactCount, uint32_t maxContacts,
               const float* __restrict__ xf, const uint4* __restrict__ shapes, float4* __restrict__ out4,
               uint32_t* __restrict__ pointCount) {
    __shared__ __align__(16) float sOut[kManTile * kManWords];     // 22.5 KB
    __shared__ float sPoly[2 * 24 * kManThreads];                  // 24 KB: two 8-vertex polygons per thread
    __shared__ uint16_t sList[kManTile];
    __shared__ uint32_t sCount, sPts;
    const uint32_t total = min(*contactCount, maxContacts);
    const int tid = threadIdx.x;
    for (uint32_t base = blockIdx.x * kManTile; base < total; base += gridDim.x * kManTile) {
        const uint32_t cnt = min((uint32_t)kManTile, total - base);
        if (tid == 0) { sCount = 0; sPts = 0; }
        __syncthreads();
        // ---- phase 1: default (single-point) records; list the box-box contacts ------------------
        for (uint32_t l = tid; l < cnt; l += kManThreads) {
            // 40-byte contact record, 8-byte aligned
            const float2* cr = reinterpret_cast<const float2*>(contacts + base + l);
            const float2 c0 = __ldg(cr), c1 = __ldg(cr + 1), c2 = __ldg(cr + 2), c3 = __ldg(cr + 3), c4 = __ldg(cr + 4);
            const uint32_t a = __float_as_uint(c0.x), b = __float_as_uint(c0.y);
            float* o = sOut + l * kManWords;
            uint32_t* ou = reinterpret_cast<uint32_t*>(o);
            ou[0] = a; ou[1] = b;
            o[2] = c2.y; o[3] = c3.x; o[4] = c3.y;
            ou[5] = 1u;
#pragma unroll
            for (int k = 0; k < 4; ++k) { o[6 + k] = 0.0f; o[10 + k] = 0.0f; o[14 + k] = 0.0f; o[18 + k] = 0.0f; }
            o[6] = c1.x; o[10] = c1.y; o[14] = c2.x; o[18] = c4.x;
            const uint32_t tyA = __ldg(&shapes[a].x), tyB = __ldg(&shapes[b].x);
            // contacts that get more than the narrowphase point: box-box and capsule-box
            const bool bb = (tyA == AXCD_SHAPE_BOX && (tyB == AXCD_SHAPE_BOX || tyB == AXCD_SHAPE_CAPSULE)) ||
                            (tyA == AXCD_SHAPE_CAPSULE && (tyB == AXCD_SHAPE_BOX || tyB == AXCD_SHAPE_CAPSULE));
            const uint32_t bal = __ballot_sync(__activemask(), bb);
            if (bb) {
                const int lane = tid & 31;
                uint32_t wbase = 0;
                const int leader = __ffs(bal) - 1;
                if (lane == leader) wbase = atomicAdd(&sCount, (uint32_t)__popc(bal));
                wbase = __shfl_sync(bal, wbase, leader);
                sList[wbase + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)l;
            }
        }
        __syncthreads();
        // ---- phase 2: clip the box-box contacts, dense over the threads --------------------------------
        const uint32_t nbb = sCount;
        uint32_t extra = 0;   // points beyond the one every contact already has
        for (uint32_t t = tid; t < nbb; t += kManThreads) {
            const uint32_t l = sList[t];
            float* o = sOut + l * kManWords;
            const uint32_t* ou = reinterpret_cast<const uint32_t*>(o);
            const uint32_t a = ou[0], b = ou[1];
            const V3 n = mk3(o[2], o[3], o[4]);
            const BodyPose ta = loadPose(xf, a), tb = loadPose(xf, b);
            const V3 origin = ta.p;
            const uint4 sa = __ldg(shapes + a), sb = __ldg(shapes + b);
            if (sa.x == AXCD_SHAPE_CAPSULE && sb.x == AXCD_SHAPE_CAPSULE) {
                V3 c0, c1, c2;
                quatToColumns(ta.q, c0, c1, c2);
                const V3 eA = c1 * ((__uint_as_float(sa.z) * 0.5f) * ta.s.y);
                quatToColumns(tb.q, c0, c1, c2);
                const V3 eB = c1 * ((__uint_as_float(sb.z) * 0.5f) * tb.s.y);
                V3 cp[2];
                float cd[2];
                const int nc = capsuleCapsuleManifold(eA, __uint_as_float(sa.y), tb.p - origin, eB, __uint_as_float(sb.y), n, cp, cd);
                if (nc == 0) continue;
                for (int k = 0; k < nc; ++k) {
                    const V3 w = cp[k] + origin;
                    o[6 + k] = w.x; o[10 + k] = w.y; o[14 + k] = w.z; o[18 + k] = cd[k];
                }
                reinterpret_cast<uint32_t*>(o)[5] = (uint32_t)nc;
                extra += (uint32_t)nc - 1u;
                continue;
            }
            if (sa.x == AXCD_SHAPE_CAPSULE || sb.x == AXCD_SHAPE_CAPSULE) {
                const bool boxIsA = sa.x == AXCD_SHAPE_BOX;
                const BoxFrame R = boxIsA ? makeBoxFrame(ta, sa, origin) : makeBoxFrame(tb, sb, origin);
                const BodyPose& tC = boxIsA ? tb : ta;
                const uint4 sC = boxIsA ? sb : sa;
                V3 c0, c1, c2;
                quatToColumns(tC.q, c0, c1, c2);
                const V3 e = c1 * ((__uint_as_float(sC.z) * 0.5f) * tC.s.y);   // half segment, as the narrowphase core
                V3 cp[2];
                float cd[2];
                const int nc = capsuleBoxManifold(R, tC.p - origin, e, __uint_as_float(sC.y), boxIsA ? n : -n, cp, cd);
                if (nc == 0) continue;
                for (int k = 0; k < nc; ++k) {
                    const V3 w = cp[k] + origin;
                    o[6 + k] = w.x; o[10 + k] = w.y; o[14 + k] = w.z; o[18 + k] = cd[k];
                }
                reinterpret_cast<uint32_t*>(o)[5] = (uint32_t)nc;
                extra += (uint32_t)nc - 1u;
                continue;
            }
            const BoxFrame A = makeBoxFrame(ta, sa, origin), B = makeBoxFrame(tb, sb, origin);
            const PolyCol poly{sPoly + tid}, tmp{sPoly + 24 * kManThreads + tid};
            int nk = 0;
            const uint32_t keep = boxBoxManifold(A, B, n, poly, tmp, &nk);
            if (nk == 0) continue;   // degenerate clip: the narrowphase point stands
            int c = 0;
            for (int k = 0; k < nk; ++k) {
                if (!((keep >> k) & 1u)) continue;
                const V3 w = poly.get(k) + origin;
                o[6 + c] = w.x; o[10 + c] = w.y; o[14 + c] = w.z; o[18 + c] = tmp.f(k);
                ++c;
            }
            reinterpret_cast<uint32_t*>(o)[5] = (uint32_t)c;
            extra += (uint32_t)c - 1u;
        }
        // contact-point total (PhysicsWorldStats::contactPointCount)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, off);
        if ((tid & 31) == 0 && extra) atomicAdd(&sPts, extra);
        __syncthreads();
        if (tid == 0) atomicAdd(pointCount, sPts + cnt);
        // ---- coalesced 128-bit stores of the tile's records (base * 88 B is a multiple of 16 B) ----------
        const uint32_t nWords = cnt * kManWords;
        const uint32_t nVec = nWords / 4;
        float4* dst = out4 + (size_t)base * kManWords / 4;
        const float4* src = reinterpret_cast<const float4*>(sOut);
        for (uint32_t k = tid; k < nVec; k += kManThreads) dst[k] = src[k];
        float* dstF = reinterpret_cast<float*>(dst);
        for (uint32_t k = nVec * 4 + tid; k < nWords; k += kManThreads) dstF[k] = sOut[k];
        __syncthreads();
    }
}

}  // namespace axcd
