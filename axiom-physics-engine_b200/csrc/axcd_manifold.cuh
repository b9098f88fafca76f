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
// triangle, farthest outside that triangle).  Every other pair class, and a box-box contact whose
// clip comes out empty, keeps the narrowphase point.  Same expression trees as the CPU oracle.
//
// One thread per contact; the 88-byte records of a 128-contact block are staged in shared memory and
// written with coalesced 128-bit stores.  Algorithmic bytes: 40 B contact + 2 x 56 B pose/shape
// gathers in, 88 B out per contact.
#pragma once

#include "axcd_narrow.cuh"

namespace axcd {

constexpr int kManThreads = 128;
constexpr int kManWords = sizeof(AxcdManifold) / 4;   // 22
static_assert(sizeof(AxcdManifold) == 88, "AxcdManifold layout");
static_assert((kManThreads * sizeof(AxcdManifold)) % 16 == 0, "a block's records are whole float4s");

struct BoxFrame {
    V3 c;        // centre relative to A's position
    V3 ax[3];    // unit axes: the columns of Quat::toMatrix
    float h[3];  // half lengths |halfExtent * scale|
};

__device__ __forceinline__ BoxFrame makeBoxFrame(const BodyPose& t, uint4 sh, V3 origin) {
    BoxFrame f;
    quatToColumns(t.q, f.ax[0], f.ax[1], f.ax[2]);
    f.c = t.p - origin;
    f.h[0] = fabsf(__uint_as_float(sh.y) * t.s.x);
    f.h[1] = fabsf(__uint_as_float(sh.z) * t.s.y);
    f.h[2] = fabsf(__uint_as_float(sh.w) * t.s.z);
    return f;
}

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

// Fills pos/dep with the clipped, kept and reduced points (A-centred frame); returns their count
// (0: the narrowphase point stands).
__device__ __noinline__ int boxBoxManifold(const BoxFrame& A, const BoxFrame& B, V3 n, V3* __restrict__ outPos,
                                           float* __restrict__ outDep) {
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
    V3 poly[8], tmp[8];
    int np = 4;
    poly[0] = (fc + eu) + ev;
    poly[1] = (fc - eu) + ev;
    poly[2] = (fc - eu) - ev;
    poly[3] = (fc + eu) - ev;
    // clip against the reference face's side planes: s * dot(p - cR, ax_w) <= h_w
    for (int side = 0; side < 4 && np > 0; ++side) {
        const int w = (i + 1 + (side >> 1)) % 3;
        const float s = (side & 1) ? -1.0f : 1.0f;
        const V3 pn = pickAxis(R, w) * s;
        const float hw = pickHalf(R, w);
        int nt = 0;
        V3 prev = poly[np - 1];
        float dprev = dot3(prev - R.c, pn) - hw;
        for (int k = 0; k < np; ++k) {
            const V3 cur = poly[k];
            const float dcur = dot3(cur - R.c, pn) - hw;
            const bool inPrev = dprev <= 0.0f, inCur = dcur <= 0.0f;
            if (inPrev != inCur) {
                const float t = dprev / (dprev - dcur);
                tmp[nt++] = prev + (cur - prev) * t;
            }
            if (inCur) tmp[nt++] = cur;
            prev = cur;
            dprev = dcur;
        }
        np = nt;
        for (int k = 0; k < np; ++k) poly[k] = tmp[k];
    }
    // keep the vertices on or below the reference face
    V3 pos[8];
    float dep[8];
    int nk = 0;
    const float hi = pickHalf(R, i);
    for (int k = 0; k < np; ++k) {
        const float sep = dot3(poly[k] - R.c, nr) - hi;
        if (sep <= 0.0f) {
            pos[nk] = poly[k] - nr * (sep * 0.5f);
            dep[nk] = -sep;
            ++nk;
        }
    }
    if (nk == 0) return 0;
    uint32_t keep = (1u << nk) - 1u;
    if (nk > 4) {
        // reduction: deepest, farthest from it, largest triangle, then the vertex farthest outside it
        int p0 = 0;
        for (int k = 1; k < nk; ++k)
            if (dep[k] > dep[p0]) p0 = k;
        int p1 = -1;
        float best = -1.0f;
        for (int k = 0; k < nk; ++k) {
            if (k == p0) continue;
            const V3 d = pos[k] - pos[p0];
            const float dd = dot3(d, d);
            if (dd > best) { best = dd; p1 = k; }
        }
        const V3 e = pos[p1] - pos[p0];
        float area[8];
        for (int k = 0; k < nk; ++k) area[k] = dot3(cross3(e, pos[k] - pos[p0]), nr);
        int p2 = -1;
        best = 0.0f;
        for (int k = 0; k < nk; ++k) {
            if (k == p0 || k == p1) continue;
            if (fabsf(area[k]) > best) { best = fabsf(area[k]); p2 = k; }
        }
        keep = (1u << p0) | (1u << p1);
        if (p2 >= 0) {
            keep |= 1u << p2;
            const float flip = (area[p2] >= 0.0f) ? -1.0f : 1.0f;
            const V3 e12 = pos[p2] - pos[p1], e20 = pos[p0] - pos[p2];
            int p3 = -1;
            best = 0.0f;
            for (int k = 0; k < nk; ++k) {
                if (k == p0 || k == p1 || k == p2) continue;
                const float o01 = area[k] * flip;
                const float o12 = dot3(cross3(e12, pos[k] - pos[p1]), nr) * flip;
                const float o20 = dot3(cross3(e20, pos[k] - pos[p2]), nr) * flip;
                float v = o01;
                if (o12 > v) v = o12;
                if (o20 > v) v = o20;
                if (v > best) { best = v; p3 = k; }
            }
            if (p3 >= 0) keep |= 1u << p3;
        }
    }
    int cnt = 0;
    for (int k = 0; k < nk; ++k) {
        if (!((keep >> k) & 1u)) continue;
        outPos[cnt] = pos[k];
        outDep[cnt] = dep[k];
        ++cnt;
    }
    return cnt;
}

__global__ void __launch_bounds__(kManThreads)
manifoldKernel(const AxcdContact* __restrict__ contacts, const uint32_t* __restrict__ contactCount, uint32_t maxContacts,
               const float* __restrict__ xf, const uint4* __restrict__ shapes, float4* __restrict__ out4,
               uint32_t* __restrict__ pointCount) {
    __shared__ __align__(16) float sOut[kManThreads * kManWords];
    __shared__ uint32_t sPts[kManThreads / 32];
    const uint32_t total = min(*contactCount, maxContacts);
    const int tid = threadIdx.x;
    for (uint32_t base = blockIdx.x * kManThreads; base < total; base += gridDim.x * kManThreads) {
        const uint32_t cnt = min((uint32_t)kManThreads, total - base);
        uint32_t myPoints = 0;
        if (tid < (int)cnt) {
            // 40-byte contact record, 8-byte aligned
            const float2* cr = reinterpret_cast<const float2*>(contacts + base + tid);
            const float2 c0 = __ldg(cr), c1 = __ldg(cr + 1), c2 = __ldg(cr + 2), c3 = __ldg(cr + 3), c4 = __ldg(cr + 4);
            const uint32_t a = __float_as_uint(c0.x), b = __float_as_uint(c0.y);
            const V3 cpos = mk3(c1.x, c1.y, c2.x);
            const V3 n = mk3(c2.y, c3.x, c3.y);
            const float cdepth = c4.x;
            const uint4 sa = __ldg(shapes + a), sb = __ldg(shapes + b);
            V3 pos[4];
            float dep[4];
            int np = 0;
            V3 origin = mk3(0.0f, 0.0f, 0.0f);
            if (sa.x == AXCD_SHAPE_BOX && sb.x == AXCD_SHAPE_BOX) {
                const BodyPose ta = loadPose(xf, a), tb = loadPose(xf, b);
                origin = ta.p;
                const BoxFrame A = makeBoxFrame(ta, sa, origin), B = makeBoxFrame(tb, sb, origin);
                np = boxBoxManifold(A, B, n, pos, dep);
            }
            float* o = sOut + tid * kManWords;
            uint32_t* ou = reinterpret_cast<uint32_t*>(o);
            ou[0] = a; ou[1] = b;
            o[2] = n.x; o[3] = n.y; o[4] = n.z;
#pragma unroll
            for (int k = 0; k < 4; ++k) { o[6 + k] = 0.0f; o[10 + k] = 0.0f; o[14 + k] = 0.0f; o[18 + k] = 0.0f; }
            if (np == 0) {
                ou[5] = 1u;
                o[6] = cpos.x; o[10] = cpos.y; o[14] = cpos.z; o[18] = cdepth;
                myPoints = 1;
            } else {
                ou[5] = (uint32_t)np;
                for (int k = 0; k < np; ++k) {
                    const V3 w = pos[k] + origin;
                    o[6 + k] = w.x; o[10 + k] = w.y; o[14 + k] = w.z; o[18 + k] = dep[k];
                }
                myPoints = (uint32_t)np;
            }
        }
        // contact-point total (PhysicsWorldStats::contactPointCount): warp sums, one atomic per block tile
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) myPoints += __shfl_xor_sync(0xffffffffu, myPoints, off);
        if ((tid & 31) == 0) sPts[tid >> 5] = myPoints;
        __syncthreads();
        if (tid == 0) {
            uint32_t s = 0;
            for (int w = 0; w < kManThreads / 32; ++w) s += sPts[w];
            if (s) atomicAdd(pointCount, s);
        }
        // coalesced 128-bit stores of the block's records (base * 88 B is a multiple of 16 B)
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
