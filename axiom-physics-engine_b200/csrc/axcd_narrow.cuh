// Stage 3: narrowphase over the sorted candidate pairs — GJK distance/overlap query on convex
// "cores" (sphere = point + radius, box, hull) and EPA for overlapping cores — on FP32 CUDA cores.
//
// Frame: everything is expressed relative to body A's position (axes world-aligned) so magnitudes
// stay O(shape size).  Rotation enters as the matrix of Quat::toMatrix (reference:
// src/math/quat.cpp:117-129 -> glm::mat4_cast); placement follows
// Transform::transformPoint's scale -> rotate -> translate order (src/math/transform.cpp:86-93).
// Contact record = debug::DebugContactPoint (include/axiom/debug/physics_debug_draw.hpp:128-132).
// The arithmetic (operation order, tie rules, tolerances) is the contract the CPU oracle restates;
// the library is built -fmad=false so every product and sum rounds separately.
#pragma once

#include <cfloat>

#include "axcd_common.cuh"

namespace axcd {

enum { CORE_POINT = 0, CORE_BOX = 1, CORE_HULL = 2, CORE_SEGMENT = 3, CORE_CYLINDER = 4 };


// CYL: whether cylinder cores can occur.  Scenes without cylinders instantiate every GJK / EPA kernel with
// CYL = false, so the cylinder's support code (two divisions and a square root per call) costs them nothing —
// inlined, or even as an out-of-line call, it cost the hull-mix scene C2 8 % of its step.
template <bool CYL>
struct CoreT {
    static constexpr bool kCyl = CYL;
    int kind;
    V3 c;            // centre relative to A's position
    V3 e0, e1, e2;   // box: rotation columns * (halfExtent*scale); hull: rotation columns
    V3 s;            // hull: scale
    const float4* verts;
    uint32_t nv;
    float r;         // sphere radius
};

struct NarrowParams {
    uint32_t gjkMaxIters, epaMaxIters, epaMaxFaces;
    float gjkTol, epaTol;
    uint32_t wantDistances;
    uint32_t boxBoxGeneric;   // AXCD_FLAG_BOXBOX_GJK_EPA: box-box pairs through GJK/EPA instead of the SAT
};

struct BodyPose {
    V3 p;
    float4 q;
    V3 s;
};

__device__ __forceinline__ BodyPose loadPose(const float* __restrict__ xf, uint32_t i) {
    // 40-byte record, 8-byte aligned: five float2 loads
    const float2* f = reinterpret_cast<const float2*>(xf + (size_t)i * 10);
    const float2 a = __ldg(f), b = __ldg(f + 1), c = __ldg(f + 2), d = __ldg(f + 3), e = __ldg(f + 4);
    BodyPose t;
    t.p = mk3(a.x, a.y, b.x);
    t.q = make_float4(b.y, c.x, c.y, d.x);
    t.s = mk3(d.y, e.x, e.y);
    return t;
}

__device__ __forceinline__ void quatToColumns(float4 q, V3& c0, V3& c1, V3& c2) {
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    const float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    c0 = mk3(1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy));
    c1 = mk3(2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx));
    c2 = mk3(2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy));
}

using Core = CoreT<false>;

template <bool CYL = false>
__device__ __forceinline__ CoreT<CYL> makeCore(const BodyPose& t, uint4 sh, const float4* __restrict__ hull, V3 origin) {
    CoreT<CYL> k;
    k.c = t.p - origin;
    k.r = 0.0f;
    k.verts = nullptr;
    k.nv = 0;
    k.s = mk3(1.f, 1.f, 1.f);
    k.e0 = k.e1 = k.e2 = mk3(0.f, 0.f, 0.f);
    const float p0 = __uint_as_float(sh.y), p1 = __uint_as_float(sh.z), p2 = __uint_as_float(sh.w);
    if (sh.x == AXCD_SHAPE_SPHERE) {
        k.kind = CORE_POINT;
        k.r = p0;
    } else if (sh.x == AXCD_SHAPE_CAPSULE) {
        // segment core along the local Y axis (half length height/2, scaled by scale.y) + radius
        k.kind = CORE_SEGMENT;
        V3 c0, c1, c2;
        quatToColumns(t.q, c0, c1, c2);
        k.e0 = c1 * ((p1 * 0.5f) * t.s.y);
        k.r = p0;
    } else if (CYL && sh.x == AXCD_SHAPE_CYLINDER) {
        // radial frame E0, E2 (rotation columns x / z scaled by scale * radius) and half axis E1 (column y scaled by
        // scale.y * height / 2): a rim point is c +- E1 + E0 * cos + E2 * sin
        k.kind = CORE_CYLINDER;
        V3 c0, c1, c2;
        quatToColumns(t.q, c0, c1, c2);
        k.e0 = c0 * (t.s.x * p0);
        k.e1 = c1 * (t.s.y * (p1 * 0.5f));
        k.e2 = c2 * (t.s.z * p0);
    } else if (sh.x == AXCD_SHAPE_BOX) {
        k.kind = CORE_BOX;
        V3 c0, c1, c2;
        quatToColumns(t.q, c0, c1, c2);
        k.e0 = c0 * (p0 * t.s.x);
        k.e1 = c1 * (p1 * t.s.y);
        k.e2 = c2 * (p2 * t.s.z);
    } else {
        k.kind = CORE_HULL;
        quatToColumns(t.q, k.e0, k.e1, k.e2);
        k.s = t.s;
        k.verts = hull + sh.y;
        k.nv = sh.z;
    }
    return k;
}

// Cylinder support.  The rim direction is quantised to 4 x 8192 directions (a 32768-gon inscribed in the rim: chord
// error 5e-9 of the radius, below float resolution) so that a support point is named by 16 bits — sign of the E0
// part, sign of the E2 part, 13 bits of the |E2 share| in the diamond parametrisation |u| + |w| = 1, cap sign —
// and is rebuilt bit for bit from that id by cylinderPoint (same expression trees as oracle/axref.cpp).
template <bool CYL>
__device__ __forceinline__ uint32_t cylinderId(const CoreT<CYL>& k, V3 d) {
    const float a = dot3(d, k.e0), b = dot3(d, k.e2);
    const float s = fabsf(a) + fabsf(b);
    uint32_t m = 0;
    if (s > 0.0f) {
        const float w = fabsf(b) / s;
        m = (uint32_t)(w * 8191.0f + 0.5f);
        if (m > 8191u) m = 8191u;
    }
    return (!(a >= 0.0f) ? 1u : 0u) | (!(b >= 0.0f) ? 2u : 0u) | (m << 2) | (!(dot3(d, k.e1) >= 0.0f) ? 0x8000u : 0u);
}
template <bool CYL>
__device__ __forceinline__ V3 cylinderPoint(const CoreT<CYL>& k, uint32_t id) {
    const float w = (float)((id >> 2) & 8191u) / 8191.0f, u = 1.0f - w;
    const float inv = 1.0f / sqrtf(u * u + w * w);
    const float fx = (id & 1u) ? -(u * inv) : (u * inv), fz = (id & 2u) ? -(w * inv) : (w * inv);
    const V3 cap = (id & 0x8000u) ? -k.e1 : k.e1;
    return ((k.e0 * fx + cap) + k.e2 * fz) + k.c;
}

// Support point of a core in world-aligned direction d (any length).  `id` names the chosen
// vertex (box: one sign bit per axis, hull: vertex index) so the point can be rebuilt later with
// pointFromId instead of being stored; both produce the same floats (same operation sequence).
template <bool CYL>
__device__ __forceinline__ V3 support(const CoreT<CYL>& k, V3 d, uint32_t& id) {
    id = 0;
    if (k.kind == CORE_POINT) return k.c;
    if (CYL && k.kind == CORE_CYLINDER) {
        id = cylinderId(k, d);
        return cylinderPoint(k, id);
    }
    if (k.kind == CORE_SEGMENT) {
        const bool n0 = !(dot3(d, k.e0) >= 0.0f);
        id = n0 ? 1u : 0u;
        return k.c + (n0 ? -k.e0 : k.e0);
    }
    if (k.kind == CORE_BOX) {
        const bool n0 = !(dot3(d, k.e0) >= 0.0f), n1 = !(dot3(d, k.e1) >= 0.0f), n2 = !(dot3(d, k.e2) >= 0.0f);
        id = (n0 ? 1u : 0u) | (n1 ? 2u : 0u) | (n2 ? 4u : 0u);
        V3 p = k.c;
        p = p + (n0 ? -k.e0 : k.e0);
        p = p + (n1 ? -k.e1 : k.e1);
        p = p + (n2 ? -k.e2 : k.e2);
        return p;
    }
    // hull: local direction = scale * (R^T d); the first maximal vertex wins
    const V3 l = mk3(dot3(d, k.e0) * k.s.x, dot3(d, k.e1) * k.s.y, dot3(d, k.e2) * k.s.z);
    float4 bv = __ldg(k.verts);
    float bestDot = dot3(l, mk3(bv.x, bv.y, bv.z));
    for (uint32_t i = 1; i < k.nv; ++i) {
        const float4 v = __ldg(k.verts + i);
        const float di = dot3(l, mk3(v.x, v.y, v.z));
        if (di > bestDot) {
            bestDot = di;
            bv = v;
            id = i;
        }
    }
    const V3 lv = mk3(bv.x * k.s.x, bv.y * k.s.y, bv.z * k.s.z);
    return ((k.e0 * lv.x + k.e1 * lv.y) + k.e2 * lv.z) + k.c;
}

template <bool CYL>
__device__ __forceinline__ V3 pointFromId(const CoreT<CYL>& k, uint32_t id) {
    if (k.kind == CORE_POINT) return k.c;
    if (CYL && k.kind == CORE_CYLINDER) return cylinderPoint(k, id);
    if (k.kind == CORE_SEGMENT) return k.c + ((id & 1u) ? -k.e0 : k.e0);
    if (k.kind == CORE_BOX) {
        V3 p = k.c;
        p = p + ((id & 1u) ? -k.e0 : k.e0);
        p = p + ((id & 2u) ? -k.e1 : k.e1);
        p = p + ((id & 4u) ? -k.e2 : k.e2);
        return p;
    }
    const float4 bv = __ldg(k.verts + id);
    const V3 lv = mk3(bv.x * k.s.x, bv.y * k.s.y, bv.z * k.s.z);
    return ((k.e0 * lv.x + k.e1 * lv.y) + k.e2 * lv.z) + k.c;
}

// Point of the Minkowski difference A - B in direction d; id = idA | idB << 16.
template <bool CYL>
__device__ __forceinline__ V3 supportDiff(const CoreT<CYL>& A, const CoreT<CYL>& B, V3 d, uint32_t& id) {
    uint32_t ia, ib;
    const V3 a = support(A, d, ia);
    const V3 b = support(B, -d, ib);
    id = ia | (ib << 16);
    return a - b;
}

struct Simplex {
    V3 y[4];          // points of the Minkowski difference A - B
    uint32_t id[4];   // which vertex of A (low 16 bits) and of B (high 16 bits) made each point
    float lam[4];
    int n;
};

constexpr float kGjkEpsAbs2 = 1e-12f;
constexpr float kDegenerateEps = 1e-12f;

// Closest point to the origin on segment [a,b]; mask bit0 = a, bit1 = b.
__device__ __forceinline__ V3 closestSegment(V3 a, V3 b, float& la, float& lb, int& mask) {
    const V3 ab = b - a;
    float t = -dot3(a, ab);
    if (t <= 0.0f) {
        la = 1.0f; lb = 0.0f; mask = 1;
        return a;
    }
    const float denom = dot3(ab, ab);
    if (t >= denom) {
        la = 0.0f; lb = 1.0f; mask = 2;
        return b;
    }
    t = t / denom;
    la = 1.0f - t; lb = t; mask = 3;
    return a + ab * t;
}

// Closest point to the origin on triangle (a,b,c), Voronoi-region walk; mask bits a=1, b=2, c=4.
__device__ __forceinline__ V3 closestTriangleInline(V3 a, V3 b, V3 c, float& la, float& lb, float& lc, int& mask) {
    const V3 ab = b - a, ac = c - a;
    const float d1 = -dot3(ab, a), d2 = -dot3(ac, a);
    if (d1 <= 0.0f && d2 <= 0.0f) {
        la = 1.0f; lb = 0.0f; lc = 0.0f; mask = 1;
        return a;
    }
    const float d3 = -dot3(ab, b), d4 = -dot3(ac, b);
    if (d3 >= 0.0f && d4 <= d3) {
        la = 0.0f; lb = 1.0f; lc = 0.0f; mask = 2;
        return b;
    }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        const float v = d1 / (d1 - d3);
        la = 1.0f - v; lb = v; lc = 0.0f; mask = 3;
        return a + ab * v;
    }
    const float d5 = -dot3(ab, c), d6 = -dot3(ac, c);
    if (d6 >= 0.0f && d5 <= d6) {
        la = 0.0f; lb = 0.0f; lc = 1.0f; mask = 4;
        return c;
    }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        const float w = d2 / (d2 - d6);
        la = 1.0f - w; lb = 0.0f; lc = w; mask = 5;
        return a + ac * w;
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        la = 0.0f; lb = 1.0f - w; lc = w; mask = 6;
        return b + (c - b) * w;
    }
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    la = (1.0f - v) - w; lb = v; lc = w; mask = 7;
    return (a + ab * v) + ac * w;
}

// The same as a call, for the places that run once per pair (EPA set-up and witness points).
__device__ __noinline__ V3 closestTriangle(V3 a, V3 b, V3 c, float& la, float& lb, float& lc, int& mask) {
    return closestTriangleInline(a, b, c, la, lb, lc, mask);
}

// Origin strictly on the far side of plane (a,b,c) from d?  A flat tetrahedron counts as outside.
__device__ __forceinline__ bool originOutside(V3 a, V3 b, V3 c, V3 d) {
    const V3 n = cross3(b - a, c - a);
    const V3 ad = d - a;
    const float sp = -dot3(a, n);
    const float sd = dot3(ad, n);
    if (sd * sd <= kDegenerateEps * (dot3(n, n) * dot3(ad, ad))) return true;
    return sp * sd < 0.0f;
}

// Reduce the simplex to the feature closest to the origin; returns that point, sets lam[].
// Every index into the simplex arrays is a compile-time constant (selects instead of s.y[f], a shift network
// instead of s.y[n++] = s.y[i]) so the simplex stays in registers, and the triangle code is instantiated once:
// a triangle is the tetrahedron loop's face 0 without the sidedness test.
__device__ __forceinline__ V3 solveSimplex(Simplex& s, bool& enclosed) {
    enclosed = false;
    int mask = 0;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    V3 v = mk3(0.f, 0.f, 0.f);
    if (s.n == 2) {
        v = closestSegment(s.y[0], s.y[1], l0, l1, mask);
    } else {
        const bool tetra = s.n == 4;
        const int nFaces = tetra ? 4 : 1;
        float best = FLT_MAX;
        bool any = !tetra;
        // faces (i,j,k | opposite o): (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0), examined in this order
#pragma unroll 1
        for (int f = 0; f < nFaces; ++f) {
            const V3 a = (f == 3) ? s.y[1] : s.y[0];
            const V3 b = (f == 0) ? s.y[1] : ((f == 1) ? s.y[2] : s.y[3]);
            const V3 c = (f == 0) ? s.y[2] : ((f == 1) ? s.y[3] : ((f == 2) ? s.y[1] : s.y[2]));
            if (tetra) {
                const V3 o = (f == 0) ? s.y[3] : ((f == 1) ? s.y[1] : ((f == 2) ? s.y[2] : s.y[0]));
                if (!originOutside(a, b, c, o)) continue;
                any = true;
            }
            float li, lj, lk;
            int m;
            const V3 q = closestTriangleInline(a, b, c, li, lj, lk, m);
            const float qq = dot3(q, q);
            if (!tetra || qq < best) {
                best = qq;
                v = q;
                const int b0 = m & 1, b1 = (m >> 1) & 1, b2 = (m >> 2) & 1;
                if (f == 0) {
                    l0 = li; l1 = lj; l2 = lk; l3 = 0.f;
                    mask = m;
                } else if (f == 1) {
                    l0 = li; l1 = 0.f; l2 = lj; l3 = lk;
                    mask = b0 | (b1 << 2) | (b2 << 3);
                } else if (f == 2) {
                    l0 = li; l1 = lk; l2 = 0.f; l3 = lj;
                    mask = b0 | (b1 << 3) | (b2 << 1);
                } else {
                    l0 = 0.f; l1 = li; l2 = lk; l3 = lj;
                    mask = (b0 << 1) | (b1 << 3) | (b2 << 2);
                }
            }
        }
        if (!any) {
            enclosed = true;
            return mk3(0.f, 0.f, 0.f);
        }
    }
    s.lam[0] = l0; s.lam[1] = l1; s.lam[2] = l2; s.lam[3] = l3;
    mask &= (1 << s.n) - 1;
    // keep the points whose bit is set, in order: drop from the top so lower indices stay put
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        if (!((mask >> i) & 1)) {
#pragma unroll
            for (int j = i; j < 3; ++j) {
                s.y[j] = s.y[j + 1];
                s.id[j] = s.id[j + 1];
                s.lam[j] = s.lam[j + 1];
            }
        }
    }
    s.n = __popc((uint32_t)mask);
    return v;
}

enum { GJK_SEPARATED = 0, GJK_OVERLAP = 1 };
struct GjkResult {
    int state;
    bool exact;
    V3 v;
    float vv;
    uint32_t status;
};

template <bool CYL>
__device__ __forceinline__ GjkResult gjk(const CoreT<CYL>& A, const CoreT<CYL>& B, const NarrowParams& cfg, float marginSum,
                                         Simplex& s) {
    GjkResult r;
    r.status = 0;
    V3 d0 = B.c - A.c;
    if (dot3(d0, d0) < 1e-12f) d0 = mk3(1.0f, 0.0f, 0.0f);
    s.y[0] = supportDiff(A, B, d0, s.id[0]);
    s.lam[0] = 1.0f;
    s.n = 1;
    V3 v = s.y[0];
    float vv = dot3(v, v);
    r.state = GJK_SEPARATED;
    r.exact = true;
    for (uint32_t it = 0;; ++it) {
        if (vv <= kGjkEpsAbs2) {
            r.state = GJK_OVERLAP;
            break;
        }
        if (it >= cfg.gjkMaxIters) {
            r.status = AXCD_ERR_GJK_NO_CONVERGE;
            break;
        }
        uint32_t wid;
        const V3 w = supportDiff(A, B, -v, wid);
        const float vw = dot3(v, w);
        if (!cfg.wantDistances && vw > 0.0f && vw * vw > vv * (marginSum * marginSum)) {
            r.exact = false;   // separating axis with a gap larger than the radii
            break;
        }
        if (vv - vw <= cfg.gjkTol * vv) break;
        bool dup = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) dup = dup || (i < s.n && same3(w, s.y[i]));
        if (dup) break;
#pragma unroll
        for (int i = 1; i < 4; ++i)   // s.n is 1..3 here (solveSimplex never leaves four points)
            if (s.n == i) {
                s.y[i] = w;
                s.id[i] = wid;
            }
        s.n++;
        bool enclosed;
        const V3 nv = solveSimplex(s, enclosed);
        if (enclosed) {
            r.state = GJK_OVERLAP;
            v = nv;
            vv = 0.0f;
            break;
        }
        const float nvv = dot3(nv, nv);
        const bool stalled = nvv >= vv;
        v = nv;
        vv = nvv;
        if (stalled) {
            if (vv <= kGjkEpsAbs2) r.state = GJK_OVERLAP;
            break;
        }
    }
    r.v = v;
    r.vv = vv;
    return r;
}

// ---- sphere against box, closed form ------------------------------------------------------------------
// The box as in makeCore: centre cX, unit axes = the columns of Quat::toMatrix, half lengths
// |halfExtent * scale|; all points relative to A's position.
//   outside: closest box point q by clamping the centre's box coordinates; gap = |q - cS| - r
//   inside : leave through the nearest face (lowest axis on ties); depth = face gap + r
// nsx = unit direction from the sphere towards the box, ps / px = witness points on the sphere / box.
struct SphereBox {
    bool contact;
    float dist, depth;
    V3 nsx, ps, px;
};
__device__ __forceinline__ SphereBox sphereBox(V3 cS, float r, V3 cX, const BodyPose& tX, uint4 sX) {
    SphereBox o;
    V3 ax[3];
    quatToColumns(tX.q, ax[0], ax[1], ax[2]);
    const float half[3] = {fabsf(__uint_as_float(sX.y) * tX.s.x), fabsf(__uint_as_float(sX.z) * tX.s.y),
                           fabsf(__uint_as_float(sX.w) * tX.s.z)};
    const V3 d = cS - cX;
    float x[3];
    bool inside = true;
    V3 q = cX;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        x[i] = dot3(d, ax[i]);
        float k = x[i];
        if (k > half[i]) {
            k = half[i];
            inside = false;
        } else if (k < -half[i]) {
            k = -half[i];
            inside = false;
        }
        q = q + ax[i] * k;
    }
    if (!inside) {
        const V3 v = q - cS;
        const float l = sqrtf(dot3(v, v));
        o.dist = l - r;
        o.depth = r - l;
        o.contact = o.depth >= 0.0f;
        o.nsx = (l > 0.0f) ? v * (1.0f / l) : mk3(1.0f, 0.0f, 0.0f);
        o.ps = cS + o.nsx * r;
        o.px = q;
        return o;
    }
    int best = 0;
    float gap = half[0] - fabsf(x[0]);
#pragma unroll
    for (int i = 1; i < 3; ++i) {
        const float g = half[i] - fabsf(x[i]);
        if (g < gap) {
            gap = g;
            best = i;
        }
    }
    const V3 axb = (best == 0) ? ax[0] : ((best == 1) ? ax[1] : ax[2]);
    const float xb = (best == 0) ? x[0] : ((best == 1) ? x[1] : x[2]);
    const V3 out = (xb >= 0.0f) ? axb : -axb;   // from the box centre's side towards the sphere
    o.contact = true;
    o.depth = gap + r;
    o.dist = -o.depth;
    o.nsx = -out;
    o.ps = cS + o.nsx * r;
    o.px = cS + out * gap;
    return o;
}

// ---- box against box, closed form: the 15-axis separating-axis test -------------------------------------
// (3 face normals of each box + 9 edge-edge cross products), as collision libraries dispatch this pair
// class.  The Minkowski difference of two boxes is a polytope whose face normals are among those 15
// directions, so the boxes are apart iff some axis has a negative overlap, and otherwise the penetration
// depth is the smallest overlap and the contact normal that axis — the answer EPA converges to, at a few
// hundred flops instead of 10-20 kflop.  Edge axes are compared un-normalised (cross-multiplied), only the
// winner pays a square root; nearly parallel edge pairs (|a_i x b_j|^2 <= 1e-5) are skipped.  Ties keep the
// earlier axis.  Witness points: face axis -> the other box's deepest vertex and its projection on the
// face; edge-edge -> closest points of the two supporting edges (as segments).  Same expression trees as
// the CPU oracle (oracle/axref.cpp boxBox).
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

constexpr float kSatParallelEps = 1e-5f;
struct BoxBox {
    bool contact;
    float depth;
    V3 n, pa, pb;   // unit normal from A to B, witness points on A and on B
};
__device__ __forceinline__ BoxBox boxBox(const BoxFrame& A, const BoxFrame& B) {
    BoxBox o;
    o.contact = false;
    o.depth = 0.0f;
    o.n = o.pa = o.pb = mk3(0.f, 0.f, 0.f);
    const V3 t = B.c - A.c;
    float R[3][3], AR[3][3], tA[3], tB[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        tA[i] = dot3(t, A.ax[i]);
        tB[i] = dot3(t, B.ax[i]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R[i][j] = dot3(A.ax[i], B.ax[j]);
            AR[i][j] = fabsf(R[i][j]);
        }
    }
    float bestOv = FLT_MAX, bestL2 = 1.0f;
    int axis = -1;
    bool apart = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // faces of A
        const float rb = (B.h[0] * AR[i][0] + B.h[1] * AR[i][1]) + B.h[2] * AR[i][2];
        const float ov = (A.h[i] + rb) - fabsf(tA[i]);
        apart = apart || (ov < 0.0f);
        if (ov < bestOv) {
            bestOv = ov;
            axis = i;
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // faces of B
        const float ra = (A.h[0] * AR[0][j] + A.h[1] * AR[1][j]) + A.h[2] * AR[2][j];
        const float ov = (ra + B.h[j]) - fabsf(tB[j]);
        apart = apart || (ov < 0.0f);
        if (ov < bestOv) {
            bestOv = ov;
            axis = 3 + j;
        }
    }
    if (apart || axis < 0) return o;   // a separating face axis, or NaN input (every compare false)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const V3 L = cross3(A.ax[i], B.ax[j]);
            const float l2 = dot3(L, L);
            const float ra = A.h[i1] * AR[i2][j] + A.h[i2] * AR[i1][j];
            const float rb = B.h[j1] * AR[i][j2] + B.h[j2] * AR[i][j1];
            const float ov = (ra + rb) - fabsf(dot3(t, L));   // scaled by |L|
            if (l2 > kSatParallelEps) {
                apart = apart || (ov < 0.0f);
                // ov / |L| < bestOv / |bestL|, cross-multiplied (all terms >= 0 unless already apart)
                if (!apart && (ov * ov) * bestL2 < (bestOv * bestOv) * l2) {
                    bestOv = ov;
                    bestL2 = l2;
                    axis = 6 + 3 * i + j;
                }
            }
        }
    }
    if (apart) return o;
    o.contact = true;
    if (axis < 3) {
        const int i = axis;
        const float tAi = (i == 0) ? tA[0] : ((i == 1) ? tA[1] : tA[2]);
        const V3 axi = (i == 0) ? A.ax[0] : ((i == 1) ? A.ax[1] : A.ax[2]);
        const bool pos = tAi >= 0.0f;
        o.n = pos ? axi : -axi;
        o.depth = bestOv;
        V3 v = B.c;   // B's deepest vertex: support of B in direction -n
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float rij = (i == 0) ? R[0][j] : ((i == 1) ? R[1][j] : R[2][j]);
            const float dj = pos ? rij : -rij;
            v = v + B.ax[j] * ((dj > 0.0f) ? -B.h[j] : B.h[j]);
        }
        o.pb = v;
        o.pa = v + o.n * o.depth;
    } else if (axis < 6) {
        const int j = axis - 3;
        const float tBj = (j == 0) ? tB[0] : ((j == 1) ? tB[1] : tB[2]);
        const V3 axj = (j == 0) ? B.ax[0] : ((j == 1) ? B.ax[1] : B.ax[2]);
        const bool pos = tBj >= 0.0f;
        o.n = pos ? axj : -axj;
        o.depth = bestOv;
        V3 v = A.c;   // A's deepest vertex: support of A in direction n
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float rij = (j == 0) ? R[i][0] : ((j == 1) ? R[i][1] : R[i][2]);
            const float di = pos ? rij : -rij;
            v = v + A.ax[i] * ((di >= 0.0f) ? A.h[i] : -A.h[i]);
        }
        o.pa = v;
        o.pb = v - o.n * o.depth;
    } else {
        const int i = (axis - 6) / 3, j = (axis - 6) % 3;
        const V3 u = (i == 0) ? A.ax[0] : ((i == 1) ? A.ax[1] : A.ax[2]);
        const V3 v = (j == 0) ? B.ax[0] : ((j == 1) ? B.ax[1] : B.ax[2]);
        const float hu = (i == 0) ? A.h[0] : ((i == 1) ? A.h[1] : A.h[2]);
        const float hv = (j == 0) ? B.h[0] : ((j == 1) ? B.h[1] : B.h[2]);
        const V3 L = cross3(u, v);
        const float invl = 1.0f / sqrtf(bestL2);
        const bool pos = dot3(t, L) >= 0.0f;
        o.n = L * (pos ? invl : -invl);
        o.depth = bestOv * invl;
        // centres of the two supporting edges
        V3 ea = A.c, eb = B.c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k != i) ea = ea + A.ax[k] * ((dot3(o.n, A.ax[k]) >= 0.0f) ? A.h[k] : -A.h[k]);
            if (k != j) eb = eb + B.ax[k] * ((dot3(o.n, B.ax[k]) > 0.0f) ? -B.h[k] : B.h[k]);
        }
        // closest points of the segments ea + u*s (|s| <= hu) and eb + v*q (|q| <= hv)
        const V3 r = ea - eb;
        const float a = dot3(u, u), e = dot3(v, v), b = dot3(u, v);
        const float c = dot3(u, r), f = dot3(v, r);
        const float denom = a * e - b * b;   // > 0: the edges are not parallel
        float s = (b * f - c * e) / denom;
        s = fminf(fmaxf(s, -hu), hu);
        float q = (b * s + f) / e;
        if (q < -hv) {
            q = -hv;
            s = fminf(fmaxf((b * q - c) / a, -hu), hu);
        } else if (q > hv) {
            q = hv;
            s = fminf(fmaxf((b * q - c) / a, -hu), hu);
        }
        o.pa = ea + u * s;
        o.pb = eb + v * q;
    }
    return o;
}

// ---- EPA ------------------------------------------------------------------------------------------
// The polytope lives in per-thread storage addressed as base[word * STRIDE]:
//   STRIDE = block size  -> a column of shared memory (the same word of all threads is contiguous,
//                           so uniform-index accesses are conflict-free)
//   STRIDE = 1           -> a private array (fallback kernel with the full caps)
// words: [0, 3V) vertex y;  [3V, 4V) vertex ids;  [4V, 4V+4F) face (n.x,n.y,n.z,d);
//        [4V+4F, 4V+5F) face indices i0 | i1<<8 | i2<<16;  then E/2 words of packed horizon edges.
constexpr int kEpaHardVerts = 40;   // full caps (fallback path) == the oracle's
constexpr int kEpaHardFaces = 64;
// shared-memory fast path; a pair that outgrows it is spilled and finished by the fallback kernel.
// Sized from the measured distribution of final polytope sizes (profiles/epa_polytope_hist.py).
#ifndef AXCD_EPA_FAST_VERTS
#define AXCD_EPA_FAST_VERTS 14   // 99.2 % of the headline's EPA runs end with <= 14 vertices / 24 faces
#define AXCD_EPA_FAST_FACES 24
#define AXCD_EPA_FAST_EDGES 20
#endif
constexpr int kEpaFastVerts = AXCD_EPA_FAST_VERTS;
constexpr int kEpaFastFaces = AXCD_EPA_FAST_FACES;
constexpr int kEpaFastEdges = AXCD_EPA_FAST_EDGES;
static_assert(kEpaFastVerts <= 16 && kEpaFastFaces <= 32 && kEpaFastEdges % 2 == 0, "fast-path caps");

template <int MAXF> struct FaceMask { using T = std::conditional_t<(MAXF <= 32), uint32_t, uint64_t>; };
__device__ __forceinline__ int lowestBit(uint32_t m) { return __ffs((int)m) - 1; }
__device__ __forceinline__ int lowestBit(uint64_t m) { return __ffsll((long long)m) - 1; }
__device__ __forceinline__ int popCount(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int popCount(uint64_t m) { return __popcll(m); }

template <int MAXV, int MAXF, int MAXE, int STRIDE>
struct Poly {
    using Mask = typename FaceMask<MAXF>::T;
    // per-vertex row of "visible directed edge a->b" bits: 16-bit rows (two per word) when
    // MAXV <= 16, else 64-bit rows (two words)
    static constexpr bool kSmallRows = MAXV <= 16;
    static constexpr int kRowWords = kSmallRows ? (MAXV + 1) / 2 : 2 * MAXV;
    static constexpr int kRowBase = 4 * MAXV + 5 * MAXF + MAXE / 2;
    static constexpr int kWords = kRowBase + kRowWords;
    float* base;
    __device__ __forceinline__ float& w(int i) const { return base[i * STRIDE]; }
    __device__ __forceinline__ uint32_t& u(int i) const { return reinterpret_cast<uint32_t*>(base)[i * STRIDE]; }
    __device__ __forceinline__ V3 y(int i) const { return mk3(w(3 * i), w(3 * i + 1), w(3 * i + 2)); }
    __device__ __forceinline__ void setY(int i, V3 v) const { w(3 * i) = v.x; w(3 * i + 1) = v.y; w(3 * i + 2) = v.z; }
    __device__ __forceinline__ uint32_t id(int i) const { return u(3 * MAXV + i); }
    __device__ __forceinline__ void setId(int i, uint32_t v) const { u(3 * MAXV + i) = v; }
    __device__ __forceinline__ V3 fn(int f) const { return mk3(w(4 * MAXV + 4 * f), w(4 * MAXV + 4 * f + 1), w(4 * MAXV + 4 * f + 2)); }
    __device__ __forceinline__ float fd(int f) const { return w(4 * MAXV + 4 * f + 3); }
    __device__ __forceinline__ void setPlane(int f, V3 n, float d) const {
        w(4 * MAXV + 4 * f) = n.x; w(4 * MAXV + 4 * f + 1) = n.y; w(4 * MAXV + 4 * f + 2) = n.z; w(4 * MAXV + 4 * f + 3) = d;
    }
    __device__ __forceinline__ uint32_t fi(int f) const { return u(4 * MAXV + 4 * MAXF + f); }
    __device__ __forceinline__ void setFi(int f, uint32_t v) const { u(4 * MAXV + 4 * MAXF + f) = v; }
    __device__ __forceinline__ uint32_t edge(int h) const {
        const uint32_t p = u(4 * MAXV + 5 * MAXF + (h >> 1));
        return (h & 1) ? (p >> 16) : (p & 0xffffu);
    }
    __device__ __forceinline__ void setEdge(int h, uint32_t e) const {
        uint32_t& r = u(4 * MAXV + 5 * MAXF + (h >> 1));
        r = (h & 1) ? ((r & 0xffffu) | (e << 16)) : ((r & 0xffff0000u) | e);
    }
    __device__ __forceinline__ uint32_t& edgeWord(int h) const { return u(4 * MAXV + 5 * MAXF + h); }   // unpacked use
    __device__ __forceinline__ void clearRows(int nv) const {
        if (kSmallRows) {
            for (int i = 0; i < (nv + 1) / 2; ++i) u(kRowBase + i) = 0u;
        } else {
            for (int i = 0; i < 2 * nv; ++i) u(kRowBase + i) = 0u;
        }
    }
    __device__ __forceinline__ void setEdgeBit(uint32_t a, uint32_t b) const {
        if (kSmallRows) u(kRowBase + (a >> 1)) |= 1u << (b + 16 * (a & 1));
        else u(kRowBase + 2 * a + (b >> 5)) |= 1u << (b & 31);
    }
    __device__ __forceinline__ bool edgeBit(uint32_t a, uint32_t b) const {
        if (kSmallRows) return (u(kRowBase + (a >> 1)) >> (b + 16 * (a & 1))) & 1u;
        return (u(kRowBase + 2 * a + (b >> 5)) >> (b & 31)) & 1u;
    }
};

template <class P>
__device__ __forceinline__ void epaSetFace(const P& e, int slot, int i0, int i1, int i2) {
    const V3 p0 = e.y(i0);
    V3 n = cross3(e.y(i1) - p0, e.y(i2) - p0);
    const float len2 = dot3(n, n);
    e.setFi(slot, (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)i2 << 16));
    if (len2 <= 1e-30f) {   // zero-area face: never the closest, never visible
        e.setPlane(slot, mk3(0.f, 0.f, 0.f), FLT_MAX);
        return;
    }
    const float inv = 1.0f / sqrtf(len2);
    n = n * inv;
    e.setPlane(slot, n, dot3(n, p0));
}

struct EpaResult {
    V3 n;
    float depth;
    V3 pa, pb;
    uint32_t status;
    bool overflow;   // fast-path caps exceeded: rerun with the full caps
};

__device__ __forceinline__ EpaResult epaTouching(V3 n, V3 pa) {
    EpaResult r;
    const float l2 = dot3(n, n);
    r.n = (l2 > 0.0f) ? n * (1.0f / sqrtf(l2)) : mk3(1.0f, 0.0f, 0.0f);
    r.depth = 0.0f;
    r.pa = pa;
    r.pb = pa;
    r.status = 0;
    r.overflow = false;
    return r;
}

// EPA from a GJK end simplex (n0 points y0[], ids id0[]), split into init / iterate / finish so a
// persistent lane can interleave pairs.  MAXV/MAXF/MAXE are storage caps; the algorithmic caps
// (cfg.epaMaxFaces, kEpaHardVerts, cfg.epaMaxIters) give status 302, while hitting a smaller
// storage cap sets `overflow` and the caller reruns the pair on the full-cap path.
template <class Mask>
struct EpaState {
    int nv, nf, best;  // best = closest alive face (lowest slot on ties), bd its plane distance
    float bd;
    Mask alive;        // bit f set = face slot f is part of the polytope
    uint32_t it, status;
    bool overflow, degenerate;
};

// Grows the simplex to an oriented tetrahedron.  Returns 0 when the expansion can start, 1 when the
// Minkowski difference is degenerate at the origin and `touching` already is the answer.
template <int MAXV, int MAXF, int MAXE, int STRIDE, bool CYL>
__device__ __forceinline__ int epaInit(const CoreT<CYL>& A, const CoreT<CYL>& B, int n0, const V3* y0, const uint32_t* id0,
                                       const Poly<MAXV, MAXF, MAXE, STRIDE>& e,
                                       EpaState<typename Poly<MAXV, MAXF, MAXE, STRIDE>::Mask>& st,
                                       EpaResult& touching) {
    int nv = n0;
    for (int i = 0; i < 4; ++i)
        if (i < n0) {
            e.setY(i, y0[i]);
            e.setId(i, id0[i]);
        }
    uint32_t tid_;
    // ---- grow the GJK simplex to a tetrahedron -------------------------------------------------
    if (nv == 1) {
        for (int k = 0; k < 6 && nv == 1; ++k) {
            const float sg = (k & 1) ? -1.0f : 1.0f;
            const V3 ax = mk3((k >> 1) == 0 ? sg : 0.f, (k >> 1) == 1 ? sg : 0.f, (k >> 1) == 2 ? sg : 0.f);
            const V3 y1 = supportDiff(A, B, ax, tid_);
            e.setY(1, y1);
            e.setId(1, tid_);
            const V3 d = y1 - e.y(0);
            if (dot3(d, d) > kDegenerateEps) nv = 2;
        }
        if (nv == 1) { touching = epaTouching(mk3(1.f, 0.f, 0.f), pointFromId(A, e.id(0) & 0xffffu)); return 1; }
    }
    if (nv == 2) {
        const V3 d = e.y(1) - e.y(0);
        V3 firstDir = mk3(0.f, 0.f, 0.f);
        for (int k = 0; k < 3 && nv == 2; ++k) {
            const V3 ax = mk3(k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f);
            const V3 dir = cross3(d, ax);
            if (dot3(dir, dir) <= kDegenerateEps) continue;
            if (dot3(firstDir, firstDir) == 0.0f) firstDir = dir;
            for (int sgn = 0; sgn < 2 && nv == 2; ++sgn) {
                const V3 y2 = supportDiff(A, B, sgn ? -dir : dir, tid_);
                e.setY(2, y2);
                e.setId(2, tid_);
                const V3 c = cross3(y2 - e.y(0), d);
                if (dot3(c, c) > kDegenerateEps) nv = 3;
            }
        }
        if (nv == 2) { touching = epaTouching(firstDir, pointFromId(A, e.id(0) & 0xffffu)); return 1; }
    }
    if (nv == 3) {
        const V3 n = cross3(e.y(1) - e.y(0), e.y(2) - e.y(0));
        const float n2 = dot3(n, n);
        if (n2 <= 1e-30f) { touching = epaTouching(mk3(1.f, 0.f, 0.f), pointFromId(A, e.id(0) & 0xffffu)); return 1; }
        for (int sgn = 0; sgn < 2 && nv == 3; ++sgn) {
            const V3 y3 = supportDiff(A, B, sgn ? -n : n, tid_);
            e.setY(3, y3);
            e.setId(3, tid_);
            const float vol = dot3(y3 - e.y(0), n);
            if (vol * vol > kDegenerateEps * n2) nv = 4;
        }
        if (nv == 3) {   // flat at the origin: touching contact along +n
            float la, lb, lc;
            int m;
            closestTriangle(e.y(0), e.y(1), e.y(2), la, lb, lc, m);
            const V3 pa = (pointFromId(A, e.id(0) & 0xffffu) * la + pointFromId(A, e.id(1) & 0xffffu) * lb) +
                          pointFromId(A, e.id(2) & 0xffffu) * lc;
            { touching = epaTouching(n, pa); return 1; }
        }
    }
    // orientation: make (0,1,2) face away from vertex 3
    if (dot3(cross3(e.y(1) - e.y(0), e.y(2) - e.y(0)), e.y(3) - e.y(0)) > 0.0f) {
        const V3 t = e.y(0);
        e.setY(0, e.y(1));
        e.setY(1, t);
        const uint32_t ti = e.id(0);
        e.setId(0, e.id(1));
        e.setId(1, ti);
    }
    epaSetFace(e, 0, 0, 1, 2);
    epaSetFace(e, 1, 0, 3, 1);
    epaSetFace(e, 2, 0, 2, 3);
    epaSetFace(e, 3, 1, 3, 2);
    st.nv = nv;
    st.nf = 4;
    st.best = -1;
    st.bd = FLT_MAX;
    for (int i = 0; i < 4; ++i) {
        const float di = e.fd(i);
        if (di < st.bd) {
            st.bd = di;
            st.best = i;
        }
    }
    st.alive = 0xf;
    st.it = 0;
    st.status = 0;
    st.overflow = false;
    st.degenerate = false;
    return 0;
}

// The per-face / per-vertex scans of one expansion step are independent iterations over shared
// memory: unrolled by 4 they give the latency-bound EPA kernel (9 warps per SM) some ILP.  Measured
// (profiles/r01_experiments.md): 1 -> 0.85 ms, 4 -> 0.80 ms, 8 -> 0.87 ms; unrolling the face
// construction loop (sqrt + divide chains) costs more in divergence than it hides.
#ifndef AXCD_EPA_UNROLL
#define AXCD_EPA_UNROLL 4
#endif
#define AXCD_STR_(x) #x
#define AXCD_UNROLL(n) _Pragma(AXCD_STR_(unroll n))
// One expansion step.  Returns true when the pair is finished (converged, capped or overflowed).
template <int MAXV, int MAXF, int MAXE, int STRIDE, bool CYL>
__device__ __forceinline__ bool epaIterate(const CoreT<CYL>& A, const CoreT<CYL>& B, const NarrowParams& cfg,
                                           const Poly<MAXV, MAXF, MAXE, STRIDE>& e,
                                           EpaState<typename Poly<MAXV, MAXF, MAXE, STRIDE>::Mask>& st) {
    using Mask = typename Poly<MAXV, MAXF, MAXE, STRIDE>::Mask;
    const Mask one = 1;
    const int maxFaces = (int)min(cfg.epaMaxFaces, (uint32_t)kEpaHardFaces);
    int& nv = st.nv;
    int& nf = st.nf;
    int& best = st.best;
    Mask& alive = st.alive;
    const float bd = st.bd;   // closest face: found while the previous step scanned the faces
    if (best < 0) {   // every face degenerate: give up on this polytope
        st.degenerate = true;
        return true;
    }
    const V3 bn = e.fn(best);
    uint32_t wid;
    const V3 w = supportDiff(A, B, bn, wid);
    const float dw = dot3(w, bn);
    const float scale = (bd > 1.0f) ? bd : 1.0f;
    if (dw - bd <= cfg.epaTol * scale) return true;
    bool dup = false;
    AXCD_UNROLL(AXCD_EPA_UNROLL)
    for (int i = 0; i < nv; ++i) dup = dup || same3(w, e.y(i));
    if (dup) return true;
    if (st.it >= cfg.epaMaxIters || nv >= kEpaHardVerts) {
        st.status = AXCD_ERR_EPA_NO_CONVERGE;
        return true;
    }
    // visible faces: w clearly in front
    const float wl = fabsf(w.x) + fabsf(w.y) + fabsf(w.z);
    const float visEps = 1e-6f * ((wl > 1.0f) ? wl : 1.0f);
    Mask vis = 0;
    int nbest = -1;          // closest face among those that survive this step
    float nbd = FLT_MAX;
    AXCD_UNROLL(AXCD_EPA_UNROLL)
    for (int i = 0; i < nf; ++i) {
        const float di = e.fd(i);
        const bool v = dot3(e.fn(i), w) - di > visEps;
        if ((alive >> i) & one) {
            if (v) {
                vis |= one << i;
            } else if (di < nbd) {
                nbd = di;
                nbest = i;
            }
        }
    }
    // directed edges of the visible faces, as one bit row per start vertex
    e.clearRows(nv);
    for (Mask m = vis; m; m &= m - 1) {
        const uint32_t fi = e.fi(lowestBit(m));
        const uint32_t v0 = fi & 0xffu, v1 = (fi >> 8) & 0xffu, v2 = (fi >> 16) & 0xffu;
        e.setEdgeBit(v0, v1);
        e.setEdgeBit(v1, v2);
        e.setEdgeBit(v2, v0);
    }
    // horizon in canonical order: visible faces by ascending slot, edges in winding order, an
    // edge is kept iff its reverse is not an edge of a visible face
    int nh = 0;
    uint64_t starts = 0, ends = 0;
    bool loopOk = true, edgeOverflow = false;
    for (Mask m = vis; m; m &= m - 1) {
        const uint32_t fi = e.fi(lowestBit(m));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t ea = (fi >> (8 * k)) & 0xffu;
            const uint32_t eb = (fi >> (8 * ((k + 1) % 3))) & 0xffu;
            if (e.edgeBit(eb, ea)) continue;   // shared by two visible faces
            if (((starts >> ea) & 1ull) || ((ends >> eb) & 1ull)) loopOk = false;
            starts |= 1ull << ea;
            ends |= 1ull << eb;
            if (nh < MAXE) e.setEdge(nh, ea | (eb << 8));
            else edgeOverflow = true;
            ++nh;
        }
    }
    if (nh < 3) loopOk = false;
    const int nalive = popCount(alive), nvis = popCount(vis);
    if (!loopOk || nalive - nvis + nh > maxFaces) {
        st.status = AXCD_ERR_EPA_NO_CONVERGE;
        return true;
    }
    if (edgeOverflow || nv >= MAXV || nalive - nvis + nh > MAXF) {   // storage caps of this path
        st.overflow = true;
        return true;
    }
    const int wi = nv;
    e.setY(wi, w);
    e.setId(wi, wid);
    nv++;
    alive &= ~vis;
    for (int h = 0; h < nh; ++h) {
        const int slot = lowestBit((Mask)~alive);   // lowest free slot
        const uint32_t ed = e.edge(h);
        epaSetFace(e, slot, (int)(ed & 0xffu), (int)(ed >> 8), wi);
        alive |= one << slot;
        nf = max(nf, slot + 1);
        const float di = e.fd(slot);   // the closest face is the lowest slot among the minima
        if (di < nbd || (di == nbd && slot < nbest)) {
            nbd = di;
            nbest = slot;
        }
    }
    best = nbest;
    st.bd = nbd;
    st.it++;
    return false;
}

template <int MAXV, int MAXF, int MAXE, int STRIDE, bool CYL>
__device__ __forceinline__ EpaResult epaFinish(const CoreT<CYL>& A, const Poly<MAXV, MAXF, MAXE, STRIDE>& e,
                                               const EpaState<typename Poly<MAXV, MAXF, MAXE, STRIDE>::Mask>& st) {
    if (st.degenerate) return epaTouching(mk3(1.f, 0.f, 0.f), pointFromId(A, e.id(0) & 0xffffu));
    EpaResult r;
    r.overflow = st.overflow;
    r.status = st.status;
    if (st.overflow) {
        r.n = mk3(1.f, 0.f, 0.f);
        r.depth = 0.f;
        r.pa = r.pb = mk3(0.f, 0.f, 0.f);
        return r;
    }
    const int best = st.best;
    const uint32_t fi = e.fi(best);
    const int i0 = fi & 0xffu, i1 = (fi >> 8) & 0xffu, i2 = (fi >> 16) & 0xffu;
    const float fdist = e.fd(best);
    r.n = e.fn(best);
    r.depth = (fdist > 0.0f) ? fdist : 0.0f;
    float la, lb, lc;
    int m;
    const V3 p = closestTriangle(e.y(i0), e.y(i1), e.y(i2), la, lb, lc, m);
    r.pa = (pointFromId(A, e.id(i0) & 0xffffu) * la + pointFromId(A, e.id(i1) & 0xffffu) * lb) +
           pointFromId(A, e.id(i2) & 0xffffu) * lc;
    r.pb = r.pa - p;
    return r;
}

template <int MAXV, int MAXF, int MAXE, int STRIDE, bool CYL>
__device__ __forceinline__ EpaResult epaRun(const CoreT<CYL>& A, const CoreT<CYL>& B, const NarrowParams& cfg, int n0,
                                            const V3* y0, const uint32_t* id0,
                                            const Poly<MAXV, MAXF, MAXE, STRIDE>& e) {
    EpaState<typename Poly<MAXV, MAXF, MAXE, STRIDE>::Mask> st;
    EpaResult touching;
    if (epaInit(A, B, n0, y0, id0, e, st, touching)) return touching;
    while (!epaIterate(A, B, cfg, e, st)) {
    }
    return epaFinish(A, e, st);
}

// ---- contact record helpers ----------------------------------------------------------------------
__device__ __forceinline__ void storeContact(AxcdContact* __restrict__ dst, uint32_t a, uint32_t b, V3 pos, V3 n,
                                             float depth, uint32_t status) {
    float2* o = reinterpret_cast<float2*>(dst);   // 40-byte record, 8-byte aligned
    o[0] = make_float2(__uint_as_float(a), __uint_as_float(b));
    o[1] = make_float2(pos.x, pos.y);
    o[2] = make_float2(pos.z, n.x);
    o[3] = make_float2(n.y, n.z);
    o[4] = make_float2(depth, __uint_as_float(status));
}

// GJK -> EPA hand-off record (80 bytes, 16-byte aligned)
struct __align__(16) EpaWork {
    uint32_t pair, unused, n, pad;   // pair index, -, simplex size (bit 31: GJK status 301), pair class
    float y[12];              // simplex points
    uint32_t id[4];           // simplex vertex ids
};
static_assert(sizeof(EpaWork) == 80, "EpaWork must be 80 bytes");

constexpr int kSpillStateWords = 8;   // nv, nf, best, bd, alive, it, (2 spare)
struct NarrowQueues {
    EpaWork* work;        // capacity maxContacts
    uint32_t* overflow;   // indices into work[] that need the full-cap path; bit 31: polytope spilled
    float* spill;         // spillCap x (fast polytope words + kSpillStateWords): state to resume from
    uint32_t spillCap;
};

// ---- kernel 1: GJK over every candidate pair -------------------------------------------------------
#ifndef AXCD_GJK_THREADS
#define AXCD_GJK_THREADS 128
#endif
constexpr int kGjkThreads = AXCD_GJK_THREADS;

// Pair class by the two core kinds, so that a warp runs ONE pair of support functions (a warp that mixes a box
// support with a 16-vertex hull scan executes both for every lane).  1-4: the classes with closed forms (sphere-
// sphere, sphere-box, box-sphere, box-box); 5: capsule against sphere / box / capsule (point, segment and box
// supports, all cheap); 6-10: the five (kind A, kind B) combinations with a convex hull; 11: anything with a cylinder.
enum { CLASS_CAPSULE_MIX = 5, CLASS_LIGHT_HULL = 6, CLASS_HULL_LIGHT = 7, CLASS_BOX_HULL = 8, CLASS_HULL_BOX = 9,
       CLASS_HULL_HULL = 10, CLASS_CYLINDER = 11 };
constexpr uint32_t kGenericClassMask = 0xfe0u;   // classes 5..11 always need GJK
static_assert(kNumClasses == 12, "classes 0..11");
__device__ __forceinline__ int pairClass(uint32_t typeA, uint32_t typeB) {
    if (typeA == AXCD_SHAPE_CYLINDER || typeB == AXCD_SHAPE_CYLINDER) return CLASS_CYLINDER;
    const bool hullA = typeA == AXCD_SHAPE_CONVEX, hullB = typeB == AXCD_SHAPE_CONVEX;
    if (hullA || hullB) {
        if (hullA && hullB) return CLASS_HULL_HULL;
        if (hullA) return (typeB == AXCD_SHAPE_BOX) ? CLASS_HULL_BOX : CLASS_HULL_LIGHT;
        return (typeA == AXCD_SHAPE_BOX) ? CLASS_BOX_HULL : CLASS_LIGHT_HULL;
    }
    if (typeA == AXCD_SHAPE_CAPSULE || typeB == AXCD_SHAPE_CAPSULE) return CLASS_CAPSULE_MIX;
    if (typeA == AXCD_SHAPE_SPHERE) return (typeB == AXCD_SHAPE_SPHERE) ? 1 : 2;   // SS, point-box
    return (typeB == AXCD_SHAPE_SPHERE) ? 3 : 4;                                   // box-point, box-box
}

// ---- kernel 0: deal the pairs into class-homogeneous chunks of 32 --------------------------------------
// A 50/50 box/sphere scene has 25 % box-box pairs that cost an order of magnitude more than the rest.
// Binning inside a block tile leaves one warp per tile grinding through them while its neighbours
// idle at the tile barrier, so the binning is global instead: this pass classifies every pair
// (sphere-sphere / sphere-box / box-sphere / box-box / other) and emits chunks of 32 pair indices of
// ONE class; gjkKernel's warps then claim chunks by ticket, with no block barrier and no mixed warps.
// Output order does not matter: every result of gjkKernel is stored by pair index.
constexpr int kClsThreads = 384;   // >= kNumClasses * 32: one thread per carry slot
constexpr int kClsItems = 4;
constexpr int kClsTile = kClsThreads * kClsItems;
constexpr uint32_t kNoPair = 0xffffffffu;
static_assert(kClsThreads >= kNumClasses * 32, "one thread per carry slot");

__host__ __device__ constexpr uint32_t classifyBlocksFor(uint32_t maxPairs) {
    const uint32_t tiles = (maxPairs + kClsTile - 1) / kClsTile;
    return tiles < (uint32_t)kNumSMs * 4 ? (tiles ? tiles : 1u) : (uint32_t)kNumSMs * 4;
}
__host__ __device__ constexpr uint32_t chunkCapFor(uint32_t maxPairs) {
    return maxPairs / 32 + classifyBlocksFor(maxPairs) * kNumClasses + 1;   // full chunks + every block's padded tails
}

// Two chunk lists share the `chunks` array: classes decided in closed form (sphere-sphere, sphere-box,
// box-box by SAT) fill it from the front (ctr->gjkChunks), the classes in `genericMask` (hulls / capsules;
// box-box when forced through GJK) from the back (ctr->gjkChunksGeneric), so closedFormKernel and gjkKernel
// each walk their own list.
__global__ void __launch_bounds__(kClsThreads)
classifyPairsKernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pairCount, uint32_t maxPairs,
                    const uint8_t* __restrict__ type8, uint32_t* __restrict__ chunks, uint32_t chunkCap,
                    uint32_t genericMask, Counters* __restrict__ ctr) {
    __shared__ uint32_t sCarry[kNumClasses][32];   // < 32 leftovers per class from the earlier tiles
    __shared__ uint32_t sCarryN[kNumClasses];
    __shared__ uint32_t sCnt[kNumClasses];         // carry + this tile, per class
    __shared__ uint32_t sBase[kNumClasses];        // start of each class in sItems
    __shared__ uint32_t sChunkStart[kNumClasses + 1];   // first chunk (within the tile) of each class
    __shared__ uint32_t sDst[kNumClasses];         // position of each class's first chunk in its list
    __shared__ uint32_t sItems[kClsTile + kNumClasses * 32];
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t npairs = min(*pairCount, maxPairs);
    // list position -> chunk index: closed list from the front, generic list from the back
    auto chunkIndex = [&](int c, uint32_t posInList) -> uint32_t {
        return ((genericMask >> c) & 1u) ? chunkCap - 1u - posInList : posInList;   // wraps past chunkCap if the list is too long
    };
    if (tid < kNumClasses) sCarryN[tid] = 0;
    __syncthreads();
    for (uint32_t tileBase = blockIdx.x * kClsTile; tileBase < npairs; tileBase += gridDim.x * kClsTile) {
        if (tid < kNumClasses) sCnt[tid] = sCarryN[tid];   // the carry takes the first slots of its class
        __syncthreads();
        int cls[kClsItems];
        uint32_t pos[kClsItems];
#pragma unroll
        for (int j = 0; j < kClsItems; ++j) {
            const uint32_t k = tileBase + j * kClsThreads + tid;
            cls[j] = -1;
            pos[j] = 0;
            if (k < npairs) {
                const uint2 pk = __ldg(pairs + k);
                cls[j] = pairClass(__ldg(type8 + pk.x), __ldg(type8 + pk.y));
            }
            // warp-aggregated slot reservation: the lanes of one class elect a leader
            const uint32_t peers = __match_any_sync(0xffffffffu, cls[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader && cls[j] >= 0) base = atomicAdd(&sCnt[cls[j]], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            pos[j] = base + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t run = 0, ch = 0, chList[2] = {0u, 0u}, rel[kNumClasses];
            for (int c = 0; c < kNumClasses; ++c) {
                sBase[c] = run;
                sChunkStart[c] = ch;
                run += sCnt[c];
                ch += sCnt[c] >> 5;
                const int g = (genericMask >> c) & 1u;
                rel[c] = chList[g];
                chList[g] += sCnt[c] >> 5;
            }
            sChunkStart[kNumClasses] = ch;
            const uint32_t base0 = chList[0] ? atomicAdd(&ctr->gjkChunks, chList[0]) : 0u;
            const uint32_t base1 = chList[1] ? atomicAdd(&ctr->gjkChunksGeneric, chList[1]) : 0u;
            for (int c = 0; c < kNumClasses; ++c) sDst[c] = (((genericMask >> c) & 1u) ? base1 : base0) + rel[c];
        }
        __syncthreads();
        if (tid < kNumClasses * 32) {
            const int c = tid >> 5, i = tid & 31;
            if ((uint32_t)i < sCarryN[c]) sItems[sBase[c] + i] = sCarry[c][i];
        }
#pragma unroll
        for (int j = 0; j < kClsItems; ++j)
            if (cls[j] >= 0) sItems[sBase[cls[j]] + pos[j]] = tileBase + j * kClsThreads + tid;
        __syncthreads();
        // full chunks go out (coalesced), the remainder of each class is carried to the next tile
        const uint32_t nOut = sChunkStart[kNumClasses] * 32;
        for (uint32_t e = tid; e < nOut; e += kClsThreads) {
            const uint32_t ch = e >> 5;
            int c = 0;
#pragma unroll
            for (int t = 1; t < kNumClasses; ++t) c += (ch >= sChunkStart[t]) ? 1 : 0;
            const uint32_t g = chunkIndex(c, sDst[c] + (ch - sChunkStart[c]));
            if (g < chunkCap) chunks[(size_t)g * 32 + (e & 31)] = sItems[sBase[c] + ((ch - sChunkStart[c]) << 5) + (e & 31)];
        }
        uint32_t carryVal = 0, rem = 0;
        if (tid < kNumClasses * 32) {
            const int c = tid >> 5, i = tid & 31;
            rem = sCnt[c] & 31u;
            if ((uint32_t)i < rem) carryVal = sItems[sBase[c] + (sCnt[c] & ~31u) + i];
        }
        __syncthreads();
        if (tid < kNumClasses * 32) {
            const int c = tid >> 5, i = tid & 31;
            if ((uint32_t)i < rem) sCarry[c][i] = carryVal;
            if (i == 0) sCarryN[c] = rem;
        }
        __syncthreads();
    }
    // what is left: one padded chunk per class that still holds pairs
    if (tid == 0) {
        uint32_t chList[2] = {0u, 0u}, rel[kNumClasses];
        for (int c = 0; c < kNumClasses; ++c) {
            const int g = (genericMask >> c) & 1u;
            rel[c] = chList[g];
            chList[g] += sCarryN[c] ? 1u : 0u;
        }
        const uint32_t base0 = chList[0] ? atomicAdd(&ctr->gjkChunks, chList[0]) : 0u;
        const uint32_t base1 = chList[1] ? atomicAdd(&ctr->gjkChunksGeneric, chList[1]) : 0u;
        for (int c = 0; c < kNumClasses; ++c) sDst[c] = (((genericMask >> c) & 1u) ? base1 : base0) + rel[c];
    }
    __syncthreads();
    if (tid < kNumClasses * 32) {
        const int c = tid >> 5, i = tid & 31;
        const uint32_t g = chunkIndex(c, sDst[c]);
        if (sCarryN[c] && g < chunkCap) chunks[(size_t)g * 32 + i] = ((uint32_t)i < sCarryN[c]) ? sCarry[c][i] : kNoPair;
    }
}

// ---- kernel 1a: the pair classes that are decided in closed form ------------------------------------------
// One lane per candidate pair, one class-homogeneous chunk of 32 pairs per warp at a time (claimed by
// ticket): sphere-sphere, sphere-box / box-sphere (clamp in the box frame) and box-box (15-axis SAT).
// Per pair it writes flag[k] (0 = no contact, 1 = contact) and the record to tmp[k]; contact slots are
// assigned afterwards, in pair order, by slotKernel.
#ifndef AXCD_CLOSED_MIN_BLOCKS
#define AXCD_CLOSED_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(kGjkThreads, AXCD_CLOSED_MIN_BLOCKS)
closedFormKernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ chunks, uint32_t chunkCap,
                 const float* __restrict__ xf, const uint4* __restrict__ shapes, uint8_t* __restrict__ flags,
                 AxcdContact* __restrict__ tmp, float* __restrict__ pairDist, Counters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    const uint32_t nChunks = min(ctr->gjkChunks, chunkCap);
    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&ctr->gjkChunkCursor, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= nChunks) break;
        const uint32_t k = __ldg(chunks + (size_t)chunk * 32 + lane);   // global pair index
        if (k == kNoPair) continue;
        const uint2 pk = __ldg(pairs + k);
        const uint32_t ia = pk.x, ib = pk.y;
        const BodyPose ta = loadPose(xf, ia), tb = loadPose(xf, ib);
        const uint4 sa = __ldg(shapes + ia), sb = __ldg(shapes + ib);
        const V3 origin = ta.p;
        bool contact = false;
        V3 n = mk3(0.f, 0.f, 0.f), pos = n;
        float depth = 0.f, dist = 0.f;
        if (sa.x == AXCD_SHAPE_SPHERE && sb.x == AXCD_SHAPE_SPHERE) {
            const float ra = __uint_as_float(sa.y), rb = __uint_as_float(sb.y);
            const V3 d = tb.p - origin;
            const float len = sqrtf(dot3(d, d));
            const float rs = ra + rb;
            depth = rs - len;
            dist = len - rs;
            if (depth >= 0.0f) {
                contact = true;
                n = (len > 0.0f) ? d * (1.0f / len) : mk3(1.0f, 0.0f, 0.0f);
                const V3 pa = n * ra;
                const V3 pb = d - n * rb;
                pos = (pa + pb) * 0.5f + origin;
            }
        } else if (sa.x == AXCD_SHAPE_SPHERE && sb.x == AXCD_SHAPE_BOX) {
            const SphereBox r = sphereBox(mk3(0.f, 0.f, 0.f), __uint_as_float(sa.y), tb.p - origin, tb, sb);
            dist = r.dist;
            if (r.contact) {
                contact = true;
                depth = r.depth;
                n = r.nsx;
                pos = (r.ps + r.px) * 0.5f + origin;
            }
        } else if (sa.x == AXCD_SHAPE_BOX && sb.x == AXCD_SHAPE_SPHERE) {
            const SphereBox r = sphereBox(tb.p - origin, __uint_as_float(sb.y), mk3(0.f, 0.f, 0.f), ta, sa);
            dist = r.dist;
            if (r.contact) {
                contact = true;
                depth = r.depth;
                n = -r.nsx;
                pos = (r.px + r.ps) * 0.5f + origin;
            }
        } else {   // box-box (the classifier sends nothing else here)
            const BoxBox r = boxBox(makeBoxFrame(ta, sa, origin), makeBoxFrame(tb, sb, origin));
            dist = FLT_MAX;   // separated: no distance in this mode (with PAIR_DISTANCES box-box goes to gjkKernel)
            if (r.contact) {
                contact = true;
                depth = r.depth;
                dist = -r.depth;
                n = r.n;
                pos = (r.pa + r.pb) * 0.5f + origin;
            }
        }
        if (pairDist) pairDist[k] = contact ? ((dist < 0.0f) ? dist : 0.0f) : dist;
        flags[k] = contact ? (uint8_t)1 : (uint8_t)0;
        if (contact) storeContact(tmp + k, ia, ib, pos, n, depth, 0u);
    }
}

// ---- the whole narrowphase in one pass, for scenes whose pairs are all decided in closed form ----------------
// (spheres and boxes only: sphere-sphere, sphere-box, box-box by SAT — the headline, C3 and C4 scenes).  A tile of
// 1024 canonical pairs per block trip:
//   1. load the pairs (coalesced), classify them by the two shape types, counting-sort the tile's local indices by
//      class in shared memory, so that every warp then works through class-homogeneous runs of 32;
//   2. evaluate the closed forms; a contact's 40-byte record is parked in shared memory under its local index;
//   3. scan the contact flags in pair order, chain the tile into the global order with a decoupled look-back
//      (tiles are claimed by ticket, so every predecessor has started), and copy the records out contiguously:
//      contacts come out in canonical pair order without per-pair flags, slots or staging records in HBM.
// It replaces classifyPairsKernel + closedFormKernel + slotKernel (and the 40 B/pair tmp records between them).
#ifndef AXCD_FUSED_THREADS
#define AXCD_FUSED_THREADS 256
#define AXCD_FUSED_MINBLOCKS 4   // 64 registers; sweep 256x3 / 256x4 / 128x6 / 128x8 / 64x12: 0.216 / 0.197 / 0.217 / 0.214 / 0.213 ms
#endif
constexpr int kFusedThreads = AXCD_FUSED_THREADS;
constexpr int kFusedItems = 4;
constexpr int kFusedTile = kFusedThreads * kFusedItems;   // 1024 pairs
constexpr int kFusedRecWords = 7;                          // position, normal, depth (ids come from the pair, status is 0)

// PROGRESS = true (a contact sink is attached, axcd_set_contact_sink): the kernel also tells the HOST how far the
// contact array is complete, through words in page-locked host memory: progress[0] = pair count + 1 (so the host
// knows how many tiles to expect), progress[1 + t] = (number of contacts of tiles 0..t, capped at the capacity) + 1,
// stored once tile t's records are in the device array (system-scope fence between the two).  The host drains the
// tiles in order while the kernel is still running and moves the finished stretch of the array with the copy engine
// (drainContactSink in axcd_api.cu): DMA bursts reach ~55 GB/s over PCIe where stores from the SMs reached ~49.
template <bool PROGRESS>
__global__ void __launch_bounds__(kFusedThreads, AXCD_FUSED_MINBLOCKS)
narrowClosedFusedKernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ pairCount, uint32_t maxPairs,
                        const uint8_t* __restrict__ type8, const float* __restrict__ xf, const uint4* __restrict__ shapes,
                        AxcdContact* __restrict__ contacts, uint32_t maxContacts, volatile uint32_t* __restrict__ progress,
                        volatile uint32_t* __restrict__ tileStatus, Counters* __restrict__ ctr) {
    __shared__ __align__(16) float sRec[kFusedTile * kFusedRecWords];   // 28 KB: contact floats by local pair index
    __shared__ uint2 sPair[kFusedTile];
    __shared__ uint16_t sIdx[kFusedTile];                  // local indices, sorted by class
    __shared__ __align__(16) uint8_t sFlag[kFusedTile];
    __shared__ uint32_t sCnt[4], sBase[4];
    __shared__ uint32_t sWarp[kFusedThreads / 32];
    __shared__ uint32_t sTile, sSlotBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t npairs = min(*pairCount, maxPairs);
    if (PROGRESS && blockIdx.x == 0 && tid == 0) progress[0] = npairs + 1u;
    uint32_t doneTile = 0, doneEnd = 0;   // thread 0: the tile this block finished last, not yet reported to the host
    bool haveDone = false;
    while (true) {
        __syncthreads();   // (PROGRESS: every thread's records of the previous tile are stored and fenced)
        if (PROGRESS && tid == 0 && haveDone) {
            progress[1u + doneTile] = doneEnd + 1u;
            haveDone = false;
        }
        if (tid == 0) sTile = atomicAdd(&ctr->gjkTicket, 1u);
        if (tid < 4) sCnt[tid] = 0;
        __syncthreads();
        const uint32_t tile = sTile;
        if ((uint64_t)tile * kFusedTile >= npairs) break;
        const uint32_t tileBase = tile * kFusedTile;
        const uint32_t tileCount = min((uint32_t)kFusedTile, npairs - tileBase);
        // ---- 1. load + classify + counting sort of the local indices by class (0 SS, 1 sphere-box either way, 2 BB) ---
        int cls[kFusedItems];
        uint32_t pos[kFusedItems];
#pragma unroll
        for (int j = 0; j < kFusedItems; ++j) {
            const uint32_t li = j * kFusedThreads + tid;
            cls[j] = -1;
            if (li < tileCount) {
                const uint2 pk = __ldg(pairs + tileBase + li);
                sPair[li] = pk;
                const uint32_t ta = __ldg(type8 + pk.x), tb = __ldg(type8 + pk.y);
                cls[j] = (ta == AXCD_SHAPE_BOX ? 1 : 0) + (tb == AXCD_SHAPE_BOX ? 1 : 0);
            }
            sFlag[li] = 0;
            const uint32_t peers = peersByBallot<2>((uint32_t)cls[j]);   // classes 0, 1, 2; -1 (no pair) reads as 3
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader && cls[j] >= 0) base = atomicAdd(&sCnt[cls[j]], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            pos[j] = base + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (tid == 0) {
            sBase[0] = 0;
            sBase[1] = sCnt[0];
            sBase[2] = sCnt[0] + sCnt[1];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kFusedItems; ++j)
            if (cls[j] >= 0) sIdx[sBase[cls[j]] + pos[j]] = (uint16_t)(j * kFusedThreads + tid);
        __syncthreads();
        // ---- 2. the closed forms; warp w takes runs w, w + 8, ... of 32 sorted entries (classes interleave evenly) -----
        for (uint32_t e = warp * 32 + lane; e < tileCount; e += kFusedThreads) {
            const uint32_t li = sIdx[e];
            const uint2 pk = sPair[li];
            const uint32_t ia = pk.x, ib = pk.y;
            const BodyPose ta = loadPose(xf, ia), tb = loadPose(xf, ib);
            const uint4 sa = __ldg(shapes + ia), sb = __ldg(shapes + ib);
            const V3 origin = ta.p;
            bool contact = false;
            V3 n = mk3(0.f, 0.f, 0.f), cpos = n;
            float depth = 0.f;
            if (sa.x == AXCD_SHAPE_SPHERE && sb.x == AXCD_SHAPE_SPHERE) {
                const float ra = __uint_as_float(sa.y), rb = __uint_as_float(sb.y);
                const V3 d = tb.p - origin;
                const float len = sqrtf(dot3(d, d));
                const float rs = ra + rb;
                depth = rs - len;
                if (depth >= 0.0f) {
                    contact = true;
                    n = (len > 0.0f) ? d * (1.0f / len) : mk3(1.0f, 0.0f, 0.0f);
                    const V3 pa = n * ra;
                    const V3 pb = d - n * rb;
                    cpos = (pa + pb) * 0.5f + origin;
                }
            } else if (sa.x == AXCD_SHAPE_SPHERE) {
                const SphereBox r = sphereBox(mk3(0.f, 0.f, 0.f), __uint_as_float(sa.y), tb.p - origin, tb, sb);
                if (r.contact) {
                    contact = true;
                    depth = r.depth;
                    n = r.nsx;
                    cpos = (r.ps + r.px) * 0.5f + origin;
                }
            } else if (sb.x == AXCD_SHAPE_SPHERE) {
                const SphereBox r = sphereBox(tb.p - origin, __uint_as_float(sb.y), mk3(0.f, 0.f, 0.f), ta, sa);
                if (r.contact) {
                    contact = true;
                    depth = r.depth;
                    n = -r.nsx;
                    cpos = (r.px + r.ps) * 0.5f + origin;
                }
            } else {
                const BoxBox r = boxBox(makeBoxFrame(ta, sa, origin), makeBoxFrame(tb, sb, origin));
                if (r.contact) {
                    contact = true;
                    depth = r.depth;
                    n = r.n;
                    cpos = (r.pa + r.pb) * 0.5f + origin;
                }
            }
            if (contact) {
                sFlag[li] = 1;
                float* o = sRec + li * kFusedRecWords;
                o[0] = cpos.x; o[1] = cpos.y; o[2] = cpos.z;
                o[3] = n.x; o[4] = n.y; o[5] = n.z;
                o[6] = depth;
            }
        }
        __syncthreads();
        // ---- 3. contact slots in pair order: block scan + decoupled look-back over the tiles --------------------------
        const uint32_t fl = *reinterpret_cast<const uint32_t*>(sFlag + tid * kFusedItems);   // 4 one-byte flags
        const uint32_t sum = __popc(fl & 0x01010101u);
        uint32_t inc = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) sWarp[warp] = inc;
        __syncthreads();
        uint32_t warpPrefix = 0, tileTotal = 0;
#pragma unroll
        for (int i = 0; i < kFusedThreads / 32; ++i) {
            warpPrefix += (i < warp) ? sWarp[i] : 0u;
            tileTotal += sWarp[i];
        }
        if (warp == 0) {
            uint32_t excl = 0;
            if (tile == 0) {
                if (lane == 0) tileStatus[0] = kFlagInclusive | tileTotal;
            } else {
                if (lane == 0) tileStatus[tile] = kFlagAggregate | tileTotal;
                excl = lookbackWide(tileStatus, tile, lane);
                if (lane == 0) tileStatus[tile] = kFlagInclusive | (excl + tileTotal);
            }
            if (lane == 0) {
                sSlotBase = excl;
                if ((uint64_t)(tile + 1) * kFusedTile >= npairs) ctr->contactCount = excl + tileTotal;   // last tile
            }
        }
        __syncthreads();
        {   // every thread stores its own records
            uint32_t slot = sSlotBase + warpPrefix + inc - sum;
#pragma unroll
            for (int i = 0; i < kFusedItems; ++i) {
                if ((fl >> (8 * i)) & 1u) {
                    if (slot < maxContacts) {
                        const uint32_t li = tid * kFusedItems + i;
                        const float* r = sRec + li * kFusedRecWords;
                        const uint2 pk = sPair[li];
                        storeContact(contacts + slot, pk.x, pk.y, mk3(r[0], r[1], r[2]), mk3(r[3], r[4], r[5]), r[6], 0u);
                    }
                    ++slot;
                }
            }
        }
        if (PROGRESS) {
            __threadfence_system();   // the records before the word the host reads (stored after the next barrier)
            if (tid == 0) {
                doneTile = tile;
                doneEnd = min(sSlotBase + tileTotal, maxContacts);
                haveDone = true;
            }
        }
    }
}

// The SAT as a call for gjkKernel (box-box reaches it only with AXCD_FLAG_PAIR_DISTANCES): keeps the
// closed form's registers out of the GJK kernel's allocation.
__device__ __noinline__ BoxBox boxBoxCall(const BodyPose& ta, uint4 sa, const BodyPose& tb, uint4 sb, V3 origin) {
    return boxBox(makeBoxFrame(ta, sa, origin), makeBoxFrame(tb, sb, origin));
}

// ---- kernel 1: GJK over the pairs of the generic classes ---------------------------------------------------
// One lane per candidate pair, one class-homogeneous chunk of 32 pairs per warp at a time (claimed
// by ticket).  Per pair it writes flag[k]: 0 = no contact, 1 = shallow contact (cores apart, radii
// overlapping; record written to tmp[k]), 2 = cores overlap (EpaWork queued; EPA writes the record).
// Contact slots are assigned afterwards, in pair order, by slotKernel.
#ifndef AXCD_GJK_MIN_BLOCKS
#define AXCD_GJK_MIN_BLOCKS 4   // 128 registers; C2 GJK stage with 3 / 4 / 5 / 6 blocks: 0.646 / 0.589 / 0.621 / 0.653 ms
#endif
template <bool CYL>
__global__ void __launch_bounds__(kGjkThreads, AXCD_GJK_MIN_BLOCKS)
gjkKernel(const uint2* __restrict__ pairs, const uint32_t* __restrict__ chunks, uint32_t chunkCap,
          const float* __restrict__ xf, const uint4* __restrict__ shapes,
          const float4* __restrict__ hull, NarrowParams cfg, uint8_t* __restrict__ flags,
          AxcdContact* __restrict__ tmp, NarrowQueues q, uint32_t queueCap, float* __restrict__ pairDist,
          Counters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    const uint32_t nChunks = min(ctr->gjkChunksGeneric, chunkCap);
    while (true) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(&ctr->gjkChunkCursorGeneric, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= nChunks) break;
    const uint32_t k = __ldg(chunks + (size_t)(chunkCap - 1u - chunk) * 32 + lane);   // the generic list grows from the back
    if (k == kNoPair) continue;

    // ---- per-pair GJK ----------------------------------------------------------------------------
    int kind = 0;
    uint32_t status = 0;
    const uint2 pk = __ldg(pairs + k);
    const uint32_t ia = pk.x, ib = pk.y;
    V3 n = mk3(0.f, 0.f, 0.f), pos = n;
    float depth = 0.f, dist = 0.f;
    Simplex s;
    s.n = 0;
    const BodyPose ta = loadPose(xf, ia), tb = loadPose(xf, ib);
    const uint4 sa = __ldg(shapes + ia), sb = __ldg(shapes + ib);
    const V3 origin = ta.p;
    // box-box without the generic flag gets here only for its distance (AXCD_FLAG_PAIR_DISTANCES): the
    // contact decision and the record still come from the SAT
    bool satApart = false;
    if (sa.x == AXCD_SHAPE_BOX && sb.x == AXCD_SHAPE_BOX && !cfg.boxBoxGeneric) {
        const BoxBox r = boxBoxCall(ta, sa, tb, sb, origin);
        if (r.contact) {
            kind = 1;
            depth = r.depth;
            dist = -r.depth;
            n = r.n;
            pos = (r.pa + r.pb) * 0.5f + origin;
        } else {
            satApart = true;
        }
    }
    if (kind == 0) {
        const CoreT<CYL> A = makeCore<CYL>(ta, sa, hull, origin);
        const CoreT<CYL> B = makeCore<CYL>(tb, sb, hull, origin);
        const float rs = A.r + B.r;
        const GjkResult g = gjk(A, B, cfg, rs, s);
        if (satApart) {
            dist = (g.state == GJK_SEPARATED) ? sqrtf(g.vv) : 0.0f;
        } else {
        status = g.status;
        if (g.state == GJK_SEPARATED) {
            const float len = sqrtf(g.vv);
            dist = len - rs;
            if (g.exact) {
                depth = rs - len;
                if (depth >= 0.0f) {
                    kind = 1;
                    n = -(g.v * (1.0f / len));
                    V3 ca = mk3(0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < s.n) ca = ca + pointFromId(A, s.id[i] & 0xffffu) * s.lam[i];
                    const V3 pa = ca + n * A.r;
                    const V3 pb = (ca - g.v) - n * B.r;
                    pos = (pa + pb) * 0.5f + origin;
                }
            }
        } else {
            kind = 2;
        }
        }
    }
    if (pairDist) pairDist[k] = (kind == 1) ? ((dist < 0.0f) ? dist : 0.0f) : dist;   // kind 2: EPA overwrites
    flags[k] = (uint8_t)kind;
    if (kind == 1) {
        storeContact(tmp + k, ia, ib, pos, n, depth, status);
        if (status == AXCD_ERR_GJK_NO_CONVERGE) atomicAdd(&ctr->gjkFailures, 1u);
    } else if (kind == 2) {
        // warp-aggregated queue push (threads of a warp are class-sorted, so runs stay homogeneous)
        const uint32_t m = __activemask();
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&ctr->epaCount, (uint32_t)__popc(m));
        base = __shfl_sync(m, base, leader);
        const uint32_t wq = base + __popc(m & ((1u << lane) - 1u));
        if (wq < queueCap) {
            EpaWork* wk = q.work + wq;
            uint4* o = reinterpret_cast<uint4*>(wk);
            o[0] = make_uint4(k, 0u, (uint32_t)s.n | (status ? 0x80000000u : 0u), (uint32_t)pairClass(sa.x, sb.x));
            float4* of = reinterpret_cast<float4*>(wk) + 1;
            of[0] = make_float4(s.y[0].x, s.y[0].y, s.y[0].z, s.y[1].x);
            of[1] = make_float4(s.y[1].y, s.y[1].z, s.y[2].x, s.y[2].y);
            of[2] = make_float4(s.y[2].z, s.y[3].x, s.y[3].y, s.y[3].z);
            o[4] = make_uint4(s.id[0], s.id[1], s.id[2], s.id[3]);
        }
    }
    }   // chunk loop
}

// ---- kernel 1b: contact slots in pair order ------------------------------------------------------------
// Exclusive scan of (flag != 0) over the pairs (single pass, decoupled look-back, ticketed tiles).
// Writes slot[k] for every contact pair, moves the shallow contacts tmp[k] -> contacts[slot], and
// leaves the total in ctr->contactCount.  EPA later writes its records to contacts[slot[k]].
constexpr int kSlotThreads = 256;
constexpr int kSlotItems = 8;
constexpr int kSlotTile = kSlotThreads * kSlotItems;
static_assert(kFusedTile <= kSlotTile, "the status buffer is sized by the smaller tile");

__global__ void __launch_bounds__(kSlotThreads)
slotKernel(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ pairCount, uint32_t maxPairs,
           const AxcdContact* __restrict__ tmp, AxcdContact* __restrict__ contacts, uint32_t maxContacts,
           uint32_t* __restrict__ slots, volatile uint32_t* __restrict__ tileStatus, Counters* __restrict__ ctr) {
    __shared__ uint32_t sWarp[kSlotThreads / 32];
    __shared__ uint32_t sSlotOut[kSlotTile];
    __shared__ uint32_t sTile, sBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t npairs = min(*pairCount, maxPairs);
    while (true) {   // persistent blocks take ticketed tiles until the pairs run out
    __syncthreads();
    if (tid == 0) sTile = atomicAdd(&ctr->gjkTicket, 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    if ((uint64_t)tile * kSlotTile >= npairs) break;
    const uint32_t base = tile * kSlotTile + tid * kSlotItems;
    // 8 one-byte flags per thread: one 8-byte load (flags buffer is padded to a tile multiple)
    const uint2 fl = (base < npairs) ? __ldg(reinterpret_cast<const uint2*>(flags + base)) : make_uint2(0u, 0u);
    uint32_t f[kSlotItems];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kSlotItems; ++i) {
        const uint32_t word = (i < 4) ? fl.x : fl.y;
        f[i] = (base + i < npairs) ? ((word >> (8 * (i & 3))) & 0xffu) : 0u;
        sum += f[i] ? 1u : 0u;
    }
    uint32_t inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    uint32_t warpPrefix = 0, tileTotal = 0;
#pragma unroll
    for (int i = 0; i < kSlotThreads / 32; ++i) {
        warpPrefix += (i < warp) ? sWarp[i] : 0u;
        tileTotal += sWarp[i];
    }
    if (warp == 0) {   // warp-parallel decoupled look-back, 128 tiles per round
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) tileStatus[0] = kFlagInclusive | tileTotal;
        } else {
            if (lane == 0) tileStatus[tile] = kFlagAggregate | tileTotal;
            excl = lookbackWide(tileStatus, tile, lane);
            if (lane == 0) tileStatus[tile] = kFlagInclusive | (excl + tileTotal);
        }
        if (lane == 0) {
            sBase = excl;
            if ((uint64_t)(tile + 1) * kSlotTile >= npairs) ctr->contactCount = excl + tileTotal;   // last tile
        }
    }
    __syncthreads();
    // hand the per-pair results to a striped mapping (consecutive lanes = consecutive pairs) so the
    // slot stores and the 40-byte record moves coalesce
    uint32_t run = sBase + warpPrefix + inc - sum;
#pragma unroll
    for (int i = 0; i < kSlotItems; ++i) {
        sSlotOut[tid * kSlotItems + i] = f[i] ? (run | (f[i] << 30)) : 0xffffffffu;   // slot < 2^30
        run += f[i] ? 1u : 0u;
    }
    __syncthreads();
    const uint32_t tileBase = tile * kSlotTile;
#pragma unroll
    for (int i = 0; i < kSlotItems; ++i) {
        const uint32_t local = i * kSlotThreads + tid;
        const uint32_t v = sSlotOut[local];
        if (v == 0xffffffffu) continue;
        const uint32_t k = tileBase + local, slot = v & 0x3fffffffu;
        slots[k] = slot;
        if ((v >> 30) == 1u && slot < maxContacts) {
            const float2* src = reinterpret_cast<const float2*>(tmp + k);
            float2* dst = reinterpret_cast<float2*>(contacts + slot);
#pragma unroll
            for (int w = 0; w < 5; ++w) dst[w] = src[w];
        }
    }
    }   // tile loop
}

// ---- kernel 2: EPA over the queued pairs -------------------------------------------------------------
#ifndef AXCD_EPA_THREADS
#define AXCD_EPA_THREADS 96      // 96 x 772 B = 72 KB of polytopes per block, three blocks per SM
#define AXCD_EPA_BLOCKS_PER_SM 3
#endif
constexpr int kEpaThreads = AXCD_EPA_THREADS;
constexpr int kEpaBlocksPerSM = AXCD_EPA_BLOCKS_PER_SM;
constexpr int kEpaSmemBytes =
    Poly<kEpaFastVerts, kEpaFastFaces, kEpaFastEdges, kEpaThreads>::kWords * kEpaThreads * (int)sizeof(float);
constexpr int kEpaChunk = 64;        // queue items a warp claims at a time
#ifndef AXCD_EPA_BATCH
#define AXCD_EPA_BATCH 16
#endif
constexpr int kEpaBatchLanes = AXCD_EPA_BATCH;    // refill / finalize once this many lanes wait for it

// What a lane carries for the pair it is expanding.
template <bool CYL>
struct EpaLaneT {
    CoreT<CYL> A, B;
    V3 origin;
    uint32_t pairIdx, ia, ib, status, queueIdx;
};

template <int MAXV, int MAXF, int MAXE, int STRIDE, bool CYL>
__device__ __forceinline__ int epaBegin(const EpaWork* __restrict__ wk, const uint2* __restrict__ pairs,
                                        const float* __restrict__ xf, const uint4* __restrict__ shapes,
                                        const float4* __restrict__ hull, const Poly<MAXV, MAXF, MAXE, STRIDE>& poly,
                                        EpaLaneT<CYL>& L, EpaState<typename Poly<MAXV, MAXF, MAXE, STRIDE>::Mask>& st,
                                        EpaResult& touching, bool skipInit = false) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(wk));
    const float4 f0 = __ldg(reinterpret_cast<const float4*>(wk) + 1), f1 = __ldg(reinterpret_cast<const float4*>(wk) + 2),
                 f2 = __ldg(reinterpret_cast<const float4*>(wk) + 3);
    const uint4 idv = __ldg(reinterpret_cast<const uint4*>(wk) + 4);
    L.pairIdx = h.x;
    const uint2 pk = __ldg(pairs + L.pairIdx);
    L.ia = pk.x;
    L.ib = pk.y;
    const int n0 = (int)(h.z & 0xffu);
    L.status = (h.z & 0x80000000u) ? (uint32_t)AXCD_ERR_GJK_NO_CONVERGE : 0u;
    const V3 y0[4] = {mk3(f0.x, f0.y, f0.z), mk3(f0.w, f1.x, f1.y), mk3(f1.z, f1.w, f2.x), mk3(f2.y, f2.z, f2.w)};
    const uint32_t id0[4] = {idv.x, idv.y, idv.z, idv.w};
    const BodyPose ta = loadPose(xf, L.ia), tb = loadPose(xf, L.ib);
    const uint4 sa = __ldg(shapes + L.ia), sb = __ldg(shapes + L.ib);
    L.origin = ta.p;
    L.A = makeCore<CYL>(ta, sa, hull, L.origin);
    L.B = makeCore<CYL>(tb, sb, hull, L.origin);
    if (skipInit) return 0;   // the caller restores a spilled polytope instead
    return epaInit(L.A, L.B, n0, y0, id0, poly, st, touching);
}

template <bool CYL>
__device__ __forceinline__ void epaEmit(const EpaLaneT<CYL>& L, const EpaResult& r, AxcdContact* __restrict__ contacts,
                                        uint32_t maxContacts, const uint32_t* __restrict__ slots,
                                        float* __restrict__ pairDist, Counters* __restrict__ ctr) {
    uint32_t status = L.status;
    if (r.status) status = r.status;
    const float rs = L.A.r + L.B.r;
    const float depth = r.depth + rs;
    const V3 pa = r.pa + r.n * L.A.r;
    const V3 pb = r.pb - r.n * L.B.r;
    const V3 pos = (pa + pb) * 0.5f + L.origin;
    const uint32_t slot = __ldg(slots + L.pairIdx);
    if (slot < maxContacts) storeContact(contacts + slot, L.ia, L.ib, pos, r.n, depth, status);
    if (pairDist) pairDist[L.pairIdx] = (depth > 0.0f) ? -depth : 0.0f;
    if (status == AXCD_ERR_GJK_NO_CONVERGE) atomicAdd(&ctr->gjkFailures, 1u);
    if (status == AXCD_ERR_EPA_NO_CONVERGE) atomicAdd(&ctr->epaFailures, 1u);
}

// Fast path: polytope in shared memory (one column per thread).  Each lane is a small state machine
// (EMPTY -> RUNNING -> DONE -> EMPTY): all RUNNING lanes of a warp execute one expansion step per
// trip, and refills / finalisations are batched (>= kEpaBatchLanes lanes, or nothing else to do), so
// a pair that needs 2 steps does not hold its lane hostage to a neighbour that needs 15.  Warps claim
// queue items in chunks; the queue length is only known on the device.
template <bool CYL>
__global__ void __launch_bounds__(kEpaThreads, kEpaBlocksPerSM)
epaKernel(NarrowQueues q, uint32_t queueCap, const uint2* __restrict__ pairs,
          const float* __restrict__ xf, const uint4* __restrict__ shapes, const float4* __restrict__ hull,
          NarrowParams cfg, AxcdContact* __restrict__ contacts, uint32_t maxContacts,
          const uint32_t* __restrict__ slots, float* __restrict__ pairDist, Counters* __restrict__ ctr) {
    extern __shared__ float sPoly[];
    using P = Poly<kEpaFastVerts, kEpaFastFaces, kEpaFastEdges, kEpaThreads>;
    P poly;
    poly.base = sPoly + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t count = min(ctr->epaCount, queueCap);
    enum { EMPTY = 0, RUNNING = 1, DONE = 2 };
    int state = EMPTY;
    EpaLaneT<CYL> L;
    EpaState<P::Mask> st;
    uint32_t cur = 0, end = 0;   // this warp's claimed queue range (warp-uniform)
    bool drained = false;        // no more items to claim (warp-uniform)
    while (true) {
        const uint32_t running = __ballot_sync(0xffffffffu, state == RUNNING);
        const uint32_t done = __ballot_sync(0xffffffffu, state == DONE);
        // ---- finalise finished pairs, batched ---------------------------------------------------
        if (done && (__popc(done) >= kEpaBatchLanes || !running)) {
            if (state == DONE) {
                const EpaResult r = epaFinish(L.A, poly, st);
                if (r.overflow) {
                    // hand the pair to the full-cap kernel together with its polytope, so that kernel
                    // continues the expansion instead of repeating it (nothing of the overflowing
                    // step has been applied yet)
                    const uint32_t o = atomicAdd(&ctr->epaOverflow, 1u);
                    const bool spilled = o < q.spillCap;
                    q.overflow[o] = L.queueIdx | (spilled ? 0x80000000u : 0u);
                    if (spilled) {
                        float* dst = q.spill + (size_t)o * (P::kWords + kSpillStateWords);
                        for (int i = 0; i < P::kWords; ++i) dst[i] = poly.w(i);
                        uint32_t* du = reinterpret_cast<uint32_t*>(dst + P::kWords);
                        du[0] = (uint32_t)st.nv; du[1] = (uint32_t)st.nf; du[2] = (uint32_t)st.best;
                        du[3] = __float_as_uint(st.bd); du[4] = (uint32_t)st.alive; du[5] = st.it;
                    }
                } else {
                    epaEmit(L, r, contacts, maxContacts, slots, pairDist, ctr);
                }
                state = EMPTY;
            }
        }
        // ---- refill empty lanes, batched ----------------------------------------------------------
        const uint32_t empty = __ballot_sync(0xffffffffu, state == EMPTY);
        if (!drained && empty && (__popc(empty) >= kEpaBatchLanes || !running)) {
            if (cur >= end) {   // claim the next chunk for the warp
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(&ctr->epaCursor, (uint32_t)kEpaChunk);
                cur = __shfl_sync(0xffffffffu, c, 0);
                end = min(cur + kEpaChunk, count);
                if (cur >= count) drained = true;
            }
            if (!drained) {
                const uint32_t mine = cur + __popc(empty & ((1u << lane) - 1u));
                if (state == EMPTY && mine < end) {
                    EpaResult touching;
                    L.queueIdx = mine;
                    if (epaBegin(q.work + mine, pairs, xf, shapes, hull, poly, L, st, touching)) {
                        epaEmit(L, touching, contacts, maxContacts, slots, pairDist, ctr);
                    } else {
                        state = RUNNING;
                    }
                }
                cur = min(cur + (uint32_t)__popc(empty), end);
            }
        }
        // ---- one expansion step for every running lane ------------------------------------------------
        const uint32_t nowRunning = __ballot_sync(0xffffffffu, state == RUNNING);
        if (!nowRunning) {
            if (drained && !__ballot_sync(0xffffffffu, state == DONE)) break;
            continue;
        }
        if (state == RUNNING) {
            if (epaIterate(L.A, L.B, cfg, poly, st)) state = DONE;
        }
    }
}

// axcd_set_contact_sink on the multi-kernel narrowphase path: the finished contact array goes to the caller's
// page-locked buffer with 8-byte stores in order (consecutive threads, consecutive addresses).
__global__ void copyContactsKernel(const float2* __restrict__ src, float2* __restrict__ dst, const uint32_t* __restrict__ count,
                                   uint32_t maxContacts) {
    const size_t words = (size_t)min(*count, maxContacts) * 5u;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < words; k += (size_t)gridDim.x * blockDim.x) dst[k] = src[k];
}

}  // namespace axcd
