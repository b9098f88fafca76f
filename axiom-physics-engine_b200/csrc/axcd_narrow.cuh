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

enum { CORE_POINT = 0, CORE_BOX = 1, CORE_HULL = 2 };

struct Core {
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

__device__ __forceinline__ Core makeCore(const BodyPose& t, uint4 sh, const float4* __restrict__ hull, V3 origin) {
    Core k;
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

// Support point of a core in world-aligned direction d (any length).
__device__ __forceinline__ V3 support(const Core& k, V3 d) {
    if (k.kind == CORE_POINT) return k.c;
    if (k.kind == CORE_BOX) {
        V3 p = k.c;
        p = p + ((dot3(d, k.e0) >= 0.0f) ? k.e0 : -k.e0);
        p = p + ((dot3(d, k.e1) >= 0.0f) ? k.e1 : -k.e1);
        p = p + ((dot3(d, k.e2) >= 0.0f) ? k.e2 : -k.e2);
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
        }
    }
    const V3 lv = mk3(bv.x * k.s.x, bv.y * k.s.y, bv.z * k.s.z);
    return ((k.e0 * lv.x + k.e1 * lv.y) + k.e2 * lv.z) + k.c;
}

struct Simplex {
    V3 y[4];   // points of the Minkowski difference A - B
    V3 a[4];   // matching support points on A
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
__device__ __noinline__ V3 closestTriangle(V3 a, V3 b, V3 c, float& la, float& lb, float& lc, int& mask) {
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
__device__ __forceinline__ V3 solveSimplex(Simplex& s, bool& enclosed) {
    enclosed = false;
    int mask = 0;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    V3 v = mk3(0.f, 0.f, 0.f);
    if (s.n == 2) {
        v = closestSegment(s.y[0], s.y[1], l0, l1, mask);
    } else if (s.n == 3) {
        v = closestTriangle(s.y[0], s.y[1], s.y[2], l0, l1, l2, mask);
    } else {
        float best = FLT_MAX;
        bool any = false;
        // faces (i,j,k | opposite o): (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0), examined in this order
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int i = (f == 3) ? 1 : 0;
            const int j = (f == 0) ? 1 : ((f == 1) ? 2 : 3);
            const int k = (f == 0) ? 2 : ((f == 1) ? 3 : ((f == 2) ? 1 : 2));
            const int o = (f == 0) ? 3 : ((f == 1) ? 1 : ((f == 2) ? 2 : 0));
            if (!originOutside(s.y[i], s.y[j], s.y[k], s.y[o])) continue;
            any = true;
            float li, lj, lk;
            int m;
            const V3 q = closestTriangle(s.y[i], s.y[j], s.y[k], li, lj, lk, m);
            const float qq = dot3(q, q);
            if (qq < best) {
                best = qq;
                v = q;
                float l[4] = {0.f, 0.f, 0.f, 0.f};
                l[i] = li; l[j] = lj; l[k] = lk;
                l0 = l[0]; l1 = l[1]; l2 = l[2]; l3 = l[3];
                mask = ((m & 1) ? (1 << i) : 0) | ((m & 2) ? (1 << j) : 0) | ((m & 4) ? (1 << k) : 0);
            }
        }
        if (!any) {
            enclosed = true;
            return mk3(0.f, 0.f, 0.f);
        }
    }
    const float l[4] = {l0, l1, l2, l3};
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i < s.n && (mask & (1 << i))) {
            s.y[n] = s.y[i];
            s.a[n] = s.a[i];
            s.lam[n] = l[i];
            ++n;
        }
    }
    s.n = n;
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

__device__ __forceinline__ GjkResult gjk(const Core& A, const Core& B, const NarrowParams& cfg, float marginSum,
                                         Simplex& s) {
    GjkResult r;
    r.status = 0;
    V3 d0 = B.c - A.c;
    if (dot3(d0, d0) < 1e-12f) d0 = mk3(1.0f, 0.0f, 0.0f);
    s.a[0] = support(A, d0);
    s.y[0] = s.a[0] - support(B, -d0);
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
        const V3 a = support(A, -v);
        const V3 w = a - support(B, v);
        const float vw = dot3(v, w);
        if (!cfg.wantDistances && vw > 0.0f && vw * vw > vv * (marginSum * marginSum)) {
            r.exact = false;   // separating axis with a gap larger than the radii
            break;
        }
        if (vv - vw <= cfg.gjkTol * vv) break;
        bool dup = false;
        for (int i = 0; i < s.n; ++i) dup = dup || same3(w, s.y[i]);
        if (dup) break;
        s.y[s.n] = w;
        s.a[s.n] = a;
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

// ---- EPA ------------------------------------------------------------------------------------------
constexpr int kEpaMaxVerts = 40;
constexpr int kEpaMaxFaces = 64;

struct EpaFace {
    V3 n;
    float d;
    uint8_t i0, i1, i2, alive;
};
struct EpaPolytope {
    V3 y[kEpaMaxVerts];
    V3 a[kEpaMaxVerts];
    EpaFace f[kEpaMaxFaces];
    int nv, nf;
};

__device__ __forceinline__ void epaSetFace(EpaPolytope& e, int slot, int i0, int i1, int i2) {
    EpaFace& F = e.f[slot];
    V3 n = cross3(e.y[i1] - e.y[i0], e.y[i2] - e.y[i0]);
    const float len2 = dot3(n, n);
    F.alive = 1;
    F.i0 = (uint8_t)i0; F.i1 = (uint8_t)i1; F.i2 = (uint8_t)i2;
    if (len2 <= 1e-30f) {   // zero-area face: never the closest, never visible
        F.n = mk3(0.f, 0.f, 0.f);
        F.d = FLT_MAX;
        return;
    }
    const float inv = 1.0f / sqrtf(len2);
    n = n * inv;
    F.n = n;
    F.d = dot3(n, e.y[i0]);
}

struct EpaResult {
    V3 n;
    float depth;
    V3 pa, pb;
    uint32_t status;
};

__device__ __forceinline__ EpaResult epaTouching(V3 n, V3 pa) {
    EpaResult r;
    const float l2 = dot3(n, n);
    r.n = (l2 > 0.0f) ? n * (1.0f / sqrtf(l2)) : mk3(1.0f, 0.0f, 0.0f);
    r.depth = 0.0f;
    r.pa = pa;
    r.pb = pa;
    r.status = 0;
    return r;
}

__device__ __noinline__ EpaResult epa(const Core& A, const Core& B, const NarrowParams& cfg, const Simplex& s0,
                                      EpaPolytope& e) {
    e.nv = s0.n;
    for (int i = 0; i < s0.n; ++i) {
        e.y[i] = s0.y[i];
        e.a[i] = s0.a[i];
    }
    // ---- grow the GJK simplex to a tetrahedron -------------------------------------------------
    if (e.nv == 1) {
        for (int k = 0; k < 6 && e.nv == 1; ++k) {
            const float sg = (k & 1) ? -1.0f : 1.0f;
            const V3 ax = mk3((k >> 1) == 0 ? sg : 0.f, (k >> 1) == 1 ? sg : 0.f, (k >> 1) == 2 ? sg : 0.f);
            e.a[1] = support(A, ax);
            e.y[1] = e.a[1] - support(B, -ax);
            const V3 d = e.y[1] - e.y[0];
            if (dot3(d, d) > kDegenerateEps) e.nv = 2;
        }
        if (e.nv == 1) return epaTouching(mk3(1.f, 0.f, 0.f), e.a[0]);
    }
    if (e.nv == 2) {
        const V3 d = e.y[1] - e.y[0];
        V3 firstDir = mk3(0.f, 0.f, 0.f);
        for (int k = 0; k < 3 && e.nv == 2; ++k) {
            const V3 ax = mk3(k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f);
            const V3 dir = cross3(d, ax);
            if (dot3(dir, dir) <= kDegenerateEps) continue;
            if (dot3(firstDir, firstDir) == 0.0f) firstDir = dir;
            for (int sgn = 0; sgn < 2 && e.nv == 2; ++sgn) {
                const V3 dd = sgn ? -dir : dir;
                e.a[2] = support(A, dd);
                e.y[2] = e.a[2] - support(B, -dd);
                const V3 c = cross3(e.y[2] - e.y[0], d);
                if (dot3(c, c) > kDegenerateEps) e.nv = 3;
            }
        }
        if (e.nv == 2) return epaTouching(firstDir, e.a[0]);
    }
    if (e.nv == 3) {
        const V3 n = cross3(e.y[1] - e.y[0], e.y[2] - e.y[0]);
        const float n2 = dot3(n, n);
        if (n2 <= 1e-30f) return epaTouching(mk3(1.f, 0.f, 0.f), e.a[0]);
        for (int sgn = 0; sgn < 2 && e.nv == 3; ++sgn) {
            const V3 dd = sgn ? -n : n;
            e.a[3] = support(A, dd);
            e.y[3] = e.a[3] - support(B, -dd);
            const float vol = dot3(e.y[3] - e.y[0], n);
            if (vol * vol > kDegenerateEps * n2) e.nv = 4;
        }
        if (e.nv == 3) {   // flat at the origin: touching contact along +n
            float la, lb, lc;
            int m;
            closestTriangle(e.y[0], e.y[1], e.y[2], la, lb, lc, m);
            const V3 pa = (e.a[0] * la + e.a[1] * lb) + e.a[2] * lc;
            return epaTouching(n, pa);
        }
    }
    // orientation: make (0,1,2) face away from vertex 3
    if (dot3(cross3(e.y[1] - e.y[0], e.y[2] - e.y[0]), e.y[3] - e.y[0]) > 0.0f) {
        V3 t = e.y[0]; e.y[0] = e.y[1]; e.y[1] = t;
        t = e.a[0]; e.a[0] = e.a[1]; e.a[1] = t;
    }
    epaSetFace(e, 0, 0, 1, 2);
    epaSetFace(e, 1, 0, 3, 1);
    epaSetFace(e, 2, 0, 2, 3);
    epaSetFace(e, 3, 1, 3, 2);
    e.nf = 4;
    const int maxFaces = (int)min(cfg.epaMaxFaces, (uint32_t)kEpaMaxFaces);

    uint32_t status = 0;
    int best = 0;
    for (uint32_t it = 0;; ++it) {
        best = -1;
        float bd = FLT_MAX;
        for (int i = 0; i < e.nf; ++i) {
            if (e.f[i].alive && e.f[i].d < bd) {
                bd = e.f[i].d;
                best = i;
            }
        }
        if (best < 0) return epaTouching(mk3(1.f, 0.f, 0.f), e.a[0]);
        const V3 bn = e.f[best].n;
        const float bdist = e.f[best].d;
        const V3 a = support(A, bn);
        const V3 w = a - support(B, -bn);
        const float dw = dot3(w, bn);
        const float scale = (bdist > 1.0f) ? bdist : 1.0f;
        if (dw - bdist <= cfg.epaTol * scale) break;
        bool dup = false;
        for (int i = 0; i < e.nv; ++i) dup = dup || same3(w, e.y[i]);
        if (dup) break;
        if (it >= cfg.epaMaxIters || e.nv >= kEpaMaxVerts) {
            status = AXCD_ERR_EPA_NO_CONVERGE;
            break;
        }
        // visible faces (w clearly in front) and the horizon edge loop
        uint8_t he0[kEpaMaxFaces * 3], he1[kEpaMaxFaces * 3];
        uint64_t visMask = 0;
        int nh = 0, nvis = 0, nalive = 0;
        const float wl = fabsf(w.x) + fabsf(w.y) + fabsf(w.z);
        const float visEps = 1e-6f * ((wl > 1.0f) ? wl : 1.0f);
        for (int i = 0; i < e.nf; ++i) {
            if (!e.f[i].alive) continue;
            ++nalive;
            if (dot3(e.f[i].n, w) - e.f[i].d > visEps) {
                visMask |= 1ull << i;
                ++nvis;
            }
        }
        for (int i = 0; i < e.nf; ++i) {
            if (!((visMask >> i) & 1ull)) continue;
            const uint8_t v0 = e.f[i].i0, v1 = e.f[i].i1, v2 = e.f[i].i2;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint8_t ea = (k == 0) ? v0 : ((k == 1) ? v1 : v2);
                const uint8_t eb = (k == 0) ? v1 : ((k == 1) ? v2 : v0);
                int found = -1;
                for (int h = 0; h < nh; ++h)
                    if (he0[h] == eb && he1[h] == ea) {
                        found = h;
                        break;
                    }
                if (found >= 0) {
                    he0[found] = he0[nh - 1];
                    he1[found] = he1[nh - 1];
                    --nh;
                } else {
                    he0[nh] = ea;
                    he1[nh] = eb;
                    ++nh;
                }
            }
        }
        bool loopOk = nh >= 3;
        for (int h = 0; h < nh && loopOk; ++h)
            for (int g = h + 1; g < nh; ++g)
                if (he0[g] == he0[h] || he1[g] == he1[h]) {
                    loopOk = false;
                    break;
                }
        if (!loopOk || nalive - nvis + nh > maxFaces) {
            status = AXCD_ERR_EPA_NO_CONVERGE;
            break;
        }
        const int wi = e.nv;
        e.y[wi] = w;
        e.a[wi] = a;
        e.nv++;
        for (int i = 0; i < e.nf; ++i)
            if ((visMask >> i) & 1ull) e.f[i].alive = 0;
        int slot = 0;
        for (int h = 0; h < nh; ++h) {
            while (slot < e.nf && e.f[slot].alive) ++slot;
            if (slot == e.nf) e.nf++;
            epaSetFace(e, slot, he0[h], he1[h], wi);
        }
    }
    const EpaFace fb = e.f[best];
    EpaResult r;
    r.n = fb.n;
    r.depth = (fb.d > 0.0f) ? fb.d : 0.0f;
    float la, lb, lc;
    int m;
    const V3 p = closestTriangle(e.y[fb.i0], e.y[fb.i1], e.y[fb.i2], la, lb, lc, m);
    r.pa = (e.a[fb.i0] * la + e.a[fb.i1] * lb) + e.a[fb.i2] * lc;
    r.pb = r.pa - p;
    r.status = status;
    return r;
}

// ---- one pair -------------------------------------------------------------------------------------
struct PairResult {
    bool contact;
    bool usedEpa;
    float dist;
    AxcdContact c;
};

__device__ __forceinline__ PairResult collidePair(uint32_t ia, uint32_t ib, const float* __restrict__ xf,
                                                  const uint4* __restrict__ shapes,
                                                  const float4* __restrict__ hull, const NarrowParams& cfg,
                                                  EpaPolytope& scratch) {
    PairResult o;
    o.contact = false;
    o.usedEpa = false;
    o.dist = 0.0f;
    o.c.a = ia;
    o.c.b = ib;
    const BodyPose ta = loadPose(xf, ia), tb = loadPose(xf, ib);
    const uint4 sa = __ldg(shapes + ia), sb = __ldg(shapes + ib);
    const V3 origin = ta.p;
    V3 n, pa, pb;
    float depth;
    uint32_t status = 0;
    if (sa.x == AXCD_SHAPE_SPHERE && sb.x == AXCD_SHAPE_SPHERE) {
        const float ra = __uint_as_float(sa.y), rb = __uint_as_float(sb.y);
        const V3 d = tb.p - origin;
        const float dist = sqrtf(dot3(d, d));
        const float rs = ra + rb;
        depth = rs - dist;
        o.dist = dist - rs;
        if (!(depth >= 0.0f)) return o;
        n = (dist > 0.0f) ? d * (1.0f / dist) : mk3(1.0f, 0.0f, 0.0f);
        pa = n * ra;
        pb = d - n * rb;
    } else {
        const Core A = makeCore(ta, sa, hull, origin);
        const Core B = makeCore(tb, sb, hull, origin);
        const float rs = A.r + B.r;
        Simplex s;
        const GjkResult g = gjk(A, B, cfg, rs, s);
        status = g.status;
        if (g.state == GJK_SEPARATED) {
            const float dist = sqrtf(g.vv);
            if (!g.exact) {
                o.dist = dist - rs;
                return o;
            }
            depth = rs - dist;
            o.dist = dist - rs;
            if (!(depth >= 0.0f)) return o;
            n = -(g.v * (1.0f / dist));
            V3 ca = mk3(0.f, 0.f, 0.f);
            for (int i = 0; i < s.n; ++i) ca = ca + s.a[i] * s.lam[i];
            pa = ca + n * A.r;
            pb = (ca - g.v) - n * B.r;
        } else {
            const EpaResult e = epa(A, B, cfg, s, scratch);
            o.usedEpa = true;
            if (e.status) status = e.status;
            n = e.n;
            depth = e.depth + rs;
            o.dist = -depth;
            pa = e.pa + n * A.r;
            pb = e.pb - n * B.r;
        }
    }
    o.contact = true;
    const V3 mid = (pa + pb) * 0.5f + origin;
    o.c.px = mid.x; o.c.py = mid.y; o.c.pz = mid.z;
    o.c.nx = n.x; o.c.ny = n.y; o.c.nz = n.z;
    o.c.depth = depth;
    o.c.status = status;
    return o;
}

// ---- kernels --------------------------------------------------------------------------------------
constexpr int kNarrowThreads = 128;

// One thread per candidate pair (pairs are (a,b)-sorted).  Writes flag[k] (1 = contact) and the
// contact record into tmp[k]; a scan + compaction pass then packs contacts in pair order.
__global__ void __launch_bounds__(kNarrowThreads)
narrowphaseKernel(const uint64_t* __restrict__ pairs, uint32_t npairs, int idxBits,
                  const float* __restrict__ xf, const uint4* __restrict__ shapes,
                  const float4* __restrict__ hull, NarrowParams cfg, uint32_t* __restrict__ flags,
                  AxcdContact* __restrict__ tmp, float* __restrict__ pairDist, Counters* __restrict__ ctr) {
    const uint32_t k = blockIdx.x * kNarrowThreads + threadIdx.x;
    if (k >= npairs) return;
    const uint64_t pk = pairs[k];
    const uint32_t a = (uint32_t)(pk >> idxBits), b = (uint32_t)(pk & ((1ull << idxBits) - 1ull));
    EpaPolytope scratch;
    const PairResult r = collidePair(a, b, xf, shapes, hull, cfg, scratch);
    flags[k] = r.contact ? 1u : 0u;
    if (pairDist) pairDist[k] = r.contact ? ((r.dist < 0.0f) ? r.dist : 0.0f) : r.dist;
    if (r.contact) {
        // 40-byte record, 8-byte aligned
        float2* o = reinterpret_cast<float2*>(tmp + k);
        o[0] = make_float2(__uint_as_float(r.c.a), __uint_as_float(r.c.b));
        o[1] = make_float2(r.c.px, r.c.py);
        o[2] = make_float2(r.c.pz, r.c.nx);
        o[3] = make_float2(r.c.ny, r.c.nz);
        o[4] = make_float2(r.c.depth, __uint_as_float(r.c.status));
        if (r.usedEpa) atomicAdd(&ctr->epaCount, 1u);
        if (r.c.status == AXCD_ERR_GJK_NO_CONVERGE) atomicAdd(&ctr->gjkFailures, 1u);
        if (r.c.status == AXCD_ERR_EPA_NO_CONVERGE) atomicAdd(&ctr->epaFailures, 1u);
    }
}

__global__ void compactContactsKernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offsets,
                                      const AxcdContact* __restrict__ tmp, uint32_t npairs,
                                      AxcdContact* __restrict__ out, uint32_t maxContacts) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs || !flags[k]) return;
    const uint32_t dst = offsets[k];
    if (dst >= maxContacts) return;
    const float2* s = reinterpret_cast<const float2*>(tmp + k);
    float2* o = reinterpret_cast<float2*>(out + dst);
#pragma unroll
    for (int i = 0; i < 5; ++i) o[i] = s[i];
}

}  // namespace axcd
