/*
 * axref.cpp — CPU ORACLE for the collision hot path (refit -> broadphase -> GJK/EPA).
 *
 * TEST INFRASTRUCTURE ONLY (see axref.h).  Scalar C++, IEEE binary32, built with
 * -ffp-contract=off so the compiler never fuses a multiply-add on its own: every expression below
 * rounds exactly as written, and the only fused operations are the explicit std::fmaf calls in
 * dot / cross / quatRotate.  The CUDA path (built -fmad=false, same explicit fmaf calls) is
 * therefore bit-identical.
 *
 * What is restated from the reference (file:line under /root/reference) and what is authored:
 *   Vec3 ops ................ include/axiom/math/vec3.hpp:15-280            (restated)
 *   Quat * Vec3 ............. src/math/quat.cpp:33-38 -> glm::quat * glm::vec3 (GLM 1.0.x scalar
 *                             formula; GLM itself is not in the snapshot)
 *   Quat -> matrix .......... src/math/quat.cpp:117-129 -> glm::mat4_cast   (Appendix B)
 *   Transform::transformPoint src/math/transform.cpp:86-93                  (restated)
 *   AABB ops ................ include/axiom/math/aabb.hpp:47,132-135,143-160,213-215 (restated)
 *   box corner order ........ src/debug/debug_draw.cpp:97-113               (restated)
 *   sphere placement ........ src/debug/physics_debug_draw.cpp:246-248      (restated)
 *   hull placement .......... src/debug/debug_draw.cpp:431-448              (restated)
 *   DeterministicRNG ........ include/axiom/math/random.hpp:19-68           (restated)
 *   broadphase, GJK, EPA .... ABSENT from the reference (src/collision/.gitkeep) — authored here
 *                             from the spec in SURVEY.md Appendix A; "parity unpinned".
 */
#include "axref.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// Vec3 (include/axiom/math/vec3.hpp)
// ------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }   // vec3.hpp:52
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }   // vec3.hpp:60
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }      // vec3.hpp:84
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }                        // vec3.hpp:100
// dot / cross: the reference writes them as plain products and sums (vec3.hpp:179-191) and builds
// with -mfma and GCC's default contraction, which turns those into fused multiply-adds in a
// compiler-chosen pattern.  This restatement pins ONE explicit pattern (std::fmaf = one rounding) — the one
// GCC 13 generates for those formulas under the reference's own flags, checked bit for bit against the
// compiled reference headers in tests/test_oracle_vs_reference.py — identical in the CUDA kernels.
inline float dot(V3 a, V3 b) { return std::fmaf(a.z, b.z, std::fmaf(a.x, b.x, a.y * b.y)); }
inline V3 cross(V3 a, V3 b) {
    return mk(std::fmaf(-a.z, b.y, a.y * b.z), std::fmaf(-a.x, b.z, a.z * b.x),
              std::fmaf(a.x, b.y, -(a.y * b.x)));
}
inline bool same(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct Q4 {
    float x, y, z, w;
};

// glm::quat * glm::vec3 (reached from src/math/quat.cpp:33-38):
//   uv = cross(u, v); uuv = cross(u, uv); v + ((uv * q.w) + uuv) * 2
inline V3 quatRotate(Q4 q, V3 v) {
    V3 u = mk(q.x, q.y, q.z);
    V3 uv = cross(u, v);
    V3 uuv = cross(u, uv);
    V3 t = mk(std::fmaf(uv.x, q.w, uuv.x), std::fmaf(uv.y, q.w, uuv.y), std::fmaf(uv.z, q.w, uuv.z));
    return mk(std::fmaf(t.x, 2.0f, v.x), std::fmaf(t.y, 2.0f, v.y), std::fmaf(t.z, 2.0f, v.z));
}

// glm::quat * glm::quat (src/math/quat.cpp:27-31)
inline Q4 quatMul(Q4 p, Q4 q) {
    Q4 r;
    r.w = p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z;
    r.x = p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y;
    r.y = p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z;
    r.z = p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x;
    return r;
}

// glm::mat3_cast (src/math/quat.cpp:117-129 via mat4_cast): columns c0,c1,c2.
struct M3 {
    V3 c0, c1, c2;
};
inline M3 quatToMat3(Q4 q) {
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    M3 m;
    m.c0 = mk(1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy));
    m.c1 = mk(2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx));
    m.c2 = mk(2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy));
    return m;
}

// axiom::math::Transform, 10 floats: position 0..2, rotation (x,y,z,w) 3..6, scale 7..9
struct Xf {
    V3 p;
    Q4 q;
    V3 s;
};
inline Xf loadXf(const float* f) {
    Xf t;
    t.p = mk(f[0], f[1], f[2]);
    t.q = Q4{f[3], f[4], f[5], f[6]};
    t.s = mk(f[7], f[8], f[9]);
    return t;
}
// Transform::transformPoint (src/math/transform.cpp:86-93)
inline V3 transformPoint(const Xf& t, V3 p) {
    V3 scaled = mk(p.x * t.s.x, p.y * t.s.y, p.z * t.s.z);
    V3 rotated = quatRotate(t.q, scaled);
    return rotated + t.p;
}

// Quat::conjugate (include/axiom/math/quat.hpp:96): negates the vector part
inline Q4 conjugate(Q4 q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
// Transform::transformDirection (src/math/transform.cpp:95-100): scale, rotate, no translation
inline V3 transformDirection(const Xf& t, V3 d) {
    V3 scaled = mk(d.x * t.s.x, d.y * t.s.y, d.z * t.s.z);
    return quatRotate(t.q, scaled);
}
// Transform::inverseTransformPoint (src/math/transform.cpp:113-122): untranslate, conjugate-rotate, multiply by 1/scale
inline V3 inverseTransformPoint(const Xf& t, V3 p) {
    V3 translated = p - t.p;
    V3 rotated = quatRotate(conjugate(t.q), translated);
    V3 inv = mk(1.0f / t.s.x, 1.0f / t.s.y, 1.0f / t.s.z);
    return mk(rotated.x * inv.x, rotated.y * inv.y, rotated.z * inv.z);
}
// Transform::inverseTransformDirection (src/math/transform.cpp:124-131)
inline V3 inverseTransformDirection(const Xf& t, V3 d) {
    V3 rotated = quatRotate(conjugate(t.q), d);
    V3 inv = mk(1.0f / t.s.x, 1.0f / t.s.y, 1.0f / t.s.z);
    return mk(rotated.x * inv.x, rotated.y * inv.y, rotated.z * inv.z);
}

// AABB (include/axiom/math/aabb.hpp)
struct Box {
    V3 lo, hi;
};
inline void expandPoint(Box& b, V3 p) {   // aabb.hpp:143-150
    b.lo.x = (p.x < b.lo.x) ? p.x : b.lo.x;
    b.lo.y = (p.y < b.lo.y) ? p.y : b.lo.y;
    b.lo.z = (p.z < b.lo.z) ? p.z : b.lo.z;
    b.hi.x = (p.x > b.hi.x) ? p.x : b.hi.x;
    b.hi.y = (p.y > b.hi.y) ? p.y : b.hi.y;
    b.hi.z = (p.z > b.hi.z) ? p.z : b.hi.z;
}
inline bool intersects(const float* a, const float* b) {   // aabb.hpp:132-135 (closed intervals)
    return a[0] <= b[3] && a[3] >= b[0] && a[1] <= b[4] && a[4] >= b[1] && a[2] <= b[5] &&
           a[5] >= b[2];
}

// gui::FilterInfo (include/axiom/gui/body_inspector.hpp:38-42): category / mask bits and a group
// index, Box2D rule: same non-zero group -> collide iff the group is positive; otherwise both masks
// must accept the other's category.  filt = n x 3 words (categoryBits, maskBits, groupIndex as int32).
inline bool shouldCollide(const uint32_t* filt, uint32_t i, uint32_t j) {
    if (!filt) return true;
    const uint32_t* a = filt + 3ull * i;
    const uint32_t* b = filt + 3ull * j;
    const int32_t ga = (int32_t)a[2], gb = (int32_t)b[2];
    if (ga == gb && ga != 0) return ga > 0;
    return (a[1] & b[0]) != 0 && (a[0] & b[1]) != 0;
}

// DeterministicRNG (include/axiom/math/random.hpp:19-68)
struct Rng {
    uint64_t state;
    explicit Rng(uint64_t seed) : state(seed | 1ULL) {
        for (int i = 0; i < 10; ++i) next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + 1442695040888963407ULL;
        uint32_t xs = static_cast<uint32_t>(((old >> 18U) ^ old) >> 27U);
        uint32_t rot = static_cast<uint32_t>(old >> 59U);
        return (xs >> rot) | (xs << ((~rot + 1U) & 31));
    }
    float nextFloat() { return static_cast<float>(next()) / static_cast<float>(0x100000000ULL); }
};

enum { SHAPE_SPHERE = 0, SHAPE_BOX = 1, SHAPE_CAPSULE = 2, SHAPE_CONVEX = 4, SHAPE_CYLINDER = 6 };

inline uint32_t fbits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

template <class F>
void parallelFor(uint64_t n, int nthreads, F&& fn) {
    if (nthreads <= 1 || n < 2048) {
        fn(0, 0, n);
        return;
    }
    std::vector<std::thread> th;
    uint64_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        uint64_t lo = std::min<uint64_t>(n, t * chunk), hi = std::min<uint64_t>(n, lo + chunk);
        th.emplace_back([&fn, t, lo, hi] { fn(t, lo, hi); });
    }
    for (auto& x : th) x.join();
}

// ------------------------------------------------------------------------------------------
// Stage 1: refit
// ------------------------------------------------------------------------------------------
// The alternative refit route through AABB::transform(Mat4) (src/math/aabb.cpp:8-35) with M = Transform::toMatrix()
// = T * R * S (src/math/transform.cpp:13-24): M's upper-left 3x3 is mat3_cast(q) with column j scaled by scale_j
// (the products with the identity's zeros and ones are exact), the last column is the position; the 8 corners
// of the LOCAL box [-h, h] in aabb.cpp's order (x fastest) go through glm's mat4 * vec4(c, 1) =
// (m[0]*x + m[1]*y) + (m[2]*z + m[3]), each product-sum one fused operation; AABB(Vec3) then expand().
Box refitBoxMat4(const Xf& t, float hx, float hy, float hz) {
    M3 r = quatToMat3(t.q);
    const V3 c0 = r.c0 * t.s.x, c1 = r.c1 * t.s.y, c2 = r.c2 * t.s.z;
    auto xform = [&](V3 c) {
        return mk(std::fmaf(c1.x, c.y, c0.x * c.x) + std::fmaf(c2.x, c.z, t.p.x),
                  std::fmaf(c1.y, c.y, c0.y * c.x) + std::fmaf(c2.y, c.z, t.p.y),
                  std::fmaf(c1.z, c.y, c0.z * c.x) + std::fmaf(c2.z, c.z, t.p.z));
    };
    Box b;
    for (int k = 0; k < 8; ++k) {
        const V3 c = mk((k & 1) ? hx : -hx, (k & 2) ? hy : -hy, (k & 4) ? hz : -hz);
        const V3 p = xform(c);
        if (k == 0) {
            b.lo = p;
            b.hi = p;
        } else {
            expandPoint(b, p);
        }
    }
    return b;
}

int refitOne(const Xf& t, const AxrefShape& s, const float* hull, uint32_t nHull, float margin,
             float* out, bool mat4Route = false) {
    Box b;
    if (s.type == SHAPE_SPHERE) {
        // AABB::fromCenterExtents(position, Vec3(r)) (aabb.hpp:213-215); rotation/scale ignored
        // like the reference's sphere placement (src/debug/physics_debug_draw.cpp:246-248).
        V3 r = mk(s.p0, s.p0, s.p0);
        b.lo = t.p - r;
        b.hi = t.p + r;
    } else if (s.type == SHAPE_BOX && mat4Route) {
        b = refitBoxMat4(t, s.p0, s.p1, s.p2);
    } else if (s.type == SHAPE_BOX) {
        // corner order of src/debug/debug_draw.cpp:99-108
        float hx = s.p0, hy = s.p1, hz = s.p2;
        const V3 c[8] = {mk(-hx, -hy, -hz), mk(hx, -hy, -hz), mk(hx, -hy, hz), mk(-hx, -hy, hz),
                         mk(-hx, hy, -hz),  mk(hx, hy, -hz),  mk(hx, hy, hz),  mk(-hx, hy, hz)};
        V3 p0 = transformPoint(t, c[0]);
        b.lo = p0;   // AABB(Vec3) (aabb.hpp:47)
        b.hi = p0;
        for (int k = 1; k < 8; ++k) expandPoint(b, transformPoint(t, c[k]));
    } else if (s.type == SHAPE_CAPSULE) {
        // p0 = radius, p1 = height.  Endpoints as the reference places them
        // (src/debug/physics_debug_draw.cpp:254-266): local (0, -+height/2, 0) through transformPoint;
        // the radius is not scaled.  Box of the two end spheres.
        V3 half = mk(0.0f, s.p1 * 0.5f, 0.0f);
        V3 start = transformPoint(t, -half);
        V3 end = transformPoint(t, half);
        b.lo = start;
        b.hi = start;
        expandPoint(b, end);
        V3 r = mk(s.p0, s.p0, s.p0);
        b.lo = b.lo - r;
        b.hi = b.hi + r;
    } else if (s.type == SHAPE_CYLINDER) {
        // gui::ShapeType::Cylinder (include/axiom/gui/body_inspector.hpp:24): p0 = radius, p1 = height, local Y axis
        // like the capsule (src/debug/physics_debug_draw.cpp:254-266), placed by transformPoint (scale, rotate,
        // translate).  Exact box of the scaled cylinder: along world axis k the half extent is the support
        // distance h/2 * |l.y| + r * |(l.x, l.z)| of the local direction l = scale * (R^T e_k).
        M3 m = quatToMat3(t.q);
        const float cx[3] = {m.c0.x, m.c0.y, m.c0.z}, cy[3] = {m.c1.x, m.c1.y, m.c1.z}, cz[3] = {m.c2.x, m.c2.y, m.c2.z};
        float ext[3];
        for (int k = 0; k < 3; ++k) {
            const float lx = cx[k] * t.s.x, ly = cy[k] * t.s.y, lz = cz[k] * t.s.z;
            ext[k] = (s.p1 * 0.5f) * std::fabs(ly) + s.p0 * std::sqrt(lx * lx + lz * lz);
        }
        b.lo = t.p - mk(ext[0], ext[1], ext[2]);
        b.hi = t.p + mk(ext[0], ext[1], ext[2]);
    } else if (s.type == SHAPE_CONVEX) {
        uint32_t first = fbits(s.p0), cnt = fbits(s.p1);
        if (cnt == 0 || static_cast<uint64_t>(first) + cnt > nHull) return 300;
        const float* v = hull + 3ull * first;
        V3 p0 = transformPoint(t, mk(v[0], v[1], v[2]));
        b.lo = p0;
        b.hi = p0;
        for (uint32_t k = 1; k < cnt; ++k)
            expandPoint(b, transformPoint(t, mk(v[3 * k], v[3 * k + 1], v[3 * k + 2])));
    } else {
        return 300;
    }
    if (margin != 0.0f) {   // AABB::expand(float) (aabb.hpp:156-160)
        V3 m = mk(margin, margin, margin);
        b.lo = b.lo - m;
        b.hi = b.hi + m;
    }
    out[0] = b.lo.x; out[1] = b.lo.y; out[2] = b.lo.z;
    out[3] = b.hi.x; out[4] = b.hi.y; out[5] = b.hi.z;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Stage 3: narrowphase.  Convex "cores" in a frame translated so body A's position is the origin
// (axes stay world-aligned).  Sphere = point core + radius margin; box and hull have no margin.
// ------------------------------------------------------------------------------------------
enum { CORE_POINT = 0, CORE_BOX = 1, CORE_HULL = 2, CORE_SEGMENT = 3, CORE_CYLINDER = 4 };
struct Core {
    int kind;
    V3 c;            // centre relative to A's position
    V3 e0, e1, e2;   // box: rotation columns * (halfExtent*scale); hull: rotation columns
    V3 s;            // hull: scale
    const float* verts;
    uint32_t nv;
    float r;         // sphere radius
};

Core makeCore(const Xf& t, const AxrefShape& sh, const float* hull, V3 origin) {
    Core k{};
    k.c = t.p - origin;
    k.r = 0.0f;
    if (sh.type == SHAPE_SPHERE) {
        k.kind = CORE_POINT;
        k.r = sh.p0;
    } else if (sh.type == SHAPE_CAPSULE) {
        // segment core along the local Y axis (half length height/2, scaled by scale.y) + radius
        k.kind = CORE_SEGMENT;
        M3 m = quatToMat3(t.q);
        k.e0 = m.c1 * ((sh.p1 * 0.5f) * t.s.y);
        k.r = sh.p0;
    } else if (sh.type == SHAPE_CYLINDER) {
        // radial frame E0, E2 (rotation columns x / z scaled by scale * radius) and half axis E1 (column y scaled by
        // scale.y * height / 2): a rim point is c +- E1 + E0 * cos + E2 * sin
        k.kind = CORE_CYLINDER;
        M3 m = quatToMat3(t.q);
        k.e0 = m.c0 * (t.s.x * sh.p0);
        k.e1 = m.c1 * (t.s.y * (sh.p1 * 0.5f));
        k.e2 = m.c2 * (t.s.z * sh.p0);
    } else if (sh.type == SHAPE_BOX) {
        k.kind = CORE_BOX;
        M3 m = quatToMat3(t.q);
        k.e0 = m.c0 * (sh.p0 * t.s.x);
        k.e1 = m.c1 * (sh.p1 * t.s.y);
        k.e2 = m.c2 * (sh.p2 * t.s.z);
    } else {
        k.kind = CORE_HULL;
        M3 m = quatToMat3(t.q);
        k.e0 = m.c0; k.e1 = m.c1; k.e2 = m.c2;
        k.s = t.s;
        k.verts = hull + 3ull * fbits(sh.p0);
        k.nv = fbits(sh.p1);
    }
    return k;
}

// Cylinder support.  The rim direction is quantised to 4 x 8192 directions (a 32768-gon inscribed in the rim:
// chord error 5e-9 of the radius, below float resolution) so that a support point is named by 16 bits — sign of
// the E0 part, sign of the E2 part, 13 bits of |E2 share| in the diamond parametrisation |u| + |w| = 1, cap sign
// — and can be rebuilt bit for bit from that id (the CUDA path keeps ids, not points, in its polytopes).
inline uint32_t cylinderId(const Core& k, V3 d) {
    const float a = dot(d, k.e0), b = dot(d, k.e2);
    const float s = std::fabs(a) + std::fabs(b);
    uint32_t m = 0;
    if (s > 0.0f) {
        const float w = std::fabs(b) / s;
        m = (uint32_t)(w * 8191.0f + 0.5f);
        if (m > 8191u) m = 8191u;
    }
    return (!(a >= 0.0f) ? 1u : 0u) | (!(b >= 0.0f) ? 2u : 0u) | (m << 2) | (!(dot(d, k.e1) >= 0.0f) ? 0x8000u : 0u);
}
inline V3 cylinderPoint(const Core& k, uint32_t id) {
    const float w = (float)((id >> 2) & 8191u) / 8191.0f, u = 1.0f - w;
    const float inv = 1.0f / std::sqrt(u * u + w * w);
    const float fx = (id & 1u) ? -(u * inv) : (u * inv), fz = (id & 2u) ? -(w * inv) : (w * inv);
    const V3 cap = (id & 0x8000u) ? -k.e1 : k.e1;
    return ((k.e0 * fx + cap) + k.e2 * fz) + k.c;
}

// Support point of a core in world-aligned direction d (need not be unit length).
V3 support(const Core& k, V3 d) {
    if (k.kind == CORE_POINT) return k.c;
    if (k.kind == CORE_CYLINDER) return cylinderPoint(k, cylinderId(k, d));
    if (k.kind == CORE_SEGMENT) return k.c + ((dot(d, k.e0) >= 0.0f) ? k.e0 : -k.e0);
    if (k.kind == CORE_BOX) {
        V3 p = k.c;
        p = p + ((dot(d, k.e0) >= 0.0f) ? k.e0 : -k.e0);
        p = p + ((dot(d, k.e1) >= 0.0f) ? k.e1 : -k.e1);
        p = p + ((dot(d, k.e2) >= 0.0f) ? k.e2 : -k.e2);
        return p;
    }
    // hull: local direction = scale * (R^T d); first maximal vertex wins
    V3 l = mk(dot(d, k.e0) * k.s.x, dot(d, k.e1) * k.s.y, dot(d, k.e2) * k.s.z);
    uint32_t best = 0;
    float bestDot = dot(l, mk(k.verts[0], k.verts[1], k.verts[2]));
    for (uint32_t i = 1; i < k.nv; ++i) {
        float di = dot(l, mk(k.verts[3 * i], k.verts[3 * i + 1], k.verts[3 * i + 2]));
        if (di > bestDot) {
            bestDot = di;
            best = i;
        }
    }
    V3 lv = mk(k.verts[3 * best] * k.s.x, k.verts[3 * best + 1] * k.s.y,
               k.verts[3 * best + 2] * k.s.z);
    return ((k.e0 * lv.x + k.e1 * lv.y) + k.e2 * lv.z) + k.c;
}

struct Simplex {
    V3 y[4];   // points of the Minkowski difference A - B
    V3 a[4];   // the matching support points on A
    float lam[4];
    int n;
};

const float GJK_EPS_ABS2 = 1e-12f;   // |v|^2 at or below this: origin is on the simplex
const float DEGENERATE_EPS = 1e-12f;

// Closest point to the origin on segment [a,b]; mask bit0=a bit1=b.
inline V3 closestSegment(V3 a, V3 b, float* la, float* lb, int* mask) {
    V3 ab = b - a;
    float t = -dot(a, ab);
    if (t <= 0.0f) {
        *la = 1.0f; *lb = 0.0f; *mask = 1;
        return a;
    }
    float denom = dot(ab, ab);
    if (t >= denom) {
        *la = 0.0f; *lb = 1.0f; *mask = 2;
        return b;
    }
    t = t / denom;
    *la = 1.0f - t; *lb = t; *mask = 3;
    return a + ab * t;
}

// Closest point to the origin on triangle (a,b,c) by Voronoi regions; mask bits a=1,b=2,c=4.
inline V3 closestTriangle(V3 a, V3 b, V3 c, float* la, float* lb, float* lc, int* mask) {
    V3 ab = b - a, ac = c - a;
    float d1 = -dot(ab, a), d2 = -dot(ac, a);
    if (d1 <= 0.0f && d2 <= 0.0f) {
        *la = 1.0f; *lb = 0.0f; *lc = 0.0f; *mask = 1;
        return a;
    }
    float d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.0f && d4 <= d3) {
        *la = 0.0f; *lb = 1.0f; *lc = 0.0f; *mask = 2;
        return b;
    }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float v = d1 / (d1 - d3);
        *la = 1.0f - v; *lb = v; *lc = 0.0f; *mask = 3;
        return a + ab * v;
    }
    float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.0f && d5 <= d6) {
        *la = 0.0f; *lb = 0.0f; *lc = 1.0f; *mask = 4;
        return c;
    }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        *la = 1.0f - w; *lb = 0.0f; *lc = w; *mask = 5;
        return a + ac * w;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        *la = 0.0f; *lb = 1.0f - w; *lc = w; *mask = 6;
        return b + (c - b) * w;
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    *la = (1.0f - v) - w; *lb = v; *lc = w; *mask = 7;
    return (a + ab * v) + ac * w;
}

// Is the origin strictly on the other side of plane (a,b,c) from d?  Degenerate (d in the plane)
// counts as outside so the face is still examined.
inline bool originOutside(V3 a, V3 b, V3 c, V3 d) {
    V3 n = cross(b - a, c - a);
    V3 ad = d - a;
    float sp = -dot(a, n);
    float sd = dot(ad, n);
    if (sd * sd <= DEGENERATE_EPS * (dot(n, n) * dot(ad, ad))) return true;   // flat tetrahedron
    return sp * sd < 0.0f;
}

// Reduce the simplex to the feature closest to the origin, set lam[], return that closest point.
// *enclosed = true when a tetrahedron contains the origin.
V3 solveSimplex(Simplex& s, bool* enclosed) {
    *enclosed = false;
    int mask = 0;
    float l[4] = {0, 0, 0, 0};
    V3 v = mk(0, 0, 0);
    if (s.n == 2) {
        v = closestSegment(s.y[0], s.y[1], &l[0], &l[1], &mask);
    } else if (s.n == 3) {
        v = closestTriangle(s.y[0], s.y[1], s.y[2], &l[0], &l[1], &l[2], &mask);
    } else {
        // tetrahedron: examine the faces the origin is outside of, keep the nearest result
        static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
        float best = FLT_MAX;
        bool any = false;
        for (int f = 0; f < 4; ++f) {
            const int i = F[f][0], j = F[f][1], k = F[f][2], o = F[f][3];
            if (!originOutside(s.y[i], s.y[j], s.y[k], s.y[o])) continue;
            any = true;
            float li, lj, lk;
            int m;
            V3 q = closestTriangle(s.y[i], s.y[j], s.y[k], &li, &lj, &lk, &m);
            float qq = dot(q, q);
            if (qq < best) {
                best = qq;
                v = q;
                l[0] = l[1] = l[2] = l[3] = 0.0f;
                l[i] = li; l[j] = lj; l[k] = lk;
                mask = ((m & 1) ? (1 << i) : 0) | ((m & 2) ? (1 << j) : 0) |
                       ((m & 4) ? (1 << k) : 0);
            }
        }
        if (!any) {
            *enclosed = true;
            return mk(0, 0, 0);
        }
    }
    int n = 0;
    for (int i = 0; i < s.n; ++i) {
        if (mask & (1 << i)) {
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
    int state;      // GJK_SEPARATED: v is the closest point of A-B to the origin (or, without
                    //   wantDistances, possibly only a separating direction); GJK_OVERLAP: cores
                    //   touch/overlap
    bool exact;     // separated && v converged (distance valid)
    V3 v;
    float vv;
    uint32_t status;
    uint32_t iters;
};

GjkResult gjk(const Core& A, const Core& B, const AxrefNarrowCfg& cfg, float marginSum,
              Simplex& s) {
    GjkResult r{};
    V3 d0 = B.c - A.c;
    if (dot(d0, d0) < 1e-12f) d0 = mk(1.0f, 0.0f, 0.0f);
    s.a[0] = support(A, d0);
    s.y[0] = s.a[0] - support(B, -d0);
    s.lam[0] = 1.0f;
    s.n = 1;
    V3 v = s.y[0];
    float vv = dot(v, v);
    r.state = GJK_SEPARATED;
    r.exact = true;
    uint32_t it = 0;
    for (;; ++it) {
        if (vv <= GJK_EPS_ABS2) {
            r.state = GJK_OVERLAP;
            break;
        }
        if (it >= cfg.gjkMaxIters) {
            r.status = 301;
            break;
        }
        V3 a = support(A, -v);
        V3 w = a - support(B, v);
        float vw = dot(v, w);
        if (!cfg.wantDistances && vw > 0.0f && vw * vw > vv * (marginSum * marginSum)) {
            r.exact = false;   // separating axis with a gap larger than the margins
            break;
        }
        if (vv - vw <= cfg.gjkTol * vv) break;   // converged: no closer point in direction -v
        bool dup = false;
        for (int i = 0; i < s.n; ++i) dup = dup || same(w, s.y[i]);
        if (dup) break;
        s.y[s.n] = w;
        s.a[s.n] = a;
        s.n++;
        bool enclosed;
        V3 nv = solveSimplex(s, &enclosed);
        if (enclosed) {
            r.state = GJK_OVERLAP;
            v = nv;
            vv = 0.0f;
            break;
        }
        float nvv = dot(nv, nv);
        bool stalled = nvv >= vv;
        v = nv;
        vv = nvv;
        if (stalled) {
            if (vv <= GJK_EPS_ABS2) r.state = GJK_OVERLAP;
            break;
        }
    }
    r.v = v;
    r.vv = vv;
    r.iters = it;
    return r;
}

// ---- EPA ------------------------------------------------------------------------------------
const int EPA_MAX_VERTS = 40;
const int EPA_MAX_FACES = 64;
struct EpaFace {
    uint8_t i0, i1, i2, alive;
    V3 n;
    float d;
};
struct Epa {
    V3 y[EPA_MAX_VERTS], a[EPA_MAX_VERTS];
    int nv;
    EpaFace f[EPA_MAX_FACES];
    int nf;   // slots in use (alive or dead)
};

inline void epaSetFace(Epa& e, int slot, int i0, int i1, int i2) {
    EpaFace& F = e.f[slot];
    V3 n = cross(e.y[i1] - e.y[i0], e.y[i2] - e.y[i0]);
    float len2 = dot(n, n);
    F.alive = 1;
    if (len2 <= 1e-30f) {   // zero-area face: never the closest, never visible
        F.i0 = (uint8_t)i0; F.i1 = (uint8_t)i1; F.i2 = (uint8_t)i2;
        F.n = mk(0, 0, 0);
        F.d = FLT_MAX;
        return;
    }
    float inv = 1.0f / std::sqrt(len2);
    n = n * inv;
    // Outward orientation comes from the winding (initial tetrahedron oriented explicitly, new
    // faces inherit the horizon edge direction); d may be a hair negative when the origin sits on
    // the boundary, and such a face is then simply the first one expanded.
    float d = dot(n, e.y[i0]);
    F.i0 = (uint8_t)i0; F.i1 = (uint8_t)i1; F.i2 = (uint8_t)i2;
    F.n = n;
    F.d = d;
}

struct EpaResult {
    V3 n;        // unit, from A to B
    float depth; // core penetration depth (>= 0)
    V3 pa;       // witness point on A's core (A-centred frame)
    V3 pb;       // witness point on B's core
    uint32_t status;
};

// Fallback for a Minkowski difference that is degenerate at the origin: zero depth along n.
inline EpaResult epaTouching(V3 n, V3 pa) {
    EpaResult r;
    float l2 = dot(n, n);
    r.n = (l2 > 0.0f) ? n * (1.0f / std::sqrt(l2)) : mk(1.0f, 0.0f, 0.0f);
    r.depth = 0.0f;
    r.pa = pa;
    r.pb = pa;
    r.status = 0;
    return r;
}

#ifdef AXREF_EPA_HIST
// Diagnostic build only (make -C oracle hist): distribution of the polytope size EPA ends with, used
// to size the CUDA fast-path storage caps (profiles/r01_experiments.md).
static std::atomic<uint64_t> gEpaHistV[EPA_MAX_VERTS + 1], gEpaHistF[EPA_MAX_FACES + 1];
inline void epaHistRecord(int nv, int nf) {
    gEpaHistV[nv].fetch_add(1, std::memory_order_relaxed);
    gEpaHistF[nf].fetch_add(1, std::memory_order_relaxed);
}
extern "C" void axref_epa_hist(uint64_t* outV, uint64_t* outF) {
    for (int i = 0; i <= EPA_MAX_VERTS; ++i) outV[i] = gEpaHistV[i].exchange(0);
    for (int i = 0; i <= EPA_MAX_FACES; ++i) outF[i] = gEpaHistF[i].exchange(0);
}
#endif

EpaResult epa(const Core& A, const Core& B, const AxrefNarrowCfg& cfg, const Simplex& s0) {
    Epa e;
    e.nv = s0.n;
    for (int i = 0; i < s0.n; ++i) {
        e.y[i] = s0.y[i];
        e.a[i] = s0.a[i];
    }
    auto addSupport = [&](V3 d, int slot) {
        e.a[slot] = support(A, d);
        e.y[slot] = e.a[slot] - support(B, -d);
    };
    // --- grow the GJK simplex to a tetrahedron -------------------------------------------------
    if (e.nv == 1) {
        const V3 ax[6] = {mk(1, 0, 0), mk(-1, 0, 0), mk(0, 1, 0), mk(0, -1, 0), mk(0, 0, 1),
                          mk(0, 0, -1)};
        for (int k = 0; k < 6 && e.nv == 1; ++k) {
            addSupport(ax[k], 1);
            V3 d = e.y[1] - e.y[0];
            if (dot(d, d) > DEGENERATE_EPS) e.nv = 2;
        }
        if (e.nv == 1) return epaTouching(mk(1, 0, 0), e.a[0]);
    }
    if (e.nv == 2) {
        V3 d = e.y[1] - e.y[0];
        const V3 ax[3] = {mk(1, 0, 0), mk(0, 1, 0), mk(0, 0, 1)};
        V3 firstDir = mk(0, 0, 0);
        for (int k = 0; k < 3 && e.nv == 2; ++k) {
            V3 dir = cross(d, ax[k]);
            if (dot(dir, dir) <= DEGENERATE_EPS) continue;
            if (dot(firstDir, firstDir) == 0.0f) firstDir = dir;
            for (int sgn = 0; sgn < 2 && e.nv == 2; ++sgn) {
                addSupport(sgn ? -dir : dir, 2);
                V3 c = cross(e.y[2] - e.y[0], d);
                if (dot(c, c) > DEGENERATE_EPS) e.nv = 3;
            }
        }
        if (e.nv == 2) return epaTouching(firstDir, e.a[0]);
    }
    if (e.nv == 3) {
        V3 n = cross(e.y[1] - e.y[0], e.y[2] - e.y[0]);
        float n2 = dot(n, n);
        if (n2 <= 1e-30f) return epaTouching(mk(1, 0, 0), e.a[0]);
        for (int sgn = 0; sgn < 2 && e.nv == 3; ++sgn) {
            addSupport(sgn ? -n : n, 3);
            float vol = dot(e.y[3] - e.y[0], n);
            if (vol * vol > DEGENERATE_EPS * n2) e.nv = 4;
        }
        if (e.nv == 3) {
            // flat at the origin: M has no thickness along n -> touching contact along +n
            float la, lb, lc;
            int m;
            closestTriangle(e.y[0], e.y[1], e.y[2], &la, &lb, &lc, &m);
            V3 pa = (e.a[0] * la + e.a[1] * lb) + e.a[2] * lc;
            return epaTouching(n, pa);
        }
    }
    // orientation: make (0,1,2) face away from vertex 3
    if (dot(cross(e.y[1] - e.y[0], e.y[2] - e.y[0]), e.y[3] - e.y[0]) > 0.0f) {
        std::swap(e.y[0], e.y[1]);
        std::swap(e.a[0], e.a[1]);
    }
    epaSetFace(e, 0, 0, 1, 2);
    epaSetFace(e, 1, 0, 3, 1);
    epaSetFace(e, 2, 0, 2, 3);
    epaSetFace(e, 3, 1, 3, 2);
    e.nf = 4;
    const int maxFaces = (int)std::min<uint32_t>(cfg.epaMaxFaces, EPA_MAX_FACES);

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
        if (best < 0) {   // every face degenerate: give up on this polytope
            return epaTouching(mk(1, 0, 0), e.a[0]);
        }
        const EpaFace fb = e.f[best];
        V3 a = support(A, fb.n);
        V3 w = a - support(B, -fb.n);
        float dw = dot(w, fb.n);
        float scale = (fb.d > 1.0f) ? fb.d : 1.0f;
        if (dw - fb.d <= cfg.epaTol * scale) break;
        bool dup = false;
        for (int i = 0; i < e.nv; ++i) dup = dup || same(w, e.y[i]);
        if (dup) break;
        if (it >= cfg.epaMaxIters || e.nv >= EPA_MAX_VERTS) {
            status = 302;
            break;
        }
        // visible faces and horizon
        uint8_t he0[EPA_MAX_FACES * 3], he1[EPA_MAX_FACES * 3];
        int nh = 0, nvis = 0, nalive = 0;
        uint8_t vis[EPA_MAX_FACES];
        // A face counts as visible from w only when w is clearly in front of it: faces coplanar
        // with w (common on boxes) must not be split by rounding noise into a ragged horizon.
        const float wl = std::fabs(w.x) + std::fabs(w.y) + std::fabs(w.z);
        const float visEps = 1e-6f * ((wl > 1.0f) ? wl : 1.0f);
        for (int i = 0; i < e.nf; ++i) {
            vis[i] = 0;
            if (!e.f[i].alive) continue;
            ++nalive;
            if (dot(e.f[i].n, w) - e.f[i].d > visEps) {
                vis[i] = 1;
                ++nvis;
            }
        }
        // Horizon, in canonical order: visible faces by ascending slot, their edges in winding
        // order, keeping a directed edge only if its reverse does not belong to a visible face.
        uint64_t visEdge[EPA_MAX_VERTS];
        for (int i = 0; i < e.nv; ++i) visEdge[i] = 0;
        for (int i = 0; i < e.nf; ++i) {
            if (!vis[i]) continue;
            visEdge[e.f[i].i0] |= 1ull << e.f[i].i1;
            visEdge[e.f[i].i1] |= 1ull << e.f[i].i2;
            visEdge[e.f[i].i2] |= 1ull << e.f[i].i0;
        }
        for (int i = 0; i < e.nf; ++i) {
            if (!vis[i]) continue;
            const uint8_t ev[3][2] = {{e.f[i].i0, e.f[i].i1}, {e.f[i].i1, e.f[i].i2},
                                      {e.f[i].i2, e.f[i].i0}};
            for (int k = 0; k < 3; ++k) {
                if ((visEdge[ev[k][1]] >> ev[k][0]) & 1ull) continue;   // shared by two visible faces
                he0[nh] = ev[k][0];
                he1[nh] = ev[k][1];
                ++nh;
            }
        }
        // the horizon must be a simple loop: every vertex starts exactly one edge
        bool loopOk = nh >= 3;
        for (int h = 0; h < nh && loopOk; ++h)
            for (int g = h + 1; g < nh; ++g)
                if (he0[g] == he0[h] || he1[g] == he1[h]) {
                    loopOk = false;
                    break;
                }
        if (!loopOk || nalive - nvis + nh > maxFaces) {
            status = 302;
            break;
        }
        const int wi = e.nv;
        e.y[wi] = w;
        e.a[wi] = a;
        e.nv++;
        for (int i = 0; i < e.nf; ++i)
            if (vis[i]) e.f[i].alive = 0;
        int slot = 0;
        for (int h = 0; h < nh; ++h) {
            while (slot < e.nf && e.f[slot].alive) ++slot;
            if (slot == e.nf) e.nf++;
            epaSetFace(e, slot, he0[h], he1[h], wi);
        }
    }
#ifdef AXREF_EPA_HIST
    epaHistRecord(e.nv, e.nf);
#endif
    const EpaFace fb = e.f[best];
    EpaResult r;
    r.n = fb.n;
    r.depth = (fb.d > 0.0f) ? fb.d : 0.0f;
    float la, lb, lc;
    int m;
    V3 p = closestTriangle(e.y[fb.i0], e.y[fb.i1], e.y[fb.i2], &la, &lb, &lc, &m);
    r.pa = (e.a[fb.i0] * la + e.a[fb.i1] * lb) + e.a[fb.i2] * lc;
    r.pb = r.pa - p;
    r.status = status;
    return r;
}

// Sphere against box in closed form (the box as in makeCore: centre cX, unit axes = the columns of
// Quat::toMatrix, half lengths |halfExtent * scale|; all points relative to A's position).
//   outside: closest box point q by clamping the centre's box coordinates; gap = |q - cS| - r
//   inside : leave through the nearest face (lowest axis on ties); depth = face gap + r
// nsx = unit direction from the sphere towards the box, ps / px = witness points on the sphere / box.
struct SphereBox {
    bool contact;
    float dist, depth;
    V3 nsx, ps, px;
};
SphereBox sphereBox(V3 cS, float r, V3 cX, const Xf& tX, const AxrefShape& sX) {
    SphereBox o{};
    M3 m = quatToMat3(tX.q);
    const V3 ax[3] = {m.c0, m.c1, m.c2};
    const float half[3] = {std::fabs(sX.p0 * tX.s.x), std::fabs(sX.p1 * tX.s.y), std::fabs(sX.p2 * tX.s.z)};
    V3 d = cS - cX;
    float x[3];
    bool inside = true;
    V3 q = cX;
    for (int i = 0; i < 3; ++i) {
        x[i] = dot(d, ax[i]);
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
        V3 v = q - cS;
        float l = std::sqrt(dot(v, v));
        o.dist = l - r;
        o.depth = r - l;
        o.contact = o.depth >= 0.0f;
        o.nsx = (l > 0.0f) ? v * (1.0f / l) : mk(1.0f, 0.0f, 0.0f);
        o.ps = cS + o.nsx * r;
        o.px = q;
        return o;
    }
    int best = 0;
    float gap = half[0] - std::fabs(x[0]);
    for (int i = 1; i < 3; ++i) {
        float g = half[i] - std::fabs(x[i]);
        if (g < gap) {
            gap = g;
            best = i;
        }
    }
    V3 out = (x[best] >= 0.0f) ? ax[best] : -ax[best];   // from the box centre's side towards the sphere
    o.contact = true;
    o.depth = gap + r;
    o.dist = -o.depth;
    o.nsx = -out;
    o.ps = cS + o.nsx * r;
    o.px = cS + out * gap;
    return o;
}


// ------------------------------------------------------------------------------------------
// Box against box in closed form: the 15-axis separating-axis test (3 face normals of each box +
// 9 edge-edge cross products), as collision libraries dispatch this pair class.  The Minkowski
// difference of two boxes is a polytope whose face normals are among those 15 directions, so
//   * the boxes are apart  <=>  some axis has a negative overlap, and
//   * the penetration depth = the smallest overlap, the contact normal = that axis
// (the answer EPA converges to; tests/test_oracle_narrow.py checks one against the other and against
// the convex hull of the Minkowski difference).  Edge-edge axes are compared un-normalised
// (overlap^2 * |L'|^2 cross-multiplied) so only the winning axis costs a square root; a pair of
// nearly parallel edges (|a_i x b_j|^2 <= 1e-5) is skipped — its axis degenerates into the face
// axes.  Ties keep the earlier axis (faces of A, faces of B, then edges i-major).
// Witness points: face axis -> the other box's deepest vertex (support-point tie rule: a zero dot
// picks the + side) and its projection onto the face; edge-edge -> the closest points of the two
// supporting edges (as segments).  All points relative to A's position.
// ------------------------------------------------------------------------------------------
struct BoxFrame {
    V3 c;          // centre relative to A's position
    V3 ax[3];      // unit axes: the columns of Quat::toMatrix
    float h[3];    // half lengths |halfExtent * scale|
};
BoxFrame makeBoxFrame(const Xf& t, const AxrefShape& sh, V3 origin) {
    BoxFrame f;
    M3 m = quatToMat3(t.q);
    f.c = t.p - origin;
    f.ax[0] = m.c0; f.ax[1] = m.c1; f.ax[2] = m.c2;
    f.h[0] = std::fabs(sh.p0 * t.s.x);
    f.h[1] = std::fabs(sh.p1 * t.s.y);
    f.h[2] = std::fabs(sh.p2 * t.s.z);
    return f;
}

const float SAT_PARALLEL_EPS = 1e-5f;
struct BoxBox {
    bool contact;
    float depth;
    V3 n, pa, pb;   // unit normal from A to B, witness points on A and on B
};
BoxBox boxBox(const BoxFrame& A, const BoxFrame& B) {
    BoxBox o{};
    const V3 t = B.c - A.c;
    float R[3][3], AR[3][3], tA[3], tB[3];
    for (int i = 0; i < 3; ++i) {
        tA[i] = dot(t, A.ax[i]);
        tB[i] = dot(t, B.ax[i]);
        for (int j = 0; j < 3; ++j) {
            R[i][j] = dot(A.ax[i], B.ax[j]);
            AR[i][j] = std::fabs(R[i][j]);
        }
    }
    // best axis so far: overlap bestOv measured along an axis of squared length bestL2
    float bestOv = FLT_MAX, bestL2 = 1.0f;
    int axis = -1;
    for (int i = 0; i < 3; ++i) {   // faces of A
        const float rb = (B.h[0] * AR[i][0] + B.h[1] * AR[i][1]) + B.h[2] * AR[i][2];
        const float ov = (A.h[i] + rb) - std::fabs(tA[i]);
        if (ov < 0.0f) return o;
        if (ov < bestOv) {
            bestOv = ov;
            axis = i;
        }
    }
    for (int j = 0; j < 3; ++j) {   // faces of B
        const float ra = (A.h[0] * AR[0][j] + A.h[1] * AR[1][j]) + A.h[2] * AR[2][j];
        const float ov = (ra + B.h[j]) - std::fabs(tB[j]);
        if (ov < 0.0f) return o;
        if (ov < bestOv) {
            bestOv = ov;
            axis = 3 + j;
        }
    }
    if (axis < 0) return o;   // NaN input: every compare was false
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; ++j) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const V3 L = cross(A.ax[i], B.ax[j]);
            const float l2 = dot(L, L);
            if (!(l2 > SAT_PARALLEL_EPS)) continue;
            const float ra = A.h[i1] * AR[i2][j] + A.h[i2] * AR[i1][j];
            const float rb = B.h[j1] * AR[i][j2] + B.h[j2] * AR[i][j1];
            const float ov = (ra + rb) - std::fabs(dot(t, L));   // scaled by |L|
            if (ov < 0.0f) return o;
            // ov / |L| < bestOv / |bestL|, cross-multiplied (all terms >= 0)
            if ((ov * ov) * bestL2 < (bestOv * bestOv) * l2) {
                bestOv = ov;
                bestL2 = l2;
                axis = 6 + 3 * i + j;
            }
        }
    }
    o.contact = true;
    if (axis < 3) {
        const int i = axis;
        const bool pos = tA[i] >= 0.0f;
        o.n = pos ? A.ax[i] : -A.ax[i];
        o.depth = bestOv;
        V3 v = B.c;   // B's deepest vertex: support of B in direction -n
        for (int j = 0; j < 3; ++j) {
            const float dj = pos ? R[i][j] : -R[i][j];
            v = v + B.ax[j] * ((dj > 0.0f) ? -B.h[j] : B.h[j]);
        }
        o.pb = v;
        o.pa = v + o.n * o.depth;
    } else if (axis < 6) {
        const int j = axis - 3;
        const bool pos = tB[j] >= 0.0f;
        o.n = pos ? B.ax[j] : -B.ax[j];
        o.depth = bestOv;
        V3 v = A.c;   // A's deepest vertex: support of A in direction n
        for (int i = 0; i < 3; ++i) {
            const float di = pos ? R[i][j] : -R[i][j];
            v = v + A.ax[i] * ((di >= 0.0f) ? A.h[i] : -A.h[i]);
        }
        o.pa = v;
        o.pb = v - o.n * o.depth;
    } else {
        const int i = (axis - 6) / 3, j = (axis - 6) % 3;
        const V3 L = cross(A.ax[i], B.ax[j]);
        const float invl = 1.0f / std::sqrt(bestL2);
        const bool pos = dot(t, L) >= 0.0f;
        o.n = L * (pos ? invl : -invl);
        o.depth = bestOv * invl;
        // centres of the two supporting edges
        V3 ea = A.c, eb = B.c;
        for (int k = 0; k < 3; ++k) {
            if (k != i) ea = ea + A.ax[k] * ((dot(o.n, A.ax[k]) >= 0.0f) ? A.h[k] : -A.h[k]);
            if (k != j) eb = eb + B.ax[k] * ((dot(o.n, B.ax[k]) > 0.0f) ? -B.h[k] : B.h[k]);
        }
        // closest points of the segments ea + u*s (|s| <= hA_i) and eb + v*q (|q| <= hB_j)
        const V3 u = A.ax[i], v = B.ax[j];
        const V3 r = ea - eb;
        const float a = dot(u, u), e = dot(v, v), b = dot(u, v);
        const float c = dot(u, r), f = dot(v, r);
        const float denom = a * e - b * b;   // > 0: the edges are not parallel
        float s = (b * f - c * e) / denom;
        s = std::fmin(std::fmax(s, -A.h[i]), A.h[i]);
        float q = (b * s + f) / e;
        if (q < -B.h[j]) {
            q = -B.h[j];
            s = std::fmin(std::fmax((b * q - c) / a, -A.h[i]), A.h[i]);
        } else if (q > B.h[j]) {
            q = B.h[j];
            s = std::fmin(std::fmax((b * q - c) / a, -A.h[i]), A.h[i]);
        }
        o.pa = ea + u * s;
        o.pb = eb + v * q;
    }
    return o;
}

struct PairOut {
    bool contact;
    bool usedEpa;
    float dist;   // signed: core distance - radii (<=0 for contacts); lower bound if !exact
    AxrefContact c;
    uint32_t iters;
};

PairOut collidePair(uint32_t ia, uint32_t ib, const Xf& ta, const AxrefShape& sa, const Xf& tb,
                    const AxrefShape& sb, const float* hull, const AxrefNarrowCfg& cfg) {
    PairOut o{};
    o.c.a = ia;
    o.c.b = ib;
    V3 origin = ta.p;
    V3 n, pa, pb;   // pa/pb: surface witness points in the A-centred frame
    float depth;
    uint32_t status = 0;
    if (sa.type == SHAPE_SPHERE && sb.type == SHAPE_SPHERE) {
        V3 d = tb.p - origin;
        float dist = std::sqrt(dot(d, d));
        float rs = sa.p0 + sb.p0;
        depth = rs - dist;
        o.dist = dist - rs;
        if (!(depth >= 0.0f)) return o;
        n = (dist > 0.0f) ? d * (1.0f / dist) : mk(1.0f, 0.0f, 0.0f);
        pa = n * sa.p0;
        pb = d - n * sb.p0;
    } else if (sa.type == SHAPE_SPHERE && sb.type == SHAPE_BOX) {
        SphereBox r = sphereBox(mk(0, 0, 0), sa.p0, tb.p - origin, tb, sb);
        o.dist = r.dist;
        if (!r.contact) return o;
        depth = r.depth;
        n = r.nsx;
        pa = r.ps;
        pb = r.px;
    } else if (sa.type == SHAPE_BOX && sb.type == SHAPE_SPHERE) {
        SphereBox r = sphereBox(tb.p - origin, sb.p0, mk(0, 0, 0), ta, sa);
        o.dist = r.dist;
        if (!r.contact) return o;
        depth = r.depth;
        n = -r.nsx;
        pa = r.px;
        pb = r.ps;
    } else {
        // box-box: closed form (15-axis SAT) unless the generic path is forced.  The contact decision and
        // the contact record always come from the SAT; with wantDistances a separated pair additionally
        // runs GJK for its exact distance.
        const bool sat = sa.type == SHAPE_BOX && sb.type == SHAPE_BOX && !(cfg.flags & AXREF_BOXBOX_GJK_EPA);
        if (sat) {
            BoxBox r = boxBox(makeBoxFrame(ta, sa, origin), makeBoxFrame(tb, sb, origin));
            if (r.contact) {
                o.dist = -r.depth;
                o.contact = true;
                V3 mid = (r.pa + r.pb) * 0.5f + origin;
                o.c.px = mid.x; o.c.py = mid.y; o.c.pz = mid.z;
                o.c.nx = r.n.x; o.c.ny = r.n.y; o.c.nz = r.n.z;
                o.c.depth = r.depth;
                o.c.status = 0;
                return o;
            }
            if (!cfg.wantDistances) {
                o.dist = FLT_MAX;   // separated; no distance is computed in this mode
                return o;
            }
        }
        Core A = makeCore(ta, sa, hull, origin);
        Core B = makeCore(tb, sb, hull, origin);
        float rs = A.r + B.r;
        Simplex s;
        GjkResult g = gjk(A, B, cfg, rs, s);
        o.iters = g.iters;
        status = g.status;
        if (sat) {   // separated by the SAT: GJK only supplies the distance
            o.dist = (g.state == GJK_SEPARATED) ? std::sqrt(g.vv) : 0.0f;
            return o;
        }
        if (g.state == GJK_SEPARATED) {
            float dist = std::sqrt(g.vv);
            if (!g.exact) {
                o.dist = dist - rs;   // not converged: only known to be separated
                return o;
            }
            depth = rs - dist;
            o.dist = dist - rs;
            if (!(depth >= 0.0f)) return o;
            // shallow contact: cores separated by less than the radii
            n = -(g.v * (1.0f / dist));
            V3 ca = mk(0, 0, 0);
            for (int i = 0; i < s.n; ++i) ca = ca + s.a[i] * s.lam[i];
            pa = ca + n * A.r;
            pb = (ca - g.v) - n * B.r;
        } else {
            EpaResult e = epa(A, B, cfg, s);
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
    V3 mid = (pa + pb) * 0.5f + origin;
    o.c.px = mid.x; o.c.py = mid.y; o.c.pz = mid.z;
    o.c.nx = n.x; o.c.ny = n.y; o.c.nz = n.z;
    o.c.depth = depth;
    o.c.status = status;
    return o;
}


// ------------------------------------------------------------------------------------------
// Stage 4: contact manifolds.  One manifold per contact, 1..4 points that share
// the contact's normal; each point expands to a debug::DebugContactPoint (position, normal,
// penetrationDepth; include/axiom/debug/physics_debug_draw.hpp:128-132) and the sum of the point
// counts is gui::PhysicsWorldStats::contactPointCount (include/axiom/gui/physics_panel.hpp:21).
// Box-box: feature clipping — reference face = the face of either box most aligned with the
// contact normal (ties: body a), incident face = the other box's face most anti-parallel to it,
// clipped (Sutherland-Hodgman) against the reference face's four side planes; vertices on or below
// the reference face are kept (position = midpoint between the vertex and its projection onto the
// reference face, depth = distance below the face) and reduced to at most four.  Capsule-box: the capsule's
// segment clipped against the facing box face (capsuleBoxManifold, 1..2 points).  Capsule-capsule with
// (nearly) parallel axes: the two ends of the overlapping stretch (capsuleCapsuleManifold, 1..2 points).
// Every other class, and a clip that comes out empty, keeps the single narrowphase point.
// ------------------------------------------------------------------------------------------
inline int argmaxAbs3(const float d[3]) {   // lowest index on ties
    int k = 0;
    float best = std::fabs(d[0]);
    if (std::fabs(d[1]) > best) { best = std::fabs(d[1]); k = 1; }
    if (std::fabs(d[2]) > best) { k = 2; }
    return k;
}

// Capsule against box: the capsule's core segment is clipped against the side planes of the box face that
// faces the capsule (the face most aligned with the contact normal); the ends of the clipped segment that
// lie within the radius of that face are the contact points (position = midpoint between the capsule's
// surface point below the end and its projection on the face, depth = radius - height above the face).
// Two points for a capsule lying on a face, one for a tilted capsule; none kept -> the narrowphase point.
void capsuleBoxManifold(const AxrefContact& c, const Xf& ta, const AxrefShape& sa, const Xf& tb, const AxrefShape& sb,
                        AxrefManifold& m) {
    const V3 origin = ta.p;
    const V3 n = mk(c.nx, c.ny, c.nz);
    const bool boxIsA = sa.type == SHAPE_BOX;
    const BoxFrame R = boxIsA ? makeBoxFrame(ta, sa, origin) : makeBoxFrame(tb, sb, origin);
    const Xf& tC = boxIsA ? tb : ta;
    const AxrefShape& sC = boxIsA ? sb : sa;
    const M3 mm = quatToMat3(tC.q);
    const V3 e = mm.c1 * ((sC.p1 * 0.5f) * tC.s.y);   // half segment, as the narrowphase core
    const V3 cc = tC.p - origin;
    const float r = sC.p0;
    const V3 p0 = cc - e, seg = e * 2.0f;
    const V3 toCap = boxIsA ? n : -n;                  // from the box towards the capsule
    const float d[3] = {dot(toCap, R.ax[0]), dot(toCap, R.ax[1]), dot(toCap, R.ax[2])};
    const int i = argmaxAbs3(d);
    const V3 nr = R.ax[i] * ((d[i] >= 0.0f) ? 1.0f : -1.0f);
    float t0 = 0.0f, t1 = 1.0f;
    for (int side = 0; side < 4; ++side) {
        const int w = (i + 1 + (side >> 1)) % 3;
        const V3 pn = R.ax[w] * ((side & 1) ? -1.0f : 1.0f);
        const float d0 = dot(p0 - R.c, pn) - R.h[w];   // <= 0 inside
        const float dd = dot(seg, pn);
        if (dd > 0.0f) {
            const float t = -d0 / dd;
            if (t < t1) t1 = t;
        } else if (dd < 0.0f) {
            const float t = -d0 / dd;
            if (t > t0) t0 = t;
        } else if (d0 > 0.0f) {
            return;   // parallel to the plane and outside it
        }
    }
    if (!(t0 <= t1)) return;
    const float tt[2] = {t0, t1};
    const int cand = (t0 == t1) ? 1 : 2;
    int cnt = 0;
    for (int k = 0; k < cand; ++k) {
        const V3 q = p0 + seg * tt[k];
        const float sep = (dot(q - R.c, nr) - R.h[i]) - r;
        if (sep <= 0.0f) {
            const V3 w = (q - nr * (r + sep * 0.5f)) + origin;
            m.px[cnt] = w.x; m.py[cnt] = w.y; m.pz[cnt] = w.z;
            m.depth[cnt] = -sep;
            ++cnt;
        }
    }
    if (cnt == 0) {   // nothing within reach of the face (edge / corner contact): the narrowphase point stands
        m.px[0] = c.px; m.py[0] = c.py; m.pz[0] = c.pz;
        m.depth[0] = c.depth;
        return;
    }
    m.count = (uint32_t)cnt;
}

// Capsule against capsule with (nearly) parallel axes — |eA x eB|^2 <= 1e-2 |eA|^2 |eB|^2, about 5.7 degrees:
// the stretch of A's segment that B's segment overlaps (B's ends in A's axis coordinate, clamped to [-1, 1])
// gives two candidate points, one at each end of the stretch; a candidate is kept when the two surfaces overlap
// there along the contact normal (position = midpoint of the two surface points, depth = overlap along n).
// Crossed capsules, a degenerate stretch, or no candidate kept: the narrowphase point stands.
void capsuleCapsuleManifold(const AxrefContact& c, const Xf& ta, const AxrefShape& sa, const Xf& tb, const AxrefShape& sb,
                            AxrefManifold& m) {
    const V3 origin = ta.p;
    const V3 n = mk(c.nx, c.ny, c.nz);
    const V3 eA = quatToMat3(ta.q).c1 * ((sa.p1 * 0.5f) * ta.s.y);
    const V3 eB = quatToMat3(tb.q).c1 * ((sb.p1 * 0.5f) * tb.s.y);
    const V3 cB = tb.p - origin;   // A's centre is the origin
    const float rA = sa.p0, rB = sb.p0;
    const float aa = dot(eA, eA), bb = dot(eB, eB);
    if (!(aa > 0.0f && bb > 0.0f)) return;
    const V3 cr = cross(eA, eB);
    if (!(dot(cr, cr) <= 1e-2f * (aa * bb))) return;
    const float tc = dot(cB, eA) / aa, te = dot(eB, eA) / aa;
    float lo = tc - te, hi = tc + te;
    if (lo > hi) { const float t = lo; lo = hi; hi = t; }
    if (lo < -1.0f) lo = -1.0f;
    if (hi > 1.0f) hi = 1.0f;
    if (!(lo < hi)) return;
    const float tt[2] = {lo, hi};
    int cnt = 0;
    for (int k = 0; k < 2; ++k) {
        const V3 pA = eA * tt[k];
        const float sB = dot(pA - cB, eB) / bb;
        const V3 pB = cB + eB * sB;
        const float sep = dot(pB - pA, n) - (rA + rB);
        if (sep <= 0.0f) {
            const V3 w = ((pA + n * rA) + (pB - n * rB)) * 0.5f + origin;
            m.px[cnt] = w.x; m.py[cnt] = w.y; m.pz[cnt] = w.z;
            m.depth[cnt] = -sep;
            ++cnt;
        }
    }
    if (cnt == 0) {
        m.px[0] = c.px; m.py[0] = c.py; m.pz[0] = c.pz;
        m.depth[0] = c.depth;
        return;
    }
    m.count = (uint32_t)cnt;
}

void buildManifold(const AxrefContact& c, const Xf& ta, const AxrefShape& sa, const Xf& tb,
                   const AxrefShape& sb, AxrefManifold& m) {
    m.a = c.a; m.b = c.b;
    m.nx = c.nx; m.ny = c.ny; m.nz = c.nz;
    m.count = 1;
    for (int k = 0; k < 4; ++k) { m.px[k] = m.py[k] = m.pz[k] = 0.0f; m.depth[k] = 0.0f; }
    m.px[0] = c.px; m.py[0] = c.py; m.pz[0] = c.pz;
    m.depth[0] = c.depth;
    if ((sa.type == SHAPE_BOX && sb.type == SHAPE_CAPSULE) || (sa.type == SHAPE_CAPSULE && sb.type == SHAPE_BOX)) {
        capsuleBoxManifold(c, ta, sa, tb, sb, m);
        return;
    }
    if (sa.type == SHAPE_CAPSULE && sb.type == SHAPE_CAPSULE) {
        capsuleCapsuleManifold(c, ta, sa, tb, sb, m);
        return;
    }
    if (sa.type != SHAPE_BOX || sb.type != SHAPE_BOX) return;
    const V3 origin = ta.p;
    const V3 n = mk(c.nx, c.ny, c.nz);
    const BoxFrame A = makeBoxFrame(ta, sa, origin), B = makeBoxFrame(tb, sb, origin);
    const float da[3] = {dot(n, A.ax[0]), dot(n, A.ax[1]), dot(n, A.ax[2])};
    const float db[3] = {dot(n, B.ax[0]), dot(n, B.ax[1]), dot(n, B.ax[2])};
    const int ia = argmaxAbs3(da), ib = argmaxAbs3(db);
    const bool refIsA = std::fabs(da[ia]) >= std::fabs(db[ib]);
    const BoxFrame& R = refIsA ? A : B;
    const BoxFrame& I = refIsA ? B : A;
    const int i = refIsA ? ia : ib;
    // outward normal of the reference face, facing the other box (n points from a to b)
    const float sgn = refIsA ? ((da[ia] >= 0.0f) ? 1.0f : -1.0f) : ((db[ib] >= 0.0f) ? -1.0f : 1.0f);
    const V3 nr = R.ax[i] * sgn;
    const float di[3] = {dot(nr, I.ax[0]), dot(nr, I.ax[1]), dot(nr, I.ax[2])};
    const int j = argmaxAbs3(di);
    const float sI = (di[j] >= 0.0f) ? -1.0f : 1.0f;
    const int ju = (j + 1) % 3, jv = (j + 2) % 3;
    const V3 fc = I.c + I.ax[j] * (sI * I.h[j]);
    const V3 eu = I.ax[ju] * I.h[ju], ev = I.ax[jv] * I.h[jv];
    V3 poly[8], tmp[8];
    int np = 4;
    poly[0] = (fc + eu) + ev;
    poly[1] = (fc - eu) + ev;
    poly[2] = (fc - eu) - ev;
    poly[3] = (fc + eu) - ev;
    // clip against the reference face's side planes: s * dot(p - cR, ax_w) <= h_w
    bool over = false;
    for (int side = 0; side < 4 && np > 0; ++side) {
        const int w = (i + 1 + (side >> 1)) % 3;
        const float s = (side & 1) ? -1.0f : 1.0f;
        const V3 pn = R.ax[w] * s;
        const float hw = R.h[w];
        int nt = 0;
        V3 prev = poly[np - 1];
        float dprev = dot(prev - R.c, pn) - hw;
        for (int k = 0; k < np; ++k) {
            const V3 cur = poly[k];
            const float dcur = dot(cur - R.c, pn) - hw;
            const bool inPrev = dprev <= 0.0f, inCur = dcur <= 0.0f;
            // never past the 8 slots (sign noise on a degenerate reference face): overflow -> narrowphase point
            if (inPrev != inCur) {
                const float t = dprev / (dprev - dcur);
                if (nt < 8) tmp[nt++] = prev + (cur - prev) * t;
                else over = true;
            }
            if (inCur) {
                if (nt < 8) tmp[nt++] = cur;
                else over = true;
            }
            prev = cur;
            dprev = dcur;
        }
        np = nt;
        for (int k = 0; k < np; ++k) poly[k] = tmp[k];
    }
    if (over) np = 0;
    // keep the vertices on or below the reference face
    V3 pos[8];
    float dep[8];
    int nk = 0;
    for (int k = 0; k < np; ++k) {
        const float sep = dot(poly[k] - R.c, nr) - R.h[i];
        if (sep <= 0.0f) {
            pos[nk] = poly[k] - nr * (sep * 0.5f);
            dep[nk] = -sep;
            ++nk;
        }
    }
    if (nk == 0) return;   // degenerate clip: the narrowphase point stands
    bool keep[8];
    for (int k = 0; k < 8; ++k) keep[k] = k < nk;
    if (nk > 4) {
        // reduction: deepest, farthest from it, largest triangle, then the vertex farthest outside it
        for (int k = 0; k < nk; ++k) keep[k] = false;
        int p0 = 0;
        for (int k = 1; k < nk; ++k)
            if (dep[k] > dep[p0]) p0 = k;
        int p1 = -1;
        float best = -1.0f;
        for (int k = 0; k < nk; ++k) {
            if (k == p0) continue;
            const V3 d = pos[k] - pos[p0];
            const float dd = dot(d, d);
            if (dd > best) { best = dd; p1 = k; }
        }
        const V3 e = pos[p1] - pos[p0];
        float area[8];
        for (int k = 0; k < nk; ++k) area[k] = dot(cross(e, pos[k] - pos[p0]), nr);
        int p2 = -1;
        best = 0.0f;
        for (int k = 0; k < nk; ++k) {
            if (k == p0 || k == p1) continue;
            if (std::fabs(area[k]) > best) { best = std::fabs(area[k]); p2 = k; }
        }
        keep[p0] = keep[p1] = true;
        if (p2 >= 0) {
            keep[p2] = true;
            // fourth point: the vertex farthest outside the triangle (p0,p1,p2), measured as the
            // largest parallelogram area beyond one of its three edges
            const float flip = (area[p2] >= 0.0f) ? -1.0f : 1.0f;
            const V3 e12 = pos[p2] - pos[p1], e20 = pos[p0] - pos[p2];
            int p3 = -1;
            best = 0.0f;
            for (int k = 0; k < nk; ++k) {
                if (k == p0 || k == p1 || k == p2) continue;
                const float o01 = area[k] * flip;
                const float o12 = dot(cross(e12, pos[k] - pos[p1]), nr) * flip;
                const float o20 = dot(cross(e20, pos[k] - pos[p2]), nr) * flip;
                float v = o01;
                if (o12 > v) v = o12;
                if (o20 > v) v = o20;
                if (v > best) { best = v; p3 = k; }
            }
            if (p3 >= 0) keep[p3] = true;
        }
    }
    int cnt = 0;
    for (int k = 0; k < nk; ++k) {
        if (!keep[k]) continue;
        const V3 w = pos[k] + origin;
        m.px[cnt] = w.x; m.py[cnt] = w.y; m.pz[cnt] = w.z;
        m.depth[cnt] = dep[k];
        ++cnt;
    }
    m.count = (uint32_t)cnt;
}


// ------------------------------------------------------------------------------------------
// Scene queries ("scene queries", CLAUDE.md:88).  The oracle answers by brute
// force over all bodies; the CUDA path answers through the LBVH and must return the same sets / the
// same closest body.
//   AABB query: every body whose AABB meets the query box under AABB::intersects (closed intervals).
//   Ray cast  : a body is hit iff the ray's slab test against its AABB passes within [0, tMax] AND the
//               shape test reports t in [0, tMax]: sphere / oriented box / capsule in closed form, convex
//               hulls by conservative advancement on the GJK distance (first t with the gap <= 1e-4).  The reported t is max(shape t, AABB entry t);
//               the closest hit is the minimum (t, body index).  Origin inside a shape: t = 0, normal 0.
// ------------------------------------------------------------------------------------------
inline bool raySlab(V3 o, V3 d, const float* lo, const float* hi, float tMax, float* tNear) {
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    float tn = 0.0f, tf = tMax;
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
    *tNear = tn;
    return true;
}

struct ShapeHit {
    bool hit;
    float t;
    V3 n;
};
inline ShapeHit raySphere(V3 m /* origin - centre */, V3 d, float r, float tMax) {
    ShapeHit h{false, 0.0f, mk(0, 0, 0)};
    const float dd = dot(d, d), b = dot(m, d), cc = dot(m, m) - r * r;
    if (cc <= 0.0f) {
        h.hit = true;
        return h;
    }
    const float disc = b * b - dd * cc;
    if (!(disc >= 0.0f)) return h;
    const float t = (-b - std::sqrt(disc)) / dd;
    if (!(t >= 0.0f && t <= tMax)) return h;
    h.hit = true;
    h.t = t;
    h.n = (m + d * t) * (1.0f / r);
    return h;
}
ShapeHit rayShape(V3 o, V3 d, float tMax, const Xf& t, const AxrefShape& sh) {
    ShapeHit h{false, 0.0f, mk(0, 0, 0)};
    const V3 m = o - t.p;
    if (sh.type == SHAPE_SPHERE) return raySphere(m, d, sh.p0, tMax);
    if (sh.type == SHAPE_BOX) {
        M3 mm = quatToMat3(t.q);
        const V3 ax[3] = {mm.c0, mm.c1, mm.c2};
        const float half[3] = {std::fabs(sh.p0 * t.s.x), std::fabs(sh.p1 * t.s.y), std::fabs(sh.p2 * t.s.z)};
        float ol[3], dl[3];
        bool inside = true;
        for (int k = 0; k < 3; ++k) {
            ol[k] = dot(m, ax[k]);
            dl[k] = dot(d, ax[k]);
            if (!(std::fabs(ol[k]) <= half[k])) inside = false;
        }
        if (inside) {
            h.hit = true;
            return h;
        }
        float tn = -3.0e38f, tf = 3.0e38f;
        int axis = -1;
        for (int k = 0; k < 3; ++k) {
            if (dl[k] == 0.0f) {
                if (!(std::fabs(ol[k]) <= half[k])) return h;
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
        h.hit = true;
        h.t = tt;
        h.n = (dl[axis] > 0.0f) ? -ax[axis] : ax[axis];
        return h;
    }
    if (sh.type == SHAPE_CAPSULE) {
        // segment pa..pb = centre -+ column1 * (height/2 * scale.y), as the narrowphase core; radius unscaled
        M3 mm = quatToMat3(t.q);
        const V3 e = mm.c1 * ((sh.p1 * 0.5f) * t.s.y);
        const float r = sh.p0;
        const V3 oa = m + e;          // origin - pa   (pa = centre - e)
        const V3 ob = m - e;          // origin - pb
        const V3 ba = e * 2.0f;
        const float baba = dot(ba, ba), baoa = dot(ba, oa);
        // origin inside: distance to the segment <= r
        {
            float sgm = (baba > 0.0f) ? baoa / baba : 0.0f;
            sgm = (sgm < 0.0f) ? 0.0f : ((sgm > 1.0f) ? 1.0f : sgm);
            const V3 q = oa - ba * sgm;
            if (dot(q, q) <= r * r) {
                h.hit = true;
                return h;
            }
        }
        bool any = false;
        float best = 0.0f;
        V3 bn = mk(0, 0, 0);
        const float dd = dot(d, d), bard = dot(ba, d), rdoa = dot(d, oa), oaoa = dot(oa, oa);
        const float a = baba * dd - bard * bard;
        if (a > 0.0f) {
            const float b = baba * rdoa - baoa * bard;
            const float c = (baba * oaoa - baoa * baoa) - (r * r) * baba;
            const float disc = b * b - a * c;
            if (disc >= 0.0f) {
                const float tc = (-b - std::sqrt(disc)) / a;
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
    return h;   // hulls are answered at AABB level by the caller
}


// ------------------------------------------------------------------------------------------
// GJK-based continuous collision detection ("GJK-based CCD", CLAUDE.md:136): time of
// impact of a pair under LINEAR motion by conservative advancement.  B moves by D = dispB - dispA relative
// to A over t in [0,1]; at each step the exact GJK distance (cores minus radii) gives a gap and the
// closest direction, gap / (approach speed along it) is a lower bound of the time to contact for convex
// shapes, and t advances by it until the gap is <= tol (hit), the shapes move apart, or t > 1 (miss).
// ------------------------------------------------------------------------------------------
const float CCD_TOL = 1e-4f;
const int CCD_MAX_ITERS = 48;
// Conservative advancement of core B (moved by D * t, t in [0,1]) against the fixed core A.
void sweepCores(const Core& A, Core B, V3 D, const AxrefNarrowCfg& cfgIn, AxrefSweep& out) {
    AxrefNarrowCfg cfg = cfgIn;
    cfg.wantDistances = 1;
    const V3 c0 = B.c;
    const float rs = A.r + B.r;
    out = AxrefSweep{0u, 1.0f, 0.0f, 0.0f, 0.0f, 0u};
    float t = 0.0f;
    V3 nLast = mk(0.0f, 0.0f, 0.0f);
    int it = 0;
    for (; it < CCD_MAX_ITERS; ++it) {
        B.c = c0 + D * t;
        Simplex s;
        const GjkResult g = gjk(A, B, cfg, rs, s);
        if (g.state == GJK_OVERLAP) {   // cores touch at t: the normal is the last closest direction (zero at t = 0)
            out.hit = 1u;
            out.toi = t;
            out.nx = nLast.x; out.ny = nLast.y; out.nz = nLast.z;
            break;
        }
        const float dist = std::sqrt(g.vv);
        const float gap = dist - rs;
        const V3 n = -(g.v * (1.0f / dist));   // from a to b
        nLast = n;
        if (gap <= CCD_TOL) {
            out.hit = 1u;
            out.toi = t;
            out.nx = n.x; out.ny = n.y; out.nz = n.z;
            break;
        }
        const float approach = dot(D, g.v) / dist;   // speed at which B closes in along the closest direction
        if (!(approach > 0.0f)) break;               // moving apart or sliding past
        t = t + gap / approach;
        if (!(t <= 1.0f)) break;
    }
    if (it == CCD_MAX_ITERS) {
        // out of iterations while still closing in within the step: the conservative answer is a hit at
        // the time reached (a "no hit" here would be a tunnelling false negative)
        out.hit = 1u;
        out.toi = t;
        out.nx = nLast.x; out.ny = nLast.y; out.nz = nLast.z;
    }
    out.iterations = (uint32_t)it;
}

void sweepPair(const Xf& ta, const AxrefShape& sa, V3 dispA, const Xf& tb, const AxrefShape& sb, V3 dispB,
               const float* hull, const AxrefNarrowCfg& cfg, AxrefSweep& out) {
    const V3 origin = ta.p;
    sweepCores(makeCore(ta, sa, hull, origin), makeCore(tb, sb, hull, origin), dispB - dispA, cfg, out);
}

// ---- CCD with rotation ---------------------------------------------------------------------------------
// Each body moves by disp * t and turns by its rotation vector w (angular velocity * dt, world frame) under the
// first-order integration physics engines use:  q(t) = normalize(q + t * 0.5 * (w, 0) (x) q)  — algebraic, so the
// CUDA path and this restatement produce the same bits (no sin / cos).  The turning rate of that path never
// exceeds |w|, so a core point at distance <= rho from its body's position moves relative to the body at a speed
// <= |w| * rho, and
//     gap / (approach speed of the origins along the closest direction + |wA| rhoA + |wB| rhoB)
// is a lower bound of the time to contact: conservative advancement as in sweepCores, with both cores rebuilt
// from the poses at t every step.
const int CCD_ANG_MAX_ITERS = 64;

inline Xf poseAt(const Xf& t0, V3 d, Q4 dq, float t) {
    Xf o = t0;
    o.p = t0.p + d * t;
    const Q4 u{t0.q.x + dq.x * t, t0.q.y + dq.y * t, t0.q.z + dq.z * t, t0.q.w + dq.w * t};
    const float len2 = (u.x * u.x + u.y * u.y) + (u.z * u.z + u.w * u.w);
    const float inv = 1.0f / std::sqrt(len2);
    o.q = Q4{u.x * inv, u.y * inv, u.z * inv, u.w * inv};
    return o;
}
// 0.5 * (w, 0) (x) q
inline Q4 halfSpin(V3 w, Q4 q) {
    const V3 qv = mk(q.x, q.y, q.z);
    const V3 v = cross(w, qv) + w * q.w;
    return Q4{v.x * 0.5f, v.y * 0.5f, v.z * 0.5f, -dot(w, qv) * 0.5f};
}
// Largest distance of a core point from its body's position.
inline float coreReach(const Core& k) {
    if (k.kind == CORE_POINT) return 0.0f;
    if (k.kind == CORE_SEGMENT) return std::sqrt(dot(k.e0, k.e0));
    if (k.kind == CORE_BOX || k.kind == CORE_CYLINDER) return std::sqrt((dot(k.e0, k.e0) + dot(k.e1, k.e1)) + dot(k.e2, k.e2));
    float best = 0.0f;
    for (uint32_t i = 0; i < k.nv; ++i) {
        const V3 lv = mk(k.verts[3 * i] * k.s.x, k.verts[3 * i + 1] * k.s.y, k.verts[3 * i + 2] * k.s.z);
        const float d2 = dot(lv, lv);
        if (d2 > best) best = d2;
    }
    return std::sqrt(best);
}

void sweepPairAngular(const Xf& ta, const AxrefShape& sa, V3 dispA, V3 rotA, const Xf& tb, const AxrefShape& sb, V3 dispB,
                      V3 rotB, const float* hull, const AxrefNarrowCfg& cfgIn, AxrefSweep& out) {
    AxrefNarrowCfg cfg = cfgIn;
    cfg.wantDistances = 1;
    const V3 origin = ta.p;
    const V3 D = dispB - dispA;                      // A's position stays put, B carries the relative translation
    const Q4 dqA = halfSpin(rotA, ta.q), dqB = halfSpin(rotB, tb.q);
    const V3 zero = mk(0.0f, 0.0f, 0.0f);
    const float spin = std::sqrt(dot(rotA, rotA)) * coreReach(makeCore(ta, sa, hull, origin)) +
                       std::sqrt(dot(rotB, rotB)) * coreReach(makeCore(tb, sb, hull, origin));
    out = AxrefSweep{0u, 1.0f, 0.0f, 0.0f, 0.0f, 0u};
    float t = 0.0f;
    V3 nLast = zero;
    int it = 0;
    for (; it < CCD_ANG_MAX_ITERS; ++it) {
        const Core A = makeCore(poseAt(ta, zero, dqA, t), sa, hull, origin);
        const Core B = makeCore(poseAt(tb, D, dqB, t), sb, hull, origin);
        const float rs = A.r + B.r;
        Simplex s;
        const GjkResult g = gjk(A, B, cfg, rs, s);
        if (g.state == GJK_OVERLAP) {
            out.hit = 1u;
            out.toi = t;
            out.nx = nLast.x; out.ny = nLast.y; out.nz = nLast.z;
            break;
        }
        const float dist = std::sqrt(g.vv);
        const float gap = dist - rs;
        const V3 n = -(g.v * (1.0f / dist));   // from a to b
        nLast = n;
        if (gap <= CCD_TOL) {
            out.hit = 1u;
            out.toi = t;
            out.nx = n.x; out.ny = n.y; out.nz = n.z;
            break;
        }
        const float approach = dot(D, g.v) / dist + spin;
        if (!(approach > 0.0f)) break;               // moving apart faster than any turning can close the gap
        t = t + gap / approach;
        if (!(t <= 1.0f)) break;
    }
    if (it == CCD_ANG_MAX_ITERS) {                   // conservative: contact at the time reached
        out.hit = 1u;
        out.toi = t;
        out.nx = nLast.x; out.ny = nLast.y; out.nz = nLast.z;
    }
    out.iterations = (uint32_t)it;
}

// Ray against a convex hull: the ray origin is a point core at rest, the hull moves by -d * tMax; the
// time of impact of that sweep is t / tMax.  Surface normal = from the hull towards the ray origin.
ShapeHit rayHull(V3 o, V3 d, float tMax, const Xf& t, const AxrefShape& sh, const float* hull, const AxrefNarrowCfg& cfg) {
    ShapeHit h{false, 0.0f, mk(0, 0, 0)};
    Core P{};
    P.kind = CORE_POINT;
    P.c = mk(0.0f, 0.0f, 0.0f);
    P.r = 0.0f;
    AxrefSweep sw;
    sweepCores(P, makeCore(t, sh, hull, o), -(d * tMax), cfg, sw);
    if (!sw.hit) return h;
    h.hit = true;
    h.t = sw.toi * tMax;
    h.n = mk(-sw.nx, -sw.ny, -sw.nz);
    return h;
}

}   // namespace

// =============================================================================================
extern "C" {

void axref_quat_rotate(const float q[4], const float v[3], float out[3]) {
    V3 r = quatRotate(Q4{q[0], q[1], q[2], q[3]}, mk(v[0], v[1], v[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void axref_quat_mul(const float p[4], const float q[4], float out[4]) {
    Q4 r = quatMul(Q4{p[0], p[1], p[2], p[3]}, Q4{q[0], q[1], q[2], q[3]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void axref_quat_to_mat3(const float q[4], float o[9]) {
    M3 m = quatToMat3(Q4{q[0], q[1], q[2], q[3]});
    o[0] = m.c0.x; o[1] = m.c0.y; o[2] = m.c0.z;
    o[3] = m.c1.x; o[4] = m.c1.y; o[5] = m.c1.z;
    o[6] = m.c2.x; o[7] = m.c2.y; o[8] = m.c2.z;
}
void axref_transform_point(const float xf[10], const float p[3], float out[3]) {
    V3 r = transformPoint(loadXf(xf), mk(p[0], p[1], p[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void axref_rng_u32(uint64_t seed, uint32_t n, uint32_t* out) {
    Rng r(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.next();
}
void axref_rng_float(uint64_t seed, uint32_t n, float* out) {
    Rng r(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.nextFloat();
}
int axref_aabb_intersects(const float a[6], const float b[6]) { return intersects(a, b) ? 1 : 0; }

int32_t axref_refit_route(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                          uint32_t nHullVerts, float margin, float* outAabb, int nthreads, int mat4Route) {
    if (n && (!xf || !shapes || !outAabb)) return 202;
    std::vector<int> err((size_t)std::max(1, nthreads), 0);
    parallelFor(n, nthreads, [&](int t, uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            int e = refitOne(loadXf(xf + 10 * i), shapes[i], hullXYZ, nHullVerts, margin,
                             outAabb + 6 * i, mat4Route != 0);
            if (e) err[(size_t)t] = e;
        }
    });
    for (int e : err)
        if (e) return e;
    return 0;
}
int32_t axref_refit(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                    uint32_t nHullVerts, float margin, float* outAabb, int nthreads) {
    return axref_refit_route(xf, shapes, n, hullXYZ, nHullVerts, margin, outAabb, nthreads, 0);
}

// ---- the remaining math restatements, exported for tests/test_oracle_vs_reference.py ----------------------
float axref_vec3_dot(const float a[3], const float b[3]) { return dot(mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2])); }
void axref_vec3_cross(const float a[3], const float b[3], float o[3]) {
    V3 c = cross(mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]));
    o[0] = c.x; o[1] = c.y; o[2] = c.z;
}
void axref_quat_conjugate(const float q[4], float o[4]) {
    Q4 c = conjugate(Q4{q[0], q[1], q[2], q[3]});
    o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
}
void axref_transform_direction(const float xf[10], const float d[3], float o[3]) {
    V3 r = transformDirection(loadXf(xf), mk(d[0], d[1], d[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void axref_inverse_transform_point(const float xf[10], const float p[3], float o[3]) {
    V3 r = inverseTransformPoint(loadXf(xf), mk(p[0], p[1], p[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void axref_inverse_transform_direction(const float xf[10], const float d[3], float o[3]) {
    V3 r = inverseTransformDirection(loadXf(xf), mk(d[0], d[1], d[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
// AABB ops as the broadphase / refit / LBVH code uses them (aabb.hpp:47,62,68,143-173,213-215)
void axref_aabb_expand_point(const float a[6], const float p[3], float o[6]) {
    Box b{mk(a[0], a[1], a[2]), mk(a[3], a[4], a[5])};
    expandPoint(b, mk(p[0], p[1], p[2]));
    o[0] = b.lo.x; o[1] = b.lo.y; o[2] = b.lo.z; o[3] = b.hi.x; o[4] = b.hi.y; o[5] = b.hi.z;
}
void axref_aabb_merge(const float a[6], const float b[6], float o[6]) {   // aabb.hpp:166-173, :223-229
    for (int k = 0; k < 3; ++k) {
        o[k] = (b[k] < a[k]) ? b[k] : a[k];
        o[k + 3] = (b[k + 3] > a[k + 3]) ? b[k + 3] : a[k + 3];
    }
}
void axref_aabb_center(const float a[6], float o[3]) {   // aabb.hpp:62: (min + max) * 0.5f
    for (int k = 0; k < 3; ++k) o[k] = (a[k] + a[k + 3]) * 0.5f;
}
void axref_aabb_expand_margin(const float a[6], float margin, float o[6]) {   // aabb.hpp:156-160
    for (int k = 0; k < 3; ++k) {
        o[k] = a[k] - margin;
        o[k + 3] = a[k + 3] + margin;
    }
}
void axref_aabb_from_center_extents(const float c[3], const float h[3], float o[6]) {   // aabb.hpp:213-215
    for (int k = 0; k < 3; ++k) {
        o[k] = c[k] - h[k];
        o[k + 3] = c[k] + h[k];
    }
}

int32_t axref_broadphase_brute_f(const float* aabb, uint32_t n, const uint32_t* worldId, const uint32_t* filt,
                                 uint32_t* outPairs, uint64_t cap, uint64_t* outCount) {
    uint64_t cnt = 0;
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j) {
            if (worldId && worldId[i] != worldId[j]) continue;
            if (!intersects(aabb + 6ull * i, aabb + 6ull * j)) continue;
            if (!shouldCollide(filt, i, j)) continue;
            if (cnt < cap) {
                outPairs[2 * cnt] = i;
                outPairs[2 * cnt + 1] = j;
            }
            ++cnt;
        }
    *outCount = cnt;
    return cnt > cap ? 601 : 0;
}

// Uniform grid keyed on the AABB centre with cell edge >= the largest AABB extent: two
// intersecting boxes then have centres at most one cell apart on every axis, so the 27-cell
// neighbourhood is complete.  The overlap decision itself is always the exact predicate above.
int32_t axref_broadphase_brute(const float* aabb, uint32_t n, const uint32_t* worldId,
                               uint32_t* outPairs, uint64_t cap, uint64_t* outCount) {
    return axref_broadphase_brute_f(aabb, n, worldId, nullptr, outPairs, cap, outCount);
}

int32_t axref_broadphase_grid_f(const float* aabb, uint32_t n, const uint32_t* worldId, const uint32_t* filt,
                                uint32_t* outPairs, uint64_t cap, uint64_t* outCount, int nthreads);
int32_t axref_broadphase_grid(const float* aabb, uint32_t n, const uint32_t* worldId,
                              uint32_t* outPairs, uint64_t cap, uint64_t* outCount,
                              int nthreads) {
    return axref_broadphase_grid_f(aabb, n, worldId, nullptr, outPairs, cap, outCount, nthreads);
}

int32_t axref_broadphase_grid_f(const float* aabb, uint32_t n, const uint32_t* worldId, const uint32_t* filt,
                                uint32_t* outPairs, uint64_t cap, uint64_t* outCount, int nthreads) {
    *outCount = 0;
    if (n == 0) return 0;
    // bodies with a NaN/inf box never intersect anything (every compare is false / handled by
    // the exact predicate for inf): leave non-finite ones out of the grid and test them brutally.
    std::vector<uint32_t> finite, odd;
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    double maxExt = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        const float* b = aabb + 6ull * i;
        bool ok = true;
        for (int k = 0; k < 6; ++k) ok = ok && std::isfinite(b[k]);
        if (!ok || b[0] > b[3] || b[1] > b[4] || b[2] > b[5]) {
            odd.push_back(i);
            continue;
        }
        finite.push_back(i);
        for (int k = 0; k < 3; ++k) {
            double c = 0.5 * ((double)b[k] + (double)b[k + 3]);
            lo[k] = std::min(lo[k], c);
            hi[k] = std::max(hi[k], c);
            maxExt = std::max(maxExt, (double)b[k + 3] - (double)b[k]);
        }
    }
    std::vector<std::vector<uint64_t>> found((size_t)std::max(1, nthreads));
    if (!finite.empty()) {
        double cell = std::max(maxExt * 1.0001, 1e-9);
        // cap the grid at ~4 cells per body
        double span = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], cell});
        double maxCells = std::max(1.0, std::cbrt(4.0 * (double)finite.size()));
        if (span / cell > maxCells) cell = span / maxCells;
        int64_t dim[3];
        for (int k = 0; k < 3; ++k) dim[k] = (int64_t)std::floor((hi[k] - lo[k]) / cell) + 1;
        auto cellOf = [&](uint32_t i, int64_t c[3]) {
            const float* b = aabb + 6ull * i;
            for (int k = 0; k < 3; ++k) {
                double ctr = 0.5 * ((double)b[k] + (double)b[k + 3]);
                int64_t v = (int64_t)std::floor((ctr - lo[k]) / cell);
                c[k] = std::min<int64_t>(std::max<int64_t>(v, 0), dim[k] - 1);
            }
        };
        // batched worlds share one coordinate frame: give every world its own copy of the grid
        const uint64_t cellsPerWorld = (uint64_t)dim[0] * dim[1] * dim[2];
        uint64_t nWorlds = 1;
        if (worldId)
            for (uint32_t i : finite) nWorlds = std::max<uint64_t>(nWorlds, (uint64_t)worldId[i] + 1);
        const uint64_t ncell = cellsPerWorld * nWorlds;
        std::vector<uint32_t> start(ncell + 1, 0), items(finite.size());
        std::vector<uint64_t> cid(finite.size());
        for (size_t k = 0; k < finite.size(); ++k) {
            int64_t c[3];
            cellOf(finite[k], c);
            cid[k] = ((uint64_t)c[2] * dim[1] + c[1]) * dim[0] + c[0] +
                     (worldId ? (uint64_t)worldId[finite[k]] * cellsPerWorld : 0);
            start[cid[k] + 1]++;
        }
        for (uint64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
        {
            std::vector<uint32_t> fill(start.begin(), start.end() - 1);
            for (size_t k = 0; k < finite.size(); ++k) items[fill[cid[k]]++] = finite[k];
        }
        parallelFor(finite.size(), nthreads, [&](int t, uint64_t a, uint64_t b) {
            auto& out = found[(size_t)t];
            for (uint64_t k = a; k < b; ++k) {
                const uint32_t i = finite[k];
                int64_t c[3];
                cellOf(i, c);
                for (int64_t z = std::max<int64_t>(c[2] - 1, 0); z <= std::min(c[2] + 1, dim[2] - 1); ++z)
                for (int64_t y = std::max<int64_t>(c[1] - 1, 0); y <= std::min(c[1] + 1, dim[1] - 1); ++y)
                for (int64_t x = std::max<int64_t>(c[0] - 1, 0); x <= std::min(c[0] + 1, dim[0] - 1); ++x) {
                    uint64_t cc = ((uint64_t)z * dim[1] + y) * dim[0] + x +
                                  (worldId ? (uint64_t)worldId[i] * cellsPerWorld : 0);
                    for (uint32_t p = start[cc]; p < start[cc + 1]; ++p) {
                        const uint32_t j = items[p];
                        if (j <= i) continue;
                        if (worldId && worldId[i] != worldId[j]) continue;
                        if (intersects(aabb + 6ull * i, aabb + 6ull * j) && shouldCollide(filt, i, j))
                            out.push_back(((uint64_t)i << 32) | j);
                    }
                }
            }
        });
    }
    // non-finite boxes: exact predicate against everybody (normally finds nothing)
    for (uint32_t i : odd)
        for (uint32_t j = 0; j < n; ++j) {
            if (j == i) continue;
            if (worldId && worldId[i] != worldId[j]) continue;
            bool jOdd = std::binary_search(odd.begin(), odd.end(), j);
            if (jOdd && j < i) continue;   // odd-odd pairs once
            if (intersects(aabb + 6ull * i, aabb + 6ull * j) && shouldCollide(filt, i, j)) {
                uint32_t a = std::min(i, j), b = std::max(i, j);
                found[0].push_back(((uint64_t)a << 32) | b);
            }
        }
    std::vector<uint64_t> all;
    size_t total = 0;
    for (auto& f : found) total += f.size();
    all.reserve(total);
    for (auto& f : found) all.insert(all.end(), f.begin(), f.end());
    std::sort(all.begin(), all.end());
    *outCount = all.size();
    for (size_t k = 0; k < all.size() && k < cap; ++k) {
        outPairs[2 * k] = (uint32_t)(all[k] >> 32);
        outPairs[2 * k + 1] = (uint32_t)all[k];
    }
    return all.size() > cap ? 601 : 0;
}

int32_t axref_narrowphase(const float* xf, const AxrefShape* shapes, uint32_t n,
                          const float* hullXYZ, uint32_t nHullVerts, const uint32_t* pairs,
                          uint64_t npairs, const AxrefNarrowCfg* cfg, AxrefContact* outContacts,
                          uint64_t cap, uint64_t* outCount, float* outDist,
                          AxrefNarrowStats* stats, int nthreads) {
    (void)nHullVerts;
    if (!cfg || !outCount) return 202;
    const int T = std::max(1, nthreads);
    std::vector<std::vector<AxrefContact>> loc((size_t)T);
    std::vector<AxrefNarrowStats> st((size_t)T, AxrefNarrowStats{});
    std::vector<int> err((size_t)T, 0);
    parallelFor(npairs, nthreads, [&](int t, uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; ++k) {
            uint32_t a = pairs[2 * k], b = pairs[2 * k + 1];
            if (a >= n || b >= n) {
                err[(size_t)t] = 601;
                continue;
            }
            PairOut o = collidePair(a, b, loadXf(xf + 10ull * a), shapes[a],
                                    loadXf(xf + 10ull * b), shapes[b], hullXYZ, *cfg);
            if (outDist) outDist[k] = o.contact ? ((o.dist < 0.0f) ? o.dist : 0.0f) : o.dist;
            st[(size_t)t].gjkIterations += o.iters;
            if (o.contact) {
                loc[(size_t)t].push_back(o.c);
                st[(size_t)t].numContacts++;
                if (o.usedEpa) st[(size_t)t].numPenetrating++;
                if (o.c.status == 301) st[(size_t)t].gjkFailures++;
                if (o.c.status == 302) st[(size_t)t].epaFailures++;
            }
        }
    });
    uint64_t cnt = 0;
    AxrefNarrowStats tot{};
    for (int t = 0; t < T; ++t) {
        for (auto& c : loc[(size_t)t]) {
            if (cnt < cap) outContacts[cnt] = c;
            ++cnt;
        }
        tot.numContacts += st[(size_t)t].numContacts;
        tot.numPenetrating += st[(size_t)t].numPenetrating;
        tot.gjkFailures += st[(size_t)t].gjkFailures;
        tot.epaFailures += st[(size_t)t].epaFailures;
        tot.gjkIterations += st[(size_t)t].gjkIterations;
    }
    *outCount = cnt;
    if (stats) *stats = tot;
    for (int e : err)
        if (e) return e;
    return cnt > cap ? 601 : 0;
}

int32_t axref_collide_pair(const float xfA[10], const AxrefShape* sa, const float xfB[10],
                           const AxrefShape* sb, const float* hullXYZ, const AxrefNarrowCfg* cfg,
                           AxrefContact* out, float* outDist, uint32_t* outUsedEpa) {
    PairOut o = collidePair(0, 1, loadXf(xfA), *sa, loadXf(xfB), *sb, hullXYZ, *cfg);
    if (out) *out = o.c;
    if (outDist) *outDist = o.dist;
    if (outUsedEpa) *outUsedEpa = o.usedEpa ? 1u : 0u;
    return o.contact ? 1 : 0;
}

int32_t axref_manifolds(const float* xf, const AxrefShape* shapes, uint32_t n, const AxrefContact* contacts,
                        uint64_t ncontacts, AxrefManifold* out, uint64_t* outPointCount, int nthreads) {
    if (!out && ncontacts) return 202;
    const int T = std::max(1, nthreads);
    std::vector<uint64_t> pts((size_t)T, 0);
    std::vector<int> err((size_t)T, 0);
    parallelFor(ncontacts, nthreads, [&](int t, uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; ++k) {
            const AxrefContact& c = contacts[k];
            if (c.a >= n || c.b >= n) {
                err[(size_t)t] = 601;
                continue;
            }
            buildManifold(c, loadXf(xf + 10ull * c.a), shapes[c.a], loadXf(xf + 10ull * c.b), shapes[c.b], out[k]);
            pts[(size_t)t] += out[k].count;
        }
    });
    uint64_t tot = 0;
    for (uint64_t p : pts) tot += p;
    if (outPointCount) *outPointCount = tot;
    for (int e : err)
        if (e) return e;
    return 0;
}


int32_t axref_query_aabbs(const float* aabb, uint32_t n, const uint32_t* worldId, const float* qboxes,
                          const uint32_t* qworld, uint32_t nq, uint32_t* outHits, uint64_t cap, uint64_t* outCount) {
    if (!outCount) return 202;
    uint64_t cnt = 0;
    for (uint32_t q = 0; q < nq; ++q) {
        for (uint32_t i = 0; i < n; ++i) {
            if (worldId && qworld && worldId[i] != qworld[q]) continue;
            if (!intersects(qboxes + 6ull * q, aabb + 6ull * i)) continue;
            if (cnt < cap) {
                outHits[2 * cnt] = q;
                outHits[2 * cnt + 1] = i;
            }
            ++cnt;
        }
    }
    *outCount = cnt;
    return cnt > cap ? 601 : 0;
}

int32_t axref_raycast(const float* xf, const AxrefShape* shapes, const float* hullXYZ, const float* aabb, uint32_t n,
                      const uint32_t* worldId, const AxrefRay* rays, uint32_t nq, AxrefRayHit* out, int nthreads) {
    const AxrefNarrowCfg cfg{32u, 32u, 64u, 1e-6f, 1e-4f, 1u, 0u};   // the library defaults (axcd_default_config)
    parallelFor(nq, nthreads, [&](int, uint64_t lo, uint64_t hi) {
        for (uint64_t q = lo; q < hi; ++q) {
            const AxrefRay& r = rays[q];
            const V3 o = mk(r.ox, r.oy, r.oz), d = mk(r.dx, r.dy, r.dz);
            AxrefRayHit best{0xffffffffu, r.tMax, 0.0f, 0.0f, 0.0f, 0u};
            bool have = false;
            for (uint32_t i = 0; i < n; ++i) {
                if (worldId && worldId[i] != r.world) continue;
                float tNear;
                if (!raySlab(o, d, aabb + 6ull * i, aabb + 6ull * i + 3, r.tMax, &tNear)) continue;
                ShapeHit h;
                const uint32_t flags = 0u;
                if (shapes[i].type != SHAPE_CONVEX && shapes[i].type != SHAPE_CYLINDER) h = rayShape(o, d, r.tMax, loadXf(xf + 10ull * i), shapes[i]);
                else h = rayHull(o, d, r.tMax, loadXf(xf + 10ull * i), shapes[i], hullXYZ, cfg);
                if (!h.hit) continue;
                const float t = (h.t > tNear) ? h.t : tNear;
                if (!have || t < best.t || (t == best.t && i < best.body)) {
                    have = true;
                    best = AxrefRayHit{i, t, h.n.x, h.n.y, h.n.z, flags};
                }
            }
            out[q] = best;
        }
    });
    return 0;
}


int32_t axref_ccd_pairs(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                        const uint32_t* pairs, uint64_t npairs, const float* disp, const AxrefNarrowCfg* cfg,
                        AxrefSweep* out, int nthreads) {
    if (!cfg || (npairs && (!pairs || !disp || !out))) return 202;
    std::vector<int> err((size_t)std::max(1, nthreads), 0);
    parallelFor(npairs, nthreads, [&](int t, uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; ++k) {
            const uint32_t a = pairs[2 * k], b = pairs[2 * k + 1];
            if (a >= n || b >= n) {
                err[(size_t)t] = 601;
                continue;
            }
            sweepPair(loadXf(xf + 10ull * a), shapes[a], mk(disp[3ull * a], disp[3ull * a + 1], disp[3ull * a + 2]),
                      loadXf(xf + 10ull * b), shapes[b], mk(disp[3ull * b], disp[3ull * b + 1], disp[3ull * b + 2]),
                      hullXYZ, *cfg, out[k]);
        }
    });
    for (int e : err)
        if (e) return e;
    return 0;
}


int32_t axref_ccd_pairs_angular(const float* xf, const AxrefShape* shapes, uint32_t n, const float* hullXYZ,
                                const uint32_t* pairs, uint64_t npairs, const float* disp, const float* rot,
                                const AxrefNarrowCfg* cfg, AxrefSweep* out, int nthreads) {
    if (!cfg || (npairs && (!pairs || !disp || !rot || !out))) return 202;
    std::vector<int> err((size_t)std::max(1, nthreads), 0);
    parallelFor(npairs, nthreads, [&](int t, uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; ++k) {
            const uint32_t a = pairs[2 * k], b = pairs[2 * k + 1];
            if (a >= n || b >= n) {
                err[(size_t)t] = 601;
                continue;
            }
            sweepPairAngular(loadXf(xf + 10ull * a), shapes[a], mk(disp[3ull * a], disp[3ull * a + 1], disp[3ull * a + 2]),
                             mk(rot[3ull * a], rot[3ull * a + 1], rot[3ull * a + 2]), loadXf(xf + 10ull * b), shapes[b],
                             mk(disp[3ull * b], disp[3ull * b + 1], disp[3ull * b + 2]),
                             mk(rot[3ull * b], rot[3ull * b + 1], rot[3ull * b + 2]), hullXYZ, *cfg, out[k]);
        }
    });
    for (int e : err)
        if (e) return e;
    return 0;
}

// Pose of a body at parameter t of the CCD-with-rotation motion model (tests sample it to validate the sweep).
void axref_ccd_pose_at(const float* xf10, const float* disp3, const float* rot3, float t, float* out10) {
    const Xf t0 = loadXf(xf10);
    const Xf o = poseAt(t0, mk(disp3[0], disp3[1], disp3[2]), halfSpin(mk(rot3[0], rot3[1], rot3[2]), t0.q), t);
    out10[0] = o.p.x; out10[1] = o.p.y; out10[2] = o.p.z;
    out10[3] = o.q.x; out10[4] = o.q.y; out10[5] = o.q.z; out10[6] = o.q.w;
    out10[7] = o.s.x; out10[8] = o.s.y; out10[9] = o.s.z;
}

}   // extern "C"
