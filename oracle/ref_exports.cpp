/*
 * ref_exports.cpp — extern "C" entry points over the reference's own math (oracle/_ref/libaxref_ref.so), for
 * tests/test_oracle_vs_reference.py to bit-compare with the oracle's restatement.  TEST INFRASTRUCTURE ONLY.
 * Compiled with the reference's own x86 flags (-mavx2 -mfma, GCC's default contraction): the header-only
 * Vec3 / AABB / RNG code inlined here is therefore built exactly as the engine builds it.  See ref_glue.cpp.
 */
#include "axiom/math/aabb.hpp"
#include "axiom/math/mat4.hpp"
#include "axiom/math/quat.hpp"
#include "axiom/math/random.hpp"
#include "axiom/math/transform.hpp"
#include "axiom/math/vec3.hpp"

#include <cstdint>

using namespace axiom::math;

static Transform loadT(const float* f) { return Transform(Vec3(f[0], f[1], f[2]), Quat(f[3], f[4], f[5], f[6]), Vec3(f[7], f[8], f[9])); }
static void put(float* o, const Vec3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static void putBox(float* o, const AABB& b) { put(o, b.min); put(o + 3, b.max); }
static AABB loadBox(const float* f) { return AABB(Vec3(f[0], f[1], f[2]), Vec3(f[3], f[4], f[5])); }

extern "C" {
// Vec3 (include/axiom/math/vec3.hpp)
float ref_vec3_dot(const float* a, const float* b) { return Vec3(a[0], a[1], a[2]).dot(Vec3(b[0], b[1], b[2])); }
void ref_vec3_cross(const float* a, const float* b, float* o) { put(o, Vec3(a[0], a[1], a[2]).cross(Vec3(b[0], b[1], b[2]))); }
float ref_vec3_length(const float* a) { return Vec3(a[0], a[1], a[2]).length(); }
void ref_vec3_normalized(const float* a, float* o) { put(o, Vec3(a[0], a[1], a[2]).normalized()); }
// Quat (include/axiom/math/quat.hpp:96)
void ref_quat_conjugate(const float* q, float* o) {
    const Quat c = Quat(q[0], q[1], q[2], q[3]).conjugate();
    o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
}
// Transform (src/math/transform.cpp:86-131, :13-24)
void ref_transform_point(const float* xf, const float* p, float* o) { put(o, loadT(xf).transformPoint(Vec3(p[0], p[1], p[2]))); }
void ref_transform_direction(const float* xf, const float* d, float* o) { put(o, loadT(xf).transformDirection(Vec3(d[0], d[1], d[2]))); }
void ref_inverse_transform_point(const float* xf, const float* p, float* o) { put(o, loadT(xf).inverseTransformPoint(Vec3(p[0], p[1], p[2]))); }
void ref_inverse_transform_direction(const float* xf, const float* d, float* o) {
    put(o, loadT(xf).inverseTransformDirection(Vec3(d[0], d[1], d[2])));
}
void ref_transform_to_matrix(const float* xf, float* out16) {
    const Mat4 m = loadT(xf).toMatrix();
    for (int i = 0; i < 16; ++i) out16[i] = m.m[i];
}
// AABB (include/axiom/math/aabb.hpp, src/math/aabb.cpp:8-35)
int ref_aabb_intersects(const float* a, const float* b) { return loadBox(a).intersects(loadBox(b)) ? 1 : 0; }
void ref_aabb_expand_point(const float* a, const float* p, float* o) { AABB b = loadBox(a); b.expand(Vec3(p[0], p[1], p[2])); putBox(o, b); }
void ref_aabb_from_point(const float* p, float* o) { putBox(o, AABB(Vec3(p[0], p[1], p[2]))); }
void ref_aabb_expand_margin(const float* a, float margin, float* o) { AABB b = loadBox(a); b.expand(margin); putBox(o, b); }
void ref_aabb_merge(const float* a, const float* b, float* o) { putBox(o, AABB::merge(loadBox(a), loadBox(b))); }
void ref_aabb_center(const float* a, float* o) { put(o, loadBox(a).center()); }
void ref_aabb_extents(const float* a, float* o) { put(o, loadBox(a).extents()); }
float ref_aabb_surface_area(const float* a) { return loadBox(a).surfaceArea(); }
void ref_aabb_from_center_extents(const float* c, const float* h, float* o) {
    putBox(o, AABB::fromCenterExtents(Vec3(c[0], c[1], c[2]), Vec3(h[0], h[1], h[2])));
}
// the Mat4 refit route: AABB(-h, h).transform(T.toMatrix())   (src/math/aabb.cpp:8-35 over transform.cpp:13-24)
void ref_aabb_transform_box(const float* xf, const float* h, float* o) {
    putBox(o, AABB(Vec3(-h[0], -h[1], -h[2]), Vec3(h[0], h[1], h[2])).transform(loadT(xf).toMatrix()));
}
// the recommended refit route, composed of reference functions only: 8 corners in the order of
// src/debug/debug_draw.cpp:99-108 through Transform::transformPoint, AABB(Vec3) then expand()
void ref_refit_box(const float* xf, const float* h, float* o) {
    const Transform t = loadT(xf);
    const float hx = h[0], hy = h[1], hz = h[2];
    const Vec3 c[8] = {Vec3(-hx, -hy, -hz), Vec3(hx, -hy, -hz), Vec3(hx, -hy, hz), Vec3(-hx, -hy, hz),
                       Vec3(-hx, hy, -hz),  Vec3(hx, hy, -hz),  Vec3(hx, hy, hz),  Vec3(-hx, hy, hz)};
    AABB b(t.transformPoint(c[0]));
    for (int k = 1; k < 8; ++k) b.expand(t.transformPoint(c[k]));
    putBox(o, b);
}
// DeterministicRNG (include/axiom/math/random.hpp:19-68)
void ref_rng_u32(uint64_t seed, uint32_t n, uint32_t* out) {
    DeterministicRNG r(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.next();
}
void ref_rng_float(uint64_t seed, uint32_t n, float* out) {
    DeterministicRNG r(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.nextFloat();
}
}
