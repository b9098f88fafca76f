/*
 * ref_glue.cpp — builds oracle/_ref/libaxref_ref.so: the reference's OWN math sources compiled where they lie
 * (/root/reference/src/math/transform.cpp, aabb.cpp and the header-only vec3 / aabb / random / quat), plus this
 * file.  TEST INFRASTRUCTURE ONLY; never part of the product.  Nothing from /root/reference is copied here.
 *
 * This file holds the GLM-backed primitives those sources need at link time.  GLM (vcpkg "glm" >= 1.0.0, vcpkg.json:9-12)
 *     is not in the snapshot, so src/math/quat.cpp and mat4.cpp cannot be compiled; the handful of functions
 *     they would provide are restated here from GLM 1.0.x's published scalar formulas,
 *     with the same explicit fused pattern the oracle and the kernels use.  Everything ELSE — transformPoint /
 *     transformDirection / inverseTransform*, Transform::toMatrix, AABB::transform, every AABB / Vec3 / RNG
 *     operation — is the reference's own compiled code.
 *  The extern "C" entry points over that code are in ref_exports.cpp (built with the reference's flags).
 */
#include "axiom/math/aabb.hpp"
#include "axiom/math/mat4.hpp"
#include "axiom/math/quat.hpp"
#include "axiom/math/random.hpp"
#include "axiom/math/transform.hpp"
#include "axiom/math/vec3.hpp"

#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace axiom::math {

// ---- GLM-backed primitives (restated; the reference defines them in quat.cpp / mat4.cpp through GLM) ----------
Mat4::Mat4() noexcept {   // identity (src/math/mat4.cpp: default constructor)
    for (int i = 0; i < 16; ++i) m[i] = 0.0f;
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

// glm::quat * glm::vec3 (src/math/quat.cpp:33-38): uv = cross(u, v); uuv = cross(u, uv); v + ((uv * w) + uuv) * 2
Vec3 Quat::operator*(const Vec3& v) const noexcept {
    auto crs = [](float ax, float ay, float az, float bx, float by, float bz, float& ox, float& oy, float& oz) {
        ox = std::fmaf(-az, by, ay * bz);   // the pattern GCC generates for Vec3::cross under the reference's flags
        oy = std::fmaf(-ax, bz, az * bx);
        oz = std::fmaf(ax, by, -(ay * bx));
    };
    float uvx, uvy, uvz, uux, uuy, uuz;
    crs(x, y, z, v.x, v.y, v.z, uvx, uvy, uvz);
    crs(x, y, z, uvx, uvy, uvz, uux, uuy, uuz);
    const float tx = std::fmaf(uvx, w, uux), ty = std::fmaf(uvy, w, uuy), tz = std::fmaf(uvz, w, uuz);
    return Vec3(std::fmaf(tx, 2.0f, v.x), std::fmaf(ty, 2.0f, v.y), std::fmaf(tz, 2.0f, v.z));
}

// glm::quat * glm::quat (src/math/quat.cpp:27-31)
Quat Quat::operator*(const Quat& q) const noexcept {
    const Quat& p = *this;
    return Quat(p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y, p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z,
                p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x, p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z);
}

// glm::mat4_cast (src/math/quat.cpp:117-129)
Mat4 Quat::toMatrix() const noexcept {
    const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
    Mat4 r;
    r.at(0, 0) = 1.0f - 2.0f * (qyy + qzz); r.at(1, 0) = 2.0f * (qxy + qwz); r.at(2, 0) = 2.0f * (qxz - qwy);
    r.at(0, 1) = 2.0f * (qxy - qwz); r.at(1, 1) = 1.0f - 2.0f * (qxx + qzz); r.at(2, 1) = 2.0f * (qyz + qwx);
    r.at(0, 2) = 2.0f * (qxz + qwy); r.at(1, 2) = 2.0f * (qyz - qwx); r.at(2, 2) = 1.0f - 2.0f * (qxx + qyy);
    return r;
}

Quat Quat::fromMatrix(const Mat4&) noexcept { std::abort(); }   // glm::quat_cast: not on the collision path, never called here

// glm::translate(mat4(1), t) / glm::scale(mat4(1), s): identity with the last column / the diagonal replaced
Mat4 Mat4::translation(const Vec3& t) noexcept {
    Mat4 r;
    r.at(0, 3) = t.x; r.at(1, 3) = t.y; r.at(2, 3) = t.z;
    return r;
}
Mat4 Mat4::scaling(const Vec3& s) noexcept {
    Mat4 r;
    r.at(0, 0) = s.x; r.at(1, 1) = s.y; r.at(2, 2) = s.z;
    return r;
}
// glm mat4 * mat4: Result[j] = A[0]*B[j][0] + A[1]*B[j][1] + A[2]*B[j][2] + A[3]*B[j][3], summed left to right
Mat4 Mat4::operator*(const Mat4& o) const noexcept {
    Mat4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r.at(i, j) = ((at(i, 0) * o.at(0, j) + at(i, 1) * o.at(1, j)) + at(i, 2) * o.at(2, j)) + at(i, 3) * o.at(3, j);
    return r;
}
// glm mat4 * vec4(v, 1): (m[0]*x + m[1]*y) + (m[2]*z + m[3]*1), each product-sum as one fused operation
Vec3 Mat4::transformPoint(const Vec3& v) const noexcept {
    float out[3];
    for (int i = 0; i < 3; ++i) out[i] = std::fmaf(at(i, 1), v.y, at(i, 0) * v.x) + std::fmaf(at(i, 2), v.z, at(i, 3));
    return Vec3(out[0], out[1], out[2]);
}

}  // namespace axiom::math

