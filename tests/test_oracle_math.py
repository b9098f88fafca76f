"""Pins the oracle's GLM-free math restatement against the reference's own test expectations.

Each case cites the reference test it mirrors (paths under /root/reference).  Tolerances are the
reference's own (1e-5 / 1e-4); nothing here is bit-exact against GLM, which is not in the snapshot.
"""
import numpy as np
import pytest

import oracle_lib as O

HALF_PI = np.float32(np.pi / 2)


def test_rng_goldens():
    # SURVEY.md Appendix C probe of include/axiom/math/random.hpp:19-68 (compiled from the header)
    assert O.rng_u32(0, 5).tolist() == [1634106851, 2870156730, 3134329460, 4102939785, 1447645584]
    assert O.rng_u32(42, 5).tolist() == [1870769882, 2612922264, 273981832, 3501727647, 2781298041]
    assert O.rng_u32(12345, 5).tolist() == [3491098128, 1545297057, 3932676696, 23160467, 3725064923]
    np.testing.assert_array_equal(O.rng_float(0, 3), np.float32([0.380470157, 0.668260455, 0.729767919]))
    np.testing.assert_array_equal(O.rng_float(42, 3), np.float32([0.435572565, 0.608368397, 0.0637913644]))
    np.testing.assert_array_equal(O.rng_float(12345, 3), np.float32([0.81283462, 0.359792501, 0.915647626]))


def test_rng_same_seed_same_sequence_and_range():
    # tests/math/random_test.cpp:15-111 (self-consistency + range)
    a, b = O.rng_float(777, 1000), O.rng_float(777, 1000)
    np.testing.assert_array_equal(a, b)
    assert a.min() >= 0.0 and a.max() <= 1.0


def test_quat_rotate_axis_cases():
    # tests/math/quat_test.cpp:71-99, 213-224
    np.testing.assert_allclose(O.quat_rotate(O.axis_angle((0, 0, 1), HALF_PI), (1, 0, 0)), (0, 1, 0), atol=1e-5)
    np.testing.assert_allclose(O.quat_rotate(O.axis_angle((1, 0, 0), HALF_PI), (0, 1, 0)), (0, 0, 1), atol=1e-5)
    np.testing.assert_allclose(O.quat_rotate(O.axis_angle((0, 0, 1), np.pi), (1, 0, 0)), (-1, 0, 0), atol=1e-5)


def test_quat_composition_order():
    # tests/math/quat_test.cpp:196-211: (qx*qz) applies qz first
    qx, qz = O.axis_angle((1, 0, 0), HALF_PI), O.axis_angle((0, 0, 1), HALF_PI)
    np.testing.assert_allclose(O.quat_rotate(O.quat_mul(qx, qz), (1, 0, 0)), (0, 0, 1), atol=1e-5)


def test_quat_vs_matrix():
    # tests/math/mat4_quat_integration_test.cpp:41-55
    q = O.axis_angle((1, 1, 1), 0.7)
    v = np.float32([1, 2, 3])
    m = O.quat_to_mat3(q)
    np.testing.assert_allclose(O.quat_rotate(q, v), m @ v, atol=1e-5)


def test_transform_point_trs_order():
    # tests/math/transform_test.cpp:137-154 : (1,0,0) -> scale 2 -> rotZ90 -> +(1,2,3) = (1,4,3)
    t = O.xf((1, 2, 3), O.axis_angle((0, 0, 1), HALF_PI), (2, 2, 2))
    np.testing.assert_allclose(O.transform_point(t, (1, 0, 0)), (1, 4, 3), atol=1e-5)
    # transform_test.cpp:257-288 identity / translate / scale
    np.testing.assert_array_equal(O.transform_point(O.xf(), (1, 2, 3)), (1, 2, 3))
    np.testing.assert_array_equal(O.transform_point(O.xf((10, 20, 30)), (1, 2, 3)), (11, 22, 33))
    np.testing.assert_array_equal(O.transform_point(O.xf(scale=(2, 3, 4)), (1, 1, 1)), (2, 3, 4))


def test_transform_hierarchy():
    # tests/math/transform_test.cpp:412-425: parent (10,0,0)/Z90/scale 2, child point (5,0,0) -> (10,10,0)
    t = O.xf((10, 0, 0), O.axis_angle((0, 0, 1), HALF_PI), (2, 2, 2))
    np.testing.assert_allclose(O.transform_point(t, (5, 0, 0)), (10, 10, 0), atol=1e-5)


def test_aabb_intersects_closed_intervals():
    # tests/math/aabb_test.cpp:179-208 (touching counts), SURVEY Appendix C probe
    assert O.aabb_intersects((0, 0, 0, 1, 1, 1), (1, 0, 0, 2, 1, 1))
    assert O.aabb_intersects((0, 0, 0, 2, 2, 2), (1, 1, 1, 3, 3, 3))
    assert not O.aabb_intersects((0, 0, 0, 1, 1, 1), (2, 2, 2, 3, 3, 3))
    assert not O.aabb_intersects((0, 0, 0, 1, 1, 1), (1.0000001, 0, 0, 2, 1, 1))
    nan = float("nan")
    assert not O.aabb_intersects((nan, 0, 0, 1, 1, 1), (0, 0, 0, 1, 1, 1))


def test_refit_body_fixtures():
    # tests/debug/physics_debug_draw_test.cpp:326-333, 350-354, 369-378 (hand-written world AABBs)
    rc, bb = O.refit([O.xf((0, 5, 0)), O.xf((0, 0, 0)), O.xf((5, 0, 0))],
                     [O.box(1, 1, 1), O.sphere(1), O.sphere(1)])
    assert rc == 0
    np.testing.assert_array_equal(bb[0], (-1, 4, -1, 1, 6, 1))
    np.testing.assert_array_equal(bb[1], (-1, -1, -1, 1, 1, 1))
    np.testing.assert_array_equal(bb[2], (4, -1, -1, 6, 1, 1))


def test_refit_rotated_cube_45deg():
    # tests/math/aabb_test.cpp:330-349: cube +-1 rotated 45 deg about Z -> +-sqrt(2), z +-1 (1e-4)
    rc, bb = O.refit([O.xf(quat=O.axis_angle((0, 0, 1), np.pi / 4))], [O.box(1, 1, 1)])
    r2 = np.sqrt(2.0)
    np.testing.assert_allclose(bb[0], (-r2, -r2, -1, r2, r2, 1), atol=1e-4)


def test_refit_sphere_ignores_rotation_and_scale():
    # src/debug/physics_debug_draw.cpp:246-248
    rc, bb = O.refit([O.xf((1, 2, 3), O.axis_angle((1, 2, 3), 1.1), (5, 6, 7))], [O.sphere(0.5)])
    np.testing.assert_array_equal(bb[0], (0.5, 1.5, 2.5, 1.5, 2.5, 3.5))


def test_refit_hull_matches_numpy_transform_point():
    # src/debug/debug_draw.cpp:431-448: hull vertices placed by transformPoint
    rng = np.random.default_rng(5)
    hull = rng.uniform(-1, 1, (16, 3)).astype(np.float32)
    t = O.xf((3, -2, 1), O.axis_angle((0.3, -0.5, 0.8), 0.9), (1.5, 0.5, 2.0))
    rc, bb = O.refit([t], [O.hull_shape(0, 16)], hull)
    pts = np.array([O.transform_point(t, v) for v in hull])
    np.testing.assert_array_equal(bb[0, :3], pts.min(0))
    np.testing.assert_array_equal(bb[0, 3:], pts.max(0))


def test_refit_margin_and_invalid_shape():
    rc, bb = O.refit([O.xf()], [O.sphere(1)], margin=0.25)
    np.testing.assert_array_equal(bb[0], (-1.25, -1.25, -1.25, 1.25, 1.25, 1.25))
    rc, _ = O.refit([O.xf()], [(3, 1.0, 1.0, 0.0)])  # Plane -> InvalidShape
    assert rc == 300
    rc, _ = O.refit([O.xf()], [O.hull_shape(0, 4)], np.zeros((2, 3), np.float32))
    assert rc == 300


def test_zero_volume_aabb_is_valid():
    # tests/math/aabb_test.cpp:385-393
    rc, bb = O.refit([O.xf((1, 1, 1))], [O.sphere(0.0)])
    np.testing.assert_array_equal(bb[0], (1, 1, 1, 1, 1, 1))
    assert O.aabb_intersects(bb[0], bb[0])
