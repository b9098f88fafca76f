"""ctypes binding of the CPU oracle (oracle/libaxref.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

SHAPE_DT = np.dtype([("type", "<u4"), ("p0", "<f4"), ("p1", "<f4"), ("p2", "<f4")])
CONTACT_DT = np.dtype([("a", "<u4"), ("b", "<u4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"),
                       ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("depth", "<f4"),
                       ("status", "<u4")])
MANIFOLD_DT = np.dtype([("a", "<u4"), ("b", "<u4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                        ("count", "<u4"), ("px", "<f4", (4,)), ("py", "<f4", (4,)), ("pz", "<f4", (4,)),
                        ("depth", "<f4", (4,))])
assert MANIFOLD_DT.itemsize == 88
RAY_DT = np.dtype([("ox", "<f4"), ("oy", "<f4"), ("oz", "<f4"), ("dx", "<f4"), ("dy", "<f4"), ("dz", "<f4"),
                   ("tMax", "<f4"), ("world", "<u4")])
RAYHIT_DT = np.dtype([("body", "<u4"), ("t", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                      ("flags", "<u4")])
NO_HIT = 0xFFFFFFFF
SWEEP_DT = np.dtype([("hit", "<u4"), ("toi", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                     ("iterations", "<u4")])


class NarrowCfg(C.Structure):
    _fields_ = [("gjkMaxIters", C.c_uint32), ("epaMaxIters", C.c_uint32),
                ("epaMaxFaces", C.c_uint32), ("gjkTol", C.c_float), ("epaTol", C.c_float),
                ("wantDistances", C.c_uint32), ("flags", C.c_uint32)]


class NarrowStats(C.Structure):
    _fields_ = [("numContacts", C.c_uint64), ("numPenetrating", C.c_uint64),
                ("gjkFailures", C.c_uint64), ("epaFailures", C.c_uint64),
                ("gjkIterations", C.c_uint64)]


BOXBOX_GJK_EPA = 1   # AxrefNarrowCfg.flags: box-box pairs through GJK/EPA instead of the closed-form SAT


def default_cfg(want_distances=False, boxbox_generic=False):
    return NarrowCfg(32, 32, 64, 1e-6, 1e-4, 1 if want_distances else 0, BOXBOX_GJK_EPA if boxbox_generic else 0)


def lib():
    global _LIB
    if _LIB is None:
        path = os.environ.get("AXREF_LIB") or os.path.join(ORACLE_DIR, "libaxref.so")   # override: diagnostic builds
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libaxref.so"])
        _LIB = C.CDLL(path)
        _LIB.axref_refit.restype = C.c_int32
        _LIB.axref_broadphase_brute.restype = C.c_int32
        _LIB.axref_broadphase_grid.restype = C.c_int32
        _LIB.axref_broadphase_brute_f.restype = C.c_int32
        _LIB.axref_broadphase_grid_f.restype = C.c_int32
        _LIB.axref_narrowphase.restype = C.c_int32
        _LIB.axref_collide_pair.restype = C.c_int32
        _LIB.axref_manifolds.restype = C.c_int32
        _LIB.axref_query_aabbs.restype = C.c_int32
        _LIB.axref_raycast.restype = C.c_int32
        _LIB.axref_ccd_pairs.restype = C.c_int32
        _LIB.axref_aabb_intersects.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def quat_rotate(q, v):
    out = np.zeros(3, np.float32)
    lib().axref_quat_rotate(_p(f32(q)), _p(f32(v)), _p(out))
    return out


def quat_mul(p, q):
    out = np.zeros(4, np.float32)
    lib().axref_quat_mul(_p(f32(p)), _p(f32(q)), _p(out))
    return out


def quat_to_mat3(q):
    out = np.zeros(9, np.float32)
    lib().axref_quat_to_mat3(_p(f32(q)), _p(out))
    return out.reshape(3, 3).T  # columns were written consecutively -> return row-major matrix


def transform_point(xf, p):
    out = np.zeros(3, np.float32)
    lib().axref_transform_point(_p(f32(xf)), _p(f32(p)), _p(out))
    return out


def rng_u32(seed, n):
    out = np.zeros(n, np.uint32)
    lib().axref_rng_u32(C.c_uint64(seed), C.c_uint32(n), _p(out))
    return out


def rng_float(seed, n):
    out = np.zeros(n, np.float32)
    lib().axref_rng_float(C.c_uint64(seed), C.c_uint32(n), _p(out))
    return out


def aabb_intersects(a, b):
    return bool(lib().axref_aabb_intersects(_p(f32(a)), _p(f32(b))))


def refit(xf, shapes, hull=None, margin=0.0, nthreads=1, mat4_route=False):
    """mat4_route: boxes through AABB::transform(Transform::toMatrix()) (row a15) instead of 8 x transformPoint."""
    xf = f32(xf).reshape(-1, 10)
    n = xf.shape[0]
    shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
    hull = f32(hull if hull is not None else np.zeros((0, 3))).reshape(-1, 3)
    out = np.zeros((n, 6), np.float32)
    lib().axref_refit_route.restype = C.c_int32
    rc = lib().axref_refit_route(_p(xf), _p(shapes), C.c_uint32(n), _p(hull), C.c_uint32(len(hull)),
                                 C.c_float(margin), _p(out), C.c_int(nthreads), C.c_int(1 if mat4_route else 0))
    return rc, out


def broadphase(aabb, world_id=None, brute=False, nthreads=1, cap=None, filters=None):
    aabb = f32(aabb).reshape(-1, 6)
    n = aabb.shape[0]
    wid = np.ascontiguousarray(world_id, dtype=np.uint32) if world_id is not None else None
    flt = (np.ascontiguousarray(np.asarray(filters, dtype=np.int64).reshape(-1, 3).astype(np.uint32))
           if filters is not None else None)   # groupIndex is signed: wrap, do not range-check
    if cap is None:
        cap = max(1024, 16 * n)
    while True:
        out = np.zeros((cap, 2), np.uint32)
        cnt = C.c_uint64(0)
        if brute:
            rc = lib().axref_broadphase_brute_f(_p(aabb), C.c_uint32(n), _p(wid), _p(flt), _p(out),
                                                C.c_uint64(cap), C.byref(cnt))
        else:
            rc = lib().axref_broadphase_grid_f(_p(aabb), C.c_uint32(n), _p(wid), _p(flt), _p(out),
                                               C.c_uint64(cap), C.byref(cnt), C.c_int(nthreads))
        if rc == 601:
            cap = int(cnt.value)
            continue
        assert rc == 0, rc
        return out[:cnt.value].copy()


def narrowphase(xf, shapes, pairs, hull=None, cfg=None, want_distances=False, nthreads=1):
    xf = f32(xf).reshape(-1, 10)
    n = xf.shape[0]
    shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
    hull = f32(hull if hull is not None else np.zeros((0, 3))).reshape(-1, 3)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    cfg = cfg or default_cfg(want_distances)
    if want_distances:
        cfg.wantDistances = 1
    npairs = pairs.shape[0]
    out = np.zeros(max(npairs, 1), CONTACT_DT)
    dist = np.zeros(max(npairs, 1), np.float32) if want_distances else None
    cnt = C.c_uint64(0)
    st = NarrowStats()
    rc = lib().axref_narrowphase(_p(xf), _p(shapes), C.c_uint32(n), _p(hull),
                                 C.c_uint32(len(hull)), _p(pairs), C.c_uint64(npairs),
                                 C.byref(cfg), _p(out), C.c_uint64(len(out)), C.byref(cnt),
                                 _p(dist), C.byref(st), C.c_int(nthreads))
    assert rc == 0, rc
    return out[:cnt.value].copy(), (dist[:npairs] if dist is not None else None), st


def manifolds(xf, shapes, contacts, nthreads=1):
    """One manifold per contact (same order); returns (records, total contact-point count)."""
    xf = f32(xf).reshape(-1, 10)
    shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
    contacts = np.ascontiguousarray(contacts, dtype=CONTACT_DT)
    out = np.zeros(max(len(contacts), 1), MANIFOLD_DT)
    tot = C.c_uint64(0)
    rc = lib().axref_manifolds(_p(xf), _p(shapes), C.c_uint32(xf.shape[0]), _p(contacts),
                               C.c_uint64(len(contacts)), _p(out), C.byref(tot), C.c_int(nthreads))
    assert rc == 0, rc
    return out[:len(contacts)].copy(), int(tot.value)


def make_rays(origins, dirs, t_max, world=0):
    origins, dirs = f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
    r = np.zeros(len(origins), RAY_DT)
    r["ox"], r["oy"], r["oz"] = origins[:, 0], origins[:, 1], origins[:, 2]
    r["dx"], r["dy"], r["dz"] = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    r["tMax"] = t_max
    r["world"] = world
    return r


def query_aabbs(aabb, qboxes, world_id=None, qworld=None):
    aabb = f32(aabb).reshape(-1, 6)
    qboxes = f32(qboxes).reshape(-1, 6)
    wid = np.ascontiguousarray(world_id, dtype=np.uint32) if world_id is not None else None
    qw = np.ascontiguousarray(qworld, dtype=np.uint32) if qworld is not None else None
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 2), np.uint32)
        cnt = C.c_uint64(0)
        rc = lib().axref_query_aabbs(_p(aabb), C.c_uint32(len(aabb)), _p(wid), _p(qboxes), _p(qw),
                                     C.c_uint32(len(qboxes)), _p(out), C.c_uint64(cap), C.byref(cnt))
        if rc == 601:
            cap = int(cnt.value)
            continue
        assert rc == 0, rc
        return out[:cnt.value].copy()


def raycast(xf, shapes, aabb, rays, world_id=None, nthreads=8, hull=None):
    xf = f32(xf).reshape(-1, 10)
    shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
    hull = f32(hull if hull is not None else np.zeros((0, 3))).reshape(-1, 3)
    aabb = f32(aabb).reshape(-1, 6)
    rays = np.ascontiguousarray(rays, dtype=RAY_DT)
    wid = np.ascontiguousarray(world_id, dtype=np.uint32) if world_id is not None else None
    out = np.zeros(max(1, len(rays)), RAYHIT_DT)
    rc = lib().axref_raycast(_p(xf), _p(shapes), _p(hull), _p(aabb), C.c_uint32(len(xf)), _p(wid), _p(rays),
                             C.c_uint32(len(rays)), _p(out), C.c_int(nthreads))
    assert rc == 0, rc
    return out[:len(rays)].copy()


def ccd_pairs(xf, shapes, pairs, disp, hull=None, cfg=None, nthreads=8, rot=None):
    xf = f32(xf).reshape(-1, 10)
    shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
    hull = f32(hull if hull is not None else np.zeros((0, 3))).reshape(-1, 3)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    disp = f32(disp).reshape(-1, 3)
    assert len(disp) == len(xf)
    cfg = cfg or default_cfg(True)
    out = np.zeros(max(1, len(pairs)), SWEEP_DT)
    if rot is None:
        rc = lib().axref_ccd_pairs(_p(xf), _p(shapes), C.c_uint32(len(xf)), _p(hull), _p(pairs), C.c_uint64(len(pairs)),
                                   _p(disp), C.byref(cfg), _p(out), C.c_int(nthreads))
    else:
        rot = f32(rot).reshape(-1, 3)
        assert len(rot) == len(xf)
        rc = lib().axref_ccd_pairs_angular(_p(xf), _p(shapes), C.c_uint32(len(xf)), _p(hull), _p(pairs),
                                           C.c_uint64(len(pairs)), _p(disp), _p(rot), C.byref(cfg), _p(out), C.c_int(nthreads))
    assert rc == 0, rc
    return out[:len(pairs)].copy()


def ccd_pose_at(xf10, disp3, rot3, t):
    """Pose of a body at parameter t of the CCD-with-rotation motion model."""
    out = np.zeros(10, np.float32)
    lib().axref_ccd_pose_at(_p(f32(xf10)), _p(f32(disp3)), _p(f32(rot3)), C.c_float(t), _p(out))
    return out


def collide_pair(xfa, sa, xfb, sb, hull=None, cfg=None, want_distances=True):
    """Returns (is_contact, contact_record, signed_distance, used_epa)."""
    sa = np.array([sa], dtype=SHAPE_DT)
    sb = np.array([sb], dtype=SHAPE_DT)
    hull = f32(hull if hull is not None else np.zeros((0, 3))).reshape(-1, 3)
    cfg = cfg or default_cfg(want_distances)
    out = np.zeros(1, CONTACT_DT)
    dist = C.c_float(0)
    used = C.c_uint32(0)
    rc = lib().axref_collide_pair(_p(f32(xfa)), _p(sa), _p(f32(xfb)), _p(sb), _p(hull),
                                  C.byref(cfg), _p(out), C.byref(dist), C.byref(used))
    return bool(rc), out[0], float(dist.value), bool(used.value)


def xf(pos=(0, 0, 0), quat=(0, 0, 0, 1), scale=(1, 1, 1)):
    return np.array(list(pos) + list(quat) + list(scale), dtype=np.float32)


def axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    s = np.float32(np.sin(np.float32(angle) * np.float32(0.5)))
    c = np.float32(np.cos(np.float32(angle) * np.float32(0.5)))
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, c], dtype=np.float32)


def sphere(r):
    return (0, r, 0.0, 0.0)


def box(hx, hy, hz):
    return (1, hx, hy, hz)


def capsule(r, height):
    return (2, r, height, 0.0)


def cylinder(r, height):
    return (6, r, height, 0.0)


def hull_shape(first, count):
    return (4, np.array([first], np.uint32).view(np.float32)[0],
            np.array([count], np.uint32).view(np.float32)[0], 0.0)
