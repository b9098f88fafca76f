"""axcd — thin ctypes binding of the C ABI in include/axcd.h (the product's public API from Python).

This module is plumbing only: it loads ``libaxcd.so`` (CUDA kernels + C ABI, built in-tree by
``make -C axiom-physics-engine_b200`` / ``__graft_entry__.build()``) and ``libaxcd_scene.so`` (the
host-only scene generator) and mirrors the reference-facing call shape
``Broadphase::update / getPairCount`` and ``Narrowphase::detectCollisions / getContactCount``
(reference: CLAUDE.md:162-178).  There is no CPU fallback: if the CUDA library is missing or no
device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# AXCD_LIB lets a tuning run point at an alternative build of the same ABI
LIB_PATH = os.environ.get("AXCD_LIB") or os.path.join(PKG_DIR, "libaxcd.so")
SCENE_LIB_PATH = os.path.join(PKG_DIR, "libaxcd_scene.so")

SHAPE_SPHERE, SHAPE_BOX, SHAPE_CAPSULE, SHAPE_PLANE, SHAPE_CONVEX, SHAPE_MESH, SHAPE_CYLINDER = range(7)
FLAG_PAIR_DISTANCES = 1
FLAG_TEMPORAL_COHERENCE = 4
FLAG_BOXBOX_GJK_EPA = 8
FLAG_NO_GRAPH = 16
FLAG_REFIT_MAT4_ROUTE = 32

SHAPE_DT = np.dtype([("type", "<u4"), ("p0", "<f4"), ("p1", "<f4"), ("p2", "<f4")])
CONTACT_DT = np.dtype([("a", "<u4"), ("b", "<u4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"),
                       ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("depth", "<f4"),
                       ("status", "<u4")])
MANIFOLD_DT = np.dtype([("a", "<u4"), ("b", "<u4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                        ("count", "<u4"), ("px", "<f4", (4,)), ("py", "<f4", (4,)), ("pz", "<f4", (4,)),
                        ("depth", "<f4", (4,))])
RAY_DT = np.dtype([("ox", "<f4"), ("oy", "<f4"), ("oz", "<f4"), ("dx", "<f4"), ("dy", "<f4"), ("dz", "<f4"),
                   ("tMax", "<f4"), ("world", "<u4")])
RAYHIT_DT = np.dtype([("body", "<u4"), ("t", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                      ("flags", "<u4")])
NO_HIT = 0xFFFFFFFF
FILTER_DT = np.dtype([("categoryBits", "<u4"), ("maskBits", "<u4"), ("groupIndex", "<i2"), ("reserved_", "<u2")])
SWEEP_DT = np.dtype([("hit", "<u4"), ("toi", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                     ("iterations", "<u4")])

# every symbol include/axcd.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "axcd_default_config", "axcd_device_count", "axcd_create", "axcd_destroy", "axcd_set_shapes",
    "axcd_set_transforms", "axcd_set_poses", "axcd_set_contact_sink", "axcd_refit", "axcd_broadphase", "axcd_narrowphase", "axcd_step", "axcd_step_async",
    "axcd_get_stats", "axcd_get_aabbs", "axcd_get_pairs", "axcd_get_pair_distances",
    "axcd_get_contacts", "axcd_error_string", "axcd_last_device_error",
    "axcd_build_manifolds", "axcd_get_manifolds", "axcd_query_aabbs", "axcd_raycast", "axcd_set_awake", "axcd_ccd_pairs", "axcd_ccd_pairs_angular", "axcd_pin_host_buffer", "axcd_unpin_host_buffer",
    "axcd_set_filters", "axcd_set_slab", "axcd_set_body_keys", "axcd_set_ghosts", "axcd_pack_ghosts",
    "axcd_set_ghosts_device", "axcd_nccl_unique_id", "axcd_slab_init", "axcd_slab_init_comm",
    "axcd_slab_step_async", "axcd_slab_step", "axcd_get_body_keys",
    "axcd_test_sort_pairs32", "axcd_test_sort_keys64", "axcd_test_sort_bench", "axcd_test_fp32_peak",
    "axcd_test_sort_morton", "axcd_test_sort_bench_morton",
]
SCENE_SYMBOLS = ["axcd_scene_generate", "axcd_scene_generate_worlds", "axcd_scene_rng_u32"]


class Config(C.Structure):
    _fields_ = [("maxBodies", C.c_uint32), ("maxPairs", C.c_uint32), ("maxContacts", C.c_uint32),
                ("maxHullVerts", C.c_uint32), ("numWorlds", C.c_uint32),
                ("aabbMargin", C.c_float), ("gjkMaxIters", C.c_uint32),
                ("epaMaxIters", C.c_uint32), ("epaMaxFaces", C.c_uint32), ("gjkTol", C.c_float),
                ("epaTol", C.c_float), ("flags", C.c_uint32), ("deviceOrdinal", C.c_int32),
                ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("numBodies", C.c_uint32), ("numPairs", C.c_uint32), ("numContacts", C.c_uint32),
                ("numPenetrating", C.c_uint32), ("gjkFailures", C.c_uint32),
                ("epaFailures", C.c_uint32), ("requiredPairs", C.c_uint32),
                ("requiredContacts", C.c_uint32), ("refitMs", C.c_float), ("sortMs", C.c_float),
                ("buildMs", C.c_float), ("pairMs", C.c_float), ("pairSortMs", C.c_float),
                ("gjkMs", C.c_float), ("epaMs", C.c_float), ("totalMs", C.c_float),
                ("broadphaseTime", C.c_float), ("narrowphaseTime", C.c_float),
                ("bytesMoved", C.c_uint64), ("kernelLaunches", C.c_uint32),
                ("contactPointCount", C.c_uint32),
                ("movedBodies", C.c_uint32), ("broadphaseSkipped", C.c_uint32), ("graphLaunched", C.c_uint32),
                ("ghostBodies", C.c_uint32), ("sortFallback", C.c_uint32), ("sortMaxBucket", C.c_uint32),
                ("exchangeMs", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SceneSpec(C.Structure):
    _fields_ = [("numBodies", C.c_uint32), ("fracBox", C.c_float), ("fracSphere", C.c_float),
                ("domain", C.c_float), ("sizeMin", C.c_float), ("sizeMax", C.c_float),
                ("hullVerts", C.c_uint32), ("seed", C.c_uint64)]


class AxcdError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__(f"{what}: error {code} ({error_string(code)}) {detail}".strip())


_lib = None
_scene = None


def load_library():
    """Loads libaxcd.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make -C {PKG_DIR}` "
                              "(there is no CPU fallback for the collision path)")
        lib = C.CDLL(LIB_PATH)
        for name in ABI_SYMBOLS:
            getattr(lib, name)
        lib.axcd_error_string.restype = C.c_char_p
        lib.axcd_last_device_error.restype = C.c_char_p
        lib.axcd_last_device_error.argtypes = [C.c_void_p]
        lib.axcd_destroy.restype = None
        lib.axcd_destroy.argtypes = [C.c_void_p]
        lib.axcd_default_config.restype = None
        for name in ("axcd_create", "axcd_set_shapes", "axcd_set_transforms", "axcd_refit",
                     "axcd_broadphase", "axcd_narrowphase", "axcd_step", "axcd_get_stats",
                     "axcd_get_aabbs", "axcd_get_pairs", "axcd_get_pair_distances",
                     "axcd_get_contacts", "axcd_test_sort_pairs32", "axcd_test_sort_keys64",
                     "axcd_test_sort_bench", "axcd_set_slab", "axcd_set_body_keys", "axcd_set_ghosts",
                     "axcd_pack_ghosts", "axcd_set_ghosts_device", "axcd_set_filters",
                     "axcd_build_manifolds", "axcd_get_manifolds", "axcd_query_aabbs", "axcd_raycast", "axcd_set_awake", "axcd_ccd_pairs", "axcd_ccd_pairs_angular", "axcd_pin_host_buffer", "axcd_unpin_host_buffer"):
            getattr(lib, name).restype = C.c_int32
        lib.axcd_set_shapes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                        C.c_uint32, C.c_void_p]
        lib.axcd_set_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.axcd_set_poses.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.axcd_set_poses.restype = C.c_int32
        lib.axcd_set_contact_sink.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.axcd_set_contact_sink.restype = C.c_int32
        for name in ("axcd_refit", "axcd_broadphase", "axcd_narrowphase"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.axcd_step.argtypes = [C.c_void_p, C.c_void_p]
        lib.axcd_step_async.restype = C.c_int32
        lib.axcd_step_async.argtypes = [C.c_void_p]
        for name in ("axcd_nccl_unique_id", "axcd_slab_init", "axcd_slab_init_comm", "axcd_slab_step_async",
                     "axcd_slab_step"):
            getattr(lib, name).restype = C.c_int32
        lib.axcd_nccl_unique_id.argtypes = [C.c_void_p]
        lib.axcd_slab_init.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.axcd_slab_init_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.axcd_slab_step_async.argtypes = [C.c_void_p]
        lib.axcd_slab_step.argtypes = [C.c_void_p, C.c_void_p]
        lib.axcd_get_body_keys.restype = C.c_int32
        lib.axcd_get_body_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.axcd_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        lib.axcd_get_aabbs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.axcd_get_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.axcd_get_pair_distances.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.axcd_get_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.axcd_build_manifolds.argtypes = [C.c_void_p]
        lib.axcd_get_manifolds.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.axcd_query_aabbs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                         C.c_void_p]
        lib.axcd_set_awake.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.axcd_pin_host_buffer.argtypes = [C.c_void_p, C.c_uint64]
        lib.axcd_unpin_host_buffer.argtypes = [C.c_void_p]
        lib.axcd_ccd_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.axcd_ccd_pairs_angular.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.axcd_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.axcd_test_sort_pairs32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                               C.c_uint32]
        lib.axcd_test_sort_keys64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.axcd_set_slab.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32]
        lib.axcd_set_body_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.axcd_set_ghosts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.axcd_set_filters.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.axcd_pack_ghosts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.axcd_set_ghosts_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.axcd_test_sort_bench.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.axcd_test_fp32_peak.restype = C.c_int32
        lib.axcd_test_fp32_peak.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = lib
    return _lib


def load_scene_library():
    global _scene
    if _scene is None:
        if not os.path.exists(SCENE_LIB_PATH):
            raise ImportError(f"{SCENE_LIB_PATH} is missing: run `make -C {PKG_DIR}`")
        lib = C.CDLL(SCENE_LIB_PATH)
        for name in SCENE_SYMBOLS:
            getattr(lib, name)
        lib.axcd_scene_generate.restype = C.c_int32
        lib.axcd_scene_generate_worlds.restype = C.c_int32
        _scene = lib
    return _scene


def error_string(code):
    try:
        return load_library().axcd_error_string(C.c_int32(code)).decode()
    except Exception:
        return "?"


def nccl_unique_id():
    """128-byte ncclUniqueId created by the library (rank 0 calls this and broadcasts the bytes)."""
    buf = (C.c_char * 128)()
    rc = load_library().axcd_nccl_unique_id(buf)
    if rc != 0:
        raise AxcdError(rc, "axcd_nccl_unique_id")
    return bytes(buf)


def default_config(**kw):
    cfg = Config()
    load_library().axcd_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------------------------------------
# scenes
# ------------------------------------------------------------------------------------------------
class Scene:
    """Host-side scene blobs: transforms (n,10) f32, shapes (n,) SHAPE_DT, hull (m,3) f32."""

    def __init__(self, xf, shapes, hull=None, world_id=None, num_worlds=1, name=""):
        self.xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(-1, 10)
        self.shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
        self.hull = np.ascontiguousarray(
            hull if hull is not None else np.zeros((0, 3)), dtype=np.float32).reshape(-1, 3)
        self.world_id = (np.ascontiguousarray(world_id, dtype=np.uint32)
                         if world_id is not None else None)
        self.num_worlds = num_worlds
        self.name = name

    @property
    def n(self):
        return self.xf.shape[0]


def generate_scene(n, seed, domain, frac_box=0.5, frac_sphere=0.5, size=(0.25, 0.5),
                   hull_verts=16, name=""):
    lib = load_scene_library()
    spec = SceneSpec(n, frac_box, frac_sphere, domain, size[0], size[1], hull_verts, seed)
    xf = np.zeros((n, 10), np.float32)
    shapes = np.zeros(n, SHAPE_DT)
    # worst case every body is a hull; allocate by expectation + slack and retry
    frac_hull = max(0.0, 1.0 - frac_box - frac_sphere)
    cap = int(n * hull_verts * min(1.0, frac_hull * 1.2 + 0.01)) + 16 * hull_verts
    while True:
        hull = np.zeros((cap, 3), np.float32)
        used = C.c_uint32(0)
        rc = lib.axcd_scene_generate(C.byref(spec), _ptr(xf), _ptr(shapes), _ptr(hull),
                                     C.c_uint32(cap), C.c_uint32(0), C.byref(used))
        if rc == 601:
            cap = n * hull_verts
            continue
        if rc != 0:
            raise RuntimeError(f"scene generation failed: {rc}")
        return Scene(xf, shapes, hull[:used.value].copy(), name=name)


def generate_worlds(num_worlds, bodies_per_world, seed0, domain, frac_box=0.5, frac_sphere=0.5,
                    size=(0.25, 0.5), hull_verts=16, name=""):
    lib = load_scene_library()
    n = num_worlds * bodies_per_world
    spec = SceneSpec(bodies_per_world, frac_box, frac_sphere, domain, size[0], size[1],
                     hull_verts, seed0)
    xf = np.zeros((n, 10), np.float32)
    shapes = np.zeros(n, SHAPE_DT)
    wid = np.zeros(n, np.uint32)
    frac_hull = max(0.0, 1.0 - frac_box - frac_sphere)
    cap = n * hull_verts if frac_hull > 0 else 1
    hull = np.zeros((cap, 3), np.float32)
    used = C.c_uint32(0)
    rc = lib.axcd_scene_generate_worlds(C.byref(spec), C.c_uint32(num_worlds), _ptr(xf),
                                        _ptr(shapes), _ptr(wid), _ptr(hull), C.c_uint32(cap),
                                        C.byref(used))
    if rc != 0:
        raise RuntimeError(f"scene generation failed: {rc}")
    return Scene(xf, shapes, hull[:used.value].copy(), wid, num_worlds, name=name)


# The named configurations of BASELINE.md / SURVEY.md 8(d)
def config_scene(name, scale=1.0):
    """scale < 1 shrinks body count at constant density (tests); 1.0 is the named size."""
    def dens(n, L):
        m = max(1, int(round(n * scale)))
        return m, float(L * (m / n) ** (1.0 / 3.0))
    if name == "C0":
        n, L = dens(1000, 10.0)
        return generate_scene(n, 1, L, name="C0")
    if name == "C1":
        n, L = dens(100_000, 46.4)
        return generate_scene(n, 2, L, name="C1")
    if name == "headline":
        n, L = dens(1_000_000, 100.0)
        return generate_scene(n, 3, L, name="headline")
    if name == "C2":
        n, L = dens(1_000_000, 100.0)
        return generate_scene(n, 4, L, frac_box=0.4, frac_sphere=0.3, name="C2")
    if name == "C3":
        w = max(1, int(round(4096 * scale)))
        return generate_worlds(w, 256, 1000, 6.35, name="C3")
    if name == "C4":
        n, L = dens(16_000_000, 252.0)
        return generate_scene(n, 5, L, name="C4")
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------
# the collision world (one context == one GPU)
# ------------------------------------------------------------------------------------------------
def pin_host_buffer(arr):
    """Page-locks a numpy array's memory (axcd_pin_host_buffer) so that copies from / to it are asynchronous and
    it can serve as a contact sink; undo with unpin_host_buffer before the array is freed."""
    rc = load_library().axcd_pin_host_buffer(C.c_void_p(arr.ctypes.data), C.c_uint64(arr.nbytes))
    if rc != 0:
        raise AxcdError(rc, "axcd_pin_host_buffer", "")


def unpin_host_buffer(arr):
    rc = load_library().axcd_unpin_host_buffer(C.c_void_p(arr.ctypes.data))
    if rc != 0:
        raise AxcdError(rc, "axcd_unpin_host_buffer", "")


class CollisionWorld:
    """Owns an AxcdContext.  Mirrors Broadphase::update/getPairCount and
    Narrowphase::detectCollisions/getContactCount."""

    def __init__(self, max_bodies, max_pairs=None, max_contacts=None, max_hull_verts=0,
                 num_worlds=1, device=0, stream=None, **cfg_kw):
        self._lib = load_library()
        max_pairs = max_pairs if max_pairs is not None else max(1024, 8 * max_bodies)
        max_contacts = max_contacts if max_contacts is not None else max_pairs
        self.cfg = default_config(maxBodies=max_bodies, maxPairs=max_pairs,
                                  maxContacts=max_contacts, maxHullVerts=max_hull_verts,
                                  numWorlds=num_worlds, deviceOrdinal=device,
                                  stream=C.c_void_p(stream) if stream else None, **cfg_kw)
        self._ctx = C.c_void_p()
        rc = self._lib.axcd_create(C.byref(self.cfg), C.byref(self._ctx))
        if rc != 0:
            raise AxcdError(rc, "axcd_create")
        self.n = 0

    @classmethod
    def for_scene(cls, scene, pairs_per_body=8, **kw):
        w = cls(scene.n, max_pairs=max(1024, pairs_per_body * scene.n),
                max_hull_verts=len(scene.hull), num_worlds=scene.num_worlds, **kw)
        w.set_shapes(scene.shapes, scene.hull, scene.world_id)
        w.set_transforms(scene.xf)
        return w

    def close(self):
        if self._ctx:
            self._lib.axcd_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise AxcdError(rc, what, self._lib.axcd_last_device_error(self._ctx).decode())

    def set_shapes(self, shapes, hull=None, world_id=None):
        shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
        hull = np.ascontiguousarray(hull if hull is not None else np.zeros((0, 3)),
                                    dtype=np.float32).reshape(-1, 3)
        wid = np.ascontiguousarray(world_id, dtype=np.uint32) if world_id is not None else None
        self.n = len(shapes)
        self._check(self._lib.axcd_set_shapes(self._ctx, _ptr(shapes), len(shapes), _ptr(hull),
                                              len(hull), _ptr(wid)), "axcd_set_shapes")

    def set_transforms(self, xf, stride=40):
        """xf: (n,10) float32 array (or any C-contiguous buffer of `stride`-byte records)."""
        if isinstance(xf, np.ndarray) and stride == 40:
            xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(-1, 10)
            n = xf.shape[0]
        else:
            n = xf.nbytes // stride
        self._check(self._lib.axcd_set_transforms(self._ctx, _ptr(xf), n, stride),
                    "axcd_set_transforms")

    def set_poses(self, poses, stride=28):
        """poses: (n,7) float32 array (position, rotation xyzw), or a buffer of `stride`-byte records; the scales
        stay as the last set_transforms left them."""
        if isinstance(poses, np.ndarray) and stride == 28:
            poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 7)
            n = poses.shape[0]
        else:
            n = poses.nbytes // stride
        self._check(self._lib.axcd_set_poses(self._ctx, _ptr(poses), n, stride), "axcd_set_poses")

    def set_contact_sink(self, host_ptr, capacity):
        """host_ptr: address of a page-locked buffer of >= maxContacts 40-byte records (None detaches); every
        narrowphase from now on also delivers its contacts there."""
        self._check(self._lib.axcd_set_contact_sink(self._ctx, C.c_void_p(host_ptr) if host_ptr else None, capacity),
                    "axcd_set_contact_sink")

    def set_poses_ptr(self, host_ptr, n, stride=28):
        self._check(self._lib.axcd_set_poses(self._ctx, C.c_void_p(host_ptr), n, stride), "axcd_set_poses")

    def set_transforms_ptr(self, host_ptr, n, stride=40):
        self._check(self._lib.axcd_set_transforms(self._ctx, C.c_void_p(host_ptr), n, stride),
                    "axcd_set_transforms")

    # Broadphase::update()
    def update(self):
        self._check(self._lib.axcd_refit(self._ctx), "axcd_refit")
        self._check(self._lib.axcd_broadphase(self._ctx), "axcd_broadphase")

    # Narrowphase::detectCollisions()
    def detect_collisions(self):
        self._check(self._lib.axcd_narrowphase(self._ctx), "axcd_narrowphase")

    def refit(self):
        self._check(self._lib.axcd_refit(self._ctx), "axcd_refit")

    def step(self):
        st = Stats()
        self._check(self._lib.axcd_step(self._ctx, C.byref(st)), "axcd_step")
        return st

    def step_async(self):
        """Enqueues the fused step (one CUDA graph launch once the configuration is stable) and returns."""
        self._check(self._lib.axcd_step_async(self._ctx), "axcd_step_async")

    def stats(self):
        st = Stats()
        self._check(self._lib.axcd_get_stats(self._ctx, C.byref(st)), "axcd_get_stats")
        return st

    def get_pair_count(self):
        return self.stats().numPairs

    def get_contact_count(self):
        return self.stats().numContacts

    def aabbs(self):
        out = np.zeros((self.n, 6), np.float32)
        self._check(self._lib.axcd_get_aabbs(self._ctx, _ptr(out), self.n), "axcd_get_aabbs")
        return out

    def pairs(self):
        cap = max(1, self.stats().numPairs)
        out = np.zeros((cap, 2), np.uint32)
        cnt = C.c_uint32(0)
        self._check(self._lib.axcd_get_pairs(self._ctx, _ptr(out), cap, C.byref(cnt)),
                    "axcd_get_pairs")
        return out[:cnt.value]

    def pair_distances(self):
        cap = max(1, self.stats().numPairs)
        out = np.zeros(cap, np.float32)
        cnt = C.c_uint32(0)
        self._check(self._lib.axcd_get_pair_distances(self._ctx, _ptr(out), cap, C.byref(cnt)),
                    "axcd_get_pair_distances")
        return out[:cnt.value]

    def contacts(self):
        cap = max(1, self.stats().numContacts)
        out = np.zeros(cap, CONTACT_DT)
        cnt = C.c_uint32(0)
        self._check(self._lib.axcd_get_contacts(self._ctx, _ptr(out), cap, C.byref(cnt)),
                    "axcd_get_contacts")
        return out[:cnt.value]

    def contacts_into(self, host_ptr, cap):
        cnt = C.c_uint32(0)
        self._check(self._lib.axcd_get_contacts(self._ctx, C.c_void_p(host_ptr), cap,
                                                C.byref(cnt)), "axcd_get_contacts")
        return cnt.value

    def build_manifolds(self):
        """Contact manifolds of the last narrowphase (asynchronous; once per narrowphase)."""
        self._check(self._lib.axcd_build_manifolds(self._ctx), "axcd_build_manifolds")

    def manifolds(self):
        """(records in contact order, total contact-point count)."""
        cap = max(1, self.stats().numContacts)
        out = np.zeros(cap, MANIFOLD_DT)
        cnt, pts = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.axcd_get_manifolds(self._ctx, _ptr(out), cap, C.byref(cnt), C.byref(pts)),
                    "axcd_get_manifolds")
        return out[:cnt.value], pts.value

    # ---- scene queries on the LBVH of the last broadphase ------------------------------------------
    def query_aabbs(self, boxes, query_world=None):
        """(query, body) index pairs, sorted, for query boxes (k,6) float32 = (min, max)."""
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        qw = np.ascontiguousarray(query_world, dtype=np.uint32) if query_world is not None else None
        cap = max(1024, 8 * len(boxes))
        while True:
            out = np.zeros((cap, 2), np.uint32)
            cnt = C.c_uint32(0)
            rc = self._lib.axcd_query_aabbs(self._ctx, _ptr(boxes), _ptr(qw), len(boxes), _ptr(out), cap,
                                            C.byref(cnt))
            if rc == 601 and cnt.value > cap:
                cap = cnt.value
                continue
            self._check(rc, "axcd_query_aabbs")
            return out[:cnt.value]

    def raycast(self, rays):
        """Closest hit per ray; rays is a RAY_DT array, the result a RAYHIT_DT array."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DT)
        out = np.zeros(max(1, len(rays)), RAYHIT_DT)
        self._check(self._lib.axcd_raycast(self._ctx, _ptr(rays), len(rays), _ptr(out)), "axcd_raycast")
        return out[:len(rays)]

    def ccd_pairs(self, pairs, displacement, rotation=None):
        """Time of impact of body pairs; displacement is (n,3) per body.  rotation (n,3): one rotation vector per
        body (angular velocity * dt) — the sweep then includes the turning (axcd_ccd_pairs_angular)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        disp = np.ascontiguousarray(displacement, dtype=np.float32).reshape(-1, 3)
        assert len(disp) == self.n
        out = np.zeros(max(1, len(pairs)), SWEEP_DT)
        if rotation is None:
            self._check(self._lib.axcd_ccd_pairs(self._ctx, _ptr(pairs), len(pairs), _ptr(disp), _ptr(out)),
                        "axcd_ccd_pairs")
        else:
            rot = np.ascontiguousarray(rotation, dtype=np.float32).reshape(-1, 3)
            assert len(rot) == self.n
            self._check(self._lib.axcd_ccd_pairs_angular(self._ctx, _ptr(pairs), len(pairs), _ptr(disp), _ptr(rot), _ptr(out)),
                        "axcd_ccd_pairs_angular")
        return out[:len(pairs)]

    def set_awake(self, awake):
        """awake: (n,) array, 0 = sleeping; None switches the rule off.  Sleeping-sleeping pairs are dropped."""
        if awake is None:
            self._check(self._lib.axcd_set_awake(self._ctx, None, 0), "axcd_set_awake")
            return
        a = np.ascontiguousarray(np.asarray(awake) != 0, dtype=np.uint8)
        self._check(self._lib.axcd_set_awake(self._ctx, _ptr(a), len(a)), "axcd_set_awake")

    def set_filters(self, filters):
        """filters: (n,3) array of (categoryBits, maskBits, groupIndex) or None to switch filtering off."""
        if filters is None:
            self._check(self._lib.axcd_set_filters(self._ctx, None, 0), "axcd_set_filters")
            return
        src = np.ascontiguousarray(filters, dtype=np.int64).reshape(-1, 3)
        f = np.zeros(len(src), FILTER_DT)   # the layout of gui::FilterInfo: uint32, uint32, int16 (+ padding)
        f["categoryBits"] = src[:, 0].astype(np.uint32)
        f["maskBits"] = src[:, 1].astype(np.uint32)
        f["groupIndex"] = src[:, 2].astype(np.int16)
        self._check(self._lib.axcd_set_filters(self._ctx, _ptr(f), len(f)), "axcd_set_filters")

    # ---- x-slab mode (one scene across several GPUs) -------------------------------------------
    def set_slab(self, x_lo, x_hi, enable=True):
        self._check(self._lib.axcd_set_slab(self._ctx, float(x_lo), float(x_hi), 1 if enable else 0),
                    "axcd_set_slab")

    def set_body_keys(self, keys, first=0):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        self._check(self._lib.axcd_set_body_keys(self._ctx, _ptr(keys), first, len(keys)),
                    "axcd_set_body_keys")

    def set_ghosts(self, n_owned, xf, shapes, keys):
        xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(-1, 10)
        shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DT)
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        self._check(self._lib.axcd_set_ghosts(self._ctx, n_owned, len(xf), _ptr(xf), _ptr(shapes),
                                              _ptr(keys)), "axcd_set_ghosts")
        self.n = n_owned + len(xf)

    def pack_ghosts(self, edges, num_ranks, my_rank):
        """Device-side ghost selection.  Returns (device pointers, counts) per destination rank."""
        e = np.ascontiguousarray(edges, dtype=np.float32)
        ptrs = (C.c_void_p * num_ranks)()
        counts = np.zeros(num_ranks, np.uint32)
        self._check(self._lib.axcd_pack_ghosts(self._ctx, _ptr(e), num_ranks, my_rank, ptrs, _ptr(counts)),
                    "axcd_pack_ghosts")
        return [p or 0 for p in ptrs], counts

    # the exchange inside the library (NCCL bound at run time)
    def slab_init(self, unique_id, rank, num_ranks, edges):
        """unique_id: 128 bytes from nccl_unique_id() of rank 0 (broadcast by the host); collective."""
        e = np.ascontiguousarray(edges, dtype=np.float32)
        assert len(e) == num_ranks + 1
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.axcd_slab_init(self._ctx, buf, rank, num_ranks, _ptr(e)), "axcd_slab_init")

    def slab_step(self):
        st = Stats()
        self._check(self._lib.axcd_slab_step(self._ctx, C.byref(st)), "axcd_slab_step")
        self.n = st.numBodies
        return st

    def slab_step_async(self):
        self._check(self._lib.axcd_slab_step_async(self._ctx), "axcd_slab_step_async")

    def body_keys(self):
        """Global ids of the local bodies (owned, then this step's ghosts)."""
        out = np.zeros(max(1, self.n), np.uint32)
        self._check(self._lib.axcd_get_body_keys(self._ctx, _ptr(out), len(out)), "axcd_get_body_keys")
        return out[:self.n]

    def set_ghosts_device(self, n_owned, n_ghosts, dev_ptr):
        self._check(self._lib.axcd_set_ghosts_device(self._ctx, n_owned, n_ghosts, C.c_void_p(dev_ptr)),
                    "axcd_set_ghosts_device")
        self.n = n_owned + n_ghosts

    # test hooks for the device primitives
    def test_sort_pairs32(self, keys, vals, key_bits=32):
        keys = np.ascontiguousarray(keys, dtype=np.uint32).copy()
        vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        self._check(self._lib.axcd_test_sort_pairs32(self._ctx, _ptr(keys), _ptr(vals),
                                                     len(keys), key_bits), "sort_pairs32")
        return keys, vals

    def sort_bench(self, n, key_bits=32, iters=10):
        """Average ms of one device-resident (key,value) radix sort of n pseudo-random keys."""
        ms = C.c_float(0)
        self._check(self._lib.axcd_test_sort_bench(self._ctx, n, key_bits, iters, C.byref(ms)), "sort_bench")
        return ms.value

    def test_sort_morton(self, keys, key_bits=32, mode=1):
        """The step's Morton sort on caller keys (identity payload).  Returns (sorted keys, permutation, fallback)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        ok, ov = np.zeros(len(keys), np.uint32), np.zeros(len(keys), np.uint32)
        fb = C.c_uint32(0)
        self._lib.axcd_test_sort_morton.restype = C.c_int32
        self._check(self._lib.axcd_test_sort_morton(self._ctx, _ptr(keys), C.c_uint32(len(keys)), C.c_uint32(key_bits),
                                                    C.c_uint32(mode), _ptr(ok), _ptr(ov), C.byref(fb)), "sort_morton")
        return ok, ov, fb.value

    def sort_bench_morton(self, n, key_bits=32, iters=10, mode=1):
        ms = C.c_float(0)
        self._lib.axcd_test_sort_bench_morton.restype = C.c_int32
        self._check(self._lib.axcd_test_sort_bench_morton(self._ctx, C.c_uint32(n), C.c_uint32(key_bits), C.c_uint32(iters),
                                                          C.c_uint32(mode), C.byref(ms)), "sort_bench_morton")
        return ms.value

    def fp32_peak(self, iters=4096):
        """Measured FP32 FMA-chain peak of the device in TFLOP/s."""
        tf = C.c_float(0)
        self._check(self._lib.axcd_test_fp32_peak(self._ctx, iters, C.byref(tf)), "fp32_peak")
        return tf.value

    def test_sort_keys64(self, keys, key_bits=64):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        self._check(self._lib.axcd_test_sort_keys64(self._ctx, _ptr(keys), len(keys), key_bits),
                    "sort_keys64")
        return keys
