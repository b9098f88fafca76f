"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/axcd.h declares; error strings are the reference's (src/core/error_code.cpp:5-62)."""
import ctypes as C
import os
import re

import pytest

import axcd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"AXCD_API\s+[\w\s\*]+?\b(axcd_\w+)\s*\(", txt)))


def test_header_symbols_are_exported():
    lib = axcd.load_library()
    names = _declared("axcd.h")
    assert sorted(names) == sorted(axcd.ABI_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_scene_header_symbols_are_exported():
    lib = axcd.load_scene_library()
    names = _declared("axcd_scene.h")
    assert sorted(names) == sorted(axcd.SCENE_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_error_strings_match_reference():
    # src/core/error_code.cpp:29-34 and tests/core/error_code_test.cpp:28-34
    assert axcd.error_string(0) == "Success"
    assert axcd.error_string(300) == "Invalid collision shape"
    assert axcd.error_string(301) == "GJK algorithm failed to converge"
    assert axcd.error_string(302) == "EPA algorithm failed to converge"
    assert axcd.error_string(505) == "GPU operation failed"
    assert axcd.error_string(600) == "Invalid parameter"
    assert axcd.error_string(601) == "Value out of range"
    assert axcd.error_string(12345) == "Unknown error"


def test_struct_layouts():
    assert C.sizeof(axcd.Config) == 64
    assert axcd.SHAPE_DT.itemsize == 16
    assert axcd.CONTACT_DT.itemsize == 40


def test_default_config():
    cfg = axcd.default_config()
    assert (cfg.gjkMaxIters, cfg.epaMaxIters, cfg.epaMaxFaces) == (32, 32, 64)
    assert cfg.gjkTol == pytest.approx(1e-6) and cfg.epaTol == pytest.approx(1e-4)
    assert cfg.numWorlds == 1 and cfg.aabbMargin == 0.0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(axcd.AxcdError) as e:
        axcd.CollisionWorld(16)
    assert e.value.code == 500


def test_bad_config_rejected():
    lib = axcd.load_library()
    cfg = axcd.default_config(maxBodies=0)
    ctx = C.c_void_p()
    assert lib.axcd_create(C.byref(cfg), C.byref(ctx)) == 600
    assert lib.axcd_create(None, C.byref(ctx)) == 202


def test_scene_generator_is_deterministic_and_matches_rng():
    import numpy as np
    a, b = axcd.config_scene("C0"), axcd.config_scene("C0")
    assert np.array_equal(a.xf, b.xf) and np.array_equal(a.shapes, b.shapes)
    out = np.zeros(5, np.uint32)
    axcd.load_scene_library().axcd_scene_rng_u32(C.c_uint64(42), C.c_uint32(5), out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == [1870769882, 2612922264, 273981832, 3501727647, 2781298041]
    q = a.xf[:, 3:7].astype(np.float64)
    np.testing.assert_allclose((q * q).sum(1), 1.0, atol=1e-5)
    assert (a.xf[:, :3] >= 0).all() and (a.xf[:, :3] <= 10).all()
    assert set(np.unique(a.shapes["type"])) == {0, 1}
    c2 = axcd.config_scene("C2", scale=0.002)
    assert set(np.unique(c2.shapes["type"])) == {0, 1, 4}
    nh = (c2.shapes["type"] == 4).sum()
    assert len(c2.hull) == 16 * nh


def test_struct_sizes_match_the_header(tmp_path):
    """sizeof() of every record in include/axcd.h, as a C compiler sees it, equals the size of the
    ctypes / numpy mirror the host side uses (guards against silent ABI drift)."""
    import subprocess
    import numpy as np
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "axcd.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(AxcdShape),sizeof(AxcdContact),sizeof(AxcdConfig),sizeof(AxcdStats),sizeof(AxcdFilter),'
                   'sizeof(AxcdManifold),sizeof(AxcdRay),sizeof(AxcdRayHit),sizeof(AxcdSweep),sizeof(void*));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [axcd.SHAPE_DT.itemsize, axcd.CONTACT_DT.itemsize, C.sizeof(axcd.Config), C.sizeof(axcd.Stats), 12,
            axcd.MANIFOLD_DT.itemsize, axcd.RAY_DT.itemsize, axcd.RAYHIT_DT.itemsize, axcd.SWEEP_DT.itemsize, 8]
    assert got == want, (got, want)
    import oracle_lib as O   # the oracle mirrors the same records
    assert O.MANIFOLD_DT == axcd.MANIFOLD_DT and O.RAY_DT == axcd.RAY_DT and O.RAYHIT_DT == axcd.RAYHIT_DT
    assert O.SWEEP_DT == axcd.SWEEP_DT and O.CONTACT_DT == axcd.CONTACT_DT
