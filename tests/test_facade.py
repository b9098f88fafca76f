"""The C++20 façade (include/axiom/collision/collision_world.hpp) compiles against the C ABI and
behaves like the documented Broadphase/Narrowphase call shape (reference: CLAUDE.md:162-178)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "axiom-physics-engine_b200")
EXE = os.path.join(PKG, "facade_smoke")


def _build():
    subprocess.check_call(["make", "-C", PKG, "facade_smoke", "bench_step"], stdout=subprocess.DEVNULL)


def test_facade_links_and_refuses_without_device():
    import torch
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        # no device: create() must fail with the GPU-init code, never fall back to a CPU path
        assert r.returncode == 77, r.stdout + r.stderr
        assert "500" in r.stdout


@pytest.mark.gpu
def test_facade_c0_counts():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.split()
    assert out[:2] == ["2875", "1381"]   # C0 fixture: candidate pairs, contacts
    assert int(out[2]) >= 1381           # contact points over all manifolds


@pytest.mark.gpu
def test_cpp_host_at_c1_size():
    """The C++20 host (tests/cpp/bench_step.cpp) through the façade at config C1: 100 k bodies, the same
    counts as the ctypes path."""
    _build()
    exe = os.path.join(os.path.dirname(EXE), "bench_step")
    r = subprocess.run([exe, "100000", "46.4", "2", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n, pairs, contacts, dev_ms, wall_ms = r.stdout.split()
    import axcd
    w = axcd.CollisionWorld.for_scene(axcd.config_scene("C1"))
    st = w.step()
    assert (int(n), int(pairs), int(contacts)) == (100000, st.numPairs, st.numContacts)
    assert 0 < float(dev_ms) < float(wall_ms)
    w.close()
